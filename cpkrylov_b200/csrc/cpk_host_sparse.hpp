// cpk_host_sparse.hpp -- MATLAB CSC -> rows, 2x2 block assembly, SELL-32-sigma builder
// Host-side part of libcpk_b200 (included by cpk_host.cu only; uses its fail() and the
// constants of cpk_device.cuh).
#pragma once

// ===========================================================================
// host sparse helpers
// ===========================================================================
struct HCsr {
    int nrows = 0, ncols = 0;
    std::vector<int64_t> ptr;
    std::vector<int> col;
    std::vector<double> val;
    int64_t nnz() const { return ptr.empty() ? 0 : ptr.back(); }
    int len(int r) const { return (int)(ptr[r + 1] - ptr[r]); }
};

static bool csc_ok(const cpk_csc *A)
{
    return A && A->nrows >= 0 && A->ncols >= 0 && A->colptr && (A->colptr[A->ncols] == 0 || (A->rowind && A->val));
}

// CSR of A (rows of A), from MATLAB CSC: a counting-sort transpose; column
// indices inside a row come out ascending.
static HCsr csr_from_csc(const cpk_csc &A)
{
    HCsr R;
    R.nrows = (int)A.nrows; R.ncols = (int)A.ncols;
    const int64_t nnz = A.colptr[A.ncols];
    R.ptr.assign((size_t)R.nrows + 1, 0);
    for (int64_t k = 0; k < nnz; ++k) R.ptr[A.rowind[k] + 1]++;
    for (int i = 0; i < R.nrows; ++i) R.ptr[i + 1] += R.ptr[i];
    R.col.resize(nnz); R.val.resize(nnz);
    std::vector<int64_t> next(R.ptr.begin(), R.ptr.end() - 1);
    for (int64_t j = 0; j < A.ncols; ++j)
        for (int64_t k = A.colptr[j]; k < A.colptr[j + 1]; ++k) {
            const int64_t p = next[A.rowind[k]]++;
            R.col[p] = (int)j; R.val[p] = A.val[k];
        }
    return R;
}
// CSR of A' : the CSC arrays read as rows (sorted by index inside each row).
static HCsr csr_of_transpose(const cpk_csc &A)
{
    HCsr R;
    R.nrows = (int)A.ncols; R.ncols = (int)A.nrows;
    const int64_t nnz = A.colptr[A.ncols];
    R.ptr.assign(A.colptr, A.colptr + A.ncols + 1);
    R.col.resize(nnz); R.val.resize(nnz);
    for (int j = 0; j < R.nrows; ++j) {
        const int64_t b = R.ptr[j], e = R.ptr[j + 1];
        std::vector<std::pair<int, double>> tmp;
        bool sorted = true;
        for (int64_t k = b; k < e; ++k) {
            if (k > b && A.rowind[k] < A.rowind[k - 1]) sorted = false;
            R.col[k] = (int)A.rowind[k]; R.val[k] = A.val[k];
        }
        if (!sorted) {
            tmp.reserve(e - b);
            for (int64_t k = b; k < e; ++k) tmp.emplace_back(R.col[k], R.val[k]);
            std::sort(tmp.begin(), tmp.end(), [](auto &x, auto &y) { return x.first < y.first; });
            for (int64_t k = b; k < e; ++k) { R.col[k] = tmp[k - b].first; R.val[k] = tmp[k - b].second; }
        }
    }
    return R;
}

// [A 0; 0 C] or general 2x2 block assembly by rows.  Blocks may be null (zero).
static HCsr block2x2(const HCsr *A11, const HCsr *A12, const HCsr *A21, const HCsr *A22, int n1, int n2)
{
    HCsr R;
    R.nrows = n1 + n2; R.ncols = n1 + n2;
    R.ptr.assign((size_t)R.nrows + 1, 0);
    auto rowlen = [](const HCsr *M, int r) { return M ? M->len(r) : 0; };
    for (int i = 0; i < n1; ++i) R.ptr[i + 1] = R.ptr[i] + rowlen(A11, i) + rowlen(A12, i);
    for (int i = 0; i < n2; ++i) R.ptr[n1 + i + 1] = R.ptr[n1 + i] + rowlen(A21, i) + rowlen(A22, i);
    R.col.resize(R.ptr.back()); R.val.resize(R.ptr.back());
    auto put = [&](const HCsr *M, int r, int off, int64_t &p) {
        if (!M) return;
        for (int64_t k = M->ptr[r]; k < M->ptr[r + 1]; ++k) { R.col[p] = M->col[k] + off; R.val[p] = M->val[k]; ++p; }
    };
    for (int i = 0; i < n1; ++i) { int64_t p = R.ptr[i]; put(A11, i, 0, p); put(A12, i, n1, p); }
    for (int i = 0; i < n2; ++i) { int64_t p = R.ptr[n1 + i]; put(A21, i, 0, p); put(A22, i, n1, p); }
    return R;
}

// ---------------------------------------------------------------------------
// SELL-32-sigma builder.  Rows are stably sorted by length inside windows of
// `sigma` rows (keeps x-locality, removes padding), cut into slices of 32.
// Rows much longer than the mean go to the CSR "long" list (warp per row).
// ---------------------------------------------------------------------------
struct HSell {
    int nrows = 0, ncols = 0, nslices = 0;
    std::vector<int> sptr, col, rowmap;
    std::vector<double> val;
    std::vector<int> lrow, lptr, lcol;
    std::vector<double> lval;
    int64_t nnz = 0;
};

static int sell_sigma()
{
    static int s = [] { const char *e = getenv("CPK_SELL_SIGMA"); int v = e ? atoi(e) : 4096; return std::max(32, v / 32 * 32); }();
    return s;
}

// Appends the rows [r0, r1) of A (column offset coff, row offset roff) as new slices.
static void sell_append(HSell &S, const HCsr &A, int r0, int r1, int coff, int roff)
{
    const int64_t nnz = A.ptr[r1] - A.ptr[r0];
    const double mean = (r1 > r0) ? (double)nnz / (r1 - r0) : 0.0;
    const int long_thr = (int)std::max(128.0, 8.0 * mean);
    std::vector<int> rows;
    rows.reserve(r1 - r0);
    for (int r = r0; r < r1; ++r) {
        if (A.len(r) > long_thr) {
            if (S.lptr.empty()) S.lptr.push_back(0);
            S.lrow.push_back(r + roff);
            for (int64_t k = A.ptr[r]; k < A.ptr[r + 1]; ++k) { S.lcol.push_back(A.col[k] + coff); S.lval.push_back(A.val[k]); }
            S.lptr.push_back((int)S.lcol.size());
        } else rows.push_back(r);
    }
    const int sigma = sell_sigma();
    if (S.sptr.empty()) S.sptr.push_back(0);
    for (size_t w0 = 0; w0 < rows.size(); w0 += sigma) {
        const size_t w1 = std::min(rows.size(), w0 + (size_t)sigma);
        std::stable_sort(rows.begin() + w0, rows.begin() + w1, [&](int a, int b) { return A.len(a) > A.len(b); });
        for (size_t s0 = w0; s0 < w1; s0 += 32) {
            const size_t s1 = std::min(w1, s0 + 32);
            int width = 0;
            for (size_t t = s0; t < s1; ++t) width = std::max(width, A.len(rows[t]));
            const size_t base = S.col.size();
            S.col.resize(base + (size_t)width * 32, 0);
            S.val.resize(base + (size_t)width * 32, 0.0);
            for (int lane = 0; lane < 32; ++lane) {
                const size_t t = s0 + lane;
                if (t < s1) {
                    const int r = rows[t];
                    S.rowmap.push_back(r + roff);
                    const int len = A.len(r);
                    int lastc = 0;
                    for (int j = 0; j < width; ++j) {
                        if (j < len) {
                            lastc = A.col[A.ptr[r] + j] + coff;
                            S.col[base + (size_t)j * 32 + lane] = lastc;
                            S.val[base + (size_t)j * 32 + lane] = A.val[A.ptr[r] + j];
                        } else S.col[base + (size_t)j * 32 + lane] = lastc;
                    }
                } else S.rowmap.push_back(-1);
            }
            S.sptr.push_back((int)S.col.size());
            S.nslices++;
        }
    }
    S.nnz += nnz;
}

static HSell build_sell(const HCsr &A)
{
    HSell S;
    S.nrows = A.nrows; S.ncols = A.ncols;
    sell_append(S, A, 0, A.nrows, 0, 0);
    if (S.sptr.empty()) S.sptr.push_back(0);
    if (S.lptr.empty()) S.lptr.push_back(0);
    return S;
}

