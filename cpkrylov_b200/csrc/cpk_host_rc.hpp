// cpk_host_rc.hpp -- row-class layout (DevRc): builder and upload
// Host-side part of libcpk_b200 (included by cpk_host.cu only; uses its fail(), DevArena and
// the constants of cpk_device.cuh).
#pragma once

// ---------------------------------------------------------------------------
// One pass (level) = its rows sorted by entry count; class w = the rows with exactly w
// entries, in ascending order of their row index, stored as a dense rectangle.  Rows with
// more than `maxw` entries go to a CSR list (one warp per row).  See cpk_device.cuh.
// ---------------------------------------------------------------------------
struct HRc {
    int nlev = 0, nfwd_lev = 0;
    std::vector<RcPiece> pieces;
    std::vector<int> levp;                  // [nlev+1]
    std::vector<int> levb;                  // [nlev] batches per level
    std::vector<int> rowmap, col;
    std::vector<double> d, val;
    std::vector<int> lptr, lcol;
    std::vector<double> lval;
    bool with_d = false;
    int64_t nrows = 0, nnz = 0, nlong = 0;
};

struct RcRowIn {
    int code;       // row code (index | flags << 28)
    int level;
    int src;        // handle passed to entries(src)
    int len;
    double d;
};

// Appends the rows of consecutive levels (level numbers relative to `level_base`: the first
// appended level becomes level R.nlev).  `entries(src)` returns the (column code, value) list
// of a row in accumulation order.  Levels without rows are kept (as empty levels) only when
// `keep_empty`; normally the caller passes dense level numbers.
template <class Entries>
static void rc_append(HRc &R, std::vector<RcRowIn> rows, int nlevels, Entries entries, int maxw)
{
    if (R.levp.empty()) R.levp.push_back(0);
    if (R.lptr.empty()) R.lptr.push_back(0);
    // class of a row = its width after padding: exact up to 12 entries, then 16 / 24 / 32 (a few
    // classes instead of one per length: every class costs each warp a pass over the piece table);
    // padding entries repeat the row's first column with value 0
    auto bw = [](int len) { return len <= 12 ? len : len <= 16 ? 16 : len <= 24 ? 24 : len <= 32 ? 32 : ((len + 15) / 16) * 16; };
    auto cls = [&](const RcRowIn &r) { return r.len <= maxw ? bw(r.len) : 1 << 30; };
    std::stable_sort(rows.begin(), rows.end(), [&](const RcRowIn &a, const RcRowIn &b) {
        if (a.level != b.level) return a.level < b.level;
        const int ca = cls(a), cb = cls(b);
        if (ca != cb) return ca < cb;
        return (a.code & RC_IDX_MASK) < (b.code & RC_IDX_MASK);
    });
    size_t i0 = 0;
    for (int lev = 0; lev < nlevels; ++lev) {
        int cum = 0;
        while (i0 < rows.size() && rows[i0].level == lev) {
            const int c = cls(rows[i0]);
            size_t i1 = i0;
            while (i1 < rows.size() && rows[i1].level == lev && cls(rows[i1]) == c) ++i1;
            const int cnt = (int)(i1 - i0);
            RcPiece pc;
            if (c < (1 << 30)) {
                const int ng = (cnt + 31) / 32, stride = ng * 32;
                pc.wn = c | (ng << 8); pc.cum = cum; pc.row_off = (int)R.rowmap.size(); pc.ent_off = (int)R.col.size();
                cum += (ng + rc_batch_of(c) - 1) / rc_batch_of(c) * rc_cost_of(c);
                R.rowmap.resize(R.rowmap.size() + stride, -1);
                if (R.with_d) R.d.resize(R.rowmap.size(), 1.0);
                R.col.resize(R.col.size() + (size_t)c * stride, 0);
                R.val.resize(R.val.size() + (size_t)c * stride, 0.0);
                for (int k = 0; k < cnt; ++k) {
                    const RcRowIn &rw = rows[i0 + k];
                    R.rowmap[pc.row_off + k] = rw.code;
                    if (R.with_d) R.d[pc.row_off + k] = rw.d;
                    const auto er = entries(rw.src);
                    for (int j = 0; j < c; ++j) {
                        const bool live = j < (int)er.size();
                        R.col[(size_t)pc.ent_off + (size_t)j * stride + k] = live ? er[j].first : er[0].first;
                        R.val[(size_t)pc.ent_off + (size_t)j * stride + k] = live ? er[j].second : 0.0;
                    }
                }
            } else {
                pc.wn = kRcLongW | (cnt << 8); pc.cum = cum; pc.row_off = (int)R.rowmap.size(); pc.ent_off = (int)R.lptr.size() - 1;
                cum += cnt * rc_cost_of(kRcLongW);
                for (int k = 0; k < cnt; ++k) {
                    const RcRowIn &rw = rows[i0 + k];
                    R.rowmap.push_back(rw.code);
                    if (R.with_d) R.d.push_back(rw.d);
                    const auto er = entries(rw.src);
                    for (auto &x : er) { R.lcol.push_back(x.first); R.lval.push_back(x.second); }
                    R.lptr.push_back((int)R.lcol.size());
                }
                // keep the class rectangles that follow aligned to 32 positions
                while (R.rowmap.size() % 32) { R.rowmap.push_back(-1); if (R.with_d) R.d.push_back(1.0); }
                R.nlong += cnt;
            }
            R.pieces.push_back(pc);
            R.nrows += cnt;
            for (size_t t = i0; t < i1; ++t) R.nnz += rows[t].len;
            i0 = i1;
        }
        R.levp.push_back((int)R.pieces.size());
        R.levb.push_back(cum);
        R.nlev++;
    }
}

// a plain matrix (rows of a CSR) as ONE level: row code = row index, column code = column index
static HRc build_rc(const HCsr &A, int maxw)
{
    HRc R;
    std::vector<RcRowIn> rows((size_t)A.nrows);
    for (int r = 0; r < A.nrows; ++r) rows[r] = RcRowIn{r, 0, r, A.len(r), 0.0};
    rc_append(R, std::move(rows), 1, [&](int r) {
        std::vector<std::pair<int, double>> er((size_t)A.len(r));
        for (int64_t k = A.ptr[r]; k < A.ptr[r + 1]; ++k) er[k - A.ptr[r]] = {A.col[k], A.val[k]};
        return er;
    }, maxw);
    return R;
}

static bool rc_fits(const HRc &R)
{
    for (auto &pc : R.pieces) if ((pc.wn >> 8) >= (1 << 22)) return false;
    return R.nlev <= kRcMaxLev && R.pieces.size() < 32000 && R.col.size() < (size_t)INT32_MAX && R.lcol.size() < (size_t)INT32_MAX &&
           R.rowmap.size() < (size_t)INT32_MAX;
}

static cudaError_t upload_rc(DevArena &ar, const HRc &h, DevRc &d)
{
    cudaError_t e;
    memset(&d, 0, sizeof d);
    d.nlev = h.nlev; d.npieces = (int)h.pieces.size(); d.nfwd_lev = h.nfwd_lev;
    for (int l = 0; l <= h.nlev && l < kRcMaxLev + 2; ++l) d.levp[l] = (short)h.levp[l];
    for (int l = 0; l < h.nlev && l < kRcMaxLev + 1; ++l) d.levb[l] = h.levb[l];
    for (int p = 0; p < d.npieces && p < kRcInline; ++p) d.inl[p] = h.pieces[p];
    if ((e = ar.upload(&d.pieces, h.pieces)) != cudaSuccess) return e;
    if ((e = ar.upload(&d.rowmap, h.rowmap)) != cudaSuccess) return e;
    if (h.with_d) { if ((e = ar.upload(&d.d, h.d)) != cudaSuccess) return e; } else d.d = nullptr;
    if ((e = ar.upload(&d.col, h.col)) != cudaSuccess) return e;
    if ((e = ar.upload(&d.val, h.val)) != cudaSuccess) return e;
    if ((e = ar.upload(&d.lptr, h.lptr)) != cudaSuccess) return e;
    if ((e = ar.upload(&d.lcol, h.lcol)) != cudaSuccess) return e;
    if ((e = ar.upload(&d.lval, h.lval)) != cudaSuccess) return e;
    return cudaSuccess;
}
