// cpk_host_sqd.hpp -- factorization plan of the device-side LDL' (symbolic analysis, level schedule)
// Host-side part of libcpk_b200 (included by cpk_host.cu only; uses its fail() and the
// constants of cpk_device.cuh).
#pragma once

// ===========================================================================
// Numeric LDL' on the device for symmetric quasi-definite K_P with a STATIC
// permutation (a sequence of interior-point systems keeps
// its pattern, opLDL2.m:81-82 re-assembles and re-factors each of them from scratch).
// Quasi-definite matrices are strongly factorizable: every symmetric permutation has
// an LDL' factorization with a diagonal D, so the permutation is chosen once for
// fill and the numeric work is a fixed dependency graph:
//     d_k  = a_kk - sum_j L_kj^2 d_j
//     L_ik = (a_ik - sum_j L_ij d_j L_kj) / d_k          (j < k, both factors nonzero)
// The host compiles that graph once per pattern ("plan"): every entry of L and D is a
// node with its list of triple products, nodes are sorted into dependency levels, and
// one CTA evaluates level after level.  The plan also records where every value sits
// in the compact-walk stream, so a refactorization rewrites the operator in place.
// ===========================================================================
struct SqdPlan {
    int N = 0;
    int64_t nnzL = 0, ne = 0, nops = 0, nvals_in = 0;
    int nlev = 0;
    std::vector<int64_t> colptr, rowind;    // strict lower triangle of L, CSC (symbolic pattern incl. fill)
    int64_t nnzA = 0, nnzB = 0, nnzC = 0;   // lengths of the value arrays a refactorization must bring
    // device copies
    DevArena ar;
    double *d_fval = nullptr, *d_vals_in = nullptr;
    const int *d_asrc = nullptr, *d_dk = nullptr, *d_exec = nullptr, *d_levptr = nullptr, *d_opptr = nullptr, *d_ops = nullptr;
    const int *d_spos = nullptr, *d_ssrc = nullptr;
    int64_t ns = 0;
};

struct SqdHost {
    std::vector<int> asrc, dk, exec, levptr, opptr, ops;
};

// symbolic analysis + plan.  K_P = [A B'; B C] (blocks as handed to cpk_ldl2_create), perm[k] =
// original index of row k of the permuted matrix.
static int sqd_plan(const cpk_csc *A, const cpk_csc *B, const cpk_csc *C, const int64_t *perm, int N, int nA,
                    SqdPlan *P, SqdHost *H)
{
    std::vector<int> invp(N, -1);
    for (int k = 0; k < N; ++k) {
        if (perm[k] < 0 || perm[k] >= N || invp[perm[k]] >= 0) return fail(CPK_ERR_ARG, "perm is not a permutation of 0..N-1");
        invp[perm[k]] = k;
    }
    // lower triangle of P' K_P P by columns: (row, index into the concatenated values [A | B | C])
    std::vector<std::vector<std::pair<int, int>>> low(N);
    std::vector<int> diagsrc(N, -1);
    int64_t base = 0;
    int64_t n_upper = 0, n_lower = 0;       // A and C must bring both triangles (as MATLAB stores symmetric matrices)
    auto add = [&](int64_t r, int64_t c, int64_t src, bool mirror_ok) {
        int pr = invp[r], pc = invp[c];
        if (!mirror_ok) { n_upper += pr < pc; n_lower += pr > pc; }
        if (pr < pc) { if (!mirror_ok) return; std::swap(pr, pc); }       // the mirror entry of A / C carries this one; B has one copy
        if (pr == pc) { if (diagsrc[pc] < 0) diagsrc[pc] = (int)src; }
        else low[pc].emplace_back(pr, (int)src);
    };
    for (int64_t j = 0; j < A->ncols; ++j) for (int64_t k = A->colptr[j]; k < A->colptr[j + 1]; ++k) add(A->rowind[k], j, base + k, false);
    base += A->colptr[A->ncols];
    for (int64_t j = 0; j < B->ncols; ++j) for (int64_t k = B->colptr[j]; k < B->colptr[j + 1]; ++k) add(nA + B->rowind[k], j, base + k, true);
    base += B->colptr[B->ncols];
    for (int64_t j = 0; j < C->ncols; ++j) for (int64_t k = C->colptr[j]; k < C->colptr[j + 1]; ++k) add(nA + C->rowind[k], nA + j, base + k, false);
    base += C->colptr[C->ncols];
    if (n_upper != n_lower) return fail(CPK_ERR_ARG, "cpk_ldl2_create_sqd: A and C must be stored with both triangles (pattern is not symmetric)");
    P->nvals_in = base;
    P->nnzA = A->colptr[A->ncols]; P->nnzB = B->colptr[B->ncols]; P->nnzC = C->colptr[C->ncols];
    for (int k = 0; k < N; ++k) if (diagsrc[k] < 0) return fail(CPK_ERR_ARG, "K_P has a structurally zero diagonal entry (row %lld): not quasi-definite", (long long)perm[k]);
    // column structures with fill (elimination tree: parent = first off-diagonal row)
    std::vector<std::vector<int>> st(N), children(N);
    std::vector<int> mark(N, -1);
    for (int k = 0; k < N; ++k) {
        std::vector<int> &sk = st[k];
        for (auto &e : low[k]) if (mark[e.first] != k) { mark[e.first] = k; sk.push_back(e.first); }
        for (int c : children[k]) for (int i : st[c]) if (i != k && mark[i] != k) { mark[i] = k; sk.push_back(i); }
        std::sort(sk.begin(), sk.end());
        if (!sk.empty()) children[sk[0]].push_back(k);
    }
    P->N = N;
    P->colptr.assign(N + 1, 0);
    for (int k = 0; k < N; ++k) P->colptr[k + 1] = P->colptr[k] + (int64_t)st[k].size();
    P->nnzL = P->colptr[N];
    P->ne = P->nnzL + N;
    if (P->ne >= INT32_MAX / 2) return fail(CPK_ERR_UNSUPPORTED, "factor too large for the device factorization plan");
    P->rowind.resize(P->nnzL);
    for (int k = 0; k < N; ++k) std::copy(st[k].begin(), st[k].end(), P->rowind.begin() + P->colptr[k]);
    const int ne = (int)P->ne, nnzL = (int)P->nnzL;
    auto id_of = [&](int i, int k) -> int {         // entry (i,k), i > k
        const auto &sk = st[k];
        return (int)(P->colptr[k] + (std::lower_bound(sk.begin(), sk.end(), i) - sk.begin()));
    };
    H->asrc.assign(ne, -1);
    H->dk.assign(ne, -1);
    for (int k = 0; k < N; ++k) {
        H->asrc[nnzL + k] = diagsrc[k];
        for (auto &e : low[k]) { const int id = id_of(e.first, k); if (H->asrc[id] < 0) H->asrc[id] = e.second; }
        for (int64_t q = P->colptr[k]; q < P->colptr[k + 1]; ++q) H->dk[q] = nnzL + k;
    }
    // rows of L: (column j, entry id) with j ascending
    std::vector<std::vector<std::pair<int, int>>> rowc(N);
    for (int j = 0; j < N; ++j) for (int64_t q = P->colptr[j]; q < P->colptr[j + 1]; ++q) rowc[P->rowind[q]].emplace_back(j, (int)q);
    // triple products of every node, j ascending
    std::vector<int> lev(ne, 0), pos(N, -1);
    std::vector<std::vector<int>> opl(ne);
    int64_t nops = 0;
    for (int k = 0; k < N; ++k) {
        pos[k] = nnzL + k;
        for (int64_t q = P->colptr[k]; q < P->colptr[k + 1]; ++q) pos[P->rowind[q]] = (int)q;
        for (auto &kj : rowc[k]) {
            const int j = kj.first, id_kj = kj.second;
            const auto &sj = st[j];
            for (size_t t = std::lower_bound(sj.begin(), sj.end(), k) - sj.begin(); t < sj.size(); ++t) {
                const int i = sj[t];
                const int id_ij = (int)(P->colptr[j] + t);
                if (i != k && (pos[i] < P->colptr[k] || pos[i] >= P->colptr[k + 1]))
                    return fail(CPK_ERR_ARG, "internal: fill pattern is not closed (column %d, row %d)", k, i);
                std::vector<int> &o = opl[pos[i]];
                o.push_back(id_ij); o.push_back(id_kj); o.push_back(nnzL + j);
                ++nops;
            }
        }
        if (nops > 60000000) return fail(CPK_ERR_UNSUPPORTED, "device factorization plan exceeds 6e7 products: keep the host factorization for this system");
        // levels: the diagonal node first, then the column below it
        int l = 0;
        for (size_t t = 0; t < opl[nnzL + k].size(); t += 3) l = std::max(l, lev[opl[nnzL + k][t]] + 1);
        lev[nnzL + k] = l;
        for (int64_t q = P->colptr[k]; q < P->colptr[k + 1]; ++q) {
            int lq = l + 1;
            for (size_t t = 0; t < opl[q].size(); t += 3) lq = std::max(lq, std::max(lev[opl[q][t]], lev[opl[q][t + 1]]) + 1);
            lev[q] = lq;
        }
    }
    P->nops = nops;
    int nlev = 0;
    for (int e = 0; e < ne; ++e) nlev = std::max(nlev, lev[e] + 1);
    P->nlev = nlev;
    H->levptr.assign(nlev + 1, 0);
    for (int e = 0; e < ne; ++e) H->levptr[lev[e] + 1]++;
    for (int l = 0; l < nlev; ++l) H->levptr[l + 1] += H->levptr[l];
    H->exec.resize(ne);
    {
        std::vector<int> nxt(H->levptr.begin(), H->levptr.end() - 1);
        for (int e = 0; e < ne; ++e) H->exec[nxt[lev[e]]++] = e;
    }
    H->opptr.assign(ne + 1, 0);
    for (int e = 0; e < ne; ++e) H->opptr[e + 1] = H->opptr[e] + (int)(opl[e].size() / 3);
    H->ops.resize((size_t)3 * nops);
    for (int e = 0; e < ne; ++e) std::copy(opl[e].begin(), opl[e].end(), H->ops.begin() + (size_t)3 * H->opptr[e]);
    return CPK_OK;
}

