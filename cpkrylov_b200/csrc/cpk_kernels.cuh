// cpk_kernels.cuh -- the phase functions of the hot path, written once against
// the Team abstraction:
//   spmv_sell      sparse mtimes  (cpcg.m:151-152, cpminres.m:187-188, opLDL2.m:175,182 ...)
//   ldl_solve      P * L^-T * D^-1 * L^-1 * P'  as ONE sync-free sweep (opLDL2.m:86)
//   ldl2_apply     opLDL2.multiply (opLDL2.m:161-188)
#pragma once
#include "cpk_device.cuh"

namespace cpk {

// ---------------------------------------------------------------------------
// phase-cycle accounting (profile=1): team thread 0 charges the cycles since the
// previous mark to a phase slot.  Shares of the solve, not absolute times.
// ---------------------------------------------------------------------------
struct PhaseClock {
    long long last;
    unsigned long long *acc;    // DevStatus::phase_cycles or nullptr
    bool on;
    __device__ void start(bool enable, unsigned long long *a) { on = enable; acc = a; if (on) last = clock64(); }
    __device__ __forceinline__ void mark(int slot) {
        if (on) { long long now = clock64(); acc[slot] += (unsigned long long)(now - last); last = now; }
    }
};

// ---------------------------------------------------------------------------
// SpMV, SELL-32: one warp per slice, one lane per row, entries of a row are
// accumulated left to right (the order a CSR/CSC CPU kernel uses).  Matrix
// arrays are read-only for the kernel lifetime -> ld.global.nc; the vector x is
// mutable across phases of the persistent kernel -> plain (coherent) loads.
// Epi is called as epi(row, sum) by the lane that owns the row.
// ---------------------------------------------------------------------------
template <class Team, class Epi>
__device__ __forceinline__ void spmv_sell(const Team &T, const DevSell &A, const double *x, Epi &&epi)
{
    for (int s = T.gwarp; s < A.nslices; s += T.nwarps) {
        const int beg = __ldg(&A.sptr[s]);
        const int end = __ldg(&A.sptr[s + 1]);
        const int row = __ldg(&A.rowmap[s * 32 + T.lane]);
        double acc = 0.0;
        int k = beg + T.lane;
        // 4 entries in flight per lane; the adds stay in row order
        for (; k + 96 < end; k += 128) {
            const int c0 = __ldg(&A.col[k]),      c1 = __ldg(&A.col[k + 32]);
            const int c2 = __ldg(&A.col[k + 64]), c3 = __ldg(&A.col[k + 96]);
            const double v0 = __ldg(&A.val[k]),      v1 = __ldg(&A.val[k + 32]);
            const double v2 = __ldg(&A.val[k + 64]), v3 = __ldg(&A.val[k + 96]);
            const double x0 = x[c0], x1 = x[c1], x2 = x[c2], x3 = x[c3];
            acc += v0 * x0; acc += v1 * x1; acc += v2 * x2; acc += v3 * x3;
        }
        for (; k < end; k += 32) acc += __ldg(&A.val[k]) * x[__ldg(&A.col[k])];
        if (row >= 0) epi(row, acc);
    }
    // long rows: one warp per row, lane-strided partial sums + butterfly
    for (int r = T.gwarp; r < A.nlong; r += T.nwarps) {
        const int beg = __ldg(&A.lptr[r]), end = __ldg(&A.lptr[r + 1]);
        double acc = 0.0;
        for (int k = beg + T.lane; k < end; k += 32) acc += __ldg(&A.lval[k]) * x[__ldg(&A.lcol[k])];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
        if (T.lane == 0) epi(__ldg(&A.lrow[r]), acc);
    }
}

// ---------------------------------------------------------------------------
// Input of an LDL solve: element i of the vector opLDL2.multiply works on.
//   value(i) = sgn(i) * z[i] - (sub ? sub[i] : 0)
// sgn(i) = -1 for i >= nA when neg_tail (the solvers pass [u; -t], e.g.
// cpminres.m:190); sub = [Aty; Cy] for the residual update (opLDL2.m:164-165).
// ---------------------------------------------------------------------------
struct VecIn {
    const double *z;
    const double *sub;
    int nA;
    bool neg_tail;
    __device__ __forceinline__ double operator()(int i) const {
        double v = z[i];
        if (neg_tail && i >= nA) v = -v;
        if (sub) v = v - sub[i];
        return v;
    }
};

// ---------------------------------------------------------------------------
// y = P * L^-T * D^-1 * L^-1 * P' * in        (opLDL2.m:86, right to left)
//
// One sweep, no barrier inside: the forward items (rows of L in dependency-
// level order) are followed by the backward items (D-solve fused into the row
// of L'); every produced value carries a ready flag (= the epoch of this solve)
// and every consumer polls the flags of what it needs.  A warp takes items in
// increasing order, all warps of the team are resident (cooperative launch or
// one CTA), and an item only depends on lower-numbered items, so the sweep
// cannot deadlock; wait loops never block a lane on another lane of its warp.
// If `accumulate`, out[p] += y (the y = y + dy of opLDL2.m:181).
// Ends WITHOUT a team barrier: the caller syncs before `out` is gathered.
// ---------------------------------------------------------------------------
template <class Team>
__device__ __noinline__ void ldl_solve(Team &T, const DevLdl &M, const VecIn &in, double *out, bool accumulate, int epoch_i)
{
    const unsigned long long epoch = (unsigned long long)(unsigned)epoch_i;
    const int nf = M.fwd.nitems, nb = M.bwd.nitems;
    for (int it = T.gwarp; it < nf + nb; it += T.nwarps) {
        const bool fwd = it < nf;
        const DevSweep &S = fwd ? M.fwd : M.bwd;
        const int s = fwd ? it : it - nf;
        const int beg = __ldg(&S.sptr[s]);
        const int end = __ldg(&S.sptr[s + 1]);
        const int slot = s * 32 + T.lane;
        const int rid = __ldg(&S.rid[slot]);
        bool done = rid < 0;
        int k = beg + T.lane;
        double acc = 0.0;
        int stage = 0;          // backward rows: 0 = still waiting for own (and partner) w
        int pidx = 0, partner = -1;
        double dd = 1.0;
        const Tagged *dep = fwd ? M.wbuf : M.ybuf;
        if (!done) {
            pidx = __ldg(&S.pidx[slot]);
            if (fwd) { acc = in(pidx); stage = 1; }
            else { partner = __ldg(&M.b_partner[slot]); dd = __ldg(&M.b_d[slot]); }
        }
        long long t0 = clock64();
        unsigned spins = 0;
        for (;;) {
            if (!done) {
                if (stage == 0) {
                    // D-solve on entry of the backward row (opLDL2.m:86, inv(op.D))
                    const Tagged w = ld_tagged(&M.wbuf[rid]);
                    if (partner < 0) {
                        if (w.tag == epoch) { acc = w.v / dd; stage = 1; }
                    } else {
                        const Tagged wp = ld_tagged(&M.wbuf[partner]);
                        if (w.tag == epoch && wp.tag == epoch) {
                            const double e = __ldg(&M.b_e[slot]);
                            const double dp = __ldg(&M.b_dp[slot]);
                            const double det = dd * dp - e * e;
                            acc = (dp * w.v - e * wp.v) / det;
                            stage = 1;
                        }
                    }
                }
                if (stage == 1) {
                    // up to 4 dependencies in flight; consumed strictly in row order
                    while (k < end) {
                        int c[4]; double v[4]; Tagged t[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) c[u] = (k + 32 * u < end) ? __ldg(&S.col[k + 32 * u]) : -1;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            if (c[u] >= 0) { v[u] = __ldg(&S.val[k + 32 * u]); t[u] = ld_tagged(&dep[c[u]]); }
                            else { v[u] = 0.0; t[u].v = 0.0; t[u].tag = epoch; }
                        }
                        bool stalled = false;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            if (stalled) continue;
                            if (c[u] < 0) { k = end; stalled = true; }
                            else if (t[u].tag == epoch) { acc -= v[u] * t[u].v; k += 32; }
                            else stalled = true;
                        }
                        if (stalled) break;
                    }
                    if (k >= end) {
                        if (fwd) {
                            st_tagged(&M.wbuf[rid], acc, epoch);
                        } else {
                            st_tagged(&M.ybuf[rid], acc, epoch);
                            if (accumulate) out[pidx] = out[pidx] + acc; else out[pidx] = acc;
                        }
                        done = true;
                    }
                }
            }
            if (__all_sync(FULL, done)) break;
            if ((++spins & 0x3f) == 0) {
                if (T.aborted()) break;
                if (clock64() - t0 > kWatchdogCycles) { T.set_abort(); break; }
            }
        }
    }
}

// r = xin - K*y with partial sums of r'r (and xin'xin): opLDL2.m:175-177,182-183
template <class Team>
__device__ __forceinline__ void resid_phase(Team &T, const DevLdl &M, const VecIn &xin, const double *y,
                                            double *r, double &rr, double &xx, bool want_xx)
{
    rr = 0.0; xx = 0.0;
    spmv_sell(T, M.KP, y, [&](int row, double s) {
        const double xi = xin(row);
        const double ri = xi - s;
        r[row] = ri;
        rr += ri * ri;
        if (want_xx) xx += xi * xi;
    });
}

// ---------------------------------------------------------------------------
// y = M * xin          opLDL2.multiply, opLDL2.m:161-188
// Entry: xin complete and visible to the team (caller synced).
// Exit : y complete and visible (ends with a team barrier).
// ---------------------------------------------------------------------------
template <class Team>
__device__ __noinline__ void ldl2_apply(Team &T, const DevLdl &M, const VecIn &xin, double *y, int &epoch,
                           DevStatus *st, PhaseClock &pc)
{
    const int n = M.nA;
    pc.mark(CPK_PH_VEC_);
    VecIn first = xin;
    if (M.residual_update) first.sub = M.atycy;                         // :164-165
    ldl_solve(T, M, first, y, false, ++epoch);
    T.sync();
    pc.mark(CPK_PH_LDL_);
    long long nsolve = 1, nres = 0;
    if (M.residual_update && M.ru_stateful) {                           // :169-172
        spmv_sell(T, M.K12, y + n, [&](int row, double s) { M.atycy[row] = s; });
        spmv_sell(T, M.K22, y + n, [&](int row, double s) { M.atycy[n + row] = s; });
        T.sync();
        pc.mark(CPK_PH_RESID_);
    }
    if (M.nitref > 0) {                                                 // :174
        double red[2];
        resid_phase(T, M, xin, y, M.rvec, red[0], red[1], true);
        T.template reduce<2>(red);
        ++nres;
        pc.mark(CPK_PH_RESID_);
        double rNorm = sqrt(red[0]);
        const double xNorm = sqrt(red[1]);
        int nit = 0;
        bool rknown = true;
        while (nit < M.nitref && (rNorm >= M.itref_tol * xNorm || M.force_itref)) {   // :179
            VecIn rin{M.rvec, nullptr, n, false};
            ldl_solve(T, M, rin, y, true, ++epoch);                     // dy = LDL*r; y = y + dy
            T.sync();
            ++nsolve;
            ++nit;
            pc.mark(CPK_PH_LDL_);
            if (nit < M.nitref || M.track_rnorm) {
                double r1[1], dummy;
                resid_phase(T, M, xin, y, M.rvec, r1[0], dummy, false);
                T.template reduce<1>(r1);
                rNorm = sqrt(r1[0]);
                ++nres;
                pc.mark(CPK_PH_RESID_);
            } else {
                rknown = false;     // the residual after the last step only feeds op.rNorm (:186)
            }
        }
        if (T.leader() && rknown) *M.rnorm_out = rNorm;
    }
    if (T.leader() && st) { st->napply += 1; st->nldlsolve += nsolve; st->nresid += nres; }
}

}  // namespace cpk
