// cpk_kernels.cuh -- the phase functions of the hot path, written once against
// the Team abstraction:
//   spmv_sell      sparse mtimes  (cpcg.m:151-152, cpminres.m:187-188, opLDL2.m:175,182 ...)
//   ldl_solve      P * L^-T * D^-1 * L^-1 * P'  over one item list: level-synchronous walk
//                  (default) or sync-free tagged walk (opLDL2.m:86)
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
// SpMV, SELL-32: one lane per row, entries of a row are accumulated left to
// right (the order a CSR/CSC CPU kernel uses).  Matrix
// arrays are read-only for the kernel lifetime -> ld_stream / ld_keep (cpk_device.cuh); the vector x is
// mutable across phases of the persistent kernel -> plain (coherent) loads.
// Epi is called as epi(row, sum) by the lane that owns the row.
// ---------------------------------------------------------------------------
// The same walk over PACKED entries (DevSell::pk: 16-bit column offset from the row + index into
// the table of distinct values, 4 bytes per entry instead of 12).  The column of an entry needs
// the row of its lane, i.e. the slice the entry belongs to: a decode cursor runs over the chunk
// ahead of the accumulating cursor (slice ends and row maps are one slice ahead in registers
// either way).
#ifndef CPK_LAZY_RVEC
#define CPK_LAZY_RVEC 1
#endif
// UP = packed entries per lane in flight.  Inside the solver kernels 4 (cfg 3 cpcg 3.015 -> 2.945 ms per
// solve against 8, 3 as good, 2 slower: profiles/r2_notes.md); the product on its own (k_matvec) keeps 8
// (cold 21 us against 23.7 us with 4).
#ifndef CPK_SPMV_UP
#define CPK_SPMV_UP 4
#endif
#ifndef CPK_SPMV_UP_ALONE
#define CPK_SPMV_UP_ALONE 8
#endif
template <int UP = CPK_SPMV_UP, class Team, class Epi>
__device__ __forceinline__ void spmv_sell_packed(Team &T, const DevSell &A, const double *x, Epi &&epi)
{
    const int lane = T.lane, gwarp = T.gwarp, nwarps = T.nwarps;
    int sa, sb;
    {
        const int *split = A.wsplit[Team::kKind];
        if (split != nullptr && A.nws[Team::kKind] == nwarps) {
            sa = ld_keep(&split[gwarp]); sb = ld_keep(&split[gwarp + 1]);
        } else {
            sa = (int)((long long)A.nslices * gwarp / nwarps);
            sb = (int)((long long)A.nslices * (gwarp + 1) / nwarps);
        }
    }
    if (sa < sb) {
        int s = sa;
        int k = ld_keep(&A.sptr[sa]);
        const int kb = ld_keep(&A.sptr[sb]);
        int send = ld_keep(&A.sptr[s + 1]);
        int send2 = (s + 2 <= sb) ? ld_keep(&A.sptr[s + 2]) : kb;
        int row = ld_keep(&A.rowmap[s * 32 + lane]);
        int row2 = (s + 1 < sb) ? ld_keep(&A.rowmap[(s + 1) * 32 + lane]) : -1;
        double acc = 0.0;
        constexpr int U = UP;
        unsigned ee[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int kk = k + 32 * u + lane;
            ee[u] = (kk < kb) ? ld_keep(&A.pk[kk]) : 0u;
        }
        while (k < kb) {
            unsigned en[U];
            double xv[U], vv[U];
            const int k2 = k + 32 * U;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int kk = k2 + 32 * u + lane;
                en[u] = (kk < kb) ? ld_keep(&A.pk[kk]) : 0u;
            }
            // decode: row of every entry of the chunk (a chunk seldom crosses more than one slice end)
            {
                int sd = s, sendd = send, send2d = send2, rowd = row, row2d = row2;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int k0 = k + 32 * u;              // warp-uniform
                    if (k0 < kb) {
                        while (k0 >= sendd) {
                            ++sd; sendd = send2d; rowd = row2d;
                            send2d = (sd + 2 <= sb) ? ld_keep(&A.sptr[sd + 2]) : kb;
                            row2d = (sd + 1 < sb) ? ld_keep(&A.rowmap[(sd + 1) * 32 + lane]) : -1;
                        }
                        const int col = max(rowd, 0) + (int)(short)(ee[u] & 0xffffu);
                        xv[u] = x[col];
                        vv[u] = __ldg(&A.dict[ee[u] >> 16]);
                    } else { xv[u] = 0.0; vv[u] = 0.0; }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int k0 = k + 32 * u;
                if (k0 < kb) {
                    while (k0 >= send) {                    // slice s is complete (possibly empty ones follow)
                        if (row >= 0) epi(row, acc);
                        acc = 0.0; ++s;
                        send = send2; row = row2;
                        send2 = (s + 2 <= sb) ? ld_keep(&A.sptr[s + 2]) : kb;
                        row2 = (s + 1 < sb) ? ld_keep(&A.rowmap[(s + 1) * 32 + lane]) : -1;
                    }
                    acc += vv[u] * xv[u];
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) ee[u] = en[u];
            k = k2;
        }
        while (s < sb) {
            if (row >= 0) epi(row, acc);
            acc = 0.0; ++s;
            row = row2;
            row2 = (s + 1 < sb) ? ld_keep(&A.rowmap[(s + 1) * 32 + lane]) : -1;
        }
    }
    for (int r = gwarp; r < A.nlong; r += nwarps) {
        const int beg = ld_keep(&A.lptr[r]), end = ld_keep(&A.lptr[r + 1]);
        double acc = 0.0;
        for (int kk = beg + lane; kk < end; kk += 32) acc += ld_keep(&A.lval[kk]) * x[ld_keep(&A.lcol[kk])];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
        if (lane == 0) epi(ld_keep(&A.lrow[r]), acc);
    }
}

template <int UP = CPK_SPMV_UP, class Team, class Epi>
__device__ __forceinline__ void spmv_sell(Team &T, const DevSell &A, const double *x, Epi &&epi)
{
    if (A.pk != nullptr) { spmv_sell_packed<UP>(T, A, x, epi); return; }
    const int lane = T.lane, gwarp = T.gwarp, nwarps = T.nwarps;      // T lives in local memory: read once
    // Each warp streams a CONTIGUOUS range of slices: the (col,val) arrays of the
    // range are one contiguous span, walked in chunks of U entries per lane with
    // the next chunk's loads issued before the current chunk's x-gathers are
    // consumed (register double buffering), so HBM latency is off the per-slice
    // critical path.  Slice ends come from sptr, prefetched one slice ahead.
    int sa, sb;
    {
        const int *split = A.wsplit[Team::kKind];
        if (split != nullptr && A.nws[Team::kKind] == nwarps) {
            sa = ld_stream(&split[gwarp]); sb = ld_stream(&split[gwarp + 1]);
        } else {
            sa = (int)((long long)A.nslices * gwarp / nwarps);
            sb = (int)((long long)A.nslices * (gwarp + 1) / nwarps);
        }
    }
    if (sa < sb) {
        int s = sa;
        int k = ld_stream(&A.sptr[sa]);
        const int kb = ld_stream(&A.sptr[sb]);
        int send = ld_stream(&A.sptr[s + 1]);
        int send2 = (s + 2 <= sb) ? ld_stream(&A.sptr[s + 2]) : kb;
        int row = ld_stream(&A.rowmap[s * 32 + lane]);
        int row2 = (s + 1 < sb) ? ld_stream(&A.rowmap[(s + 1) * 32 + lane]) : -1;
        double acc = 0.0;
#ifndef CPK_SPMV_U
#define CPK_SPMV_U 4
#endif
        constexpr int U = CPK_SPMV_U;   // entries per lane in flight (x2: current + prefetched chunk)
        int cc[U]; double vv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int kk = k + 32 * u + lane;
            if (kk < kb) { cc[u] = ld_stream(&A.col[kk]); vv[u] = ld_stream(&A.val[kk]); } else { cc[u] = 0; vv[u] = 0.0; }
        }
        while (k < kb) {
            int cn[U]; double vn[U], xv[U];
            const int k2 = k + 32 * U;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int kk = k2 + 32 * u + lane;
                if (kk < kb) { cn[u] = ld_stream(&A.col[kk]); vn[u] = ld_stream(&A.val[kk]); } else { cn[u] = 0; vn[u] = 0.0; }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) xv[u] = (k + 32 * u < kb) ? x[cc[u]] : 0.0;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int k0 = k + 32 * u;                  // warp-uniform
                if (k0 < kb) {
                    while (k0 >= send) {                    // slice s is complete (possibly empty ones follow)
                        if (row >= 0) epi(row, acc);
                        acc = 0.0; ++s;
                        send = send2; row = row2;
                        send2 = (s + 2 <= sb) ? ld_stream(&A.sptr[s + 2]) : kb;
                        row2 = (s + 1 < sb) ? ld_stream(&A.rowmap[(s + 1) * 32 + lane]) : -1;
                    }
                    acc += vv[u] * xv[u];
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) { cc[u] = cn[u]; vv[u] = vn[u]; }
            k = k2;
        }
        while (s < sb) {
            if (row >= 0) epi(row, acc);
            acc = 0.0; ++s;
            row = row2;
            row2 = (s + 1 < sb) ? ld_stream(&A.rowmap[(s + 1) * 32 + lane]) : -1;
        }
    }
    // long rows: one warp per row, lane-strided partial sums + butterfly
    for (int r = gwarp; r < A.nlong; r += nwarps) {
        const int beg = ld_stream(&A.lptr[r]), end = ld_stream(&A.lptr[r + 1]);
        double acc = 0.0;
        for (int k = beg + lane; k < end; k += 32) acc += ld_stream(&A.lval[k]) * x[ld_stream(&A.lcol[k])];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
        if (lane == 0) epi(ld_stream(&A.lrow[r]), acc);
    }
}

// ---------------------------------------------------------------------------
// blkdiag(A, C) * V for the solver loops (the sparse mtimes call sites cpcg.m:151-152,
// cpminres.m:187-188, ...): epi(row, sum) for every row of [A*v; C*q].
// Explicit A: ONE streaming pass over the SELL form of blkdiag(H, C).
// Matrix-free A (DevHostOp): A*v comes from the host through the mailbox, C*q from the
// device.  `with_c` = false: only the A rows (rhs shift of reg_cpkrylov.m:157).
// Entry: V complete and visible to the team (every call site follows a team barrier).
// `seq` counts the requests of this launch and is carried by every thread.
// ---------------------------------------------------------------------------
template <class Team, class Epi>
__device__ __noinline__ void hostop_mult(Team &T, const DevSystem &S, int &seq, const double *V, bool with_c, Epi &&epi)
{
    const DevHostOp &Hh = S.hop;
    ++seq;
    if (T.leader()) {
        __threadfence_system();                 // V (written by other SMs, ordered by the team barrier) before the request
        *Hh.req_addr = (long long)V;
        __threadfence_system();
        *Hh.req = seq;
    }
    if (T.cta_leader()) {
        const long long t0 = clock64();
        unsigned spins = 0;
        while (ld_acquire_sys(Hh.ack) != seq) {
            __nanosleep(500);
            if ((++spins & 0x3f) == 0) {
                if (T.aborted()) break;
                if (clock64() - t0 > Hh.timeout) { T.set_abort(); break; }
            }
        }
    }
    T.cta_sync();
    if (T.aborted()) return;
    TEAM_FOR(T, i, S.n) epi(i, ld_cg(&Hh.u[i]));       // written by the copy engine: never through L1
    if (with_c) {
        const int n = S.n;
        spmv_sell(T, S.Cm, V + n, [&](int row, double s) { epi(n + row, s); });
    }
}

template <class Team, class Epi>
__device__ __forceinline__ void kkt_mult(Team &T, const DevSystem &S, int &seq, const double *V, Epi &&epi)
{
    if (S.hop.on) hostop_mult(T, S, seq, V, true, epi);
    else spmv_sell(T, S.HC, V, epi);
}

// ---------------------------------------------------------------------------
// Input of an LDL solve: element i of the vector opLDL2.multiply works on.
//   value(i) = sgn(i) * z[i] - (sub ? sub[i] : 0)
// sgn(i) = -1 for i >= nA when neg_tail (the solvers pass [u; -t], e.g.
// cpminres.m:190); sub = [Aty; Cy] for the residual update (opLDL2.m:164-165).
// ---------------------------------------------------------------------------
// `add`/`scale`: the vector is z + scale*add evaluated on the fly (a solver's pending
// axpy, e.g. g + alpha*Ap of cpcg.m:163-164, folded into the gather); `wb` asks the
// apply to write the evaluated vector back into z during its residual pass.
struct VecIn {
    const double *z;
    const double *sub;
    int nA;
    bool neg_tail;
    const double *add = nullptr;
    double scale = 0.0;
    double *wb = nullptr;
    __device__ __forceinline__ double operator()(int i) const {
        double v = z[i];
        if (add) v = v + scale * add[i];
        if (neg_tail && i >= nA) v = -v;
        if (sub) v = v - sub[i];
        return v;
    }
};

}  // namespace cpk
#include "cpk_rc.cuh"
namespace cpk {

struct ItemMeta { int beg, end, slot, rid, pidx, flags; double d; };
struct ItemChunk { int c[4]; double v[4]; };

__device__ __forceinline__ void item_load(const DevSweep &S, int t, int lane, ItemMeta &m, ItemChunk &r)
{
    m.beg = ld_stream(&S.sptr[t]);
    m.end = ld_stream(&S.sptr[t + 1]);
    m.slot = item_slot(S, t, lane);
    m.rid = ld_stream(&S.rid[m.slot]);
    m.pidx = ld_stream(&S.pidx[m.slot]);
    m.flags = ld_stream(&S.flags[m.slot]);
    m.d = ld_stream(&S.d[m.slot]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int kk = m.beg + 32 * u + lane;
        if (kk < m.end) { r.c[u] = ld_stream(&S.col[kk]); r.v[u] = ld_stream(&S.val[kk]); }
        else { r.c[u] = -1; r.v[u] = 0.0; }
    }
}

// Spin until the tagged value carries this solve's epoch (lanes of one item never
// depend on one another, so a per-lane blocking wait cannot deadlock).
template <class Team>
__device__ __forceinline__ double wait_tagged(const Team &T, const Tagged *p, unsigned long long epoch)
{
    Tagged tg = ld_tagged(p);
    if (tg.tag != epoch) {
        long long t0 = clock64();
        unsigned spins = 0, ns = 32;
        do {
            __nanosleep(ns);
            if (ns < 256) ns <<= 1;
            tg = ld_tagged(p);
            if ((++spins & 0x3f) == 0) {
                if (T.aborted()) break;
                if (clock64() - t0 > kWatchdogCycles) { T.set_abort(); break; }
            }
        } while (tg.tag != epoch);
    }
    return tg.v;
}

// One item.  TAGGED: sync-free walk (values carry their epoch and are polled);
// otherwise level-synchronous walk (plain values, levels separated by barriers).
template <bool TAGGED, class Team>
__device__ __forceinline__ void item_process(Team &T, const DevLdl &M, const VecIn in, double *out, bool accumulate,
                                             unsigned long long epoch, const ItemMeta &m, const ItemChunk &r)
{
    const DevSweep &S = M.sw;
    const int Nn = M.N;                 // backward rows: code c >= N addresses the forward result w[c - N]
    const int f0 = __shfl_sync(FULL, m.flags, 0);
    const bool warprow = (f0 & F_WARPROW) != 0;
    const bool isfwd = (f0 & F_FWD) != 0;       // uniform over an item
    const bool live = m.rid >= 0;

    // value of dependency code c (c >= 0), blocking until published when TAGGED
    auto dep_value = [&](int c) -> double {
        if (TAGGED) {
            const Tagged *src = (c >= Nn) ? &M.wbuf[c - Nn] : (isfwd ? &M.wbuf[c] : &M.ybuf[c]);
            return wait_tagged(T, src, epoch);
        } else {
            return (c >= Nn) ? M.wv[c - Nn] : (isfwd ? M.wv[c] : M.yv[c]);
        }
    };
    // ---- sum over the row's entries, 4 per lane in flight, first 4 prefetched --------
    double sum = 0.0;                   // = - sum_j val_j * value_j, accumulated in storage order
    {
        int c[4]; double v[4], x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { c[u] = r.c[u]; v[u] = r.v[u]; }
        int base = m.beg;
        for (;;) {
            if (TAGGED) {
                // issue all tagged loads first, wait (re-poll) afterwards
                Tagged tg[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (c[u] >= 0) tg[u] = ld_tagged((c[u] >= Nn) ? &M.wbuf[c[u] - Nn] : (isfwd ? &M.wbuf[c[u]] : &M.ybuf[c[u]]));
                    else { tg[u].v = (c[u] <= -2) ? in(-c[u] - 2) : 0.0; tg[u].tag = epoch; }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) x[u] = (tg[u].tag == epoch) ? tg[u].v : dep_value(c[u]);
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) x[u] = (c[u] >= 0) ? dep_value(c[u]) : ((c[u] <= -2) ? in(-c[u] - 2) : 0.0);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (c[u] != -1) sum -= v[u] * x[u];
            base += 128;
            if (base >= m.end) break;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int kk = base + 32 * u + T.lane;
                if (kk < m.end) { c[u] = ld_stream(&S.col[kk]); v[u] = ld_stream(&S.val[kk]); } else { c[u] = -1; v[u] = 0.0; }
            }
        }
    }
    if (warprow) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
        if (T.lane != 0) return;
    } else if (!live) return;
    // ---- base value: (P'z)_i for a forward row, D-solve on entry of a backward row ----
    double acc;
    if (isfwd) acc = in(m.pidx);
    else {
        double w;
        if (m.flags & F_WDIRECT) w = in(m.pidx);
        else w = TAGGED ? wait_tagged(T, &M.wbuf[m.rid], epoch) : M.wv[m.rid];
        if (!(m.flags & F_PARTNER)) acc = w / m.d;                          // opLDL2.m:86, inv(op.D)
        else {
            const int partner = ld_stream(&S.partner[m.slot]);
            const double wp = TAGGED ? wait_tagged(T, &M.wbuf[partner], epoch) : M.wv[partner];
            const double e = ld_stream(&S.e[m.slot]);
            const double dp = ld_stream(&S.dp[m.slot]);
            const double det = m.d * dp - e * e;
            acc = (dp * w - e * wp) / det;
        }
    }
    acc += sum;
    // ---- publish -------------------------------------------------------------------------
    if (isfwd && !(m.flags & F_FUSED)) {
        if (TAGGED) st_tagged(&M.wbuf[m.rid], acc, epoch); else M.wv[m.rid] = acc;
    } else {
        if (isfwd) acc = acc / m.d;                 // F_FUSED: column of L is empty, y_i = w_i / d_i
        if (isfwd || (m.flags & F_STORE)) { if (TAGGED) st_tagged(&M.ybuf[m.rid], acc, epoch); else M.yv[m.rid] = acc; }
        if (accumulate) out[m.pidx] = out[m.pidx] + acc; else out[m.pidx] = acc;
    }
}

// The lone rows of a solve (DevSweep::lone_*): out[p] = in[p] / d, all loads of four rows per thread in flight.
// They depend on the input vector only and nobody gathers them: run at the start of a walk, no barrier needed.
template <class Team>
__device__ __forceinline__ void ldl_lone_rows(Team &T, const DevSweep &S, const VecIn &in, double *out, bool accumulate)
{
    const int nl = S.nlone;
    const int tid = T.tid, nth = T.nthreads;
    for (int i0 = tid; i0 < nl; i0 += 4 * nth) {
        int pi[4]; double dd[4], zz[4], oo[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * nth;
            pi[u] = (i < nl) ? ld_stream(&S.lone_pidx[i]) : -1;
            dd[u] = (i < nl) ? ld_stream(&S.lone_d[i]) : 1.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) { zz[u] = (pi[u] >= 0) ? in(pi[u]) : 0.0; oo[u] = (accumulate && pi[u] >= 0) ? out[pi[u]] : 0.0; }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (pi[u] >= 0) { const double y = zz[u] / dd[u]; out[pi[u]] = accumulate ? oo[u] + y : y; }
    }
}

// Sync-free walk: no barrier inside.  Every produced value is published with
// the epoch of this solve in one 128-bit store; consumers re-poll until the tag
// matches.  Items only depend on lower-numbered items, every warp walks its
// items in increasing order and all warps of the team are resident, so the walk
// cannot deadlock.  Used when a sweep is too deep for one barrier per level.
template <class Team>
__device__ __noinline__ void ldl_solve_syncfree(Team &T, const DevLdl &M, const VecIn in, double *out, bool accumulate, int epoch_i)
{
    const unsigned long long epoch = (unsigned long long)(unsigned)epoch_i;
    const DevSweep &S = M.sw;
    ldl_lone_rows(T, S, in, out, accumulate);
    for (int g = 0; g < S.nseg; ++g) {
        const int sbeg = ld_stream(&S.seg[3 * g]), send = ld_stream(&S.seg[3 * g + 1]), blk = ld_stream(&S.seg[3 * g + 2]);
        const int nblk = (send - sbeg + blk - 1) / blk;
        for (int b = T.gwarp; b < nblk; b += T.nwarps) {
            int t = sbeg + b * blk;
            const int tend = min(send, t + blk);
            ItemMeta m0; ItemChunk r0;
            item_load(S, t, T.lane, m0, r0);
            for (; t < tend; ++t) {
                ItemMeta m1; ItemChunk r1;
                if (t + 1 < tend) item_load(S, t + 1, T.lane, m1, r1);
                item_process<true>(T, M, in, out, accumulate, epoch, m0, r0);
                m0 = m1; r0 = r1;
            }
        }
    }
}

// Level-synchronous walk of the item list: one team barrier per dependency
// level, plain 8-byte values, no polling.  Inside a level every warp owns a
// contiguous range of items and walks it in batches of B: all row data of the
// batch is requested first, then all gathers, then the arithmetic -- three
// memory round trips per batch instead of three per item.  Items that are not
// "simple" (rows longer than 2 entries, warp-rows, 2x2 pivots) take the generic
// per-item path.
template <bool DYN, class Team>
__device__ __noinline__ void ldl_solve_levels_impl(Team &T, const DevLdl &M, const VecIn in, double *out, bool accumulate,
                                                   PhaseClock *dbg)
{
    const DevSweep &S = M.sw;
#ifndef CPK_SWEEP_B
#define CPK_SWEEP_B 2
#endif
    constexpr int B = CPK_SWEEP_B;
    const int Nn = M.N;
    const int lane = T.lane, gwarp = T.gwarp, nwarps = T.nwarps;      // T lives in local memory: read once
    // the lone rows ride on the first BACKWARD level (the lighter one of a shallow sweep: the forward level
    // holds the long merged rows)
    bool lone_done = S.nlone == 0;
    // every CTA walks ITS share of a level (cost-balanced at set-up); inside the CTA the warps take
    // batches off the share through a counter in shared memory, so a warp that drew cheap items simply
    // draws more of them.  Two counters, used alternately: the one of the next level is reset while this
    // level runs (the team barrier between the levels orders the reset before its first use).
    const int cta = T.cta(), nctas = T.nctas();
    constexpr bool dyn = DYN;                     // large shallow systems (S.ctasplit); otherwise: a contiguous range of equal count per warp
    const bool exact = dyn && S.nsplit == nctas;   // the shares were cut for this team shape (else: equal counts per CTA)
    int *wq = T.sh->wq;
    if (dyn) {
        if (threadIdx.x == 0) { wq[0] = 0; wq[1] = 0; }
        T.cta_sync();
    }
    for (int g = 0; g < S.nlev; ++g) {
        const int a = ld_stream(&S.levptr[g]), b = ld_stream(&S.levptr[g + 1]);
        if (!lone_done && a >= S.nfwd) { ldl_lone_rows(T, S, in, out, accumulate); lone_done = true; }
        const long long cnt = b - a;
        int t, tend, cs = 0;
        int *queue = &wq[g & 1];
        if (dyn) {
            if (exact) {
                cs = ld_stream(&S.ctasplit[(size_t)g * (nctas + 1) + cta]);
                tend = ld_stream(&S.ctasplit[(size_t)g * (nctas + 1) + cta + 1]);
            } else {
                cs = a + (int)(cnt * cta / nctas);
                tend = a + (int)(cnt * (cta + 1) / nctas);
            }
            if (threadIdx.x == 0) wq[(g + 1) & 1] = 0;
            t = 0;
            if (lane == 0) t = atomicAdd(queue, B);
            t = __shfl_sync(FULL, t, 0) + cs;
        } else {
            t = a + (int)(cnt * gwarp / nwarps);
            tend = a + (int)(cnt * (gwarp + 1) / nwarps);
        }
        while (t < tend) {
            const int nb = min(B, tend - t);
            // ---- stage A: row data of the whole batch -------------------------------
            int beg[B], wid[B], rid[B], pidx[B], flg[B], c0[B], c1[B];
            double dd[B], v0[B], v1[B];
#pragma unroll
            for (int q = 0; q < B; ++q) {
                if (q < nb) {
                    beg[q] = ld_stream(&S.sptr[t + q]);
                    wid[q] = ld_stream(&S.sptr[t + q + 1]) - beg[q];
                    const int slot = item_slot(S, t + q, lane);
                    rid[q] = ld_stream(&S.rid[slot]); pidx[q] = ld_stream(&S.pidx[slot]);
                    flg[q] = ld_stream(&S.flags[slot]); dd[q] = ld_stream(&S.d[slot]);
                } else { beg[q] = 0; wid[q] = 0; rid[q] = -1; pidx[q] = 0; flg[q] = 0; dd[q] = 1.0; }
            }
            bool simple = true;
#pragma unroll
            for (int q = 0; q < B; ++q) simple = simple && wid[q] <= 64 && !(flg[q] & (F_PARTNER | F_WARPROW));
            simple = __all_sync(FULL, simple);
            if (simple) {
#pragma unroll
                for (int q = 0; q < B; ++q) {
                    c0[q] = -1; c1[q] = -1; v0[q] = 0.0; v1[q] = 0.0;
                    if (wid[q] >= 32) { c0[q] = ld_stream(&S.col[beg[q] + lane]); v0[q] = ld_stream(&S.val[beg[q] + lane]); }
                    if (wid[q] >= 64) { c1[q] = ld_stream(&S.col[beg[q] + 32 + lane]); v1[q] = ld_stream(&S.val[beg[q] + 32 + lane]); }
                }
                // ---- stage B: gathers ---------------------------------------------------
                double base[B], x0[B], x1[B];
#pragma unroll
                for (int q = 0; q < B; ++q) {
                    const bool isfwd = (flg[q] & F_FWD) != 0;
                    base[q] = 0.0; x0[q] = 0.0; x1[q] = 0.0;
                    if (rid[q] >= 0) {
                        base[q] = (isfwd || (flg[q] & F_WDIRECT)) ? in(pidx[q]) : ld_tmp(&M.wv[rid[q]]);
                        const double *depv = isfwd ? M.wv : M.yv;
                        if (c0[q] >= 0) x0[q] = (c0[q] >= Nn) ? M.wv[c0[q] - Nn] : depv[c0[q]];
                        else if (c0[q] <= -2) x0[q] = in(-c0[q] - 2);
                        if (c1[q] >= 0) x1[q] = (c1[q] >= Nn) ? M.wv[c1[q] - Nn] : depv[c1[q]];
                        else if (c1[q] <= -2) x1[q] = in(-c1[q] - 2);
                    }
                }
                // ---- stage C: arithmetic and publication -----------------------------------
#pragma unroll
                for (int q = 0; q < B; ++q) {
                    if (rid[q] < 0) continue;
                    const bool isfwd = (flg[q] & F_FWD) != 0;
                    double acc = isfwd ? base[q] : base[q] / dd[q];     // D-solve on entry of a backward row
                    double sum = 0.0;
                    if (c0[q] != -1) sum -= v0[q] * x0[q];
                    if (c1[q] != -1) sum -= v1[q] * x1[q];
                    acc += sum;
                    if (isfwd && !(flg[q] & F_FUSED)) st_tmp(&M.wv[rid[q]], acc);
                    else {
                        if (isfwd) acc = acc / dd[q];
                        if (isfwd || (flg[q] & F_STORE)) st_tmp(&M.yv[rid[q]], acc);
                        if (accumulate) out[pidx[q]] = out[pidx[q]] + acc; else out[pidx[q]] = acc;
                    }
                }
            } else {
                // batch of warp-rows (one long row per item, <= 64 entries, 1x1 pivot): same
                // three round trips for the whole batch, one butterfly per row
                bool wr = true;
#pragma unroll
                for (int q = 0; q < B; ++q) {
                    const int f0 = __shfl_sync(FULL, flg[q], 0);
                    wr = wr && (q >= nb || ((f0 & F_WARPROW) && !(f0 & F_PARTNER) && wid[q] <= 64));
                }
                if (wr) {
                    int rid0[B], pidx0[B], flg0[B];
                    double d0[B];
#pragma unroll
                    for (int q = 0; q < B; ++q) {
                        rid0[q] = __shfl_sync(FULL, rid[q], 0); pidx0[q] = __shfl_sync(FULL, pidx[q], 0);
                        flg0[q] = __shfl_sync(FULL, flg[q], 0); d0[q] = __shfl_sync(FULL, dd[q], 0);
                        c0[q] = -1; c1[q] = -1; v0[q] = 0.0; v1[q] = 0.0;
                        if (q < nb && wid[q] >= 32) { c0[q] = ld_stream(&S.col[beg[q] + lane]); v0[q] = ld_stream(&S.val[beg[q] + lane]); }
                        if (q < nb && wid[q] >= 64) { c1[q] = ld_stream(&S.col[beg[q] + 32 + lane]); v1[q] = ld_stream(&S.val[beg[q] + 32 + lane]); }
                    }
                    double base[B], x0[B], x1[B];
#pragma unroll
                    for (int q = 0; q < B; ++q) {
                        base[q] = 0.0; x0[q] = 0.0; x1[q] = 0.0;
                        if (q < nb) {
                            const bool isfwd = (flg0[q] & F_FWD) != 0;
                            if (lane == 0) base[q] = (isfwd || (flg0[q] & F_WDIRECT)) ? in(pidx0[q]) : ld_tmp(&M.wv[rid0[q]]);
                            const double *depv = isfwd ? M.wv : M.yv;
                            if (c0[q] >= 0) x0[q] = (c0[q] >= Nn) ? M.wv[c0[q] - Nn] : depv[c0[q]];
                            else if (c0[q] <= -2) x0[q] = in(-c0[q] - 2);
                            if (c1[q] >= 0) x1[q] = (c1[q] >= Nn) ? M.wv[c1[q] - Nn] : depv[c1[q]];
                            else if (c1[q] <= -2) x1[q] = in(-c1[q] - 2);
                        }
                    }
#pragma unroll
                    for (int q = 0; q < B; ++q) {
                        double sum = 0.0;
                        if (c0[q] != -1) sum -= v0[q] * x0[q];
                        if (c1[q] != -1) sum -= v1[q] * x1[q];
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
                        if (q < nb && lane == 0) {
                            const bool isfwd = (flg0[q] & F_FWD) != 0;
                            double acc = isfwd ? base[q] : base[q] / d0[q];
                            acc += sum;
                            if (isfwd && !(flg0[q] & F_FUSED)) st_tmp(&M.wv[rid0[q]], acc);
                            else {
                                if (isfwd) acc = acc / d0[q];
                                if (isfwd || (flg0[q] & F_STORE)) st_tmp(&M.yv[rid0[q]], acc);
                                if (accumulate) out[pidx0[q]] = out[pidx0[q]] + acc; else out[pidx0[q]] = acc;
                            }
                        }
                    }
                } else {
                    for (int q = 0; q < nb; ++q) {
                        ItemMeta m; ItemChunk r;
                        item_load(S, t + q, lane, m, r);
                        item_process<false>(T, M, in, out, accumulate, 0ull, m, r);
                    }
                }
            }
            if (dyn) {
                int tn = 0;
                if (lane == 0) tn = atomicAdd(queue, B);
                t = __shfl_sync(FULL, tn, 0) + cs;
            } else t += nb;
        }
        if (dbg) dbg->mark(min(2 * g, 6));
        if (g + 1 < S.nlev) T.sync();       // the caller syncs after the last level
        if (dbg) dbg->mark(min(2 * g + 1, 7));
    }
    if (!lone_done) ldl_lone_rows(T, S, in, out, accumulate);
}

template <class Team>
__device__ __forceinline__ void ldl_solve_levels(Team &T, const DevLdl &M, const VecIn in, double *out, bool accumulate,
                                                 PhaseClock *dbg = nullptr)
{
    if (M.sw.ctasplit != nullptr) ldl_solve_levels_impl<true>(T, M, in, out, accumulate, dbg);
    else ldl_solve_levels_impl<false>(T, M, in, out, accumulate, dbg);
}

// ---------------------------------------------------------------------------
// Compact walk (one-CTA team): see DevCompact.  Shared-memory region:
//   [ring: kCwStages x kCwBlock][mbarrier x kCwStages][ring entries consumed so far][sv: N or 2N doubles]
// The ring never stops: entry k holds stream block k mod nblk, and as soon as the
// CTA is done with an entry the block kCwStages further on is requested, also
// across the end of a solve -- the next solve finds its first blocks in place.
// ---------------------------------------------------------------------------
struct CwSmem {
    unsigned char *ring;
    unsigned long long *full;
    unsigned *count;
    double *sv;
    __device__ __forceinline__ explicit CwSmem(int off) {
        ring = reinterpret_cast<unsigned char *>(g_dsm) + off;
        full = reinterpret_cast<unsigned long long *>(ring + (size_t)kCwStages * kCwBlock);
        count = reinterpret_cast<unsigned *>(full + kCwStages);
        sv = reinterpret_cast<double *>(ring + (size_t)kCwStages * kCwBlock + 8 * kCwStages + 16);
    }
    // thread 0: request the stream block of ring entry k
    __device__ __forceinline__ void issue(const DevCompact &C, unsigned k) const {
        unsigned long long *bar = &full[k % kCwStages];
        mbar_expect_tx(bar, kCwBlock);
        bulk_g2s(ring + (size_t)(k % kCwStages) * kCwBlock, C.stream + (size_t)(k % (unsigned)C.nblk) * kCwBlock, kCwBlock, bar);
    }
    template <class Team>
    __device__ __forceinline__ bool wait(const Team &T, unsigned k) const {
        unsigned long long *bar = &full[k % kCwStages];
        const unsigned parity = (k / kCwStages) & 1u;
        if (mbar_try_wait(bar, parity)) return true;
        const long long t0 = clock64();
        while (!mbar_try_wait(bar, parity))
            if (clock64() - t0 > kWatchdogCycles) { T.set_abort(); return false; }
        return true;
    }
};

// once per kernel, before the first solve: barriers + the first ring entries
template <class Team>
__device__ __forceinline__ void compact_init(Team &T, const DevLdl &M)
{
    if (Team::kKind != 1 || M.cw.smem_off < 0) return;
    CwSmem W(M.cw.smem_off);
    if (T.tid == 0) {
        for (int s = 0; s < kCwStages; ++s) mbar_init(&W.full[s], 1);
        *W.count = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (unsigned k = 0; k < (unsigned)kCwStages; ++k) W.issue(M.cw, k);
    }
    __syncthreads();
}
// once per kernel, after the last solve: no bulk copy may be in flight when the CTA exits
template <class Team>
__device__ __forceinline__ void compact_drain(Team &T, const DevLdl &M)
{
    if (Team::kKind != 1 || M.cw.smem_off < 0) return;
    CwSmem W(M.cw.smem_off);
    __syncthreads();
    if (T.tid == 0) {
        const unsigned k0 = *W.count;
        for (unsigned k = k0; k < k0 + (unsigned)kCwStages; ++k) W.wait(T, k);
    }
    __syncthreads();
}

// What a warp does in one step, held in registers.  It is fetched right after the
// previous step's arithmetic, before the barrier, so that only gather -> arithmetic ->
// store sit between two barriers.  Only the common shapes (<= 2 entries per lane) are
// fetched ahead; wider items read their entries in place -- the interpreter is kept small
// on purpose (instruction fetch and issue of ONE warp bound a level).
struct CwTask {
    int kind, width, stride, z;
    bool barrier;
    const unsigned char *data;
    int t;                  // target index in sv, -1 = idle lane
    int c0, c1;
    double v0, v1;
};

__device__ __forceinline__ void cw_fetch(CwTask &I, const unsigned char *blk, const int4 tk, int lane)
{
    I.kind = tk.y & 15; I.barrier = (tk.y & CW_BARRIER) != 0;
    I.width = (tk.y >> 8) & 0xffff; I.stride = (tk.y >> 24) & 0xff;
    I.z = tk.z;
    I.data = blk + tk.x;
    I.t = -1; I.c0 = -1; I.c1 = -1; I.v0 = 0.0; I.v1 = 0.0;
    if (I.kind == CW_ROWS2) {
        if (lane < I.stride) {
            const int4 ri = *reinterpret_cast<const int4 *>(I.data + 32 * lane);
            const double2 rv = *reinterpret_cast<const double2 *>(I.data + 32 * lane + 16);
            I.t = ri.x; I.c0 = ri.y; I.c1 = ri.z; I.v0 = rv.x; I.v1 = rv.y;
        }
    } else if (I.kind == CW_WARPROW) {
        if (lane == 0) I.t = tk.z;
        if (I.width <= 2) {
            const double *val = reinterpret_cast<const double *>(I.data);
            const int *col = reinterpret_cast<const int *>(I.data + (size_t)256 * I.width);
            I.c0 = col[lane]; I.v0 = val[lane];
            if (I.width == 2) { I.c1 = col[32 + lane]; I.v1 = val[32 + lane]; }
        }
    } else if (I.kind == CW_ROWS) {
        if (lane < I.stride) I.t = reinterpret_cast<const int *>(I.data)[lane];
        I.data += 4 * I.stride;
    }
}

// sv[t] += -sum_k val_k * sv[col_k], entries in storage order (warp-row: lane partials, butterfly)
__device__ __forceinline__ void cw_rows(const CwTask &I, double *sv, int lane)
{
    const double base = (I.t >= 0) ? sv[I.t] : 0.0;
    double sum = 0.0;
    if (I.width <= 2) {
        const double x0 = (I.c0 >= 0) ? sv[I.c0] : 0.0;
        const double x1 = (I.c1 >= 0) ? sv[I.c1] : 0.0;
        if (I.c0 >= 0) sum -= I.v0 * x0;
        if (I.c1 >= 0) sum -= I.v1 * x1;
    } else if (lane < I.stride) {
        const double *val = reinterpret_cast<const double *>(I.data);
        const int *col = reinterpret_cast<const int *>(I.data + (size_t)8 * I.stride * I.width);
        for (int k = 0; k < I.width; ++k) {
            const int c = col[k * I.stride + lane];
            if (c >= 0) sum -= val[k * I.stride + lane] * sv[c];
        }
    }
    if (I.kind == CW_WARPROW) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
    }
    if (I.t >= 0) sv[I.t] = base + sum;
}

// D chunk, rows 32*warp + lane of it: y_i = w_i / d_i, or the 2x2 solve with the partner
// row (opLDL2.m:86, inv(op.D))
__device__ __forceinline__ void cw_dchunk(const CwTask &I, double *sv, int yoff, int warp, int lane)
{
    const int r = 32 * warp + lane, cnt = I.width;
    if (r >= cnt) return;
    const double *dd = reinterpret_cast<const double *>(I.data);
    const double w = sv[I.z + r];
    double y;
    if (!I.stride) y = w / dd[r];
    else {
        const double *e = dd + cnt, *dp = e + cnt;
        const int *partner = reinterpret_cast<const int *>(dp + cnt);
        if (partner[r] < 0) y = w / dd[r];
        else {
            const double wp = sv[partner[r]];
            const double det = dd[r] * dp[r] - e[r] * e[r];
            y = (dp[r] * w - e[r] * wp) / det;
        }
    }
    sv[yoff + I.z + r] = y;
}

template <class Team>
__device__ __noinline__ void ldl_solve_compact(Team &T, const DevLdl &M, const VecIn in, double *out, bool accumulate,
                                               unsigned long long *dbg = nullptr)
{
    long long tq = dbg ? clock64() : 0;
    unsigned long long cyc[4] = {0, 0, 0, 0};       // gather, wait for the ring, steps, scatter
    auto lap = [&](int slot) { if (dbg) { const long long t = clock64(); cyc[slot] += (unsigned long long)(t - tq); tq = t; } };
    const DevCompact &C = M.cw;
    const int N = M.N;
    if (T.aborted()) return;
    CwSmem W(C.smem_off);
    const unsigned k0 = *W.count;
    const int nblk = C.nblk;
    const int warp = T.gwarp, lane = T.lane;
    // w = P' * in: all index loads of a batch first, then all gathers
    for (int i0 = T.tid; i0 < N; i0 += 8 * T.nthreads) {
        int pi[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int i = i0 + u * T.nthreads; pi[u] = (i < N) ? ld_stream(&C.perm[i]) : -1; }
        double zv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) zv[u] = (pi[u] >= 0) ? in(pi[u]) : 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int i = i0 + u * T.nthreads; if (i < N) W.sv[i] = zv[u]; }
    }
    // staging slots of the merged chains: zero when a solve starts (C.yoff + N = end of the sweep values)
    for (int i = T.tid; i < kCwStage; i += T.nthreads) W.sv[C.yoff + N + i] = 0.0;
    __syncthreads();
    lap(0);
    // The walk itself is done by the first kCwWarps warps (a level is one warp's dependent
    // instruction stream: more warps only add barrier and issue pressure); they synchronise on a
    // named barrier of their own, the other warps wait for them at the CTA barrier below.
    auto walkers_sync = [] { asm volatile("bar.sync 1, %0;" ::"n"(32 * kCwWarps) : "memory"); };
    if (warp < kCwWarps) {
        for (int b = 0; b < nblk; ++b) {
            const unsigned k = k0 + (unsigned)b;
            if (!W.wait(T, k)) break;           // watchdog fired: the abort flag is set, results are void
            lap(1);
            const unsigned char *blk = W.ring + (size_t)(k % kCwStages) * kCwBlock;
            const int nsteps = *reinterpret_cast<const int *>(blk);
            const int4 *slot = reinterpret_cast<const int4 *>(blk + 16) + warp;
            CwTask cur;
            cw_fetch(cur, blk, slot[0], lane);
            bool synced = false;
            for (int st = 0; st < nsteps; ++st) {
                const int4 tk = slot[min(st + 1, nsteps - 1) * kCwWarps];   // next slot: its latency hides behind the arithmetic
                if (cur.kind == CW_DCHUNK) cw_dchunk(cur, W.sv, C.yoff, warp, lane);
                else if (cur.kind != 0) cw_rows(cur, W.sv, lane);
                synced = cur.barrier;
                if (st + 1 < nsteps) cw_fetch(cur, blk, tk, lane);
                if (synced) walkers_sync();
            }
            if (!synced) walkers_sync();        // every walker is done with this ring entry
            if (T.tid == 0) W.issue(C, k + (unsigned)kCwStages);
            lap(2);
        }
    }
    __syncthreads();
    // out = P * y
    for (int i0 = T.tid; i0 < N; i0 += 8 * T.nthreads) {
        int pi[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int i = i0 + u * T.nthreads; pi[u] = (i < N) ? ld_stream(&C.perm[i]) : -1; }
        if (accumulate) {
            double ov[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) ov[u] = (pi[u] >= 0) ? out[pi[u]] : 0.0;
#pragma unroll
            for (int u = 0; u < 8; ++u) if (pi[u] >= 0) out[pi[u]] = ov[u] + W.sv[C.yoff + i0 + u * T.nthreads];
        } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) if (pi[u] >= 0) out[pi[u]] = W.sv[C.yoff + i0 + u * T.nthreads];
        }
    }
    if (T.tid == 0) *W.count = k0 + (unsigned)nblk;
    lap(3);
    if (dbg && T.tid == 0) for (int q = 0; q < 4; ++q) dbg[q] = cyc[q];
}

// y = P * L^-T * D^-1 * L^-1 * P' * in   (opLDL2.m:86, right to left).
// If `accumulate`, out[p] += y (the y = y + dy of opLDL2.m:181).  Ends WITHOUT a
// team barrier: the caller syncs before `out` is gathered.
template <class Team>
__device__ __forceinline__ void ldl_solve(Team &T, const DevLdl &M, const VecIn in, double *out, bool accumulate, int epoch)
{
    if (Team::kKind == 1 && M.cw.smem_off >= 0) ldl_solve_compact(T, M, in, out, accumulate);
    else if (M.use_rc) ldl_solve_rc(T, M, in, out, accumulate);
    else if (M.sync_free) ldl_solve_syncfree(T, M, in, out, accumulate, epoch);
    else ldl_solve_levels(T, M, in, out, accumulate, (PhaseClock *)nullptr);
}

// ---------------------------------------------------------------------------
// y = M * xin          opLDL2.multiply, opLDL2.m:161-188
// Entry: xin complete and visible to the team (caller synced).
// Exit : y complete and visible (ends with a team barrier).
// ---------------------------------------------------------------------------
// Returns true when `*rider_sum` holds the rider's sum over the FINAL y (i.e. the
// first residual pass was also the last touch of y); false otherwise.
template <class Team, class Rider>
__device__ __noinline__ bool ldl2_apply(Team &T, const DevLdl &M, const VecIn xin0, double *y, int &epoch,
                           DevStatus *st, PhaseClock &pc, Rider rider, double *rider_sum)
{
    bool rider_valid = false;
    VecIn xin = xin0;
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
        double red[3];
        // The residual VECTOR is only read when a refinement step follows (rare with the reference's
        // defaults): the first pass then keeps the norms only, and the vector is produced -- same
        // operations, same bits -- by a second pass in front of the step (CPK_LAZY_RVEC).
        const bool lazy_r = CPK_LAZY_RVEC && !M.force_itref;
        bool r_stored = !lazy_r;
        resid_phase(T, M, xin, y, lazy_r ? (double *)nullptr : M.rvec, red[0], red[1], true, rider, red[2]);
        if (Rider::kActive) T.template reduce<3>(red); else { double r2[2] = {red[0], red[1]}; T.template reduce<2>(r2); red[0] = r2[0]; red[1] = r2[1]; }
        rider_valid = Rider::kActive;
        if (rider_sum) *rider_sum = red[2];
        if (xin.wb) { xin.add = nullptr; xin.wb = nullptr; }           // z now holds the evaluated vector
        ++nres;
        pc.mark(CPK_PH_RESID_);
        double rNorm = sqrt(red[0]);
        const double xNorm = sqrt(red[1]);
        int nit = 0;
        bool rknown = true;
        while (nit < M.nitref && (rNorm >= M.itref_tol * xNorm || M.force_itref)) {   // :179
            if (!r_stored) {
                double d0, d1, d2;
                NoRider nr;
                resid_phase(T, M, xin, y, M.rvec, d0, d1, false, nr, d2);
                T.sync();
                r_stored = true;
                pc.mark(CPK_PH_RESID_);
            }
            VecIn rin{M.rvec, nullptr, n, false};
            rider_valid = false;                                        // y changes: the rider's sum is stale
            ldl_solve(T, M, rin, y, true, ++epoch);                     // dy = LDL*r; y = y + dy
            T.sync();
            ++nsolve;
            ++nit;
            pc.mark(CPK_PH_LDL_);
            if (nit < M.nitref || M.track_rnorm) {
                double r1[1], dummy, dummy2;
                NoRider nr;
                resid_phase(T, M, xin, y, M.rvec, r1[0], dummy, false, nr, dummy2);
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
    return rider_valid;
}

template <class Team>
__device__ __forceinline__ void ldl2_apply(Team &T, const DevLdl &M, const VecIn xin, double *y, int &epoch,
                                           DevStatus *st, PhaseClock &pc)
{
    ldl2_apply(T, M, xin, y, epoch, st, pc, NoRider(), (double *)nullptr);
}

}  // namespace cpk
