// cpk_rc.cuh -- passes over matrices in row-class form (DevRc, cpk_device.cuh): the levels of a
// shallow LDL' sweep, the refinement residual x - K_P*y and the `divide` product K_P*b.
// Included by cpk_kernels.cuh after VecIn.
//
// What bounds such a pass is the number of DEPENDENT memory round trips a warp makes, so the
// code is written to a strict discipline:
//   1. every load of a batch is issued before the first loaded value is used.  A load that sits
//      behind a data-dependent branch breaks this (the warp waits at the first use inside the
//      branch before it reaches the next row's loads), so all loads are unconditional or simple
//      selects -- dead lanes read element 0, absent operands read a substitute -- and only the
//      stores are predicated;
//   2. everything a load needs besides the streamed data (vector base pointers, modes) lives in
//      registers: the op structs are built inside the force-inlined pass, never reached through
//      a pointer to local memory;
//   3. the shape of the input vector (plain / + scale*add / - sub) is a template parameter, so a
//      gather is one load in the common case.
#pragma once

namespace cpk {

// ---------------------------------------------------------------------------
// input vector of an apply, evaluated in two steps: issue (loads) and value (arithmetic)
//   value(i) = sgn(i) * (z[i] + scale*add[i]) - sub[i]        (see VecIn)
// MODE bit 0: `add` present, bit 1: `sub` present
// ---------------------------------------------------------------------------
template <int MODE> struct VecRaw { double z, a, s; };
template <> struct VecRaw<0> { double z; };
template <> struct VecRaw<1> { double z, a; };

template <int MODE>
struct VecEval {
    const double *z, *add, *sub;
    double scale;
    int nA;
    bool neg, has_sub;
    // an absent operand of a wider mode is replaced by z itself (a valid address) and switched off
    // in the arithmetic: scale = 0 for `add`, has_sub = false for `sub`
    __device__ __forceinline__ explicit VecEval(const VecIn &v)
        : z(v.z), add(v.add ? v.add : v.z), sub(v.sub ? v.sub : v.z), scale(v.add ? v.scale : 0.0), nA(v.nA), neg(v.neg_tail),
          has_sub(v.sub != nullptr) {}
    // `on`: the value will be used (false: the operand comes from elsewhere; only z is read, at a valid index)
    __device__ __forceinline__ VecRaw<MODE> issue(int i, bool on = true) const {
        VecRaw<MODE> r;
        r.z = z[i];
        if constexpr ((MODE & 1) != 0) r.a = on ? add[i] : 0.0;
        if constexpr ((MODE & 2) != 0) r.s = on ? sub[i] : 0.0;
        return r;
    }
    __device__ __forceinline__ double value(int i, const VecRaw<MODE> &r) const {
        double v = r.z;
        if constexpr ((MODE & 1) != 0) v = v + scale * r.a;
        if (neg && i >= nA) v = -v;
        if constexpr ((MODE & 2) != 0) { if (has_sub) v = v - r.s; }
        return v;
    }
};
__device__ __forceinline__ int vec_mode(const VecIn &v) { return (v.add ? 1 : 0) | (v.sub ? 2 : 0); }

// ---------------------------------------------------------------------------
// The pass.  `Op` supplies
//     struct Pre, struct Raw                     row inputs / one gathered operand, as loaded
//     Pre    pre(int pos, int code)              loads only (code >= 0)
//     Raw    gather(int colcode)                 loads only
//     double value(int colcode, const Raw &)     arithmetic
//     void   fin(int pos, int code, const Pre &, double s)     s = sum_j val_j * value_j, storage order
// ---------------------------------------------------------------------------
// streamed loads as volatile asm: the compiler may neither sink them into a later data-dependent
// branch nor reorder them -- they are issued where they are written, all at the top of a batch
__device__ __forceinline__ int ldg_stream(const int *p) {
    int v;
    asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ldg_stream(const double *p) {
    double v;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

struct RcPtr {                      // the streamed arrays, in registers
    const int *rowmap, *col;
    const double *val;
    __device__ __forceinline__ explicit RcPtr(const DevRc &R) : rowmap(R.rowmap), col(R.col), val(R.val) {}
};

// the streamed data of one batch of B groups of width W, in registers
template <int W, int B>
struct RcStream {
    int code[B], cc[B][W > 0 ? W : 1];
    double vv[B][W > 0 ? W : 1];
    __device__ __forceinline__ void load(const RcPtr A, int row_off, int ent_off, int stride, int ng, int batch, int lane) {
        const int ga = batch * B;
#pragma unroll
        for (int q = 0; q < B; ++q) {
            const bool on = ga + q < ng;                    // warp-uniform; a surplus slot re-reads the batch's first group
            const int g = on ? ga + q : ga;
            code[q] = ldg_stream(&A.rowmap[row_off + g * 32 + lane]);
            if (!on) code[q] = -1;
#pragma unroll
            for (int j = 0; j < W; ++j) {
                cc[q][j] = ldg_stream(&A.col[ent_off + j * stride + g * 32 + lane]);
                vv[q][j] = ldg_stream(&A.val[ent_off + j * stride + g * 32 + lane]);
            }
        }
    }
};

// batches [a, b) of a piece of width W, B groups of 32 rows each.  The streamed loads of batch
// j+1 are issued before the row inputs and gathers of batch j are consumed (software pipeline
// in registers): one exposed round trip per batch instead of two.
template <int W, int B, class Op>
__device__ __forceinline__ void rc_run_fixed(const RcPtr A, const RcPiece pc, int a, int b, int lane, Op &op)
{
    constexpr int WW = W > 0 ? W : 1;
    const int ng = pc.wn >> 8, stride = ng * 32;
    RcStream<W, B> cur;
    cur.load(A, pc.row_off, pc.ent_off, stride, ng, a, lane);
    for (int j = a; j < b; ++j) {
        RcStream<W, B> nxt;
        if (j + 1 < b) nxt.load(A, pc.row_off, pc.ent_off, stride, ng, j + 1, lane);
        else nxt = cur;
        // ---- row inputs and gathers of the whole batch, loads only -------------------------
        typename Op::Pre P[B];
        typename Op::Raw G[B][WW];
#pragma unroll
        for (int q = 0; q < B; ++q) {
            const int g = j * B + q < ng ? j * B + q : j * B;
            P[q] = op.pre(pc.row_off + g * 32 + lane, cur.code[q] < 0 ? 0 : cur.code[q]);      // dead lanes read row 0 (never finished)
#pragma unroll
            for (int w = 0; w < W; ++w) G[q][w] = op.gather(cur.cc[q][w]);     // padding lanes hold column 0: a valid address
        }
        // ---- arithmetic and stores -------------------------------------------------------------
#pragma unroll
        for (int q = 0; q < B; ++q) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < W; ++w) s += cur.vv[q][w] * op.value(cur.cc[q][w], G[q][w]);
            if (cur.code[q] >= 0) op.fin(pc.row_off + (j * B + q) * 32 + lane, cur.code[q], P[q], s);
        }
        cur = nxt;
    }
}

// any width: groups [a, b), entries in chunks of CH per lane; the columns of the next chunk are
// requested as soon as the gathers of the current one are issued
template <class Op>
__device__ __forceinline__ void rc_run_wide(const RcPtr A, const RcPiece pc, int a, int b, int lane, Op &op)
{
    constexpr int CH = CPK_RC_CH;
    const int W = pc.wn & 255, stride = (pc.wn >> 8) * 32;
    for (int g = a; g < b; ++g) {
        const int pos = pc.row_off + g * 32 + lane;
        const int e0 = pc.ent_off + g * 32 + lane;
        const int code = ldg_stream(&A.rowmap[pos]);
        int c[CH]; double v[CH];
#pragma unroll
        for (int u = 0; u < CH; ++u) {
            const int j = u < W ? u : 0;                    // W >= 4 here
            c[u] = ldg_stream(&A.col[e0 + j * stride]); v[u] = ldg_stream(&A.val[e0 + j * stride]);
        }
        const typename Op::Pre P = op.pre(pos, code < 0 ? 0 : code);
        double s = 0.0;
        for (int j0 = 0; j0 < W; j0 += CH) {
            typename Op::Raw G[CH];
            int cn[CH]; double vn[CH];
#pragma unroll
            for (int u = 0; u < CH; ++u) G[u] = op.gather(c[u]);
#pragma unroll
            for (int u = 0; u < CH; ++u) {
                const int j = j0 + CH + u < W ? j0 + CH + u : 0;
                cn[u] = ldg_stream(&A.col[e0 + j * stride]); vn[u] = ldg_stream(&A.val[e0 + j * stride]);
            }
#pragma unroll
            for (int u = 0; u < CH; ++u) if (j0 + u < W) s += v[u] * op.value(c[u], G[u]);
#pragma unroll
            for (int u = 0; u < CH; ++u) { c[u] = cn[u]; v[u] = vn[u]; }
        }
        if (code >= 0) op.fin(pos, code, P, s);
    }
}

// long rows [a, b): one warp per row, lane-strided partial sums + butterfly
template <class Op>
__device__ __forceinline__ void rc_run_long(const DevRc &R, const RcPiece pc, int a, int b, int lane, Op &op)
{
    for (int r = a; r < b; ++r) {
        const int pos = pc.row_off + r;
        const int code = ld_stream(&R.rowmap[pos]);             // warp-uniform, >= 0
        const int beg = ld_stream(&R.lptr[pc.ent_off + r]), end = ld_stream(&R.lptr[pc.ent_off + r + 1]);
        const typename Op::Pre P = op.pre(pos, code);
        double s = 0.0;
        for (int k = beg + lane; k < end; k += 64) {
            const bool two = k + 32 < end;
            const int c0 = ld_stream(&R.lcol[k]); const double v0 = ld_stream(&R.lval[k]);
            const int c1 = ld_stream(&R.lcol[two ? k + 32 : k]); const double v1 = ld_stream(&R.lval[two ? k + 32 : k]);
            const typename Op::Raw g0 = op.gather(c0), g1 = op.gather(c1);
            s += v0 * op.value(c0, g0);
            if (two) s += v1 * op.value(c1, g1);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
        if (lane == 0) op.fin(pos, code, P, s);
    }
}

// One level.  The batches of all its pieces form one sequence, each weighted by its cost in round
// trips (prefix sums in the piece table); warp w takes the batches that start inside the cost range
// [T*w/nwarps, T*(w+1)/nwarps): equal numbers of round trips -- not bytes -- for every warp, and
// consecutive batches of one piece, so that rc_run_fixed can pipeline them.
template <class Op>
__device__ __forceinline__ void rc_level(const DevRc &R, int lev, int gwarp, int nwarps, int lane, Op &op)
{
    const RcPtr A(R);
    const unsigned T = (unsigned)R.levb[lev];
    // floor(T * w / nwarps) without 64-bit division (remainder * w < 2^26)
    const unsigned q = T / (unsigned)nwarps, rem = T % (unsigned)nwarps;
    const int lo = (int)(q * (unsigned)gwarp + rem * (unsigned)gwarp / (unsigned)nwarps);
    const int hi = (int)(q * (unsigned)(gwarp + 1) + rem * (unsigned)(gwarp + 1) / (unsigned)nwarps);
    if (lo >= hi) return;
    const int p1 = R.levp[lev + 1];
    for (int p = R.levp[lev]; p < p1; ++p) {
        const RcPiece pc = R.piece(p);
        if (pc.cum >= hi) break;
        const int width = pc.wn & 255, ng = pc.wn >> 8;
        const int B = rc_batch_of(width), c = rc_cost_of(width);
        const int nb = (ng + B - 1) >> (B >> 1);            // B = 1, 2, 4
        // batch j of the piece starts at cost cum + j*c: the warp takes those that start in [lo, hi)
        const int a = lo > pc.cum ? (lo - pc.cum + c - 1) / c : 0;
        const int b = min((hi - pc.cum + c - 1) / c, nb);
        if (a >= b) continue;
        switch (width) {
            case 0: rc_run_fixed<0, rc_batch_of(0)>(A, pc, a, b, lane, op); break;
            case 1: rc_run_fixed<1, rc_batch_of(1)>(A, pc, a, b, lane, op); break;
            case 2: rc_run_fixed<2, rc_batch_of(2)>(A, pc, a, b, lane, op); break;
            case 3: rc_run_fixed<3, rc_batch_of(3)>(A, pc, a, b, lane, op); break;
            case kRcLongW: rc_run_long(R, pc, a, b, lane, op); break;
            default: rc_run_wide(A, pc, a, b, lane, op);
        }
    }
}

// ---------------------------------------------------------------------------
// Shallow sweeps in row-class form (DevLdl::rc): one rc_level per dependency level, a team
// barrier between levels.  wv / yv / out / the input are all indexed by the USER index of a
// row (its code), so there is no row-id indirection; D is stored per position.
//   forward row : w = (P'z)_i - sum;  fused: y = w / d -> out (and yv when a backward row needs it)
//   backward row: y = w / d - sum     (w from wv, or straight from the input for rows without
//                                      forward work) -> out
// Arithmetic as in the item walks (opLDL2.m:86, right to left).
// ---------------------------------------------------------------------------
template <bool FWD, bool ACC, int MODE>
struct SweepOp {
    VecEval<MODE> in;
    const double *dpos;             // D per position
    double *wy, *out;               // [w (N) | y (N)], both halves indexed by the user index
    int N;
    struct Raw { VecRaw<MODE> r; };
    struct Pre { VecRaw<MODE> b; double dd, o; };
    __device__ __forceinline__ Raw gather(int c) const {
        const bool isin = (c >> RC_IDX_BITS) != 0;
        const int i = c & RC_IDX_MASK;
        Raw g;
        const double *p = isin ? in.z : wy;
        g.r.z = p[i];
        if constexpr ((MODE & 1) != 0) g.r.a = isin ? in.add[i] : 0.0;
        if constexpr ((MODE & 2) != 0) g.r.s = isin ? in.sub[i] : 0.0;
        return g;
    }
    __device__ __forceinline__ double value(int c, const Raw &g) const {
        return (c >> RC_IDX_BITS) != 0 ? in.value(c & RC_IDX_MASK, g.r) : g.r.z;
    }
    __device__ __forceinline__ Pre pre(int pos, int code) const {
        const int i = code & RC_IDX_MASK, fl = code >> RC_IDX_BITS;
        Pre P;
        if (FWD) P.b = in.issue(i);
        else {
            const bool direct = (fl & RC_F_WDIRECT) != 0;
            const double *p = direct ? in.z : wy;
            P.b.z = p[i];
            if constexpr ((MODE & 1) != 0) P.b.a = direct ? in.add[i] : 0.0;
            if constexpr ((MODE & 2) != 0) P.b.s = direct ? in.sub[i] : 0.0;
        }
        P.dd = ld_stream(&dpos[pos]);
        P.o = ACC ? out[i] : 0.0;
        return P;
    }
    __device__ __forceinline__ void fin(int, int code, const Pre &P, double s) const {
        const int i = code & RC_IDX_MASK, fl = code >> RC_IDX_BITS;
        if (FWD) {
            double acc = in.value(i, P.b) - s;
            if (fl & RC_F_FUSED) {
                acc = acc / P.dd;                       // column of L is empty: y_i = w_i / d_i
                if (fl & RC_F_STOREY) wy[N + i] = acc;
                out[i] = ACC ? P.o + acc : acc;
            } else wy[i] = acc;
        } else {
            const double w = (fl & RC_F_WDIRECT) ? in.value(i, P.b) : P.b.z;
            double acc = w / P.dd;                      // opLDL2.m:86, inv(op.D)
            acc = acc - s;
            if (fl & RC_F_STOREY) wy[N + i] = acc;
            out[i] = ACC ? P.o + acc : acc;
        }
    }
};

template <bool ACC, int MODE, class Team>
__device__ __forceinline__ void ldl_solve_rc_mode(Team &T, const DevLdl &M, const VecIn &in, double *out)
{
    const DevRc &R = M.rc;
    const int lane = T.lane, gwarp = T.gwarp, nwarps = T.nwarps;      // T lives in local memory: read once
    const int nlev = R.nlev, nfwd = R.nfwd_lev;
    for (int lev = 0; lev < nlev; ++lev) {
        if (lev < nfwd) {
            SweepOp<true, ACC, MODE> op{VecEval<MODE>(in), R.d, M.wv, out, M.N};
            rc_level(R, lev, gwarp, nwarps, lane, op);
        } else {
            SweepOp<false, ACC, MODE> op{VecEval<MODE>(in), R.d, M.wv, out, M.N};
            rc_level(R, lev, gwarp, nwarps, lane, op);
        }
        if (lev + 1 < nlev) T.sync();           // the caller syncs after the last level
    }
}

template <class Team>
__device__ __noinline__ void ldl_solve_rc(Team &T, const DevLdl &M, const VecIn in, double *out, bool accumulate)
{
    const int mode = vec_mode(in);
    if (accumulate) {
        // the correction solve of the refinement: its input is the plain residual vector
        if (mode == 0) ldl_solve_rc_mode<true, 0>(T, M, in, out);
        else ldl_solve_rc_mode<true, 3>(T, M, in, out);
    } else {
        if (mode == 0) ldl_solve_rc_mode<false, 0>(T, M, in, out);
        else if (mode == 1) ldl_solve_rc_mode<false, 1>(T, M, in, out);
        else ldl_solve_rc_mode<false, 3>(T, M, in, out);
    }
}

// ---------------------------------------------------------------------------
// r = xin - K*y with partial sums of r'r (and xin'xin): opLDL2.m:175-177,182-183
// A rider lets the caller piggy-back one more sum over the rows on this pass (e.g. the
// P-inner product a solver needs right after the apply): `pre(row)` loads what the rider
// needs for the row (loads only: it runs with the other row inputs, before the gathers are
// consumed), `fin(row, xin_row, y_row, pre)` returns the row's contribution (and may store).
// ---------------------------------------------------------------------------
struct NoRider {
    static constexpr bool kActive = false;
    struct Pre {};
    __device__ __forceinline__ Pre pre(int) const { return Pre(); }
    __device__ __forceinline__ double fin(int, double, double, const Pre &) const { return 0.0; }
};

template <class Rider, int MODE>
struct ResidOp {
    VecEval<MODE> xin;
    const double *y;
    double *r, *wb;
    bool want_xx;
    Rider rider;
    double rr, xx, extra;
    struct Raw { double v; };
    struct Pre { VecRaw<MODE> x; double yi; typename Rider::Pre rp; };
    __device__ __forceinline__ Raw gather(int c) const { return Raw{y[c]}; }
    __device__ __forceinline__ double value(int, const Raw &g) const { return g.v; }
    __device__ __forceinline__ Pre pre(int, int row) const {
        Pre P;
        P.x = xin.issue(row);
        P.yi = Rider::kActive ? y[row] : 0.0;
        P.rp = rider.pre(row);
        return P;
    }
    __device__ __forceinline__ void fin(int, int row, const Pre &P, double s) {
        const double xi = xin.value(row, P.x);
        if (wb) wb[row] = xi;                   // pending axpy of the caller lands in memory here
        const double ri = xi - s;
        if (r) r[row] = ri;
        rr += ri * ri;
        if (want_xx) xx += xi * xi;
        if (Rider::kActive) extra += rider.fin(row, xi, P.yi, P.rp);
    }
};

template <int MODE, class Team, class Rider>
__device__ __forceinline__ void resid_phase_mode(Team &T, const DevLdl &M, const VecIn &xin, const double *y,
                                                 double *r, double &rr, double &xx, bool want_xx, Rider &rider, double &extra)
{
    ResidOp<Rider, MODE> op{VecEval<MODE>(xin), y, r, xin.wb, want_xx, rider, 0.0, 0.0, 0.0};
    rc_level(M.KP, 0, T.gwarp, T.nwarps, T.lane, op);
    rr = op.rr; xx = op.xx; extra = op.extra;
}
template <class Team, class Rider>
__device__ __forceinline__ void resid_phase(Team &T, const DevLdl &M, const VecIn xin, const double *y,
                                            double *r, double &rr, double &xx, bool want_xx, Rider &rider, double &extra)
{
    if (!M.resid_rc) {
        // K_P in SELL form: the streaming loop of the sparse mat-vec, row work in its epilogue
        rr = 0.0; xx = 0.0; extra = 0.0;
        spmv_sell(T, M.KPs, y, [&](int row, double s) {
            const double xi = xin(row);
            if (xin.wb) xin.wb[row] = xi;       // pending axpy of the caller lands in memory here
            const double ri = xi - s;
            if (r) r[row] = ri;                 // r == nullptr: norms only (ldl2_apply stores r when a step is taken)
            rr += ri * ri;
            if (want_xx) xx += xi * xi;
            if (Rider::kActive) extra += rider.fin(row, xi, y[row], rider.pre(row));
        });
        return;
    }
    const int mode = vec_mode(xin);
    if (mode == 0) resid_phase_mode<0>(T, M, xin, y, r, rr, xx, want_xx, rider, extra);
    else if (mode == 1) resid_phase_mode<1>(T, M, xin, y, r, rr, xx, want_xx, rider, extra);
    else resid_phase_mode<3>(T, M, xin, y, r, rr, xx, want_xx, rider, extra);
}

// y = K*x for a matrix in row-class form (the `divide` path of opLDL2, opLDL2.m:193-195)
struct MatvecOp {
    const double *x;
    double *y;
    struct Raw { double v; };
    struct Pre {};
    __device__ __forceinline__ Raw gather(int c) const { return Raw{x[c]}; }
    __device__ __forceinline__ double value(int, const Raw &g) const { return g.v; }
    __device__ __forceinline__ Pre pre(int, int) const { return Pre(); }
    __device__ __forceinline__ void fin(int, int row, const Pre &, double s) const { y[row] = s; }
};

}  // namespace cpk
