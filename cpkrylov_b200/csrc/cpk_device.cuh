// cpk_device.cuh -- device-side data structures, the "team" abstraction
// (whole cooperative grid, or one CTA) and the memory-model primitives used by
// the persistent solver kernels.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cpk {

// Threads per CTA (1 CTA per SM).  Every streaming phase of the persistent kernel is bound by the
// warps an SM has in flight, not by registers per thread: 896 threads (28 warps, 72 registers) with
// small per-thread batches (SpMV 4 entries per lane in flight, sweep batches of 2 items) measured
// 13 % faster on cfg 3 than 512 threads (128 registers, 8 entries, batches of 4); 768 / 832 / 960 /
// 1024 threads and other batch sizes are slower (profiles/r1_notes.md).
#ifndef CPK_BLOCK
#define CPK_BLOCK 896
#endif
constexpr int kBlock      = CPK_BLOCK;
constexpr int kWarpsPerCta = kBlock / 32;
constexpr int kCtasPerSm  = kBlock >= 512 ? 1 : 512 / kBlock;
constexpr int kRedMax     = 8;      // values reduced together in one team reduction
constexpr unsigned FULL   = 0xffffffffu;

// phase slots of DevStatus::phase_cycles (mirror CPK_PH_* of cpk_b200.h)
constexpr int CPK_PH_SPMV_ = 0, CPK_PH_LDL_ = 1, CPK_PH_RESID_ = 2, CPK_PH_VEC_ = 3, CPK_PH_OTHER_ = 4;
// device-side copies of cpk_status codes
constexpr int CPK_ERR_INDEFINITE_ = -5, CPK_ERR_BREAKDOWN_ = -6, CPK_ERR_TIMEOUT_ = -7;

// ---------------------------------------------------------------------------
// Cache hints.  A cfg-3 iteration streams ~130 MB of matrix and item arrays past ~60 MB of Krylov
// vectors that the NEXT phase gathers from; with plain loads the streams push the vectors out of
// the L2 and every gather round trip goes to DRAM.
//   ld_stream  streamed, read-once-per-pass arrays (entries, row maps, item data): ld.global.cs,
//              evict-first in L1 and L2 (CPK_STREAM_HINT=0: ld.global.nc)
//   ld_keep    a matrix small enough to survive from one iteration to the next (packed stencil
//              entries of H, 4 bytes per entry): ld.global.nc, normal priority
//   st_tmp / ld_tmp   sweep intermediates (w of the forward sweep, y of the backward sweep):
//              written once, read once by the next level (CPK_TMP_HINT=0: plain accesses)
// Measured on cfg 3 (profiles/r2_notes.md, "cache hints"): 3.164 -> 3.03 ms per solve together
// with the lazy residual vector; an evict-last policy on H, evict-first on H, and evict-first
// loads of vectors at their last use were each slower.
// ---------------------------------------------------------------------------
#ifndef CPK_STREAM_HINT
#define CPK_STREAM_HINT 1
#endif
#if CPK_STREAM_HINT
template <class V> __device__ __forceinline__ V ld_stream(const V *p) { return __ldcs(p); }
#else
template <class V> __device__ __forceinline__ V ld_stream(const V *p) { return __ldg(p); }
#endif
template <class V> __device__ __forceinline__ V ld_keep(const V *p) { return __ldg(p); }
#ifndef CPK_TMP_HINT
#define CPK_TMP_HINT 1
#endif
#if CPK_TMP_HINT
__device__ __forceinline__ void st_tmp(double *p, double v) { __stcs(p, v); }
__device__ __forceinline__ double ld_tmp(const double *p) { return __ldcs(p); }
#else
__device__ __forceinline__ void st_tmp(double *p, double v) { *p = v; }
__device__ __forceinline__ double ld_tmp(const double *p) { return *p; }
#endif

// ---------------------------------------------------------------------------
// matrices
// ---------------------------------------------------------------------------
// SELL-32 (sliced ELLPACK, slice height = warp) with a per-lane row map, plus a
// CSR remainder for rows too long to pad ("warp per row").  Immutable for the
// lifetime of a kernel, so loads go through the non-coherent path.
struct DevSell {
    int nrows, ncols;
    int nslices;
    const int    *sptr;     // [nslices+1] element offset of each slice (multiple of 32)
    const int    *col;      // padded entries: val = 0, col = a valid column of that row (or 0)
    const double *val;
    const int    *rowmap;   // [nslices*32] output row of each lane, -1 = no row
    // contiguous slice range of every warp, balanced by padded entries, for the
    // two team shapes (0: cooperative grid, 1: one CTA): [nws[k]+1] each
    const int    *wsplit[2];
    int nws[2];
    int nlong;              // rows handled warp-per-row
    const int    *lrow;     // [nlong] row index
    const int    *lptr;     // [nlong+1]
    const int    *lcol;
    const double *lval;
    // Packed form of the padded entries (nullptr: not packed, col / val above are used; otherwise
    // col / val are NOT uploaded): one 32-bit word per entry,
    //     bits 0-15   column - row, signed (banded matrices: |column - row| < 32768)
    //     bits 16-31  index into `dict`, the table of the matrix' distinct values (< 65536 of them)
    // 4 bytes per entry instead of 12: lossless, chosen at upload when the matrix allows it
    // (stencil matrices such as the H of the BASELINE configs: 2 resp. 7 distinct values).
    const unsigned *pk;
    const double   *dict;
    int ndict;
};

// ---------------------------------------------------------------------------
// Row-class layout ("RC") for passes over SHORT rows with a heavy per-row epilogue: the levels
// of a shallow LDL' sweep and the refinement residual x - K_P*y.  The rows of one pass (level)
// are sorted by their entry count; class w holds the rows with exactly w entries as a dense
// rectangle
//     rowmap[ngroups*32], d[ngroups*32], col[w][ngroups*32], val[w][ngroups*32]
// (rows of a class in ascending row order, last group padded with rowmap = -1).  No slice
// pointers and no padding entries: every address a row needs follows from its position, so a
// thread issues ALL streamed loads of a batch of rows at once (one round trip), then all
// gathers (second round trip) -- the walk over SELL items needs three dependent trips and
// 20 bytes of per-lane item metadata.  Rows with more than `maxw` entries form a CSR list of
// their own (one warp per row).
// A level is a run of pieces (one per class present).  Its batches (a few consecutive groups of
// one piece) form one sequence, cut into contiguous ranges of equal COST for the warps of the team
// from the prefix sums in the piece table -- no per-team split arrays.  Such a pass is bound by
// dependent round trips per warp, so the cost of a batch is its number of round trips, not bytes.
// A warp's consecutive batches mostly lie in one piece, so the streamed loads of the next batch
// are issued while the gathers of the current one are in flight.
// ---------------------------------------------------------------------------
struct RcPiece { int wn, cum, row_off, ent_off; };  // wn = width | ngroups << 8; cum = cost of the level's batches before this piece
constexpr int kRcLongW = 255;       // width code of the long-row list (one row per "group", ent_off = first row in lptr)
constexpr int kRcInline = 48;       // pieces kept inside the descriptor (shared-memory copy); the rest is read from global memory
constexpr int kRcMaxLev = 48;
constexpr int RC_IDX_BITS = 28;     // row / column codes: index in the low 28 bits
constexpr int RC_IDX_MASK = (1 << RC_IDX_BITS) - 1;
// column codes of a sweep: bit 28 set = element of the INPUT vector; otherwise an index into the
// sweep's value buffer [w (N) | y (N)], both halves indexed by the user index of a row
constexpr int RC_SRC_IN = 1;
// row codes of a sweep: user index + flags in bits 28-29 (-1 = padding lane)
constexpr int RC_F_FUSED = 1;       // forward row whose column of L is empty: y = w/d, written to the output
constexpr int RC_F_WDIRECT = 1;     // backward row without forward work: w = (P'z)_i read from the input
constexpr int RC_F_STOREY = 2;      // y_i is gathered by a later level: keep it in the value buffer
// groups of 32 rows per batch for the widths 0, 1, 2, 3 (one digit each, powers of two; wider rows: 1).
// Host (batch prefix sums of the piece table) and device must agree, hence one table for all passes.
#ifndef CPK_RC_B
#define CPK_RC_B 4221
#endif
__host__ __device__ constexpr int rc_batch_of(int width)
{
    return width == 0 ? CPK_RC_B / 1000 % 10 : width == 1 ? CPK_RC_B / 100 % 10 : width == 2 ? CPK_RC_B / 10 % 10 : width == 3 ? CPK_RC_B % 10 : 1;
}
static_assert(((CPK_RC_B / 1000 % 10) | (CPK_RC_B / 100 % 10) | (CPK_RC_B / 10 % 10) | (CPK_RC_B % 10)) <= 7, "CPK_RC_B digits: 1, 2 or 4");
// entries per lane a wide row (more than 3 entries) has in flight at a time
#ifndef CPK_RC_CH
#define CPK_RC_CH 4
#endif
// Cost of one batch in dependent round trips, the unit the warps' shares of a level are measured
// in: a batch of short rows is one round trip once the pipeline of its piece runs, a group of wide
// rows one per chunk of CPK_RC_CH entries plus the row codes, a long row three.
__host__ __device__ constexpr int rc_cost_of(int width)
{
    return width <= 3 ? 1 : width == kRcLongW ? 3 : 1 + (width + CPK_RC_CH - 1) / CPK_RC_CH;
}

struct DevRc {
    int nlev, npieces;
    int nfwd_lev;                   // sweeps: levels [0, nfwd_lev) are forward levels, the rest backward
    short levp[kRcMaxLev + 2];      // [nlev+1] first piece of every level
    int   levb[kRcMaxLev + 1];      // [nlev] cost of every level (sum of rc_cost_of over its batches)
    RcPiece inl[kRcInline];
    const RcPiece *pieces;          // [npieces]
    const int    *rowmap;           // row code per position
    const double *d;                // sweeps: D(i,i) per position; unused (nullptr) for plain matrices
    const int    *col;
    const double *val;
    const int    *lptr;             // long rows
    const int    *lcol;
    const double *lval;
    __device__ __forceinline__ RcPiece piece(int p) const {
        if (p < kRcInline) return inl[p];
        const int4 v = __ldg(reinterpret_cast<const int4 *>(pieces) + p);
        return RcPiece{v.x, v.y, v.z, v.w};
    }
};

// A value and the tag (solve epoch) that says it is ready.  16-byte aligned so
// that ld/st.relaxed.gpu.b128 moves both in one single-copy-atomic access: no
// fence is needed between "data" and "flag" because they are the same word.
struct __align__(16) Tagged {
    double v;
    unsigned long long tag;
};

// The LDL' solve as ONE ordered list of "items".  An item is a SELL slice of
// <= 32 rows of the same dependency level, processed by one warp; forward items
// (rows of L) come first, backward items (rows of L', D-solve fused in) follow.
// Rows that need no work are compiled away at setup:
//   * a row of L with no off-diagonal entry has w_i = (P'z)_i: it gets NO forward
//     item, and whoever needs w_i reads the input vector directly (col <= -2);
//   * a row whose column of L is empty has y_i = w_i/d_i: finished inside its
//     forward item (F_FUSED), it gets NO backward item.
// Column indices >= 0 address the tagged buffer of the sweep (LDL row ids),
// -1 is padding, c <= -2 means input element -c-2.
constexpr int F_FWD = 1, F_FUSED = 2, F_STORE = 4, F_WDIRECT = 8, F_PARTNER = 16;
// F_WARPROW (lane 0 of an item): the item is ONE long row, its entries spread over the 32 lanes
constexpr int F_WARPROW = 32;

struct DevSweep {
    int nitems, nfwd;
    int nseg;
    const int    *seg;      // [3*nseg] (first item, end item, items per warp block): sync-free walk
    int nlev;
    const int    *levptr;   // [nlev+1] first item of every dependency level: level-synchronous walk
    const int    *sptr;     // [nitems+1]
    const int    *col;
    const double *val;
    // Row data of the items, one slot per (item, lane) -- `mptr` == nullptr: slot = item*32 + lane.
    // A sweep made mostly of warp-rows (merged levels of a filled factor) stores ONE slot for such
    // an item instead of 32 (31 of them idle lanes): mptr [nitems+1] is then the first slot of every
    // item, and an item with a single slot is read by all its lanes (item_slot below).
    const int    *mptr;
    const int    *rid;      // [slots] LDL row id, -1 = idle lane
    const int    *pidx;     // [nitems*32] index into the user vector (perm[rid])
    const int    *flags;    // [nitems*32] F_* bits
    const double *d;        // [nitems*32] own diagonal entry of D
    const int    *partner;  // [nitems*32] 2x2 partner row id (F_PARTNER)
    const double *e;        // off-diagonal of the 2x2 block
    const double *dp;       // partner's diagonal entry
    // rows without any work in either sweep and without dependents (y_i = z_i / d_i): more than half
    // of the backward rows of a fill-free factor.  One streamed pass instead of items.
    int nlone;
    const int    *lone_pidx;    // [nlone] index into the user vector, ascending
    const double *lone_d;       // [nlone] D(i,i)
    // Level walk: the items of a level are cut into one contiguous share per CTA, balanced by estimated
    // COST (a generic item costs several batched ones), and the warps of a CTA take batches off the share
    // through a shared-memory counter.  ctasplit [nlev][nsplit+1] = first item of every CTA's share;
    // used when the team has exactly nsplit CTAs (otherwise: equal item counts per CTA).
    int nsplit;
    const int    *ctasplit;
};
__device__ __forceinline__ int item_slot(const DevSweep &S, int t, int lane)
{
    if (S.mptr == nullptr) return t * 32 + lane;
    const int mb = __ldg(&S.mptr[t]);
    return (__ldg(&S.mptr[t + 1]) - mb == 1) ? mb : mb + lane;
}

// One-CTA walk of the LDL' solve ("compact walk", small systems): the sweep values
// w and y live in SHARED memory and the factor is streamed through a shared-memory
// ring by the bulk-copy engine, so the dependent chain of a level is a shared-memory
// gather, a few flops and __syncthreads instead of three L2 round trips.  The
// interpreter is kept to a handful of instructions per step, because one warp's
// dependent instruction stream is what bounds a level.  The stream is a sequence of
// fixed-size blocks:
//   int4  {steps in the block, 0, 0, 0}
//   int4  slot[steps][kCwWarps]   what warp w does in step s:
//           {data offset, kind | CW_BARRIER | width << 8 | stride << 24, z, 0}
//           kind CW_ROWS2   : `stride` rows, one per lane, <= 2 entries each
//                             data: per lane {int tgt, col0, col1, 0; double val0, val1}
//           kind CW_ROWS    : `stride` rows, one per lane, `width` <= 8 entries each
//                             data: int tgt[S]; double val[width][S]; int col[width][S]
//           kind CW_WARPROW : one long row spread over the 32 lanes, z = target
//                             data: double val[width][32]; int col[width][32]
//           kind CW_DCHUNK  : rows z .. z+width-1 of D^-1 (width <= 512; stride = 1: the chunk has 2x2 blocks);
//                             warp w takes rows 32w..32w+31 of the chunk
//                             data: double d[n]; (double e[n]; double dp[n]; int partner[n])
//           kind 0          : nothing
//           CW_BARRIER (same in all slots of a step): __syncthreads after the step
//                             -- a dependency level ends
// tgt/col index the shared vector sv (w = sv[0..N), y = sv[yoff..yoff+N), see DevCompact.yoff),
// col -1 = padding.
constexpr int kCwBlock = 16384;     // bytes per stream block
constexpr int kCwStages = 3;        // ring depth
constexpr int kCwStage = 512;       // doubles of staging area behind the sweep values (chain merging, cpk_host_compact.hpp)
#ifndef CPK_CW_WARPS
#define CPK_CW_WARPS 12
#endif
constexpr int kCwWarps = CPK_CW_WARPS;      // warps that walk the stream = slots per step (the CTA may have more)
static_assert(kCwWarps <= kWarpsPerCta, "compact walk: not enough warps in a CTA");
#ifndef CPK_BATCH_BLOCK
#define CPK_BATCH_BLOCK 448
#endif
constexpr int kBatchBlock = CPK_BATCH_BLOCK;   // threads per CTA of a batch launch with more systems than SMs: two CTAs share an SM
constexpr int CW_ROWS2 = 1, CW_ROWS = 2, CW_WARPROW = 3, CW_DCHUNK = 4;
constexpr int CW_BARRIER = 16;
struct DevCompact {
    int nblk;                       // 0: no stream was built
    int smem_off;                   // byte offset of the walk's region in dynamic shared memory, < 0: walk off
    int yoff;                       // y lives at sv[yoff + i]: 0 = ONE shared vector, updated in place (w, then
                                    // D^-1 w, then y: possible when D is diagonal), N = separate w and y (2x2 pivots)
    const unsigned char *stream;    // [nblk * kCwBlock]
    const int *perm;                // [N] w_i = z[perm_i]
};
__host__ __device__ inline size_t cw_smem_bytes(int N, bool one_vector)
{
    return (size_t)kCwStages * kCwBlock + 8 * kCwStages + 16 + (size_t)(one_vector ? 8 : 16) * N + (size_t)8 * kCwStage;
}

struct DevLdl {
    int N, nA, nC;
    DevSweep sw;
    DevCompact cw;
    // sync-free state (row-id indexed): value + the epoch of the solve that
    // produced it, read and written as ONE 128-bit atomic access
    Tagged *wbuf, *ybuf;
    int    *epoch;          // [1] device-resident solve counter
    // level-synchronous state: plain values, a team barrier separates the levels
    double *wv, *yv;
    int    sync_free;       // 1: tagged sync-free walk, 0: one barrier per level
    // row-class form of a shallow sweep (levels walked with one barrier each; wv / yv are then
    // indexed by the USER index of a row, not by its LDL row id)
    int    use_rc;          // 1: walk `rc` instead of the item list
    DevRc  rc;
    // K_P = [A B'; B C] for the refinement residual and `divide`: SELL (streamed like H) or, when
    // resid_rc is set, row-class layout (one level); only the chosen form is built
    int     resid_rc;
    DevSell KPs;
    DevRc   KP;
    DevSell K12, K22;       // B' (nA x nC) and C (nC x nC) for the stateful residual update
    double *atycy;          // [N] = [Aty; Cy]
    double *rvec;           // [N] refinement residual
    // options (opLDL2.m:45-50)
    int    nitref;
    double itref_tol;
    int    force_itref;
    int    residual_update;
    int    ru_stateful;
    int    track_rnorm;
    double *rnorm_out;      // [1] op.rNorm
};

struct DevStatus {
    long long niters;
    int   solved;
    int   status;
    int   err;              // cpk_status (0 ok)
    int   err_iter;
    int   err_second;
    int   shifted;
    double err_value;
    long long hist_len;
    long long napply, nldlsolve, nresid;
    unsigned long long phase_cycles[8];
};

// Matrix-free A (reg_cpkrylov.m:40, "A may be a matrix or a linear operator"): the product A*v is
// computed by a HOST callback while the persistent kernel stays resident.  Mailbox protocol, one
// request at a time: thread 0 of the team's first CTA publishes the device address of v and a
// sequence number in mapped host memory; the host copies v out, runs the callback, copies A*v
// into `u` and then the sequence number into `ack` on ONE copy stream (so `ack` lands after `u`);
// thread 0 of every CTA polls `ack`.
struct DevHostOp {
    int on;                 // 0: A is the explicit matrix in HC / Hn
    volatile int *req;      // mapped host memory: [0] sequence number of the pending request
    volatile long long *req_addr;   // mapped host memory: device address of v
    const int *ack;         // device memory: sequence number of the last answered request
    const double *u;        // [n] device memory: A*v of that request
    long long timeout;      // SM cycles a CTA waits for the host before it gives up (CPK_ERR_TIMEOUT)
};

struct DevSystem {
    int n, m, N;
    DevSell HC;             // blkdiag(H, C), N x N
    DevSell Hn;             // H alone (n x n) and
    DevSell Cm;             // C alone (m x m): stand-alone matvec
    DevLdl  M;
    DevHostOp hop;
};

struct SolveArgs {
    int     solver;
    int     reg_mode;       // 1: rhs shift / un-shift of reg_cpkrylov.m:153-173
    const double *b;        // n (solve) or N (reg) entries
    double *x;              // N entries out: [x; y]
    double  atol, rtol, btol;
    long long itmax;
    int     restart, mem;
    int     profile;
    double *work;           // workspace
    long long work_len;     // doubles
    double *hist;           // 3 rows
    long long hist_cap;
    double *gs;             // GMRES/DQGMRES scalar scratch in global memory
    DevStatus *status;
    int     cw_off;         // byte offset of the compact-walk region in dynamic shared memory, < 0: off
};

// ---------------------------------------------------------------------------
// memory-model primitives
// ---------------------------------------------------------------------------
__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_add(unsigned *p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// L2-coherent data accesses for buffers produced by other SMs inside a phase
__device__ __forceinline__ double ld_cg(const double *p) {
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_cg(double *p, double v) {
    asm volatile("st.global.cg.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ Tagged ld_tagged(const Tagged *p) {
    unsigned long long lo, hi;
    asm volatile("{\n\t.reg .b128 t;\n\tld.relaxed.gpu.global.b128 t, [%2];\n\tmov.b128 {%0, %1}, t;\n\t}"
                 : "=l"(lo), "=l"(hi) : "l"(p) : "memory");
    Tagged r;
    r.v = __longlong_as_double((long long)lo);
    r.tag = hi;
    return r;
}
__device__ __forceinline__ void st_tagged(Tagged *p, double v, unsigned long long tag) {
    const unsigned long long lo = (unsigned long long)__double_as_longlong(v);
    asm volatile("{\n\t.reg .b128 t;\n\tmov.b128 t, {%0, %1};\n\tst.relaxed.gpu.global.b128 [%2], t;\n\t}"
                 :: "l"(lo), "l"(tag), "l"(p) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int *p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_volatile(const int *p) {
    int v;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// shared-memory barrier + bulk copy (global -> shared) of the compact walk's ring
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

// dynamic shared memory of the solver kernels: [solver scratch][compact-walk region]
extern __shared__ double g_dsm[];

// Team-shared control block in global memory (one per launch team).
struct TeamCtl {
    unsigned bar;           // monotonically increasing arrival counter
    int      abort;         // set by the watchdog; every wait loop polls it
    unsigned pad[30];
};

constexpr long long kWatchdogCycles = 4000000000LL;   // ~2 s of SM clocks

// ---------------------------------------------------------------------------
// Team = the set of threads that cooperates on ONE system.
//   GridTeam: all CTAs of a cooperative launch (large systems)
//   CtaTeam : a single CTA (small systems; a batch launch runs many of them)
// Both expose: tid/nthreads, gwarp/nwarps/lane, sync(), reduce<K>().
// ---------------------------------------------------------------------------
struct TeamShared {
    double red[kWarpsPerCta][kRedMax];
    double out[kRedMax];
    int wq[2];              // work queues of the level walk (items of the CTA's share handed out in batches)
};

struct GridTeam {
    static constexpr int kKind = 0;
    int tid, nthreads, gwarp, nwarps, lane;
    TeamCtl *ctl;
    double  *partials;      // [2][kRedMax][gridDim.x]
    TeamShared *sh;
    unsigned bar_target;
    unsigned red_parity;
    // a team is `nblk` consecutive CTAs of the cooperative grid (the whole grid, or one of
    // several sub-teams that each solve their own system): bid = CTA index inside the team,
    // base = first CTA of the team in the grid
    int bid, nblk, base;

    __device__ void init(TeamCtl *c, double *p, TeamShared *s, int team_ctas = 0) {
        nblk = team_ctas > 0 ? team_ctas : (int)gridDim.x;
        bid = (int)blockIdx.x % nblk;
        base = (int)blockIdx.x - bid;
        tid = bid * blockDim.x + threadIdx.x;
        nthreads = nblk * blockDim.x;
        gwarp = tid >> 5;
        nwarps = nthreads >> 5;
        lane = threadIdx.x & 31;
        ctl = c; partials = p; sh = s;
        bar_target = 0; red_parity = 0;
        wide = nullptr; wide_cols = 0; wide_parity = 0;
    }
    __device__ __forceinline__ bool aborted() const { return ld_volatile(&ctl->abort) != 0; }
    __device__ __forceinline__ void set_abort() const { atomicExch(&ctl->abort, 1); }
    __device__ __forceinline__ bool leader() const { return tid == 0; }
    __device__ __forceinline__ bool cta_leader() const { return threadIdx.x == 0; }
    __device__ __forceinline__ void cta_sync() const { __syncthreads(); }
    __device__ __forceinline__ int cta() const { return bid; }
    __device__ __forceinline__ int nctas() const { return nblk; }

    __device__ void sync() {
        __syncthreads();
        if (threadIdx.x == 0) {
            bar_target += (unsigned)nblk;
            red_release_add(&ctl->bar, 1u);
            unsigned spins = 0;
            long long t0 = clock64();
            while ((int)(ld_acquire(&ctl->bar) - bar_target) < 0) {
                if ((++spins & 0xff) == 0) {
                    if (aborted()) break;
                    if (clock64() - t0 > kWatchdogCycles) { set_abort(); break; }
                }
            }
        }
        __syncthreads();
    }

    // All-reduce (sum) of K per-thread values; every thread of the team gets the
    // same bits: fixed butterfly inside the warp, fixed warp order in the CTA,
    // fixed CTA order across the grid.  Contains one grid barrier.
    template <int K>
    __device__ void reduce(double (&v)[K]) {
        static_assert(K <= kRedMax, "reduce width");
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(FULL, v[k], o);
        const int w = threadIdx.x >> 5;
        if (lane == 0)
#pragma unroll
            for (int k = 0; k < K; ++k) sh->red[w][k] = v[k];
        __syncthreads();
        if (w == 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                double s = (lane < kWarpsPerCta) ? sh->red[lane][k] : 0.0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
                if (lane == 0)
                    st_cg(&partials[((size_t)red_parity * kRedMax + k) * gridDim.x + blockIdx.x], s);      // blockIdx.x = base + bid
            }
        }
        sync();
        if (w < K) {
            const double *p = &partials[((size_t)red_parity * kRedMax + w) * gridDim.x + base];
            double s = 0.0;
            for (int b = lane; b < nblk; b += 32) s += ld_cg(&p[b]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
            if (lane == 0) sh->out[w] = s;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = sh->out[k];
        __syncthreads();            // sh->out / sh->red are free again
        red_parity ^= 1u;
    }

    // Multi-column reductions (Arnoldi coefficients): block-reduce K values and
    // park them as columns col0..col0+K-1 of a wide partial array [cols][grid];
    // after ONE team barrier wide_collect() finishes all columns at once.
    double  *wide;          // [2][wide_cols][gridDim.x]
    int      wide_cols;
    unsigned wide_parity;
    template <int K>
    __device__ void wide_store(double (&v)[K], int col0, int nvalid, double * /*dst*/) {
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(FULL, v[k], o);
        const int w = threadIdx.x >> 5;
        __syncthreads();            // sh->red free (previous chunk consumed)
        if (lane == 0)
#pragma unroll
            for (int k = 0; k < K; ++k) sh->red[w][k] = v[k];
        __syncthreads();
        if (w == 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                double s = (lane < kWarpsPerCta) ? sh->red[lane][k] : 0.0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
                if (lane == 0 && k < nvalid)
                    st_cg(&wide[((size_t)wide_parity * wide_cols + col0 + k) * gridDim.x + blockIdx.x], s);
            }
        }
    }
    __device__ void wide_collect(int ncols, double *dst /*shared*/) {
        sync();
        const int w = threadIdx.x >> 5;
        for (int c = w; c < ncols; c += kWarpsPerCta) {
            const double *p = &wide[((size_t)wide_parity * wide_cols + c) * gridDim.x + base];
            double s = 0.0;
            for (int b = lane; b < nblk; b += 32) s += ld_cg(&p[b]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
            if (lane == 0) dst[c] = s;
        }
        __syncthreads();
        wide_parity ^= 1u;
    }
};

struct CtaTeam {
    static constexpr int kKind = 1;
    int tid, nthreads, gwarp, nwarps, lane;
    TeamCtl *ctl;
    TeamShared *sh;

    __device__ void init(TeamCtl *c, double *, TeamShared *s) {
        tid = threadIdx.x;
        nthreads = blockDim.x;
        gwarp = tid >> 5;
        nwarps = nthreads >> 5;
        lane = tid & 31;
        ctl = c; sh = s;
    }
    __device__ __forceinline__ bool aborted() const { return ld_volatile(&ctl->abort) != 0; }
    __device__ __forceinline__ void set_abort() const { atomicExch(&ctl->abort, 1); }
    __device__ __forceinline__ bool leader() const { return tid == 0; }

    // CTA barrier + make global writes of the CTA visible to its own later loads
    // (L1 is per SM, the team never leaves it; __syncthreads orders them).
    __device__ __forceinline__ void sync() { __syncthreads(); }
    __device__ __forceinline__ bool cta_leader() const { return threadIdx.x == 0; }
    __device__ __forceinline__ void cta_sync() const { __syncthreads(); }
    __device__ __forceinline__ int cta() const { return 0; }
    __device__ __forceinline__ int nctas() const { return 1; }

    template <int K>
    __device__ void wide_store(double (&v)[K], int col0, int nvalid, double *dst /*shared*/) {
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(FULL, v[k], o);
        const int w = threadIdx.x >> 5;
        __syncthreads();
        if (lane == 0)
#pragma unroll
            for (int k = 0; k < K; ++k) sh->red[w][k] = v[k];
        __syncthreads();
        if (w < K && w < nvalid) {
            double s = (lane < nwarps) ? sh->red[lane][w] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
            if (lane == 0) dst[col0 + w] = s;
        }
    }
    __device__ void wide_collect(int, double *) { __syncthreads(); }

    template <int K>
    __device__ void reduce(double (&v)[K]) {
        static_assert(K <= kRedMax, "reduce width");
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(FULL, v[k], o);
        const int w = threadIdx.x >> 5;
        if (lane == 0)
#pragma unroll
            for (int k = 0; k < K; ++k) sh->red[w][k] = v[k];
        __syncthreads();
        if (w < K) {
            double s = (lane < nwarps) ? sh->red[lane][w] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
            if (lane == 0) sh->out[w] = s;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = sh->out[k];
        __syncthreads();
    }
};

// (the team object lives in the kernel's frame and is passed by reference through non-inlined
// calls, so its fields are local-memory loads: read the stride once per loop)
#define TEAM_FOR(T, i, N) for (int i = (T).tid, i##_stride = (T).nthreads; i < (N); i += i##_stride)

// Element-wise team loop over N elements with NIN input vectors:
//   body(i, v) with v[k] = src[k][i]
// Loads are 16-byte (two elements) when N is even and every vector is 16-byte
// aligned, and a thread has 4 such loads per vector in flight before the first
// body runs: ~90 KB per SM must be in flight to cover the HBM latency.
#ifndef CPK_TEAM_MAP_B
#define CPK_TEAM_MAP_B 1
#endif
constexpr int kTmB = CPK_TEAM_MAP_B;    // batches of loads a thread has in flight in team_map
// CS: bit k set = source k is dead after this pass (CPK_DEAD_HINT: loaded evict-first, so that it
// leaves the L2 before the vectors the next phase gathers from)
#ifndef CPK_DEAD_HINT
#define CPK_DEAD_HINT 1
#endif
__device__ __forceinline__ double2 ld_dead(const double2 *p) { return CPK_DEAD_HINT ? __ldcs(p) : *p; }
__device__ __forceinline__ double ld_dead(const double *p) { return CPK_DEAD_HINT ? __ldcs(p) : *p; }
__device__ __forceinline__ void st_dead(double *p, double v) { if (CPK_DEAD_HINT) __stcs(p, v); else *p = v; }
template <int NIN, unsigned CS = 0u, class Team, class Body>
__device__ __forceinline__ void team_map(const Team &T, int N, const double *const (&src)[NIN], Body &&body)
{
    const int nt = T.nthreads;
#ifdef CPK_TEAM_MAP_SCALAR
    bool vec = false;
#else
    bool vec = (N & 1) == 0;
#endif
#pragma unroll
    for (int k = 0; k < NIN; ++k) vec = vec && ((reinterpret_cast<unsigned long long>(src[k]) & 15ull) == 0);
    if (vec) {
        const int N2 = N >> 1;
        for (int j0 = T.tid; j0 < N2; j0 += kTmB * nt) {
            double2 v[kTmB][NIN];
#pragma unroll
            for (int u = 0; u < kTmB; ++u) {
                const int j = j0 + u * nt;
                if (j < N2) {
#pragma unroll
                    for (int k = 0; k < NIN; ++k)
                        v[u][k] = ((CS >> k) & 1u) ? ld_dead(reinterpret_cast<const double2 *>(src[k]) + j)
                                                   : reinterpret_cast<const double2 *>(src[k])[j];
                }
            }
#pragma unroll
            for (int u = 0; u < kTmB; ++u) {
                const int j = j0 + u * nt;
                if (j < N2) {
                    double a[NIN], b[NIN];
#pragma unroll
                    for (int k = 0; k < NIN; ++k) { a[k] = v[u][k].x; b[k] = v[u][k].y; }
                    body(2 * j, a);
                    body(2 * j + 1, b);
                }
            }
        }
    } else {
        for (int i0 = T.tid; i0 < N; i0 += kTmB * nt) {
            double v[kTmB][NIN];
#pragma unroll
            for (int u = 0; u < kTmB; ++u) {
                const int i = i0 + u * nt;
                if (i < N) {
#pragma unroll
                    for (int k = 0; k < NIN; ++k) v[u][k] = ((CS >> k) & 1u) ? ld_dead(&src[k][i]) : src[k][i];
                }
            }
#pragma unroll
            for (int u = 0; u < kTmB; ++u) {
                const int i = i0 + u * nt;
                if (i < N) body(i, v[u]);
            }
        }
    }
}

}  // namespace cpk
