// cpk_host.cu -- host side of libcpk_b200.so: handle registry, one-time analysis
// and upload of the operators (SELL-32 matrices, level-ordered LDL' sweeps),
// kernel entry points and the C ABI of include/cpk_b200.h.
//
// Nothing here runs per Krylov iteration: a solve is ONE kernel launch.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <chrono>
#include <mutex>
#include <numeric>
#include <queue>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/cpk_b200.h"
#include "cpk_solvers.cuh"

using namespace cpk;

// ===========================================================================
// errors
// ===========================================================================
static thread_local std::string g_err;
static long long g_launches = 0;

static int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CUDA_TRY(expr)                                                                    \
    do {                                                                                  \
        cudaError_t e__ = (expr);                                                         \
        if (e__ != cudaSuccess)                                                           \
            return fail(CPK_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__));   \
    } while (0)

#include "cpk_host_sparse.hpp"
#include "cpk_host_sweep.hpp"

// FNV-1a over the pattern arrays (colptr, rowind) of a CSC matrix: "same sparsity pattern, entry
// order included" is what the value-refresh entry points (cpk_ldl2_refactor, cpk_system_update)
// rely on -- they copy new values into layouts compiled from the pattern seen at create time.
static uint64_t pattern_hash(const cpk_csc *A)
{
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](const void *p, size_t bytes) {
        const unsigned char *b = static_cast<const unsigned char *>(p);
        for (size_t i = 0; i < bytes; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    };
    mix(&A->nrows, sizeof A->nrows); mix(&A->ncols, sizeof A->ncols);
    mix(A->colptr, sizeof(int64_t) * (size_t)(A->ncols + 1));
    const int64_t nnz = A->colptr[A->ncols];
    if (nnz > 0) mix(A->rowind, sizeof(int64_t) * (size_t)nnz);
    return h;
}

// ===========================================================================
// device objects
// ===========================================================================
#include "cpk_host_compact.hpp"
constexpr int kRcSweepMaxW = 32;    // same for the rows of a sweep (level merging makes rows of tens of entries)
constexpr int kRcMatMaxW = 16;      // widest row class of a plain matrix in row-class form (wider rows: one warp per row)

// Device memory of one object.  Arrays come from the device's stream-ordered pool
// (cudaMallocAsync on the legacy stream, release threshold = keep everything): a small system
// needs ~50 arrays and plain cudaMalloc was half of its set-up time, while large chunks
// from cudaMalloc proved erratic (3 .. 200 ms).  Every allocation is followed by a synchronous
// copy or by the synchronization at the end of the create call before anything uses it.
struct DevArena {
    std::vector<void *> ptrs;
    int device = 0;
    ~DevArena() { release(); }
    void release() {
        if (ptrs.empty()) return;
        cudaSetDevice(device);
        for (void *p : ptrs) cudaFreeAsync(p, (cudaStream_t)0);
        ptrs.clear();
    }
    template <class T>
    cudaError_t alloc(T **out, size_t count, bool zero = false) {
        void *p = nullptr;
        const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
        cudaError_t e = cudaMallocAsync(&p, bytes, (cudaStream_t)0);
        if (e != cudaSuccess) return e;
        ptrs.push_back(p);
        if (zero) { e = cudaMemsetAsync(p, 0, bytes, (cudaStream_t)0); if (e != cudaSuccess) return e; }
        *out = (T *)p;
        return cudaSuccess;
    }
    template <class T>
    cudaError_t upload(const T **out, const std::vector<T> &v) {
        T *p = nullptr;
        cudaError_t e = alloc(&p, v.size());
        if (e != cudaSuccess) return e;
        if (!v.empty()) e = cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, (cudaStream_t)0);
        *out = p;
        return e;
    }
};

#include "cpk_host_rc.hpp"

// contiguous slice ranges per warp, balanced by padded entries (+ a per-slice cost)
static std::vector<int> warp_split(const std::vector<int> &sptr, int nslices, int nwarps)
{
    std::vector<int> out(nwarps + 1, nslices);
    out[0] = 0;
    const long long total = (nslices > 0 ? (long long)sptr[nslices] : 0) + 64LL * nslices;
    int s = 0;
    for (int w = 1; w < nwarps; ++w) {
        const long long target = total * w / nwarps;
        while (s < nslices && (long long)sptr[s] + 64LL * s < target) ++s;
        out[w] = s;
    }
    out[nwarps] = nslices;
    return out;
}

static int g_grid_warps_hint = 148 * kWarpsPerCta;

// Packed entries of a SELL matrix (DevSell::pk): false when a column offset or the number of
// distinct values does not fit.  Values are compared bit for bit (-0.0 and 0.0 are two values).
static bool sell_pack(const HSell &h, std::vector<unsigned> &pk, std::vector<double> &dict)
{
    if (getenv("CPK_SELL_PACK") && atoi(getenv("CPK_SELL_PACK")) == 0) return false;
    std::unordered_map<uint64_t, unsigned> index;
    pk.assign(h.col.size(), 0u);
    dict.clear();
    for (int sl = 0; sl < h.nslices; ++sl) {
        const int b = h.sptr[sl], en = h.sptr[sl + 1];
        for (int k = b; k < en; ++k) {
            const int lane = (k - b) & 31;
            const int row = std::max(h.rowmap[(size_t)sl * 32 + lane], 0);      // idle lanes: entries (col 0, val 0) read x[0]
            const int delta = h.col[k] - row;
            if (delta < -32768 || delta > 32767) return false;
            uint64_t bits;
            memcpy(&bits, &h.val[k], 8);
            auto it = index.find(bits);
            unsigned id;
            if (it == index.end()) {
                if (dict.size() >= 256) return false;       // beyond a few values the table stops living in L1
                id = (unsigned)dict.size();
                index.emplace(bits, id);
                dict.push_back(h.val[k]);
            } else id = it->second;
            pk[k] = ((unsigned)delta & 0xffffu) | (id << 16);
        }
    }
    return true;
}

static cudaError_t upload_sell(DevArena &ar, const HSell &h, DevSell &d, bool try_pack = false)
{
    cudaError_t e;
    d.nrows = h.nrows; d.ncols = h.ncols; d.nslices = h.nslices;
    d.pk = nullptr; d.dict = nullptr; d.ndict = 0;
    std::vector<unsigned> pk;
    std::vector<double> dict;
    const bool packed = try_pack && sell_pack(h, pk, dict);
    {
        const int nw[2] = {g_grid_warps_hint, kWarpsPerCta};
        for (int kdx = 0; kdx < 2; ++kdx) {
            std::vector<int> sp = warp_split(h.sptr, h.nslices, nw[kdx]);
            if ((e = ar.upload(&d.wsplit[kdx], sp)) != cudaSuccess) return e;
            d.nws[kdx] = nw[kdx];
        }
    }
    if ((e = ar.upload(&d.sptr, h.sptr)) != cudaSuccess) return e;
    if (packed) {
        d.col = nullptr; d.val = nullptr;
        if ((e = ar.upload(&d.pk, pk)) != cudaSuccess) return e;
        if ((e = ar.upload(&d.dict, dict)) != cudaSuccess) return e;
        d.ndict = (int)dict.size();
        if (getenv("CPK_VERBOSE"))
            fprintf(stderr, "[cpk] SELL matrix %d x %d packed: %zu entries, %zu distinct values, 4 instead of 12 bytes per entry\n",
                    h.nrows, h.ncols, pk.size(), dict.size());
    } else {
        if ((e = ar.upload(&d.col, h.col)) != cudaSuccess) return e;
        if ((e = ar.upload(&d.val, h.val)) != cudaSuccess) return e;
    }
    if ((e = ar.upload(&d.rowmap, h.rowmap)) != cudaSuccess) return e;
    d.nlong = (int)h.lrow.size();
    if ((e = ar.upload(&d.lrow, h.lrow)) != cudaSuccess) return e;
    if ((e = ar.upload(&d.lptr, h.lptr)) != cudaSuccess) return e;
    if ((e = ar.upload(&d.lcol, h.lcol)) != cudaSuccess) return e;
    if ((e = ar.upload(&d.lval, h.lval)) != cudaSuccess) return e;
    return cudaSuccess;
}
static cudaError_t upload_sweep(DevArena &ar, const HSweep &h, DevSweep &d)
{
    cudaError_t e;
    d.nitems = h.nitems; d.nfwd = h.nfwd; d.nseg = (int)h.seg.size() / 3;
    if ((e = ar.upload(&d.seg, h.seg)) != cudaSuccess) return e;
    {
        std::vector<int> lp = h.levptr;
        lp.push_back(h.nitems);
        d.nlev = (int)h.levptr.size();
        if ((e = ar.upload(&d.levptr, lp)) != cudaSuccess) return e;
    }
    if ((e = ar.upload(&d.sptr, h.sptr)) != cudaSuccess) return e;
    if ((e = ar.upload(&d.col, h.col)) != cudaSuccess) return e;
    if ((e = ar.upload(&d.val, h.val)) != cudaSuccess) return e;
    d.nsplit = h.nsplit;
    d.ctasplit = nullptr;
    if (h.ctasplit.size() == (size_t)(h.nsplit + 1) * h.levptr.size() && h.nsplit > 1)
        if ((e = ar.upload(&d.ctasplit, h.ctasplit)) != cudaSuccess) return e;
    d.nlone = (int)h.lone_pidx.size();
    if ((e = ar.upload(&d.lone_pidx, h.lone_pidx)) != cudaSuccess) return e;
    if ((e = ar.upload(&d.lone_d, h.lone_d)) != cudaSuccess) return e;
    d.mptr = nullptr;
    static const bool no_compact_meta = getenv("CPK_LDL_DENSE_META") != nullptr;
    if (h.nitems >= 4096 && h.n_warprow * 4 > (int64_t)h.nitems && !no_compact_meta) {
        // mostly warp-rows: one slot of row data per warp-row item instead of 32 (see DevSweep::mptr)
        std::vector<int> mptr((size_t)h.nitems + 1, 0), rid, pidx, flags, partner;
        std::vector<double> dd, ee, dp;
        for (int t = 0; t < h.nitems; ++t) {
            const size_t b = (size_t)t * 32;
            const int cnt = (h.flags[b] & F_WARPROW) ? 1 : 32;
            for (int l = 0; l < cnt; ++l) {
                rid.push_back(h.rid[b + l]); pidx.push_back(h.pidx[b + l]); flags.push_back(h.flags[b + l]); partner.push_back(h.partner[b + l]);
                dd.push_back(h.d[b + l]); ee.push_back(h.e[b + l]); dp.push_back(h.dp[b + l]);
            }
            mptr[t + 1] = (int)rid.size();
        }
        if ((e = ar.upload(&d.mptr, mptr)) != cudaSuccess) return e;
        if ((e = ar.upload(&d.rid, rid)) != cudaSuccess) return e;
        if ((e = ar.upload(&d.pidx, pidx)) != cudaSuccess) return e;
        if ((e = ar.upload(&d.flags, flags)) != cudaSuccess) return e;
        if ((e = ar.upload(&d.d, dd)) != cudaSuccess) return e;
        if ((e = ar.upload(&d.partner, partner)) != cudaSuccess) return e;
        if ((e = ar.upload(&d.e, ee)) != cudaSuccess) return e;
        if ((e = ar.upload(&d.dp, dp)) != cudaSuccess) return e;
        return cudaSuccess;
    }
    if ((e = ar.upload(&d.rid, h.rid)) != cudaSuccess) return e;
    if ((e = ar.upload(&d.pidx, h.pidx)) != cudaSuccess) return e;
    if ((e = ar.upload(&d.flags, h.flags)) != cudaSuccess) return e;
    if ((e = ar.upload(&d.d, h.d)) != cudaSuccess) return e;
    if ((e = ar.upload(&d.partner, h.partner)) != cudaSuccess) return e;
    if ((e = ar.upload(&d.e, h.e)) != cudaSuccess) return e;
    if ((e = ar.upload(&d.dp, h.dp)) != cudaSuccess) return e;
    return cudaSuccess;
}

enum ObjKind { OBJ_LDL2 = 1, OBJ_SYSTEM = 2 };

struct Object {
    ObjKind kind;
    int device = 0;
    virtual ~Object() {}
};

// per-device launch context: stream, events, team control block, partials
struct DeviceCtx {
    int device = -1;
    int num_sms = 0;
    int grid_blocks = 0;            // co-resident CTAs of the solver kernel
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    TeamCtl *ctl = nullptr;         // [kMaxBatch]
    double *partials = nullptr;     // [2][kRedMax][grid]
    double *wide = nullptr;         // [2][wide_cols][grid]
    int wide_cols = 0;
    DevStatus *h_status = nullptr;  // pinned
    int h_status_cap = 0;
    int max_dsm = 0;                // dynamic shared memory every solver / apply kernel may ask for
    // batch launches: pinned staging for right-hand sides / solutions, descriptor arrays
    double *h_stage = nullptr, *d_stage = nullptr; size_t h_stage_cap = 0;      // doubles
    DevSystem *d_bsys = nullptr; SolveArgs *d_bargs = nullptr; DevStatus *d_bstatus = nullptr; size_t d_batch_cap = 0;
};
constexpr int kMaxBatch = 4096;
static std::mutex g_mu;
// Every entry point that launches serialises on this lock: the per-device launch context (team
// control block, partials, pinned status, staging buffers, stream, events) is shared by all
// handles of a device, and a cooperative kernel owns every SM anyway.
static std::recursive_mutex g_launch_mu;
#define CPK_LAUNCH_LOCK() std::lock_guard<std::recursive_mutex> launch_lock__(g_launch_mu)
static std::unordered_map<int, std::unique_ptr<DeviceCtx>> g_dev;
static std::unordered_map<uint64_t, std::unique_ptr<Object>> g_obj;
static uint64_t g_next = 1;

struct Ldl2 : Object {
    DevArena ar;
    DevLdl d{};
    int64_t nnz_off = 0, lev_f = 0, lev_b = 0, n2x2 = 0;
    DevSystem *d_sys_alone = nullptr;   // DevSystem with only M filled (stand-alone apply)
    double *d_z = nullptr, *d_y = nullptr;
    DevStatus *d_status = nullptr;
    bool in_system = false;
    bool prefer_grid = false;           // team of a single-system launch (cost model at create time)
    int walk_deep = 0;                  // effective sweep depth > 24: a grid team takes the sync-free walk
    uint64_t patA = 0, patB = 0, patC = 0;  // pattern hashes of the blocks of K_P at create time
    int64_t kp_nval = 0, kp_nlval = 0; // value counts of K_P in row-class form (pattern check of the refactorization)
    bool has_items = true;              // the item list of the sweeps was built (level / sync-free walks)
    bool has_rc = false;                // the row-class form of the sweeps was built (shallow sweeps, diagonal D)
    int sync_free_env = -1;             // CPK_LDL_SYNCFREE at create time (-1: not set)
    // device-side numeric factorization (cpk_ldl2_create_sqd / cpk_ldl2_refactor)
    std::unique_ptr<struct SqdPlan> plan;
    bool sweep_stale = false;           // after a refactorization only the compact stream holds the new factor
    Ldl2() { kind = OBJ_LDL2; }
    ~Ldl2();
};

struct System : Object {
    DevArena ar;
    Ldl2 *M = nullptr;
    uint64_t M_handle = 0;
    DevSystem h{};                  // host copy (device pointers inside)
    DevSystem *d_sys = nullptr;
    SolveArgs *d_args = nullptr;
    DevStatus *d_status = nullptr;
    double *d_b = nullptr, *d_x = nullptr;
    // grow-only buffers
    DevArena war;
    double *d_work = nullptr; long long work_len = 0;
    double *d_hist = nullptr; long long hist_cap = 0;
    double *d_gs = nullptr; long long gs_len = 0;
    uint64_t patH = 0, patC = 0;        // pattern hashes of H and C at create time
    // matrix-free A (cpk_system_create_op): host callback + mailbox (DevHostOp)
    cpk_matvec_fn aop = nullptr;
    void *aop_ctx = nullptr;
    int *h_mail = nullptr;              // mapped pinned: int seq @0, long long address @8, copy sources: int ack @16, int 1 @24
    double *h_v = nullptr, *h_u = nullptr;      // pinned n-vectors
    int *d_ack = nullptr;
    double *d_u = nullptr;
    cudaStream_t copy_stream = nullptr;
    int hop_failed = 0;                 // the callback returned nonzero during the last solve
    System() { kind = OBJ_SYSTEM; }
    ~System() override {
        if (h_mail) cudaFreeHost(h_mail);
        if (h_v) cudaFreeHost(h_v);
        if (h_u) cudaFreeHost(h_u);
        if (copy_stream) cudaStreamDestroy(copy_stream);
    }
};

// forward decls of kernels
extern "C" {
const void *cpk_kernel_cpcg(int grid);
const void *cpk_kernel_cpcglanczos(int grid);
const void *cpk_kernel_cpminres(int grid);
const void *cpk_kernel_cpsymmlq(int grid);
const void *cpk_kernel_cpgmres(int grid);
const void *cpk_kernel_cpdqgmres(int grid);
}
static const void *solver_kernel(int solver, bool grid)
{
    switch (solver) {
        case CPK_CPCG: return cpk_kernel_cpcg(grid);
        case CPK_CPCGLANCZOS: return cpk_kernel_cpcglanczos(grid);
        case CPK_CPMINRES: return cpk_kernel_cpminres(grid);
        case CPK_CPSYMMLQ: return cpk_kernel_cpsymmlq(grid);
        case CPK_CPGMRES: return cpk_kernel_cpgmres(grid);
        case CPK_CPDQGMRES: return cpk_kernel_cpdqgmres(grid);
    }
    return nullptr;
}
template <bool GRID> __global__ void k_apply(const DevSystem *sys, const double *z, double *y, DevStatus *st,
                                             TeamCtl *ctl, double *partials, int cw_off);
template <bool GRID> __global__ void k_matvec(DevSell A, const double *x, double *y);
template <bool GRID> __global__ void k_matvec_rc(const __grid_constant__ DevRc A, const double *x, double *y);

static int get_device_ctx(int device, DeviceCtx **out)
{
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_dev.find(device);
    if (it != g_dev.end()) { *out = it->second.get(); return CPK_OK; }
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) {
        cudaGetLastError();
        return fail(CPK_ERR_CUDA, "no CUDA device available (libcpk_b200 has no CPU fallback)");
    }
    if (device < 0 || device >= cnt) return fail(CPK_ERR_ARG, "device %d out of range (%d visible)", device, cnt);
    CUDA_TRY(cudaSetDevice(device));
    auto c = std::make_unique<DeviceCtx>();
    c->device = device;
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(CPK_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    c->num_sms = prop.multiProcessorCount;
    int coop = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
    if (!coop) return fail(CPK_ERR_CUDA, "device does not support cooperative launch");
    // allow the largest dynamic shared memory the solver kernels may ask for
    int per_sm = 1;
    int optin = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    c->max_dsm = optin;
    for (int sv = 0; sv < 7; ++sv)
        for (int g = 0; g < 2; ++g) {
            const void *k = sv < 6 ? solver_kernel(sv, g) : (g ? (const void *)k_apply<true> : (const void *)k_apply<false>);
            cudaFuncAttributes fa;
            CUDA_TRY(cudaFuncGetAttributes(&fa, k));
            int dyn = (optin - (int)fa.sharedSizeBytes) & ~1023;
            if (g) {
                // grid kernels only ever need the GMRES/DQGMRES scratch
                static const int grid_kb = [] { const char *e = getenv("CPK_GRID_DSM_KB"); return e ? atoi(e) : 200; }();
                dyn = std::min(dyn, grid_kb * 1024);
            } else c->max_dsm = std::min(c->max_dsm, dyn);
            CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kBlock, 0));
            if (per_sm < 1) return fail(CPK_ERR_CUDA, "solver kernel %d does not fit on an SM", sv);
        }
    c->grid_blocks = c->num_sms * kCtasPerSm;      // one CTA per SM (kCtasPerSm > 1 only for experimental block sizes below 512)
    if (const char *e = getenv("CPK_GRID_BLOCKS")) { int v = atoi(e); if (v >= 1 && v <= c->num_sms * per_sm) c->grid_blocks = v; }
    g_grid_warps_hint = c->grid_blocks * kWarpsPerCta;
    {
        cudaMemPool_t pool;
        CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, device));
        unsigned long long keep = ~0ull;        // freed arrays stay in the pool for the next system
        CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&c->ev0));
    CUDA_TRY(cudaEventCreate(&c->ev1));
    CUDA_TRY(cudaMalloc(&c->ctl, sizeof(TeamCtl) * kMaxBatch));
    CUDA_TRY(cudaMalloc(&c->partials, sizeof(double) * 2 * kRedMax * c->grid_blocks));
    CUDA_TRY(cudaMallocHost(&c->h_status, sizeof(DevStatus) * kMaxBatch));
    c->h_status_cap = kMaxBatch;
    *out = c.get();
    g_dev[device] = std::move(c);
    return CPK_OK;
}

static int ensure_wide(DeviceCtx *dc, int cols)
{
    if (cols <= dc->wide_cols) return CPK_OK;
    if (dc->wide) cudaFree(dc->wide);
    dc->wide = nullptr; dc->wide_cols = 0;
    CUDA_TRY(cudaMalloc(&dc->wide, sizeof(double) * 2 * (size_t)cols * dc->grid_blocks));
    dc->wide_cols = cols;
    return CPK_OK;
}

template <class T>
static T *lookup(cpk_handle h, ObjKind kind)
{
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_obj.find(h);
    if (it == g_obj.end() || it->second->kind != kind) return nullptr;
    return static_cast<T *>(it->second.get());
}
static cpk_handle register_obj(std::unique_ptr<Object> o)
{
    std::lock_guard<std::mutex> lk(g_mu);
    const cpk_handle h = g_next++;
    g_obj[h] = std::move(o);
    return h;
}

// team selection: a system this small is solved by ONE CTA (barriers become
// __syncthreads); larger ones by the whole cooperative grid.
static bool use_grid(int N)
{
    static int thr = [] { const char *e = getenv("CPK_CTA_MAX_N"); return e ? atoi(e) : 24576; }();
    if (const char *e = getenv("CPK_TEAM")) {
        if (!strcmp(e, "grid")) return true;
        if (!strcmp(e, "cta")) return false;
    }
    return N > thr;
}

// Team of a SINGLE-system launch.  Measured per-iteration times (cpminres, example options,
// kkt_lap3d patterns, `scripts/team_crossover.py`, round-2 build with level merging):
//   shallow sweeps (k=2, merged to 1+1 levels):  grid flat 58-66 us for N = 1 250 .. 24 600;
//     one CTA 44 / 53 / 82 / 116 / 194 / 284 / 420 us at N = 1 250 / 2 160 / 4 218 / 7 290 / 11 576 / 17 280 / 24 603
//   deep factors (k=6 windowed, 91 / 134 / 257 levels per sweep): one CTA with the compact walk 206 / 459 / 997 us,
//     grid over the merged sweeps 101 / 140 / 246 us (round 1, unmerged sync-free walk: 932 / 1 376 / 2 633 us)
// Batches keep the size rule (use_grid): there one SM per system is the point.
// Cost model of one iteration (two LDL' solves), us, fitted to measurements on the example systems and the
// kkt_lap3d family (`scripts/compact_probe.py`, `scripts/team_crossover.py`):
//   `eff`   depth of the two sweeps after level merging (what the global-memory walks pay per level),
//   `raw`   depth of the factor itself (what the compact walk pays: its stream is not merged),
//   `items` length of the merged item list (one SM walks it at ~0.05 us per item).
// cvxqp1 (N = 5 500, 70+70 -> 4+5 levels, 2 886 items): one CTA level walk 378, compact walk 202, grid 138;
// cvxqp2 (N = 725, 10+10 -> 1+1 levels, 66 items): one CTA level walk 47, compact walk 54, grid 80.
static double model_cta_level(int N, int eff, int items) { return 20.0 + 0.016 * N + 2.0 * (0.05 * items + 2.3 * std::max(eff - 2, 0)); }
static double model_cta_compact(int N, int raw) { return 20.0 + 0.016 * N + 1.2 * raw; }
// the walk of a ONE-CTA team (single small system, or one system of a batch)
static bool model_prefers_compact(int eff, int raw, int items) { return model_cta_compact(0, raw) < model_cta_level(0, eff, items); }
// the team of a SINGLE-system launch (batches keep the size rule, use_grid: there one SM per system is the point)
static bool model_prefers_grid(int N, int eff, int raw, int items, bool compact_allowed)
{
    if (const char *e = getenv("CPK_TEAM")) {
        if (!strcmp(e, "grid")) return true;
        if (!strcmp(e, "cta")) return false;
    }
    if (use_grid(N)) return true;
    const double cta = std::min(model_cta_level(N, eff, items), compact_allowed ? model_cta_compact(N, raw) : 1e300);
    const double grid = 55.0 + 5.0 * std::max(eff - 2, 0);
    return grid < cta;
}
static bool team_is_grid(const Ldl2 *M);

// ===========================================================================
// kernels
// ===========================================================================
template <bool GRID>
__global__ void __launch_bounds__(kBlock, kCtasPerSm)
k_apply(const DevSystem *sys, const double *z, double *y, DevStatus *st, TeamCtl *ctl, double *partials, int cw_off)
{
    __shared__ TeamShared sh;
    __shared__ DevLdl s_M;
    {
        const int *src = reinterpret_cast<const int *>(&sys->M);
        int *dst = reinterpret_cast<int *>(&s_M);
        for (int i = threadIdx.x; i < (int)(sizeof(DevLdl) / sizeof(int)); i += blockDim.x) dst[i] = src[i];
        __syncthreads();
        if (threadIdx.x == 0) s_M.cw.smem_off = GRID ? -1 : cw_off;
        __syncthreads();
    }
    PhaseClock pc; pc.start(false, nullptr);
    const DevLdl &M = s_M;
    VecIn in{z, nullptr, M.nA, false};
    int epoch = *M.epoch;
    if (GRID && M.track_rnorm == 2) {
        // debug timing of the level walk: per-level work / barrier cycles of CTA 0
        GridTeam T; T.init(ctl, partials, &sh);
        PhaseClock dbg; dbg.start(T.leader(), st->phase_cycles);
        ldl_solve_levels(T, M, in, y, false, &dbg);
        T.sync();
        return;
    }
    if (!GRID && M.track_rnorm == 2 && M.cw.smem_off >= 0) {
        // debug timing of the compact walk: cycles of thread 0 in gather / ring wait / steps / scatter
        CtaTeam T; T.init(ctl, nullptr, &sh);
        compact_init(T, M);
        ldl_solve_compact(T, M, in, y, false, st->phase_cycles);
        T.sync();
        compact_drain(T, M);
        return;
    }
    if (!GRID && M.track_rnorm == 2) {
        CtaTeam T; T.init(ctl, nullptr, &sh);
        PhaseClock dbg; dbg.start(T.leader(), st->phase_cycles);
        ldl_solve_levels(T, M, in, y, false, &dbg);
        T.sync();
        return;
    }
    if (GRID) {
        GridTeam T; T.init(ctl, partials, &sh);
        ldl2_apply(T, M, in, y, epoch, st, pc);
        T.sync();
        if (T.leader()) { *M.epoch = epoch; if (T.aborted()) st->err = CPK_ERR_TIMEOUT_; }
    } else {
        CtaTeam T; T.init(ctl, nullptr, &sh);
        compact_init(T, M);
        ldl2_apply(T, M, in, y, epoch, st, pc);
        T.sync();
        compact_drain(T, M);
        if (T.leader()) { *M.epoch = epoch; if (T.aborted()) st->err = CPK_ERR_TIMEOUT_; }
    }
}

template <bool GRID>
__global__ void __launch_bounds__(kBlock, kCtasPerSm)
k_matvec(DevSell A, const double *x, double *y)
{
    if (GRID) {
        GridTeam T; T.init(nullptr, nullptr, nullptr);
        spmv_sell<CPK_SPMV_UP_ALONE>(T, A, x, [&](int row, double s) { y[row] = s; });
    } else {
        CtaTeam T; T.init(nullptr, nullptr, nullptr);
        spmv_sell<CPK_SPMV_UP_ALONE>(T, A, x, [&](int row, double s) { y[row] = s; });
    }
}

template <bool GRID>
__global__ void __launch_bounds__(kBlock, kCtasPerSm)
k_matvec_rc(const __grid_constant__ DevRc A, const double *x, double *y)
{
    __shared__ DevRc s_A;
    {
        const int *src = reinterpret_cast<const int *>(&A);
        int *dst = reinterpret_cast<int *>(&s_A);
        for (int i = threadIdx.x; i < (int)(sizeof(DevRc) / sizeof(int)); i += blockDim.x) dst[i] = src[i];
        __syncthreads();
    }
    MatvecOp op{x, y};
    if (GRID) {
        GridTeam T; T.init(nullptr, nullptr, nullptr);
        rc_level(s_A, 0, T.gwarp, T.nwarps, T.lane, op);
    } else {
        CtaTeam T; T.init(nullptr, nullptr, nullptr);
        rc_level(s_A, 0, T.gwarp, T.nwarps, T.lane, op);
    }
}

// debug: latency of the team barrier and of a 1-value team reduction
__global__ void __launch_bounds__(kBlock, kCtasPerSm) k_barrier_bench(TeamCtl *ctl, double *partials, int iters, long long *out)
{
    __shared__ TeamShared sh;
    GridTeam T; T.init(ctl, partials, &sh);
    T.sync();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) T.sync();
    long long t1 = clock64();
    double v[1] = {1.0};
    for (int i = 0; i < iters; ++i) { T.template reduce<1>(v); v[0] = v[0] * 1e-3; }
    long long t2 = clock64();
    if (T.leader()) { out[0] = t1 - t0; out[1] = t2 - t1; out[2] = (long long)v[0]; }
}

extern "C" int cpk_debug_barrier_cycles(int device, int iters, double *sync_cycles, double *reduce_cycles)
{
    DeviceCtx *dc;
    int rc = get_device_ctx(device, &dc);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(device));
    long long *d_out = nullptr;
    CUDA_TRY(cudaMalloc(&d_out, 3 * sizeof(long long)));
    CUDA_TRY(cudaMemsetAsync(dc->ctl, 0, sizeof(TeamCtl), dc->stream));
    void *params[] = {(void *)&dc->ctl, (void *)&dc->partials, (void *)&iters, (void *)&d_out};
    CUDA_TRY(cudaLaunchCooperativeKernel((const void *)k_barrier_bench, dim3(dc->grid_blocks), dim3(kBlock), params, 0, dc->stream));
    CUDA_TRY(cudaStreamSynchronize(dc->stream));
    long long h[3];
    CUDA_TRY(cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost));
    cudaFree(d_out);
    if (sync_cycles) *sync_cycles = (double)h[0] / iters;
    if (reduce_cycles) *reduce_cycles = (double)h[1] / iters;
    return CPK_OK;
}

// debug / test hook: util/SymGivens.m evaluated by the device function the GMRES / DQGMRES
// loops call (all five branches are reachable: b == 0 & a == 0, b == 0, a == 0, |b| > |a|, else)
__global__ void k_sym_givens(int n, const double *a, const double *b, double *c, double *s, double *d)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sym_givens(a[i], b[i], c[i], s[i], d[i]);
}
extern "C" int cpk_debug_sym_givens(int device, int n, const double *a, const double *b, double *c, double *s, double *d)
{
    DeviceCtx *dc;
    int rc = get_device_ctx(device, &dc);
    if (rc) return rc;
    if (n <= 0 || !a || !b || !c || !s || !d) return fail(CPK_ERR_ARG, "cpk_debug_sym_givens: bad argument");
    CUDA_TRY(cudaSetDevice(device));
    double *buf = nullptr;
    CUDA_TRY(cudaMalloc(&buf, sizeof(double) * 5 * (size_t)n));
    CUDA_TRY(cudaMemcpy(buf, a, sizeof(double) * n, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(buf + n, b, sizeof(double) * n, cudaMemcpyHostToDevice));
    k_sym_givens<<<(n + 127) / 128, 128, 0, dc->stream>>>(n, buf, buf + n, buf + 2 * (size_t)n, buf + 3 * (size_t)n, buf + 4 * (size_t)n);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(dc->stream));
    CUDA_TRY(cudaMemcpy(c, buf + 2 * (size_t)n, sizeof(double) * n, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(s, buf + 3 * (size_t)n, sizeof(double) * n, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(d, buf + 4 * (size_t)n, sizeof(double) * n, cudaMemcpyDeviceToHost));
    cudaFree(buf);
    return CPK_OK;
}

#include "cpk_host_sqd.hpp"
Ldl2::~Ldl2() {}

// one CTA: node values from the new matrix entries, then level after level, then the scatter
// of the factor into the compact-walk stream (ns = 0: no scatter)
__global__ void __launch_bounds__(kBlock, kCtasPerSm)
k_sqd_factor(int ne, int nlev, const double *vals_in, const int *asrc, const int *dk, const int *exec, const int *levptr,
             const int *opptr, const int *ops, double *fval, int ns, const int *spos, const int *ssrc, double *stream, int *bad)
{
    for (int e = threadIdx.x; e < ne; e += blockDim.x) fval[e] = asrc[e] >= 0 ? vals_in[asrc[e]] : 0.0;
    __syncthreads();
    for (int l = 0; l < nlev; ++l) {
        const int a = levptr[l], b = levptr[l + 1];
        for (int q = a + (int)threadIdx.x; q < b; q += blockDim.x) {
            const int e = exec[q];
            double v = fval[e];
            for (int t = opptr[e]; t < opptr[e + 1]; ++t)
                v = v - (fval[ops[3 * t]] * fval[ops[3 * t + 2]]) * fval[ops[3 * t + 1]];      // L_ij d_j L_kj
            const int d = dk[e];
            if (d >= 0) v = v / fval[d];
            else if (v == 0.0 || v != v) atomicExch(bad, e + 1);        // zero or NaN pivot
            fval[e] = v;
        }
        __syncthreads();
    }
    for (int t = threadIdx.x; t < ns; t += blockDim.x) stream[spos[t]] = fval[ssrc[t]];
}

static bool team_is_grid(const Ldl2 *M)
{
    if (const char *e = getenv("CPK_TEAM")) {
        if (!strcmp(e, "grid")) return true;
        if (!strcmp(e, "cta")) return false;
    }
    if (M->plan) return false;          // device-factorized operators live in the compact-walk stream
    return use_grid(M->d.N) || M->prefer_grid;
}
// LDL' walk of one launch: a grid team facing a deep sweep polls tagged values instead of paying a
// grid barrier per level (k=6 windowed stress system, 278+261 levels: 29 ms vs 54 ms per solve)
static void set_walk(DevLdl &d, const Ldl2 *M, bool grid)
{
    d.sync_free = M->sync_free_env >= 0 ? M->sync_free_env : (grid && M->walk_deep ? 1 : 0);
    if (!M->has_items) d.sync_free = 0;
    d.use_rc = M->has_rc && !d.sync_free;
}

// ===========================================================================
// C ABI
// ===========================================================================
// (C linkage comes from the declarations in cpk_b200.h)

// no C++ exception crosses the C ABI: the entry points that allocate on the host run inside this
template <class F>
static int guarded(F &&f)
{
    try { return f(); }
    catch (const std::bad_alloc &) { return fail(CPK_ERR_ALLOC, "out of host memory"); }
    catch (const std::exception &e) { return fail(CPK_ERR_ARG, "internal error: %s", e.what()); }
    catch (...) { return fail(CPK_ERR_ARG, "internal error"); }
}

int cpk_version(void) { return 100; }

int cpk_device_count(void)
{
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess) { cudaGetLastError(); return 0; }
    return cnt;
}

int cpk_last_error(char *buf, int64_t buflen)
{
    if (!buf || buflen <= 0) return CPK_ERR_ARG;
    snprintf(buf, (size_t)buflen, "%s", g_err.c_str());
    return CPK_OK;
}

int64_t cpk_launch_count(void) { return g_launches; }

// Host view of the factors handed to cpk_ldl2_create: permutation vector, D as
// (d, e, partner), the strict lower triangle of L by rows and by columns, and the
// dependency level of every row in the two sweeps.
struct HostLdl {
    std::vector<int64_t> p;
    std::vector<double> d, e;
    std::vector<int> partner;
    int64_t n2 = 0;
    HCsr Lrows, Lcols;
    std::vector<int> lf, lb;
    int nlf = 0, nlb = 0;
};

static int parse_ldl(const cpk_csc *L, const cpk_csc *D, const int64_t *perm, int N, HostLdl *H)
{
    // ---- permutation
    H->p.assign(perm, perm + N);
    std::vector<int64_t> &p = H->p;
    {
        std::vector<char> seen(N, 0);
        for (int k = 0; k < N; ++k) {
            if (p[k] < 0 || p[k] >= N || seen[p[k]]) return fail(CPK_ERR_ARG, "perm is not a permutation of 0..N-1");
            seen[p[k]] = 1;
        }
    }
    // ---- D: 1x1 / 2x2 blocks
    H->d.assign(N, 0.0); H->e.assign(N, 0.0);
    std::vector<double> &d = H->d, &e = H->e;
    for (int64_t j = 0; j < N; ++j)
        for (int64_t k = D->colptr[j]; k < D->colptr[j + 1]; ++k) {
            const int64_t i = D->rowind[k];
            if (i == j) d[j] = D->val[k];
            else if (i == j + 1) e[j] = D->val[k];
            else if (i == j - 1) { /* symmetric twin */ }
            else if (D->val[k] != 0.0) return fail(CPK_ERR_ARG, "D is not block diagonal with 1x1/2x2 blocks");
        }
    H->partner.assign(N, -1);
    std::vector<int> &partner = H->partner;
    int64_t &n2 = H->n2;
    n2 = 0;
    for (int i = 0; i + 1 < N; ++i)
        if (e[i] != 0.0) {
            if (partner[i] >= 0) return fail(CPK_ERR_ARG, "overlapping 2x2 pivots in D");
            partner[i] = i + 1; partner[i + 1] = i; ++n2;
        }
    // ---- L: strict lower triangle, rows (forward deps) and columns (backward deps)
    HCsr &Lrows = H->Lrows, &Lcols = H->Lcols;
    {
        // strip diagonal / reject upper entries
        std::vector<int64_t> cp(N + 1, 0), ri; std::vector<double> vv;
        const int64_t nnzL = L->colptr[N];
        ri.reserve(nnzL); vv.reserve(nnzL);
        for (int64_t j = 0; j < N; ++j) {
            for (int64_t k = L->colptr[j]; k < L->colptr[j + 1]; ++k) {
                const int64_t i = L->rowind[k];
                if (i < j) { if (L->val[k] != 0.0) return fail(CPK_ERR_ARG, "L is not lower triangular"); continue; }
                if (i == j) continue;       // unit diagonal
                ri.push_back(i); vv.push_back(L->val[k]);
            }
            cp[j + 1] = (int64_t)ri.size();
        }
        cpk_csc Ls{N, N, cp.data(), ri.data(), vv.data()};
        Lrows = csr_from_csc(Ls);
        Lcols = csr_of_transpose(Ls);
    }
    // ---- dependency levels
    H->lf.assign(N, 0); H->lb.assign(N, 0);
    std::vector<int> &lf = H->lf, &lb = H->lb;
    int &nlf = H->nlf, &nlb = H->nlb;
    nlf = 0; nlb = 0;
    for (int i = 0; i < N; ++i) {
        int lv = 0;
        for (int64_t k = Lrows.ptr[i]; k < Lrows.ptr[i + 1]; ++k) lv = std::max(lv, lf[Lrows.col[k]] + 1);
        lf[i] = lv; nlf = std::max(nlf, lv + 1);
    }
    for (int i = N - 1; i >= 0; --i) {
        int lv = 0;
        for (int64_t k = Lcols.ptr[i]; k < Lcols.ptr[i + 1]; ++k) lv = std::max(lv, lb[Lcols.col[k]] + 1);
        lb[i] = lv; nlb = std::max(nlb, lv + 1);
    }
    return CPK_OK;
}

// ---------------------------------------------------------------------------
// The two sweeps of the LDL' solve, compiled from the parsed factors: row classification
// (trivial / fused rows), level merging, then the item list (level / sync-free walks) and/or
// the row-class form (shallow sweeps with a diagonal D).  No device needed.
// ---------------------------------------------------------------------------
// substituted rows of the level merging, by LDL row (a dense table: hash maps of a million
// small vectors were most of the set-up time of a large system)
struct ExpMap {
    std::vector<EncRow> rows;
    std::vector<char> has;
    size_t count = 0;
    explicit ExpMap(int N) : rows((size_t)N), has((size_t)N, 0) {}
    const EncRow *find(int i) const { return has[i] ? &rows[i] : nullptr; }
    const EncRow &at(int i) const { return rows[i]; }
    void put(int i, EncRow &&r) { if (!has[i]) { has[i] = 1; ++count; } rows[i] = std::move(r); }
    void erase(int i) { if (has[i]) { has[i] = 0; --count; EncRow().swap(rows[i]); } }
    size_t size() const { return count; }
};

struct SweepBuild {
    HSweep W;
    HRc RC;
    bool have_rc = false, want_items = true, walk_deep = false;
};
static int build_sweeps(HostLdl &HL, int N, bool compact_walk, int grid_warps, SweepBuild *SB)
{
    std::vector<int64_t> &p = HL.p;
    std::vector<double> &d = HL.d, &e = HL.e;
    std::vector<int> &partner = HL.partner;
    const int64_t n2 = HL.n2;
    HCsr &Lrows = HL.Lrows, &Lcols = HL.Lcols;
    HSweep &W = SB->W;
    HRc &RC = SB->RC;
    // ---- classify rows and build the unified item list
    std::vector<char> hasR(N), hasC(N), triv(N), fused(N);
    for (int i = 0; i < N; ++i) {
        hasR[i] = Lrows.len(i) > 0; hasC[i] = Lcols.len(i) > 0;
        triv[i] = !hasR[i] && partner[i] < 0;                   // w_i = (P'z)_i, no forward item
        fused[i] = hasR[i] && !hasC[i] && partner[i] < 0;       // y_i = w_i/d_i inside the forward item
    }
    if (getenv("CPK_LDL_NO_SHORTCUTS")) { std::fill(triv.begin(), triv.end(), 0); std::fill(fused.begin(), fused.end(), 0); }
    // Depth reduction ("level merging"): a level costs one cross-SM hop (a team barrier, or a tagged
    // round trip in the sync-free walk) whatever its size, and the deep part of a filled factor is a
    // long chain of levels of a few rows each (k=6 windowed stress system: 578 of 629 levels hold
    // fewer than 32 rows).  Consecutive levels are therefore merged into GROUPS: the rows of a group
    // are rewritten by substituting their in-group dependencies (an explicit inverse of the group's
    // small unit-triangular block, built row by row), so the whole group becomes ONE level that
    // depends only on earlier groups and on the input vector.  Groups grow greedily, level by level,
    // while the substitution stays cheap and tame: entries of the group <= tail_fill_max x the
    // original ones (or, for small groups, original + 40 000 per absorbed level, see tail_slack), coefficients <= 10^3 x the scale of L (L'/D), rows <= max(tail_len_max, 3 x their
    // original length).  A fill-free forest-shaped factor (cfg 3 / cfg 4: 19+19 levels) collapses
    // to 1+1 levels; the stress factor to a few dozen.  A level with a 2x2 pivot closes the group.
    // A system that will walk the compact stream keeps the global sweep only as a fallback
    // (a solver scratch too large for the stream's shared memory): no merging for it,
    // which is most of the host-side set-up time of a small system.
    static const long long tail_rows_max = [] { const char *e = getenv("CPK_LDL_TAIL_ROWS"); return e ? atoll(e) : 4194304LL; }();
    static const double tail_fill_max = [] { const char *e = getenv("CPK_LDL_TAIL_FILL"); return e ? atof(e) : 6.0; }();
    static const size_t tail_len_max = [] { const char *e = getenv("CPK_LDL_TAIL_MAXLEN"); return (size_t)(e ? atoll(e) : 512LL); }();
    // what a level costs in entries: a hop of ~5 us is worth tens of thousands of entries of streamed work on the grid (40 000: measured best across sizes), so a group
    // may also grow by that much per level it absorbs (up to tail_fill_cap x the original entries) -- small systems
    // (cvxqp1: 7 876 entries in 70 levels) merge far beyond the relative bound, the wide levels of a large one do not
    static const double tail_slack = [] { const char *e = getenv("CPK_LDL_TAIL_SLACK"); return e ? atof(e) : 40000.0; }();
    static const double tail_fill_cap = [] { const char *e = getenv("CPK_LDL_TAIL_CAP"); return e ? atof(e) : 64.0; }();
    static const int merge_lmin = (getenv("CPK_LDL_CUT0") && atoi(getenv("CPK_LDL_CUT0")) == 0) ? 1 : 0;
    const bool merging = !getenv("CPK_LDL_NO_TAIL") && !compact_walk;
    auto add_to = [](std::vector<std::pair<int, double>> &acc, int code, double v) { acc.emplace_back(code, v); };
    auto compress = [](EncRow &r) {         // merge equal codes, keep first-appearance order
        if (r.size() < 2) return;
        std::vector<std::pair<int, double>> out;
        std::unordered_map<int, size_t> pos;
        for (auto &x : r) {
            auto it = pos.find(x.first);
            if (it == pos.end()) { pos[x.first] = out.size(); out.push_back(x); }
            else out[it->second].second += x.second;
        }
        r.swap(out);
    };
    double maxL = 1.0;
    for (double v : Lrows.val) maxL = std::max(maxL, std::fabs(v));
    // The greedy grouping, shared by the two sweeps.  `lev` (levels of the rows not in `skip`) is
    // rewritten to group numbers; `expd` receives the substituted rows of every group of more than
    // one level.  expand(t, grp, gid, acc, orig): encoded dependency list of row t with the
    // dependencies inside group gid substituted (reads expd of rows handled before), `orig` +=
    // its original entry count.  scale(t): what the coefficients of row t are judged against.
    auto merge_levels = [&](std::vector<int> &lev, const std::vector<char> &skip, bool descending, ExpMap &expd,
                            auto &&expand, auto &&blocker, auto &&scale_of, auto &&orig_len) {
        int maxlev = -1;
        for (int i = 0; i < N; ++i) if (!skip[i]) maxlev = std::max(maxlev, lev[i]);
        if (maxlev < 0) return;
        std::vector<std::vector<int>> by_lev((size_t)maxlev + 1);
        if (descending) { for (int i = N - 1; i >= 0; --i) if (!skip[i]) by_lev[lev[i]].push_back(i); }
        else { for (int i = 0; i < N; ++i) if (!skip[i]) by_lev[lev[i]].push_back(i); }
        std::vector<int> grp((size_t)N, -1);
        int gid = -1, g_levels = 0;
        long long g_orig = 0, g_fill = 0, g_nrows = 0;
        double g_scale = maxL;
        bool g_closed = true;
        std::vector<int> g_rows;
        auto close_group = [&] {
            if (g_levels == 1) for (int t : g_rows) expd.erase(t);      // a lone level keeps its original rows
            g_rows.clear();
        };
        std::vector<std::pair<int, EncRow>> cand;
        for (int l = 0; l <= maxlev; ++l) {
            const std::vector<int> &rows = by_lev[l];
            bool blk = false;
            for (int t : rows) if (blocker(t)) { blk = true; break; }
            bool joined = false;
            if (merging && !g_closed && !blk && l >= merge_lmin + 1 && g_nrows + (long long)rows.size() <= tail_rows_max) {
                cand.clear();
                long long lorig = 0, lfill = 0;
                double growth = 0.0, sc = g_scale;
                bool ok = true;
                for (int t : rows) {
                    EncRow acc;
                    expand(t, grp, gid, acc, lorig);
                    compress(acc);
                    lfill += (long long)acc.size();
                    for (auto &x : acc) growth = std::max(growth, std::fabs(x.second));
                    sc = std::max(sc, scale_of(t));
                    const double o = (double)std::max<long long>(g_orig + lorig, 1);
                    const double allowed = std::max(tail_fill_max * o, std::min(tail_fill_cap * o, o + tail_slack * (double)g_levels)) + 1024.0;
                    if (acc.size() > std::max(tail_len_max, (size_t)3 * (size_t)orig_len(t)) || (double)(g_fill + lfill) > allowed) { ok = false; break; }
                    cand.emplace_back(t, std::move(acc));
                }
                if (ok && !(growth <= 1e3 * sc)) ok = false;
                if (ok) {
                    for (auto &kv : cand) { expd.put(kv.first, std::move(kv.second)); grp[kv.first] = gid; g_rows.push_back(kv.first); }
                    g_orig += lorig; g_fill += lfill; g_nrows += (long long)rows.size(); g_scale = sc; ++g_levels;
                    joined = true;
                }
            }
            if (!joined) {
                close_group();
                ++gid; g_levels = 1; g_orig = 0; g_fill = 0; g_nrows = (long long)rows.size(); g_scale = maxL;
                g_closed = blk || !merging || l < merge_lmin;
                for (int t : rows) {
                    grp[t] = gid;
                    if (!g_closed) {
                        EncRow acc;
                        expand(t, grp, -2, acc, g_orig);        // no in-group dependency yet: the encoded original row
                        g_fill += (long long)acc.size();
                        g_scale = std::max(g_scale, scale_of(t));
                        expd.put(t, std::move(acc));
                        g_rows.push_back(t);
                    }
                }
            }
        }
        close_group();
        for (int i = 0; i < N; ++i) if (!skip[i]) lev[i] = grp[i];
    };
    std::vector<int> levf(N, 0), levb(N, 0);
    ExpMap tailf(N), tailb(N);
    std::vector<SweepRow> rowsF, rowsB;
    const auto tb0 = std::chrono::steady_clock::now();
    auto tb_ms = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tb0).count(); };
    double tb_f = 0.0, tb_b = 0.0;
    {
        // ---------------- forward ----------------
        for (int i = 0; i < N; ++i) {
            if (triv[i]) continue;
            int lv = 0;
            for (int64_t k = Lrows.ptr[i]; k < Lrows.ptr[i + 1]; ++k) { const int jx = Lrows.col[k]; if (!triv[jx]) lv = std::max(lv, levf[jx] + 1); }
            levf[i] = lv;
        }
        merge_levels(levf, triv, false, tailf,
            [&](int t, const std::vector<int> &grp, int gid, EncRow &acc, long long &orig) {
                for (int64_t k = Lrows.ptr[t]; k < Lrows.ptr[t + 1]; ++k) {
                    const int jx = Lrows.col[k]; const double a = Lrows.val[k];
                    ++orig;
                    if (triv[jx]) add_to(acc, (int)(-(p[jx]) - 2), a);
                    else if (grp[jx] == gid) {
                        add_to(acc, (int)(-(p[jx]) - 2), a);            // the z_j part of w_j
                        for (auto &x : tailf.at(jx)) add_to(acc, x.first, -a * x.second);
                    } else add_to(acc, jx, a);
                }
            },
            [&](int) { return false; },
            [&](int) { return maxL; },
            [&](int t) { return Lrows.len(t); });
        for (int i = 0; i < N; ++i) {
            if (triv[i]) continue;
            const EncRow *it = tailf.find(i);
            rowsF.push_back({i, levf[i], it ? (int)it->size() : Lrows.len(i)});
        }
        W.lev_f_eff = 0;
        for (int i = 0; i < N; ++i) if (!triv[i]) W.lev_f_eff = std::max(W.lev_f_eff, levf[i] + 1);
    }
    tb_f = tb_ms();
    {
        // ---------------- backward: everything not finished in the forward sweep ----------------
        for (int i = N - 1; i >= 0; --i) {
            if (fused[i]) continue;
            int lv = 0;
            for (int64_t k = Lcols.ptr[i]; k < Lcols.ptr[i + 1]; ++k) { const int r = Lcols.col[k]; if (!fused[r]) lv = std::max(lv, levb[r] + 1); }
            levb[i] = lv;
        }
        merge_levels(levb, fused, true, tailb,
            [&](int t, const std::vector<int> &grp, int gid, EncRow &acc, long long &orig) {
                for (int64_t k = Lcols.ptr[t]; k < Lcols.ptr[t + 1]; ++k) {
                    const int r = Lcols.col[k]; const double a = Lcols.val[k];
                    ++orig;
                    if (!fused[r] && grp[r] == gid) {
                        // y_r = w_r/d_r - sum(in-group deps of r):  a*y_r
                        add_to(acc, triv[r] ? (int)(-(p[r]) - 2) : N + r, a / d[r]);
                        for (auto &x : tailb.at(r)) add_to(acc, x.first, -a * x.second);
                    } else add_to(acc, r, a);
                }
            },
            [&](int t) { return partner[t] >= 0; },
            // growth is judged against the scale of L'/D entries actually present
            [&](int t) { return std::max(maxL, maxL / std::max(std::fabs(d[t]), 1e-300)); },
            [&](int t) { return Lcols.len(t); });
        for (int i = 0; i < N; ++i) {
            if (fused[i]) continue;
            const EncRow *it = tailb.find(i);
            rowsB.push_back({i, levb[i], it ? (int)it->size() : Lcols.len(i)});
        }
        W.lev_b_eff = 0;
        for (int i = 0; i < N; ++i) if (!fused[i]) W.lev_b_eff = std::max(W.lev_b_eff, levb[i] + 1);
        W.tail_f = (long long)tailf.size(); W.tail_b = (long long)tailb.size();
    }
    tb_b = tb_ms();
    for (int i = 0; i < N; ++i) { W.n_trivial += triv[i]; W.n_fused += fused[i]; }
    // encoded dependency lists of a row in the two sweeps (item-list codes: c >= 0 LDL row id,
    // c >= N forward result of row c - N, c <= -2 input element -c-2)
    EncRow scratchF, scratchB;
    auto entriesF = [&](int r) -> const EncRow & {
        if (const EncRow *it = tailf.find(r)) return *it;
        scratchF.clear();
        for (int64_t k = Lrows.ptr[r]; k < Lrows.ptr[r + 1]; ++k) {
            const int jx = Lrows.col[k];
            scratchF.emplace_back(triv[jx] ? (int)(-(p[jx]) - 2) : jx, Lrows.val[k]);
        }
        return scratchF;
    };
    auto entriesB = [&](int r) -> const EncRow & {
        if (const EncRow *it = tailb.find(r)) return *it;
        scratchB.clear();
        for (int64_t k = Lcols.ptr[r]; k < Lcols.ptr[r + 1]; ++k) scratchB.emplace_back(Lcols.col[k], Lcols.val[k]);
        return scratchB;
    };
    // Which form(s) of the sweeps to build.  Shallow sweeps without 2x2 pivots are walked in
    // row-class form (one pass per level) when CPK_LDL_RC=1 asks for it; the item list is then only
    // built when a walk that needs it can still be chosen: the sync-free walk (deep sweeps on a grid
    // team, or forced through CPK_LDL_SYNCFREE), 2x2 pivots.  Default: the item list (measured
    // 5-8 % faster inside the solver loop on cfg 3, profiles/r2_notes.md).
    const char *rcenv = getenv("CPK_LDL_RC");
    const bool rc_ok = n2 == 0 && N < (1 << (RC_IDX_BITS - 1)) && W.lev_f_eff + W.lev_b_eff <= kRcMaxLev && rcenv && atoi(rcenv) == 1;
    const bool walk_deep = W.lev_f_eff + W.lev_b_eff > 24;
    const bool want_items = !rc_ok || walk_deep || getenv("CPK_LDL_SYNCFREE") != nullptr || getenv("CPK_LDL_ITEMS") != nullptr;
    bool have_rc = false;
    if (rc_ok) {
        // codes -> user indices.  y of a row is kept in yv only when a later backward row gathers it.
        std::vector<char> yref(N, 0);
        for (auto &rw : rowsB) for (auto &x : entriesB(rw.row)) if (x.first >= 0 && x.first < N) yref[x.first] = 1;
        RC.with_d = true;
        std::vector<RcRowIn> in;
        in.reserve(rowsF.size());
        // rows without any work in either sweep (y_i = z_i / d_i) ride on the forward pass, as
        // fused rows without entries: it is the lighter of the two
        for (auto &rw : rowsF)
            in.push_back(RcRowIn{(int)p[rw.row] | ((fused[rw.row] ? RC_F_FUSED : 0) | (fused[rw.row] && yref[rw.row] ? RC_F_STOREY : 0)) << RC_IDX_BITS,
                                 rw.level, rw.row, rw.len, d[rw.row]});
        std::vector<char> lone(N, 0);
        for (auto &rw : rowsB)
            if (triv[rw.row] && rw.len == 0 && !yref[rw.row]) {
                lone[rw.row] = 1;
                in.push_back(RcRowIn{(int)p[rw.row] | (RC_F_FUSED << RC_IDX_BITS), 0, rw.row, 0, d[rw.row]});
            }
        const int nlf_eff = std::max(W.lev_f_eff, 1);
        rc_append(RC, std::move(in), nlf_eff, [&](int r) {
            EncRow er = lone[r] ? EncRow() : entriesF(r);
            for (auto &x : er) x.first = x.first >= 0 ? (int)p[x.first] : ((RC_SRC_IN << RC_IDX_BITS) | (-x.first - 2));       // w_j | input
            return er;
        }, kRcSweepMaxW);
        RC.nfwd_lev = RC.nlev;
        std::vector<RcRowIn> inb;
        inb.reserve(rowsB.size());
        for (auto &rw : rowsB)
            if (!lone[rw.row])
                inb.push_back(RcRowIn{(int)p[rw.row] | ((triv[rw.row] ? RC_F_WDIRECT : 0) | (yref[rw.row] ? RC_F_STOREY : 0)) << RC_IDX_BITS,
                                      rw.level, rw.row, rw.len, d[rw.row]});
        rc_append(RC, std::move(inb), W.lev_b_eff, [&](int r) {
            EncRow er = entriesB(r);
            for (auto &x : er)
                x.first = x.first >= N ? (int)p[x.first - N]                                                   // w_r
                        : (x.first >= 0 ? N + (int)p[x.first] : ((RC_SRC_IN << RC_IDX_BITS) | (-x.first - 2)));       // y_r | input
            return er;
        }, kRcSweepMaxW);
        have_rc = rc_fits(RC);
    }
    if (want_items || !have_rc) {
        sweep_append(W, rowsF, p, d, e, partner, grid_warps, entriesF,
                     [&](int r) { return F_FWD | (fused[r] ? F_FUSED : 0); });
        W.nfwd = W.nitems;
        // lone rows (trivial in the forward sweep, no entry in the backward one; a trivial row has no row
        // entries, so no backward row gathers its y): a streamed pass of their own instead of items
        if (!getenv("CPK_LDL_NO_LONE")) {
            std::vector<SweepRow> keep;
            std::vector<std::pair<int, double>> lone;
            keep.reserve(rowsB.size());
            for (auto &rw : rowsB) {
                if (triv[rw.row] && rw.len == 0) lone.emplace_back((int)p[rw.row], d[rw.row]);
                else keep.push_back(rw);
            }
            if (lone.size() >= 1024) {
                std::sort(lone.begin(), lone.end());
                for (auto &x : lone) { W.lone_pidx.push_back(x.first); W.lone_d.push_back(x.second); }
                rowsB.swap(keep);
            }
        }
        sweep_append(W, rowsB, p, d, e, partner, grid_warps, entriesB,
                     [&](int r) { return (triv[r] ? F_WDIRECT : 0) | (hasR[r] ? F_STORE : 0) | (partner[r] >= 0 ? F_PARTNER : 0); });
    }
    if (getenv("CPK_VERBOSE"))
        fprintf(stderr, "[cpk] build_sweeps ms: merge forward %.1f, merge backward %.1f, item lists %.1f\n", tb_f, tb_b - tb_f, tb_ms() - tb_b);
    if ((int64_t)W.col.size() >= INT32_MAX) return fail(CPK_ERR_UNSUPPORTED, "padded L exceeds int32 indexing");
    SB->have_rc = have_rc; SB->want_items = want_items || !have_rc; SB->walk_deep = walk_deep;
    return CPK_OK;
}

// ---------------------------------------------------------------------------
static thread_local bool g_force_compact = false;   // set by cpk_ldl2_create_sqd around its call of cpk_ldl2_create

static int cpk_ldl2_create_impl(cpk_handle *out, const cpk_csc *A, const cpk_csc *B, const cpk_csc *C,
                    const cpk_csc *L, const cpk_csc *D, const int64_t *perm, int device)
{
    if (!out) return fail(CPK_ERR_ARG, "cpk_ldl2_create: null output handle");
    if (!csc_ok(A) || !csc_ok(B) || !csc_ok(C) || !csc_ok(L) || !csc_ok(D) || !perm)
        return fail(CPK_ERR_ARG, "Invalid number of arguments.");                   // opLDL2.m:61-63
    if (A->nrows != A->ncols || C->nrows != C->ncols)
        return fail(CPK_ERR_DIM, "First and last arguments must be square.");       // opLDL2.m:68-70
    if (B->ncols != A->nrows || B->nrows != C->nrows)
        return fail(CPK_ERR_DIM, "Incompatible dimensions.");                       // opLDL2.m:73-75
    const int64_t nA = A->nrows, nC = C->nrows, N64 = nA + nC;
    if (N64 >= (int64_t)1 << 30) return fail(CPK_ERR_UNSUPPORTED, "N = %lld exceeds int32 indexing", (long long)N64);
    if (L->nrows != N64 || L->ncols != N64 || D->nrows != N64 || D->ncols != N64)
        return fail(CPK_ERR_DIM, "LDL factors must be %lld x %lld", (long long)N64, (long long)N64);
    const int N = (int)N64;
    DeviceCtx *dc;
    int rc = get_device_ctx(device, &dc);
    if (rc) return rc;

    const auto t_begin = std::chrono::steady_clock::now();
    auto ms_since = [&](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    };
    HostLdl HL;
    rc = parse_ldl(L, D, perm, N, &HL);
    if (rc) return rc;
    const double t_parse = ms_since(t_begin);
    std::vector<int64_t> &p = HL.p;
    std::vector<double> &d = HL.d, &e = HL.e;
    std::vector<int> &partner = HL.partner;
    const int64_t n2 = HL.n2;
    HCsr &Lrows = HL.Lrows, &Lcols = HL.Lcols;
    std::vector<int> &lf = HL.lf, &lb = HL.lb;
    const int nlf = HL.nlf, nlb = HL.nlb;
    // ---- the sweeps: item list and / or row-class form.  Level merging runs first (it is what
    // decides how deep the sweeps really are); team and walk are chosen from the EFFECTIVE depth.
    const char *cenv = getenv("CPK_LDL_COMPACT");
    const bool compact_fits = !use_grid(N) && cw_smem_bytes(N, n2 == 0) <= (size_t)dc->max_dsm;
    SweepBuild SB;
    rc = build_sweeps(HL, N, g_force_compact, g_grid_warps_hint, &SB);
    if (rc) return rc;
    const int eff_levels = SB.W.lev_f_eff + SB.W.lev_b_eff;
    const bool compact_allowed = compact_fits && !(cenv && atoi(cenv) == 0);
    const bool prefer_grid = !g_force_compact && model_prefers_grid(N, eff_levels, nlf + nlb, SB.W.nitems, compact_allowed);
    // the compact stream is built whenever a one-CTA team would want it -- also for an operator whose
    // single-system launches go to the grid: the same operator may be one system of a batch
    const bool compact_walk = g_force_compact ||
                              (compact_allowed && (model_prefers_compact(eff_levels, nlf + nlb, SB.W.nitems) || (cenv && atoi(cenv) == 1)));
    HSweep &W = SB.W;
    HRc &RC = SB.RC;
    const bool have_rc = SB.have_rc, want_items = SB.want_items, walk_deep = SB.walk_deep;
    const double t_sweeps = ms_since(t_begin);
    // ---- K_P = [A B'; B C] by rows
    HCsr Ar = csr_from_csc(*A), Br = csr_from_csc(*B), Bt = csr_of_transpose(*B), Cr = csr_from_csc(*C);
    HCsr KP = block2x2(&Ar, &Bt, &Br, &Cr, (int)nA, (int)nC);
    HSell sK12 = build_sell(Bt), sK22 = build_sell(Cr);
    // K_P itself in ONE of two forms (CPK_RESID_RC=1: row-class passes; default: SELL, streamed like H)
    const bool resid_rc = [] { const char *e = getenv("CPK_RESID_RC"); return e && atoi(e) == 1; }();
    HRc rKP;
    HSell sKP;
    if (resid_rc) {
        rKP = build_rc(KP, kRcMatMaxW);
        if (!rc_fits(rKP)) return fail(CPK_ERR_UNSUPPORTED, "K_P exceeds int32 indexing");
    } else {
        sKP = build_sell(KP);
        if ((int64_t)sKP.col.size() >= INT32_MAX) return fail(CPK_ERR_UNSUPPORTED, "K_P exceeds int32 indexing");
    }

    const double t_sell = ms_since(t_begin);
    // ---- upload
    auto o = std::make_unique<Ldl2>();
    o->device = device; o->ar.device = device;
    CUDA_TRY(cudaSetDevice(device));
    DevLdl &m = o->d;
    m.N = N; m.nA = (int)nA; m.nC = (int)nC;
    CUDA_TRY(upload_sweep(o->ar, W, m.sw));
    o->has_items = W.nitems > 0 || !have_rc;
    o->has_rc = have_rc;
    m.use_rc = 0;
    if (have_rc) CUDA_TRY(upload_rc(o->ar, RC, m.rc)); else memset(&m.rc, 0, sizeof m.rc);
    CUDA_TRY(o->ar.alloc(&m.wbuf, N, true));        // tag 0 = never produced; epochs start at 1
    CUDA_TRY(o->ar.alloc(&m.ybuf, N, true));
    CUDA_TRY(o->ar.alloc(&m.epoch, 1, true));
    CUDA_TRY(o->ar.alloc(&m.wv, 2 * (size_t)N, true));     // [w | y]: one buffer (the row-class passes address both halves through one base)
    m.yv = m.wv + N;
    // walk selection: one barrier per level is cheapest for shallow sweeps (and keeps the
    // one-CTA team free of polling); a grid team facing a deep sweep uses the sync-free walk
    // (measured on the k=6 windowed stress system, 278+261 levels: 29 ms vs 54 ms per solve)
    // (decided per launch, see set_walk: the same operator may run on either team)
    o->prefer_grid = prefer_grid;
    o->walk_deep = walk_deep;
    o->sync_free_env = getenv("CPK_LDL_SYNCFREE") ? atoi(getenv("CPK_LDL_SYNCFREE")) : -1;
    m.sync_free = 0;
    // one-CTA team: compact walk (sweep values in shared memory, factor streamed through a
    // shared-memory ring) whenever its region can fit next to a solver's scratch
    m.cw.nblk = 0; m.cw.smem_off = -1; m.cw.stream = nullptr; m.cw.perm = nullptr;
    m.cw.yoff = n2 == 0 ? 0 : N;        // diagonal D: one shared vector, updated in place
    CwStream cws;
    {
        // (below ~24 levels the level walk's two barriers-with-L2-round-trips per level cost less
        // than the compact walk's gather / scatter of the whole vector; CPK_LDL_COMPACT=1 forces it)
        if (compact_walk) {
            // (a device-factorized operator rewrites the factor values inside the stream: no merged chains there)
            CwStage stg;
            if (!g_force_compact && !getenv("CPK_CW_NO_CHAINS")) stg.base = m.cw.yoff + N;
            cw_sweep(cws, N, Lrows, lf, nlf, 0, 0, &stg);         // w_i -= L(i,:) w        target w_i,  deps w
            cw_dpass(cws, N, d, e, partner);                // y = D^-1 w
            cw_sweep(cws, N, Lcols, lb, nlb, m.cw.yoff, m.cw.yoff, &stg);      // y_i -= L(:,i)' y       target y_i,  deps y
            cws.flush();
            std::vector<int> p32(N);
            for (int i = 0; i < N; ++i) p32[i] = (int)p[i];
            CUDA_TRY(o->ar.upload(&m.cw.stream, cws.bytes));
            CUDA_TRY(o->ar.upload(&m.cw.perm, p32));
            m.cw.nblk = cws.nblk;
        }
    }
    m.resid_rc = resid_rc ? 1 : 0;
    memset(&m.KP, 0, sizeof m.KP); memset(&m.KPs, 0, sizeof m.KPs);
    if (resid_rc) CUDA_TRY(upload_rc(o->ar, rKP, m.KP)); else CUDA_TRY(upload_sell(o->ar, sKP, m.KPs));
    o->kp_nval = (int64_t)(resid_rc ? rKP.val.size() : sKP.val.size()); o->kp_nlval = (int64_t)(resid_rc ? rKP.lval.size() : sKP.lval.size());
    o->patA = pattern_hash(A); o->patB = pattern_hash(B); o->patC = pattern_hash(C);
    CUDA_TRY(upload_sell(o->ar, sK12, m.K12));
    CUDA_TRY(upload_sell(o->ar, sK22, m.K22));
    CUDA_TRY(o->ar.alloc(&m.atycy, N, true));                                   // opLDL2.m:90-91
    CUDA_TRY(o->ar.alloc(&m.rvec, N, true));
    CUDA_TRY(o->ar.alloc(&m.rnorm_out, 1, true));
    m.nitref = 3; m.itref_tol = 1.0e-8; m.force_itref = 0; m.residual_update = 0;   // opLDL2.m:46-49
    m.ru_stateful = 0; m.track_rnorm = 0;
    CUDA_TRY(o->ar.alloc(&o->d_sys_alone, 1, true));
    CUDA_TRY(o->ar.alloc(&o->d_z, N));
    CUDA_TRY(o->ar.alloc(&o->d_y, N));
    CUDA_TRY(o->ar.alloc(&o->d_status, 1, true));
    o->nnz_off = Lrows.nnz(); o->lev_f = nlf; o->lev_b = nlb; o->n2x2 = n2;
    if (getenv("CPK_VERBOSE"))
        fprintf(stderr, "[cpk] LDL sweep: N=%d trivial=%lld fused=%lld items=%d (fwd %d) segments=%d levels L %d/%d -> effective %d/%d, tail rows inverted %lld/%lld, warp-rows %lld (longest row %lld), padded entries %zu\n",
                N, (long long)W.n_trivial, (long long)W.n_fused, W.nitems, W.nfwd, (int)W.seg.size() / 3, nlf, nlb, W.lev_f_eff, W.lev_b_eff,
                (long long)W.tail_f, (long long)W.tail_b, (long long)W.n_warprow, (long long)W.max_len, W.col.size());
    if (getenv("CPK_VERBOSE") && have_rc) {
        fprintf(stderr, "[cpk] row-class sweeps: %d+%d levels, %zu pieces, %lld rows, %lld entries, %lld long rows; item list %s\n",
                RC.nfwd_lev, RC.nlev - RC.nfwd_lev, RC.pieces.size(), (long long)RC.nrows, (long long)RC.nnz, (long long)RC.nlong,
                want_items ? "built too" : "not built");
        for (int l = 0; l < RC.nlev; ++l)
            for (int q = RC.levp[l]; q < RC.levp[l + 1]; ++q)
                fprintf(stderr, "[cpk]   level %d %s: width %d, %d groups, first batch %d of %d\n", l, l < RC.nfwd_lev ? "fwd" : "bwd",
                        RC.pieces[q].wn & 255, RC.pieces[q].wn >> 8, RC.pieces[q].cum, RC.levb[l]);
    }
    if (getenv("CPK_VERBOSE"))
        fprintf(stderr, "[cpk] set-up ms: parse+levels %.2f, sweeps %.2f, K_P SELL %.2f, uploads+stream %.2f\n",
                t_parse, t_sweeps - t_parse, t_sell - t_sweeps, ms_since(t_begin) - t_sell);
    if (getenv("CPK_VERBOSE")) {
        // item shapes per direction: width (entries per lane) histogram, warp-rows counted apart
        for (int dir = 0; dir < 2; ++dir) {
            long long hist[10] = {0}, wr = 0, rows = 0;
            for (int t = dir ? W.nfwd : 0; t < (dir ? W.nitems : W.nfwd); ++t) {
                const int wdt = (W.sptr[t + 1] - W.sptr[t]) / 32;
                if (W.flags[(size_t)t * 32] & F_WARPROW) { ++wr; continue; }
                hist[std::min(wdt, 9)]++;
                for (int l = 0; l < 32; ++l) rows += W.rid[(size_t)t * 32 + l] >= 0;
            }
            fprintf(stderr, "[cpk]   %s items by width 0..8,9+: %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld; warp-rows %lld; rows in lane items %lld\n",
                    dir ? "bwd" : "fwd", hist[0], hist[1], hist[2], hist[3], hist[4], hist[5], hist[6], hist[7], hist[8], hist[9], wr, rows);
        }
    }
    if (getenv("CPK_VERBOSE") && m.cw.nblk)
        fprintf(stderr, "[cpk] compact walk: %d blocks of %d B, %lld levels in %lld steps, %lld items, shared memory %zu B\n",
                m.cw.nblk, kCwBlock, cws.n_levels, cws.n_steps, cws.n_items, cw_smem_bytes(N, n2 == 0));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)0));      // uploads done (pageable sources are still alive here)
    *out = register_obj(std::move(o));
    return CPK_OK;
}

// debug / test hook (no device needed): the compact-walk stream of a factorization,
// so that the host-side builder can be checked on a machine without a GPU
extern "C" int cpk_debug_cw_stream(const cpk_csc *L, const cpk_csc *D, const int64_t *perm, unsigned char *buf, int64_t cap,
                                   int64_t *nbytes)
{
    if (!csc_ok(L) || !csc_ok(D) || !perm || !nbytes) return fail(CPK_ERR_ARG, "cpk_debug_cw_stream: bad argument");
    const int N = (int)L->nrows;
    HostLdl HL;
    int rc = parse_ldl(L, D, perm, N, &HL);
    if (rc) return rc;
    CwStream cws;
    const int yoff = HL.n2 == 0 ? 0 : N;
    CwStage stg;
    if (!getenv("CPK_CW_NO_CHAINS")) stg.base = yoff + N;
    cw_sweep(cws, N, HL.Lrows, HL.lf, HL.nlf, 0, 0, &stg);
    cw_dpass(cws, N, HL.d, HL.e, HL.partner);
    cw_sweep(cws, N, HL.Lcols, HL.lb, HL.nlb, yoff, yoff, &stg);
    cws.flush();
    *nbytes = (int64_t)cws.bytes.size();
    if (buf && cap >= *nbytes) memcpy(buf, cws.bytes.data(), cws.bytes.size());
    return CPK_OK;
}

// debug / test hook (no device needed): the row-class form of the sweeps of a factorization
// (and, with L == NULL, of the plain matrix A as one level), so that the host-side builder can
// be checked on a machine without a GPU.
//   sizes = {have_rc, nlev, nfwd_lev, npieces, len(rowmap), len(col), len(lptr), len(lcol)}
// Call with the arrays NULL to get the sizes.
extern "C" int cpk_debug_rc(const cpk_csc *A, const cpk_csc *L, const cpk_csc *D, const int64_t *perm, int64_t *sizes,
                            int *pieces, int *levp, int *levb, int *rowmap, double *d, int *col, double *val, int *lptr, int *lcol, double *lval)
{
    if (!sizes) return fail(CPK_ERR_ARG, "cpk_debug_rc: bad argument");
    return guarded([&] {
        SweepBuild SB;
        HRc plain;
        const HRc *R = nullptr;
        if (L) {
            if (!csc_ok(L) || !csc_ok(D) || !perm) return fail(CPK_ERR_ARG, "cpk_debug_rc: bad argument");
            const int N = (int)L->nrows;
            HostLdl HL;
            int rc = parse_ldl(L, D, perm, N, &HL);
            if (rc) return rc;
            rc = build_sweeps(HL, N, false, 148 * kWarpsPerCta, &SB);
            if (rc) return rc;
            R = &SB.RC;
            sizes[0] = SB.have_rc;
        } else {
            if (!csc_ok(A)) return fail(CPK_ERR_ARG, "cpk_debug_rc: bad argument");
            plain = build_rc(csr_from_csc(*A), kRcMatMaxW);
            R = &plain;
            sizes[0] = rc_fits(plain);
        }
        sizes[1] = R->nlev; sizes[2] = R->nfwd_lev; sizes[3] = (int64_t)R->pieces.size(); sizes[4] = (int64_t)R->rowmap.size();
        sizes[5] = (int64_t)R->col.size(); sizes[6] = (int64_t)R->lptr.size(); sizes[7] = (int64_t)R->lcol.size();
        if (pieces) memcpy(pieces, R->pieces.data(), sizeof(RcPiece) * R->pieces.size());
        if (levp) std::copy(R->levp.begin(), R->levp.end(), levp);
        if (levb) std::copy(R->levb.begin(), R->levb.end(), levb);
        if (rowmap) std::copy(R->rowmap.begin(), R->rowmap.end(), rowmap);
        if (d && R->with_d) std::copy(R->d.begin(), R->d.end(), d);
        if (col) std::copy(R->col.begin(), R->col.end(), col);
        if (val) std::copy(R->val.begin(), R->val.end(), val);
        if (lptr) std::copy(R->lptr.begin(), R->lptr.end(), lptr);
        if (lcol) std::copy(R->lcol.begin(), R->lcol.end(), lcol);
        if (lval) std::copy(R->lval.begin(), R->lval.end(), lval);
        return (int)CPK_OK;
    });
}

// debug / test hook (no device needed): the item list of the two sweeps (DevSweep) as build_sweeps
// compiles it for a grid team -- trivial / fused rows, level merging, warp-rows -- so that a numpy
// walk can check it against a direct solve.  sizes = {nitems, nfwd, nlev, entries, effective forward
// levels, effective backward levels, rows rewritten by the merging (fwd), (bwd), lone rows}; the arrays may be
// null (sizes only).
extern "C" int cpk_debug_sweep(const cpk_csc *L, const cpk_csc *D, const int64_t *perm, int64_t *sizes, int32_t *levptr, int32_t *sptr,
                               int32_t *col, double *val, int32_t *rid, int32_t *pidx, int32_t *flags, double *d, int32_t *partner,
                               double *e, double *dp, int32_t *lone_pidx, double *lone_d)
{
    return guarded([&]() -> int {
        if (!csc_ok(L) || !csc_ok(D) || !perm || !sizes) return fail(CPK_ERR_ARG, "cpk_debug_sweep: bad argument");
        const int N = (int)L->nrows;
        HostLdl HL;
        int rc = parse_ldl(L, D, perm, N, &HL);
        if (rc) return rc;
        SweepBuild SB;
        rc = build_sweeps(HL, N, false, 148 * kWarpsPerCta, &SB);
        if (rc) return rc;
        const HSweep &W = SB.W;
        sizes[0] = W.nitems; sizes[1] = W.nfwd; sizes[2] = (int64_t)W.levptr.size(); sizes[3] = (int64_t)W.col.size();
        sizes[4] = W.lev_f_eff; sizes[5] = W.lev_b_eff; sizes[6] = W.tail_f; sizes[7] = W.tail_b; sizes[8] = (int64_t)W.lone_pidx.size();
        if (lone_pidx) std::copy(W.lone_pidx.begin(), W.lone_pidx.end(), lone_pidx);
        if (lone_d) std::copy(W.lone_d.begin(), W.lone_d.end(), lone_d);
        if (levptr) { std::copy(W.levptr.begin(), W.levptr.end(), levptr); levptr[W.levptr.size()] = W.nitems; }
        if (sptr) std::copy(W.sptr.begin(), W.sptr.end(), sptr);
        if (col) std::copy(W.col.begin(), W.col.end(), col);
        if (val) std::copy(W.val.begin(), W.val.end(), val);
        if (rid) std::copy(W.rid.begin(), W.rid.end(), rid);
        if (pidx) std::copy(W.pidx.begin(), W.pidx.end(), pidx);
        if (flags) std::copy(W.flags.begin(), W.flags.end(), flags);
        if (d) std::copy(W.d.begin(), W.d.end(), d);
        if (partner) std::copy(W.partner.begin(), W.partner.end(), partner);
        if (e) std::copy(W.e.begin(), W.e.end(), e);
        if (dp) std::copy(W.dp.begin(), W.dp.end(), dp);
        return (int)CPK_OK;
    });
}

// positions (in doubles) of every factor value inside a compact-walk stream, read off a twin
// stream that was built from the same pattern with "value = node id + 1"
static void cw_value_map(const std::vector<unsigned char> &ids, std::vector<int> &spos, std::vector<int> &ssrc)
{
    const size_t nblk = ids.size() / kCwBlock;
    auto take = [&](size_t byte) {
        double idv;
        memcpy(&idv, ids.data() + byte, 8);
        if (idv != 0.0) { spos.push_back((int)(byte / 8)); ssrc.push_back((int)idv - 1); }
    };
    for (size_t b = 0; b < nblk; ++b) {
        const unsigned char *blk = ids.data() + b * kCwBlock;
        const int nsteps = *reinterpret_cast<const int *>(blk);
        const int *slot = reinterpret_cast<const int *>(blk + 16);
        std::vector<int> seen;              // D chunks are shared by the 16 slots of their step
        for (int t = 0; t < nsteps * kCwWarps; ++t) {
            const int off = slot[4 * t], y = slot[4 * t + 1];
            const int kind = y & 15, width = (y >> 8) & 0xffff, stride = (y >> 24) & 0xff;
            if (kind == 0) continue;
            const size_t base = b * kCwBlock + (size_t)off;
            if (kind == CW_DCHUNK) {
                if (std::find(seen.begin(), seen.end(), off) != seen.end()) continue;
                seen.push_back(off);
                for (int r = 0; r < width; ++r) take(base + 8 * (size_t)r);
            } else if (kind == CW_ROWS2) {
                for (int q = 0; q < stride; ++q) { take(base + 32 * (size_t)q + 16); take(base + 32 * (size_t)q + 24); }
            } else {
                const size_t v0 = base + (kind == CW_ROWS ? 4 * (size_t)stride : 0);
                for (int q = 0; q < width * stride; ++q) take(v0 + 8 * (size_t)q);
            }
        }
    }
}

static int sqd_launch(DeviceCtx *dc, SqdPlan *P, double *stream)
{
    int *d_bad = nullptr;
    CUDA_TRY(cudaMallocAsync(&d_bad, sizeof(int), dc->stream));
    CUDA_TRY(cudaMemsetAsync(d_bad, 0, sizeof(int), dc->stream));
    const int ns = stream ? (int)P->ns : 0;
    k_sqd_factor<<<1, kBlock, 0, dc->stream>>>((int)P->ne, P->nlev, P->d_vals_in, P->d_asrc, P->d_dk, P->d_exec, P->d_levptr, P->d_opptr,
                                                P->d_ops, P->d_fval, ns, P->d_spos, P->d_ssrc, stream, d_bad);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    int bad = 0;
    CUDA_TRY(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, dc->stream));
    CUDA_TRY(cudaStreamSynchronize(dc->stream));
    cudaFreeAsync(d_bad, dc->stream);
    if (bad) return fail(CPK_ERR_BREAKDOWN, "device LDL': zero or NaN pivot at row %d of the permuted matrix (K_P is not quasi-definite?)",
                         bad - 1 - (int)P->nnzL);
    return CPK_OK;
}

static int sqd_upload_values(DeviceCtx *dc, SqdPlan *P, const cpk_csc *A, const cpk_csc *B, const cpk_csc *C)
{
    if (A->colptr[A->ncols] != P->nnzA || B->colptr[B->ncols] != P->nnzB || C->colptr[C->ncols] != P->nnzC)
        return fail(CPK_ERR_DIM, "refactorization needs the sparsity pattern the operator was created with");
    std::vector<double> v((size_t)P->nvals_in);
    if (P->nnzA) memcpy(v.data(), A->val, sizeof(double) * P->nnzA);
    if (P->nnzB) memcpy(v.data() + P->nnzA, B->val, sizeof(double) * P->nnzB);
    if (P->nnzC) memcpy(v.data() + P->nnzA + P->nnzB, C->val, sizeof(double) * P->nnzC);
    CUDA_TRY(cudaMemcpyAsync(P->d_vals_in, v.data(), sizeof(double) * v.size(), cudaMemcpyHostToDevice, dc->stream));
    CUDA_TRY(cudaStreamSynchronize(dc->stream));
    return CPK_OK;
}

static int cpk_ldl2_create_sqd_impl(cpk_handle *out, const cpk_csc *A, const cpk_csc *B, const cpk_csc *C, const int64_t *perm, int device)
{
    CPK_LAUNCH_LOCK();
    if (!out) return fail(CPK_ERR_ARG, "cpk_ldl2_create_sqd: null output handle");
    if (!csc_ok(A) || !csc_ok(B) || !csc_ok(C) || !perm) return fail(CPK_ERR_ARG, "Invalid number of arguments.");
    if (A->nrows != A->ncols || C->nrows != C->ncols) return fail(CPK_ERR_DIM, "First and last arguments must be square.");
    if (B->ncols != A->nrows || B->nrows != C->nrows) return fail(CPK_ERR_DIM, "Incompatible dimensions.");
    const int64_t N64 = A->nrows + C->nrows;
    DeviceCtx *dc;
    int rc = get_device_ctx(device, &dc);
    if (rc) return rc;
    if (N64 >= INT32_MAX || use_grid((int)N64) || cw_smem_bytes((int)N64, true) > (size_t)dc->max_dsm)
        return fail(CPK_ERR_UNSUPPORTED, "device factorization serves the one-CTA team (N <= ~22000); larger systems bring host factors to cpk_ldl2_create");
    const int N = (int)N64, nA = (int)A->nrows;
    CUDA_TRY(cudaSetDevice(device));
    auto P = std::make_unique<SqdPlan>();
    P->ar.device = device;
    SqdHost H;
    rc = sqd_plan(A, B, C, perm, N, nA, P.get(), &H);
    if (rc) return rc;
    CUDA_TRY(P->ar.alloc(&P->d_fval, (size_t)P->ne));
    CUDA_TRY(P->ar.alloc(&P->d_vals_in, (size_t)P->nvals_in));
    CUDA_TRY(P->ar.upload(&P->d_asrc, H.asrc));
    CUDA_TRY(P->ar.upload(&P->d_dk, H.dk));
    CUDA_TRY(P->ar.upload(&P->d_exec, H.exec));
    CUDA_TRY(P->ar.upload(&P->d_levptr, H.levptr));
    CUDA_TRY(P->ar.upload(&P->d_opptr, H.opptr));
    CUDA_TRY(P->ar.upload(&P->d_ops, H.ops));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)0));
    rc = sqd_upload_values(dc, P.get(), A, B, C);
    if (rc) return rc;
    rc = sqd_launch(dc, P.get(), nullptr);
    if (rc) return rc;
    // factor back to the host once: the regular create builds every structure from it
    std::vector<double> fval((size_t)P->ne);
    CUDA_TRY(cudaMemcpy(fval.data(), P->d_fval, sizeof(double) * fval.size(), cudaMemcpyDeviceToHost));
    std::vector<int64_t> dptr(N + 1), dind(N);
    for (int i = 0; i <= N; ++i) dptr[i] = i;
    for (int i = 0; i < N; ++i) dind[i] = i;
    cpk_csc Lc{N, N, P->colptr.data(), P->rowind.data(), fval.data()};
    cpk_csc Dc{N, N, dptr.data(), dind.data(), fval.data() + P->nnzL};
    cpk_handle h = 0;
    g_force_compact = true;
    rc = cpk_ldl2_create_impl(&h, A, B, C, &Lc, &Dc, perm, device);
    g_force_compact = false;
    if (rc) return rc;
    Ldl2 *M = lookup<Ldl2>(h, OBJ_LDL2);
    if (!M || M->d.cw.nblk == 0) { cpk_destroy(h); return fail(CPK_ERR_UNSUPPORTED, "device factorization needs the compact walk"); }
    // where every node value sits in the stream: twin stream built with ids as values
    {
        std::vector<double> idv((size_t)P->ne);
        for (int64_t e = 0; e < P->ne; ++e) idv[e] = (double)(e + 1);
        cpk_csc Li{N, N, P->colptr.data(), P->rowind.data(), idv.data()};
        cpk_csc Di{N, N, dptr.data(), dind.data(), idv.data() + P->nnzL};
        HostLdl HL;
        rc = parse_ldl(&Li, &Di, perm, N, &HL);
        if (rc) { cpk_destroy(h); return rc; }
        CwStream ids;
        cw_sweep(ids, N, HL.Lrows, HL.lf, HL.nlf, 0, 0);
        cw_dpass(ids, N, HL.d, HL.e, HL.partner);
        cw_sweep(ids, N, HL.Lcols, HL.lb, HL.nlb, 0, 0);        // diagonal D: one shared vector
        ids.flush();
        if ((int)(ids.bytes.size() / kCwBlock) != M->d.cw.nblk) { cpk_destroy(h); return fail(CPK_ERR_ARG, "internal: twin stream differs in size"); }
        std::vector<int> spos, ssrc;
        cw_value_map(ids.bytes, spos, ssrc);
        if ((int64_t)spos.size() != 2 * P->nnzL + N) { cpk_destroy(h); return fail(CPK_ERR_ARG, "internal: stream holds %zu factor values, expected %lld", spos.size(), (long long)(2 * P->nnzL + N)); }
        P->ns = (int64_t)spos.size();
        CUDA_TRY(P->ar.upload(&P->d_spos, spos));
        CUDA_TRY(P->ar.upload(&P->d_ssrc, ssrc));
        CUDA_TRY(cudaStreamSynchronize((cudaStream_t)0));
    }
    M->plan = std::move(P);
    *out = h;
    return CPK_OK;
}

// new values, same patterns: numeric factorization on the device, written into the operator in place
static int cpk_ldl2_refactor_impl(cpk_handle h, const cpk_csc *A, const cpk_csc *B, const cpk_csc *C)
{
    CPK_LAUNCH_LOCK();
    Ldl2 *M = lookup<Ldl2>(h, OBJ_LDL2);
    if (!M) return fail(CPK_ERR_ARG, "cpk_ldl2_refactor: not an opLDL2 handle");
    if (!M->plan) return fail(CPK_ERR_ARG, "cpk_ldl2_refactor: the operator was not created by cpk_ldl2_create_sqd");
    if (!csc_ok(A) || !csc_ok(B) || !csc_ok(C)) return fail(CPK_ERR_ARG, "Invalid number of arguments.");
    if (A->nrows != M->d.nA || A->ncols != M->d.nA || C->nrows != M->d.nC || C->ncols != M->d.nC || B->nrows != M->d.nC || B->ncols != M->d.nA)
        return fail(CPK_ERR_DIM, "Incompatible dimensions.");
    if (pattern_hash(A) != M->patA || pattern_hash(B) != M->patB || pattern_hash(C) != M->patC)
        return fail(CPK_ERR_DIM, "refactorization needs the sparsity pattern the operator was created with "
                                 "(same colptr and rowind arrays, entry order included)");
    DeviceCtx *dc;
    int rc = get_device_ctx(M->device, &dc);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(M->device));
    rc = sqd_upload_values(dc, M->plan.get(), A, B, C);
    if (rc) return rc;
    rc = sqd_launch(dc, M->plan.get(), reinterpret_cast<double *>(const_cast<unsigned char *>(M->d.cw.stream)));
    if (rc) return rc;
    M->sweep_stale = true;
    // K_P, B', C for the refinement residual and the residual update: same layout, new values
    const int nA = M->d.nA, nC = M->d.nC;
    HCsr Ar = csr_from_csc(*A), Br = csr_from_csc(*B), Bt = csr_of_transpose(*B), Cr = csr_from_csc(*C);
    HCsr KP = block2x2(&Ar, &Bt, &Br, &Cr, nA, nC);
    if (M->d.resid_rc) {
        const HRc rk = build_rc(KP, kRcMatMaxW);
        if ((int)rk.pieces.size() != M->d.KP.npieces || (int64_t)rk.val.size() != M->kp_nval || (int64_t)rk.lval.size() != M->kp_nlval)
            return fail(CPK_ERR_DIM, "refactorization needs the sparsity pattern the operator was created with");
        if (!rk.val.empty()) CUDA_TRY(cudaMemcpyAsync(const_cast<double *>(M->d.KP.val), rk.val.data(), sizeof(double) * rk.val.size(), cudaMemcpyHostToDevice, dc->stream));
        if (!rk.lval.empty()) CUDA_TRY(cudaMemcpyAsync(const_cast<double *>(M->d.KP.lval), rk.lval.data(), sizeof(double) * rk.lval.size(), cudaMemcpyHostToDevice, dc->stream));
        CUDA_TRY(cudaStreamSynchronize(dc->stream));      // rk goes out of scope
    } else {
        const HSell sk = build_sell(KP);
        if (sk.nslices != M->d.KPs.nslices || (int64_t)sk.val.size() != M->kp_nval || (int64_t)sk.lval.size() != M->kp_nlval)
            return fail(CPK_ERR_DIM, "refactorization needs the sparsity pattern the operator was created with");
        if (!sk.val.empty()) CUDA_TRY(cudaMemcpyAsync(const_cast<double *>(M->d.KPs.val), sk.val.data(), sizeof(double) * sk.val.size(), cudaMemcpyHostToDevice, dc->stream));
        if (!sk.lval.empty()) CUDA_TRY(cudaMemcpyAsync(const_cast<double *>(M->d.KPs.lval), sk.lval.data(), sizeof(double) * sk.lval.size(), cudaMemcpyHostToDevice, dc->stream));
        CUDA_TRY(cudaStreamSynchronize(dc->stream));
    }
    const HSell hs[2] = {build_sell(Bt), build_sell(Cr)};
    const DevSell *ds[2] = {&M->d.K12, &M->d.K22};
    for (int q = 0; q < 2; ++q) {
        if (hs[q].nslices != ds[q]->nslices || (int)hs[q].lrow.size() != ds[q]->nlong)       // (cannot differ once the hashes agree)
            return fail(CPK_ERR_DIM, "refactorization needs the sparsity pattern the operator was created with");
        if (!hs[q].val.empty()) CUDA_TRY(cudaMemcpyAsync(const_cast<double *>(ds[q]->val), hs[q].val.data(), sizeof(double) * hs[q].val.size(), cudaMemcpyHostToDevice, dc->stream));
        if (!hs[q].lval.empty()) CUDA_TRY(cudaMemcpyAsync(const_cast<double *>(ds[q]->lval), hs[q].lval.data(), sizeof(double) * hs[q].lval.size(), cudaMemcpyHostToDevice, dc->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(dc->stream));
    return CPK_OK;
}

// pattern (strict lower triangle, CSC) and values of the factor held on the device
int cpk_ldl2_get_factor(cpk_handle h, int64_t *nnz, int64_t *colptr, int64_t *rowind, double *Lval, double *d)
{
    Ldl2 *M = lookup<Ldl2>(h, OBJ_LDL2);
    if (!M || !M->plan) return fail(CPK_ERR_ARG, "cpk_ldl2_get_factor: the operator was not created by cpk_ldl2_create_sqd");
    SqdPlan *P = M->plan.get();
    if (nnz) *nnz = P->nnzL;
    if (colptr) std::copy(P->colptr.begin(), P->colptr.end(), colptr);
    if (rowind) std::copy(P->rowind.begin(), P->rowind.end(), rowind);
    CUDA_TRY(cudaSetDevice(M->device));
    if (Lval && P->nnzL) CUDA_TRY(cudaMemcpy(Lval, P->d_fval, sizeof(double) * P->nnzL, cudaMemcpyDeviceToHost));
    if (d) CUDA_TRY(cudaMemcpy(d, P->d_fval + P->nnzL, sizeof(double) * P->N, cudaMemcpyDeviceToHost));
    return CPK_OK;
}

// new values of H and C (same patterns) for the mat-vecs of the solver loops
static int cpk_system_update_impl(cpk_handle h, const cpk_csc *A, const cpk_csc *C)
{
    CPK_LAUNCH_LOCK();
    System *S = lookup<System>(h, OBJ_SYSTEM);
    if (!S) return fail(CPK_ERR_ARG, "cpk_system_update: not a system handle");
    if (S->aop) return fail(CPK_ERR_UNSUPPORTED, "cpk_system_update: the system has a matrix-free A (create a new system)");
    if (!csc_ok(A) || !csc_ok(C)) return fail(CPK_ERR_ARG, "cpk_system_update: bad matrix");
    const int n = S->h.n, m = S->h.m;
    if (A->nrows != n || A->ncols != n || C->nrows != m || C->ncols != m) return fail(CPK_ERR_DIM, "Incompatible dimensions.");
    if (pattern_hash(A) != S->patH || pattern_hash(C) != S->patC)
        return fail(CPK_ERR_DIM, "cpk_system_update needs the sparsity pattern the system was created with "
                                 "(same colptr and rowind arrays, entry order included)");
    DeviceCtx *dc;
    int rc = get_device_ctx(S->device, &dc);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(S->device));
    HCsr Hr = csr_from_csc(*A), Cr = csr_from_csc(*C);
    HSell s;
    s.nrows = n + m; s.ncols = n + m;
    sell_append(s, Hr, 0, n, 0, 0);
    sell_append(s, Cr, 0, m, n, n);
    HSell sc = build_sell(Cr);
    const HSell *hs[2] = {&s, &sc};
    const DevSell *ds[2] = {&S->h.HC, &S->h.Cm};
    for (int q = 0; q < 2; ++q) {
        if (hs[q]->nslices != ds[q]->nslices || (int)hs[q]->lrow.size() != ds[q]->nlong)
            return fail(CPK_ERR_DIM, "cpk_system_update needs the sparsity pattern the system was created with");
        if (ds[q]->pk) {
            // packed entries: the new values must still fit the value table
            std::vector<unsigned> pk;
            std::vector<double> dict;
            if (!sell_pack(*hs[q], pk, dict) || (int)dict.size() > ds[q]->ndict)
                return fail(CPK_ERR_UNSUPPORTED, "cpk_system_update: the system was created with packed matrix entries (few distinct values) and the "
                                                 "new values do not fit its value table; create the system with CPK_SELL_PACK=0");
            CUDA_TRY(cudaMemcpyAsync(const_cast<unsigned *>(ds[q]->pk), pk.data(), sizeof(unsigned) * pk.size(), cudaMemcpyHostToDevice, dc->stream));
            CUDA_TRY(cudaMemcpyAsync(const_cast<double *>(ds[q]->dict), dict.data(), sizeof(double) * dict.size(), cudaMemcpyHostToDevice, dc->stream));
            if (!hs[q]->lval.empty()) CUDA_TRY(cudaMemcpyAsync(const_cast<double *>(ds[q]->lval), hs[q]->lval.data(), sizeof(double) * hs[q]->lval.size(), cudaMemcpyHostToDevice, dc->stream));
            CUDA_TRY(cudaStreamSynchronize(dc->stream));      // pk / dict go out of scope
            continue;
        }
        if (!hs[q]->val.empty()) CUDA_TRY(cudaMemcpyAsync(const_cast<double *>(ds[q]->val), hs[q]->val.data(), sizeof(double) * hs[q]->val.size(), cudaMemcpyHostToDevice, dc->stream));
        if (!hs[q]->lval.empty()) CUDA_TRY(cudaMemcpyAsync(const_cast<double *>(ds[q]->lval), hs[q]->lval.data(), sizeof(double) * hs[q]->lval.size(), cudaMemcpyHostToDevice, dc->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(dc->stream));
    return CPK_OK;
}

// debug / test hook (no device needed): the factorization plan of cpk_ldl2_create_sqd.
// sizes = {ne, nnzL, nops, nlev, nvals_in}; call with the arrays NULL to get the sizes.
extern "C" int cpk_debug_sqd_plan(const cpk_csc *A, const cpk_csc *B, const cpk_csc *C, const int64_t *perm, int64_t *sizes,
                                  int64_t *colptr, int64_t *rowind, int *asrc, int *dk, int *exec, int *levptr, int *opptr, int *ops)
{
    if (!csc_ok(A) || !csc_ok(B) || !csc_ok(C) || !perm || !sizes) return fail(CPK_ERR_ARG, "cpk_debug_sqd_plan: bad argument");
    const int N = (int)(A->nrows + C->nrows);
    SqdPlan P;
    SqdHost H;
    int rc = sqd_plan(A, B, C, perm, N, (int)A->nrows, &P, &H);
    if (rc) return rc;
    sizes[0] = P.ne; sizes[1] = P.nnzL; sizes[2] = P.nops; sizes[3] = P.nlev; sizes[4] = P.nvals_in;
    if (colptr) std::copy(P.colptr.begin(), P.colptr.end(), colptr);
    if (rowind) std::copy(P.rowind.begin(), P.rowind.end(), rowind);
    if (asrc) std::copy(H.asrc.begin(), H.asrc.end(), asrc);
    if (dk) std::copy(H.dk.begin(), H.dk.end(), dk);
    if (exec) std::copy(H.exec.begin(), H.exec.end(), exec);
    if (levptr) std::copy(H.levptr.begin(), H.levptr.end(), levptr);
    if (opptr) std::copy(H.opptr.begin(), H.opptr.end(), opptr);
    if (ops) std::copy(H.ops.begin(), H.ops.end(), ops);
    return CPK_OK;
}

static int ldl2_set(cpk_handle h, int which, double v)
{
    Ldl2 *M = lookup<Ldl2>(h, OBJ_LDL2);
    if (!M) return fail(CPK_ERR_ARG, "not an opLDL2 handle");
    switch (which) {
        case 0: M->d.nitref = (int)std::max(0.0, std::nearbyint(v)); break;         // opLDL2.m:97-99
        case 1: M->d.itref_tol = v; break;     // the reference's setter is mis-spelt (opLDL2.m:101) and never runs: value taken as is
        case 2: M->d.force_itref = (v != 0.0 && v != 1.0) ? 0 : (int)v; break;      // opLDL2.m:105-111
        case 3: M->d.residual_update = v != 0.0; break;
        case 4: M->d.ru_stateful = v != 0.0; break;
        case 5: M->d.track_rnorm = (int)v; break;
    }
    return CPK_OK;
}
int cpk_ldl2_set_nitref(cpk_handle M, double v) { return ldl2_set(M, 0, v); }
int cpk_ldl2_set_itref_tol(cpk_handle M, double v) { return ldl2_set(M, 1, v); }
int cpk_ldl2_set_force_itref(cpk_handle M, int v) { return ldl2_set(M, 2, v); }
int cpk_ldl2_set_residual_update(cpk_handle M, int v) { return ldl2_set(M, 3, v); }
int cpk_ldl2_set_ru_stateful(cpk_handle M, int v) { return ldl2_set(M, 4, v); }
int cpk_ldl2_set_track_rnorm(cpk_handle M, int v) { return ldl2_set(M, 5, v); }

int cpk_ldl2_get_rnorm(cpk_handle h, double *rnorm)
{
    Ldl2 *M = lookup<Ldl2>(h, OBJ_LDL2);
    if (!M || !rnorm) return fail(CPK_ERR_ARG, "not an opLDL2 handle");
    CUDA_TRY(cudaSetDevice(M->device));
    CUDA_TRY(cudaMemcpy(rnorm, M->d.rnorm_out, sizeof(double), cudaMemcpyDeviceToHost));
    return CPK_OK;
}

int cpk_ldl2_size(cpk_handle h, int64_t *N, int64_t *nA, int64_t *nC)
{
    Ldl2 *M = lookup<Ldl2>(h, OBJ_LDL2);
    if (!M) return fail(CPK_ERR_ARG, "not an opLDL2 handle");
    if (N) *N = M->d.N;
    if (nA) *nA = M->d.nA;
    if (nC) *nC = M->d.nC;
    return CPK_OK;
}

int cpk_ldl2_info(cpk_handle h, int64_t *nnz_L_off, int64_t *levels_fwd, int64_t *levels_bwd, int64_t *n_2x2)
{
    Ldl2 *M = lookup<Ldl2>(h, OBJ_LDL2);
    if (!M) return fail(CPK_ERR_ARG, "not an opLDL2 handle");
    if (nnz_L_off) *nnz_L_off = M->nnz_off;
    if (levels_fwd) *levels_fwd = M->lev_f;
    if (levels_bwd) *levels_bwd = M->lev_b;
    if (n_2x2) *n_2x2 = M->n2x2;
    return CPK_OK;
}

static void fill_stats(cpk_stats *s, const DevStatus &d, float ms, int launches)
{
    if (!s) return;
    memset(s, 0, sizeof *s);
    s->niters = d.niters; s->solved = d.solved; s->status = d.status;
    s->error_iter = d.err_iter; s->error_second = d.err_second; s->error_value = d.err_value;
    s->hist_len = d.hist_len; s->napply = d.napply; s->nldlsolve = d.nldlsolve; s->nresid = d.nresid;
    s->shifted = d.shifted; s->launches = launches; s->t_solve_ms = ms;
    for (int i = 0; i < CPK_NPHASE; ++i) s->phase_cycles[i] = (double)d.phase_cycles[i];
}

static int status_to_rc(const DevStatus &d, int solver)
{
    if (d.err == 0) return CPK_OK;
    if (d.err == CPK_ERR_INDEFINITE) {
        const char *what = (solver == CPK_CPCGLANCZOS) ? "preconditioner not second-order sufficient"     // cpcglanczos.m:161,251
                                                       : "preconditioner does not behave as a spd matrix."; // cpminres.m:138,198
        return fail(CPK_ERR_INDEFINITE, "Iter %d, %sbeta (before sqrt) = %.5g : %s", d.err_iter,
                    d.err_second ? "2nd Lanczos vec, " : "", d.err_value, what);
    }
    if (d.err == CPK_ERR_BREAKDOWN)
        return fail(CPK_ERR_BREAKDOWN, "Iter %d, P-inner product = %.5g is negative under a square root "
                    "(complex in the reference): preconditioner is not positive definite on the constraint null space",
                    d.err_iter, d.err_value);
    if (d.err == CPK_ERR_TIMEOUT) return fail(CPK_ERR_TIMEOUT, "device watchdog fired inside a wait loop");
    return fail(d.err, "device error %d", d.err);
}

// launches a kernel for one team (grid: cooperative; cta: ordinary) and times it
// dynamic shared memory of a one-CTA launch: [solver scratch `dsm`][compact-walk region].
// Returns the region's offset (or -1 when the walk is off / does not fit) and grows *total.
static int cw_place(const DeviceCtx *dc, const DevLdl &m, bool grid, size_t dsm, size_t *total)
{
    if (grid || m.cw.nblk == 0) return -1;
    const size_t off = (dsm + 127) & ~(size_t)127;
    const size_t need = off + cw_smem_bytes(m.N, m.cw.yoff == 0);
    if (need > (size_t)dc->max_dsm) return -1;
    *total = std::max(*total, need);
    return (int)off;
}

template <class KG, class KC>
static int launch_team(DeviceCtx *dc, bool grid, KG kgrid, KC kcta, void **params, size_t dsm, float *ms,
                       const std::function<int()> &service = nullptr)
{
    CUDA_TRY(cudaMemsetAsync(dc->ctl, 0, sizeof(TeamCtl), dc->stream));
    CUDA_TRY(cudaEventRecord(dc->ev0, dc->stream));
    if (grid)
        CUDA_TRY(cudaLaunchCooperativeKernel((const void *)kgrid, dim3(dc->grid_blocks), dim3(kBlock), params, dsm, dc->stream));
    else
        CUDA_TRY(cudaLaunchKernel((const void *)kcta, dim3(1), dim3(kBlock), params, dsm, dc->stream));
    ++g_launches;
    CUDA_TRY(cudaEventRecord(dc->ev1, dc->stream));
    if (service) { const int rc = service(); if (rc) return rc; }      // matrix-free A: answer the kernel's requests until it ends
    CUDA_TRY(cudaStreamSynchronize(dc->stream));
    if (ms) CUDA_TRY(cudaEventElapsedTime(ms, dc->ev0, dc->ev1));
    return CPK_OK;
}

static int cpk_ldl2_apply_impl(cpk_handle h, const double *z, double *y, cpk_mem mem, cpk_stats *stats)
{
    CPK_LAUNCH_LOCK();
    Ldl2 *M = lookup<Ldl2>(h, OBJ_LDL2);
    if (!M || !z || !y) return fail(CPK_ERR_ARG, "cpk_ldl2_apply: bad handle or null vector");
    DeviceCtx *dc;
    int rc = get_device_ctx(M->device, &dc);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(M->device));
    const int N = M->d.N;
    const bool grid = team_is_grid(M);
    DevSystem hs{};
    hs.n = M->d.nA; hs.m = M->d.nC; hs.N = N; hs.M = M->d;
    set_walk(hs.M, M, grid);
    CUDA_TRY(cudaMemcpyAsync(M->d_sys_alone, &hs, sizeof hs, cudaMemcpyHostToDevice, dc->stream));
    const double *dz = z; double *dy = y;
    if (mem == CPK_MEM_HOST) {
        CUDA_TRY(cudaMemcpyAsync(M->d_z, z, sizeof(double) * N, cudaMemcpyHostToDevice, dc->stream));
        dz = M->d_z; dy = M->d_y;
    }
    DevStatus *d_st = M->d_status;
    CUDA_TRY(cudaMemsetAsync(d_st, 0, sizeof(DevStatus), dc->stream));
    const DevSystem *ps = M->d_sys_alone;
    size_t dsm = 0;
    int cw_off = cw_place(dc, M->d, grid, 0, &dsm);
    if (M->sweep_stale && cw_off < 0) return fail(CPK_ERR_UNSUPPORTED, "a refactorized operator only lives in the compact-walk stream, which does not fit this launch");
    void *params[] = {(void *)&ps, (void *)&dz, (void *)&dy, (void *)&d_st, (void *)&dc->ctl, (void *)&dc->partials, (void *)&cw_off};
    float ms = 0.f;
    rc = launch_team(dc, grid, k_apply<true>, k_apply<false>, params, dsm, &ms);
    if (rc) return rc;
    if (mem == CPK_MEM_HOST) CUDA_TRY(cudaMemcpy(y, M->d_y, sizeof(double) * N, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(dc->h_status, d_st, sizeof(DevStatus), cudaMemcpyDeviceToHost));
    fill_stats(stats, dc->h_status[0], ms, 1);
    return status_to_rc(dc->h_status[0], -1);
}

static int run_matvec(int device, const DevSell &A, const double *x, double *y, cpk_mem mem, cpk_stats *stats,
                      double *stage_x, double *stage_y)
{
    CPK_LAUNCH_LOCK();
    DeviceCtx *dc;
    int rc = get_device_ctx(device, &dc);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(device));
    const double *dx = x; double *dy = y;
    if (mem == CPK_MEM_HOST) {
        CUDA_TRY(cudaMemcpyAsync(stage_x, x, sizeof(double) * A.ncols, cudaMemcpyHostToDevice, dc->stream));
        dx = stage_x; dy = stage_y;
    }
    CUDA_TRY(cudaEventRecord(dc->ev0, dc->stream));
    const bool grid = use_grid(A.nrows);
    if (grid) k_matvec<true><<<dc->grid_blocks, kBlock, 0, dc->stream>>>(A, dx, dy);
    else      k_matvec<false><<<1, kBlock, 0, dc->stream>>>(A, dx, dy);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(dc->ev1, dc->stream));
    CUDA_TRY(cudaStreamSynchronize(dc->stream));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, dc->ev0, dc->ev1));
    if (mem == CPK_MEM_HOST) CUDA_TRY(cudaMemcpy(y, stage_y, sizeof(double) * A.nrows, cudaMemcpyDeviceToHost));
    if (stats) { memset(stats, 0, sizeof *stats); stats->t_solve_ms = ms; stats->launches = 1; }
    return CPK_OK;
}

static int run_matvec_rc(int device, const DevRc &A, int n, const double *x, double *y, cpk_mem mem, cpk_stats *stats,
                         double *stage_x, double *stage_y)
{
    CPK_LAUNCH_LOCK();
    DeviceCtx *dc;
    int rc = get_device_ctx(device, &dc);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(device));
    const double *dx = x; double *dy = y;
    if (mem == CPK_MEM_HOST) {
        CUDA_TRY(cudaMemcpyAsync(stage_x, x, sizeof(double) * n, cudaMemcpyHostToDevice, dc->stream));
        dx = stage_x; dy = stage_y;
    }
    CUDA_TRY(cudaEventRecord(dc->ev0, dc->stream));
    const bool grid = use_grid(n);
    if (grid) k_matvec_rc<true><<<dc->grid_blocks, kBlock, 0, dc->stream>>>(A, dx, dy);
    else      k_matvec_rc<false><<<1, kBlock, 0, dc->stream>>>(A, dx, dy);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(dc->ev1, dc->stream));
    CUDA_TRY(cudaStreamSynchronize(dc->stream));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, dc->ev0, dc->ev1));
    if (mem == CPK_MEM_HOST) CUDA_TRY(cudaMemcpy(y, stage_y, sizeof(double) * n, cudaMemcpyDeviceToHost));
    if (stats) { memset(stats, 0, sizeof *stats); stats->t_solve_ms = ms; stats->launches = 1; }
    return CPK_OK;
}

int cpk_ldl2_matvec(cpk_handle h, const double *b, double *y, cpk_mem mem, cpk_stats *stats)
{
    Ldl2 *M = lookup<Ldl2>(h, OBJ_LDL2);
    if (!M || !b || !y) return fail(CPK_ERR_ARG, "cpk_ldl2_matvec: bad handle or null vector");
    if (!M->d.resid_rc) return run_matvec(M->device, M->d.KPs, b, y, mem, stats, M->d_z, M->d_y);
    return run_matvec_rc(M->device, M->d.KP, M->d.N, b, y, mem, stats, M->d_z, M->d_y);
}

// ---------------------------------------------------------------------------
static int cpk_system_create_impl(cpk_handle *out, const cpk_csc *A, const cpk_csc *C, cpk_handle Mh)
{
    if (!out) return fail(CPK_ERR_ARG, "cpk_system_create: null output handle");
    Ldl2 *M = lookup<Ldl2>(Mh, OBJ_LDL2);
    if (!M) return fail(CPK_ERR_ARG, "cpk_system_create: M is not an opLDL2 handle");
    if (!csc_ok(A) || !csc_ok(C)) return fail(CPK_ERR_ARG, "cpk_system_create: bad matrix");
    if (A->nrows != A->ncols || C->nrows != C->ncols || A->nrows != M->d.nA || C->nrows != M->d.nC)
        return fail(CPK_ERR_DIM, "Incompatible dimensions.");
    if (M->in_system) return fail(CPK_ERR_ARG, "this opLDL2 handle already belongs to a system (its sweep buffers are not shareable)");
    const int n = M->d.nA, m = M->d.nC;
    DeviceCtx *dc;
    int rc = get_device_ctx(M->device, &dc);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(M->device));
    HCsr Hr = csr_from_csc(*A), Cr = csr_from_csc(*C);
    // blkdiag(H, C) as one SELL matrix whose first slices are exactly H
    HSell s;
    s.nrows = n + m; s.ncols = n + m;
    sell_append(s, Hr, 0, n, 0, 0);
    const int h_slices = s.nslices, h_long = (int)s.lrow.size();
    sell_append(s, Cr, 0, m, n, n);
    if (s.sptr.empty()) s.sptr.push_back(0);
    if (s.lptr.empty()) s.lptr.push_back(0);
    if ((int64_t)s.col.size() >= INT32_MAX) return fail(CPK_ERR_UNSUPPORTED, "H exceeds int32 indexing");
    HSell sc = build_sell(Cr);
    auto o = std::make_unique<System>();
    o->device = M->device; o->ar.device = M->device; o->war.device = M->device;
    o->M = M; o->M_handle = Mh;
    DevSystem &d = o->h;
    d.n = n; d.m = m; d.N = n + m;
    CUDA_TRY(upload_sell(o->ar, s, d.HC, true));           // packed entries when H, C allow it (stencil matrices)
    d.Hn = d.HC; d.Hn.nrows = n; d.Hn.ncols = n; d.Hn.nslices = h_slices; d.Hn.nlong = h_long;
    {
        const int nw[2] = {g_grid_warps_hint, kWarpsPerCta};
        for (int kdx = 0; kdx < 2; ++kdx) {
            std::vector<int> sp = warp_split(s.sptr, h_slices, nw[kdx]);
            CUDA_TRY(o->ar.upload(&d.Hn.wsplit[kdx], sp));
            d.Hn.nws[kdx] = nw[kdx];
        }
    }
    CUDA_TRY(upload_sell(o->ar, sc, d.Cm));
    d.M = M->d;
    CUDA_TRY(o->ar.alloc(&o->d_sys, 1));
    CUDA_TRY(o->ar.alloc(&o->d_args, 1));
    CUDA_TRY(o->ar.alloc(&o->d_status, 1, true));
    CUDA_TRY(o->ar.alloc(&o->d_b, d.N));
    CUDA_TRY(o->ar.alloc(&o->d_x, d.N));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)0));
    o->patH = pattern_hash(A); o->patC = pattern_hash(C);
    M->in_system = true;
    *out = register_obj(std::move(o));
    return CPK_OK;
}

int cpk_system_matvec(cpk_handle h, int which, const double *x, double *y, cpk_mem mem, cpk_stats *stats)
{
    System *S = lookup<System>(h, OBJ_SYSTEM);
    if (!S || !x || !y || which < 0 || which > 1) return fail(CPK_ERR_ARG, "cpk_system_matvec: bad argument");
    if (which == 0 && S->aop) {
        if (mem != CPK_MEM_HOST) return fail(CPK_ERR_UNSUPPORTED, "cpk_system_matvec: a matrix-free A is a host callback (host vectors only)");
        if (stats) memset(stats, 0, sizeof *stats);
        const int cb = S->aop(S->aop_ctx, x, y, S->h.n);
        return cb ? fail(CPK_ERR_ARG, "the operator callback for A*v failed (returned %d)", cb) : CPK_OK;
    }
    return run_matvec(S->device, which == 0 ? S->h.Hn : S->h.Cm, x, y, mem, stats, S->d_b, S->d_x);
}

void cpk_opts_default(cpk_opts *o, int solver, int64_t n, int64_t m)
{
    if (!o) return;
    memset(o, 0, sizeof *o);
    o->atol = 1.0e-6; o->rtol = 1.0e-6; o->btol = 0.0;
    o->itmax = (solver == CPK_CPGMRES || solver == CPK_CPDQGMRES) ? n + m : n;
    o->restart = 50; o->mem = 50; o->profile = 0;
}

int64_t cpk_hist_capacity(int solver, const cpk_opts *o)
{
    if (!o) return 0;
    int64_t cap = o->itmax + 2;
    if (solver == CPK_CPGMRES) {
        const int64_t R = std::max(1, o->restart);
        cap = ((o->itmax + R - 1) / R) * R + 2;                 // cpgmres.m:148: may overshoot itmax
    }
    return std::max<int64_t>(cap, 2);
}

struct Plan {
    int nvec;               // N-vectors of workspace incl. the 2 reserved
    size_t dsm;             // dynamic shared bytes
    long long gs;           // global scalar scratch doubles
    int wide_cols;
    int restart, mem;
};

static int make_plan(int solver, const cpk_opts *o, Plan *p)
{
    p->dsm = 0; p->gs = 0; p->wide_cols = 0; p->restart = 1; p->mem = 1;
    switch (solver) {
        case CPK_CPCG: p->nvec = 4; break;
        case CPK_CPCGLANCZOS: p->nvec = 6; break;
        case CPK_CPMINRES: p->nvec = 8; break;
        case CPK_CPSYMMLQ: p->nvec = 6; break;
        case CPK_CPGMRES: {
            if (o->restart < 1) return fail(CPK_ERR_ARG, "restart must be >= 1");
            const int R = o->restart;
            if (R > 2048) return fail(CPK_ERR_UNSUPPORTED, "restart > 2048 is not supported");
            p->restart = R;
            p->nvec = 2 + R + 1;
            p->dsm = sizeof(double) * (5 * (size_t)R + 2) + sizeof(int) * ((size_t)R + 2);
            p->gs = (long long)(R + 1) * R + R;
            p->wide_cols = R + 1;
            break;
        }
        case CPK_CPDQGMRES: {
            int mem = std::max(1, o->mem);                                          // cpdqgmres.m:117
            mem = (int)std::max<int64_t>(1, std::min<int64_t>(mem, o->itmax));      // cpdqgmres.m:125
            p->mem = mem;
            p->nvec = 2 + 2 * (mem + 1);
            p->dsm = sizeof(double) * (5 * (size_t)mem + 3 + (size_t)(mem + 2) * (mem + 3)) + sizeof(int) * ((size_t)mem + 2);
            if (p->dsm > 190 * 1024) return fail(CPK_ERR_UNSUPPORTED, "mem = %d needs %zu B of shared memory for the H band (max ~150)", mem, p->dsm);
            p->wide_cols = mem + 1;
            break;
        }
        default: return fail(CPK_ERR_ARG, "unknown solver id %d", solver);
    }
    p->nvec += 2;
    p->dsm = (p->dsm + 15) & ~(size_t)15;
    return CPK_OK;
}

static int ensure_buffers(System *S, const Plan &p, int64_t hist_cap)
{
    const long long need = (long long)p.nvec * S->h.N;
    if (need > S->work_len || hist_cap * 3 > S->hist_cap || p.gs > S->gs_len) {
        S->war.release();
        S->d_work = nullptr; S->d_hist = nullptr; S->d_gs = nullptr;
        S->work_len = std::max(need, S->work_len);
        S->hist_cap = std::max<long long>(hist_cap * 3, S->hist_cap);
        S->gs_len = std::max<long long>(p.gs, S->gs_len);
        CUDA_TRY(S->war.alloc(&S->d_work, (size_t)S->work_len));
        CUDA_TRY(S->war.alloc(&S->d_hist, (size_t)S->hist_cap));
        CUDA_TRY(S->war.alloc(&S->d_gs, (size_t)std::max<long long>(S->gs_len, 1)));
        CUDA_TRY(cudaStreamSynchronize((cudaStream_t)0));   // the launch stream does not wait for the legacy stream
    }
    return CPK_OK;
}

// Host side of the DevHostOp mailbox: runs on the calling thread between the launch of the
// persistent kernel and the end of it.  One request = copy v out, callback, copy A*v in, then the
// sequence number into `ack` on the same copy stream (so the kernel sees `u` complete).
static int serve_host_op(System *S, DeviceCtx *dc)
{
    const int n = S->h.n;
    volatile int *seq_p = S->h_mail;
    volatile long long *addr_p = reinterpret_cast<volatile long long *>(reinterpret_cast<char *>(S->h_mail) + 8);
    int *ack_src = reinterpret_cast<int *>(reinterpret_cast<char *>(S->h_mail) + 16);
    int *one_src = reinterpret_cast<int *>(reinterpret_cast<char *>(S->h_mail) + 24);
    *one_src = 1;
    int served = 0;
    S->hop_failed = 0;
    for (;;) {
        const int seq = *seq_p;
        if (seq != served) {
            const long long addr = *addr_p;
            CUDA_TRY(cudaMemcpyAsync(S->h_v, reinterpret_cast<const void *>(addr), sizeof(double) * n, cudaMemcpyDeviceToHost, S->copy_stream));
            CUDA_TRY(cudaStreamSynchronize(S->copy_stream));
            int cb = 1;
            try { cb = S->aop(S->aop_ctx, S->h_v, S->h_u, n); } catch (...) { cb = 1; }
            if (cb != 0) {
                // stop the kernel: its wait loops and barriers poll the abort flag
                S->hop_failed = cb;
                CUDA_TRY(cudaMemcpyAsync(&dc->ctl->abort, one_src, sizeof(int), cudaMemcpyHostToDevice, S->copy_stream));
            } else {
                CUDA_TRY(cudaMemcpyAsync(S->d_u, S->h_u, sizeof(double) * n, cudaMemcpyHostToDevice, S->copy_stream));
            }
            *ack_src = seq;
            CUDA_TRY(cudaMemcpyAsync(S->d_ack, ack_src, sizeof(int), cudaMemcpyHostToDevice, S->copy_stream));
            CUDA_TRY(cudaStreamSynchronize(S->copy_stream));      // ack_src is reused by the next request
            served = seq;
            continue;
        }
        const cudaError_t q = cudaEventQuery(dc->ev1);
        if (q == cudaSuccess) break;
        if (q != cudaErrorNotReady) { cudaGetLastError(); return fail(CPK_ERR_CUDA, "CUDA error while serving a matrix-free A: %s", cudaGetErrorString(q)); }
        std::this_thread::yield();
    }
    return CPK_OK;
}

static int cpk_system_create_op_impl(cpk_handle *out, int64_t n_in, cpk_matvec_fn Aop, void *ctx, const cpk_csc *C, cpk_handle Mh)
{
    if (!out) return fail(CPK_ERR_ARG, "cpk_system_create_op: null output handle");
    if (!Aop) return fail(CPK_ERR_ARG, "cpk_system_create_op: null operator callback");
    Ldl2 *M = lookup<Ldl2>(Mh, OBJ_LDL2);
    if (!M) return fail(CPK_ERR_ARG, "cpk_system_create_op: M is not an opLDL2 handle");
    if (!csc_ok(C)) return fail(CPK_ERR_ARG, "cpk_system_create_op: bad matrix");
    if (C->nrows != C->ncols || n_in != M->d.nA || C->nrows != M->d.nC) return fail(CPK_ERR_DIM, "Incompatible dimensions.");
    if (M->in_system) return fail(CPK_ERR_ARG, "this opLDL2 handle already belongs to a system (its sweep buffers are not shareable)");
    const int n = M->d.nA, m = M->d.nC;
    DeviceCtx *dc;
    int rc = get_device_ctx(M->device, &dc);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(M->device));
    HCsr Cr = csr_from_csc(*C);
    HSell sc = build_sell(Cr);
    auto o = std::make_unique<System>();
    o->device = M->device; o->ar.device = M->device; o->war.device = M->device;
    o->M = M; o->M_handle = Mh;
    DevSystem &d = o->h;
    d.n = n; d.m = m; d.N = n + m;
    memset(&d.HC, 0, sizeof d.HC); memset(&d.Hn, 0, sizeof d.Hn);      // no explicit A: the kernels never walk them (hop.on)
    d.HC.nrows = d.HC.ncols = n + m; d.Hn.nrows = d.Hn.ncols = n;
    CUDA_TRY(upload_sell(o->ar, sc, d.Cm));
    d.M = M->d;
    CUDA_TRY(o->ar.alloc(&o->d_sys, 1));
    CUDA_TRY(o->ar.alloc(&o->d_args, 1));
    CUDA_TRY(o->ar.alloc(&o->d_status, 1, true));
    CUDA_TRY(o->ar.alloc(&o->d_b, d.N));
    CUDA_TRY(o->ar.alloc(&o->d_x, d.N));
    CUDA_TRY(o->ar.alloc(&o->d_u, std::max(n, 1)));
    CUDA_TRY(o->ar.alloc(&o->d_ack, 1, true));
    CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&o->h_mail), 64, cudaHostAllocMapped));
    memset(o->h_mail, 0, 64);
    CUDA_TRY(cudaMallocHost(reinterpret_cast<void **>(&o->h_v), sizeof(double) * std::max(n, 1)));
    CUDA_TRY(cudaMallocHost(reinterpret_cast<void **>(&o->h_u), sizeof(double) * std::max(n, 1)));
    CUDA_TRY(cudaStreamCreateWithFlags(&o->copy_stream, cudaStreamNonBlocking));
    o->aop = Aop; o->aop_ctx = ctx;
    int *mail_dev = nullptr;
    CUDA_TRY(cudaHostGetDevicePointer(reinterpret_cast<void **>(&mail_dev), o->h_mail, 0));
    d.hop.on = 1;
    d.hop.req = mail_dev;
    d.hop.req_addr = reinterpret_cast<volatile long long *>(reinterpret_cast<char *>(mail_dev) + 8);
    d.hop.ack = o->d_ack;
    d.hop.u = o->d_u;
    {
        // how long a CTA waits for one answer (SM cycles ~ 2 GHz): CPK_HOSTOP_TIMEOUT_S seconds, 120 by default
        const char *e = getenv("CPK_HOSTOP_TIMEOUT_S");
        const double sec = e ? atof(e) : 120.0;
        d.hop.timeout = (long long)(std::max(sec, 0.001) * 2.0e9);
    }
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)0));
    o->patC = pattern_hash(C);
    M->in_system = true;
    *out = register_obj(std::move(o));
    return CPK_OK;
}

static int do_solve(cpk_handle h, int solver, const double *b, const cpk_opts *opts, double *x_out, double *dy_out,
                    cpk_mem mem, cpk_stats *stats, double *hist, int64_t hist_cap, int reg_mode)
{
    CPK_LAUNCH_LOCK();
    System *S = lookup<System>(h, OBJ_SYSTEM);
    if (!S) return fail(CPK_ERR_ARG, "not a system handle");
    if (!b || !x_out || !opts) return fail(CPK_ERR_ARG, "reg_cpkrylov: not enough inputs");       // reg_cpkrylov.m:122-125
    if (opts->itmax < 0) return fail(CPK_ERR_ARG, "itmax must be >= 0");
    Plan plan;
    int rc = make_plan(solver, opts, &plan);
    if (rc) return rc;
    DeviceCtx *dc;
    rc = get_device_ctx(S->device, &dc);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(S->device));
    const int n = S->h.n, m = S->h.m, N = S->h.N;
    const int64_t cap = cpk_hist_capacity(solver, opts);
    rc = ensure_buffers(S, plan, cap);
    if (rc) return rc;
    const bool grid = team_is_grid(S->M);
    if (grid && plan.wide_cols) { rc = ensure_wide(dc, plan.wide_cols); if (rc) return rc; }

    S->h.M = S->M->d;           // pick up option changes made through the opLDL2 setters
    set_walk(S->h.M, S->M, grid);
    CUDA_TRY(cudaMemcpyAsync(S->d_sys, &S->h, sizeof(DevSystem), cudaMemcpyHostToDevice, dc->stream));
    SolveArgs a{};
    a.solver = solver; a.reg_mode = reg_mode;
    const int nb = reg_mode ? N : n;
    if (mem == CPK_MEM_HOST) {
        CUDA_TRY(cudaMemcpyAsync(S->d_b, b, sizeof(double) * nb, cudaMemcpyHostToDevice, dc->stream));
        a.b = S->d_b; a.x = S->d_x;
    } else {
        a.b = b;
        // the kernel writes all N entries [x; y]: straight into the caller's buffer only when that
        // buffer has them (reg mode, or dy directly behind dx); otherwise through the internal
        // N-vector, of which dx (n entries) and, if asked for, dy (m entries) are copied out
        a.x = (reg_mode || (dy_out && dy_out == x_out + n)) ? x_out : S->d_x;
    }
    a.atol = opts->atol; a.rtol = opts->rtol; a.btol = opts->btol; a.itmax = opts->itmax;
    a.restart = plan.restart; a.mem = plan.mem; a.profile = opts->profile;
    a.work = S->d_work; a.work_len = (long long)plan.nvec * N;
    a.hist = S->d_hist; a.hist_cap = cap; a.gs = S->d_gs; a.status = S->d_status;
    size_t dsm_total = plan.dsm;
    a.cw_off = cw_place(dc, S->h.M, grid, plan.dsm, &dsm_total);
    if (S->M->sweep_stale && a.cw_off < 0) return fail(CPK_ERR_UNSUPPORTED, "a refactorized operator only lives in the compact-walk stream, which does not fit next to this solver's shared-memory scratch");
    CUDA_TRY(cudaMemcpyAsync(S->d_args, &a, sizeof a, cudaMemcpyHostToDevice, dc->stream));
    CUDA_TRY(cudaMemsetAsync(S->d_status, 0, sizeof(DevStatus), dc->stream));
    const DevSystem *ps = S->d_sys; const SolveArgs *pa = S->d_args;
    double *wide = dc->wide; int wide_cols = dc->wide_cols;
    int team_ctas = 0;          // the whole grid is one team
    void *params[] = {(void *)&ps, (void *)&pa, (void *)&dc->ctl, (void *)&dc->partials, (void *)&wide, (void *)&wide_cols, (void *)&team_ctas};
    float ms = 0.f;
    if (S->aop) {
        // a launch numbers its requests from 1: clear the mailbox and the acknowledged number
        memset(S->h_mail, 0, 64);
        CUDA_TRY(cudaMemsetAsync(S->d_ack, 0, sizeof(int), dc->stream));
        rc = launch_team(dc, grid, solver_kernel(solver, true), solver_kernel(solver, false), params, dsm_total, &ms,
                         [&] { return serve_host_op(S, dc); });
        if (!rc && S->hop_failed) return fail(CPK_ERR_ARG, "the operator callback for A*v failed (returned %d)", S->hop_failed);
    } else
        rc = launch_team(dc, grid, solver_kernel(solver, true), solver_kernel(solver, false), params, dsm_total, &ms);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpy(dc->h_status, S->d_status, sizeof(DevStatus), cudaMemcpyDeviceToHost));
    const DevStatus &st = dc->h_status[0];
    if (mem == CPK_MEM_HOST) {
        if (reg_mode) CUDA_TRY(cudaMemcpy(x_out, S->d_x, sizeof(double) * N, cudaMemcpyDeviceToHost));
        else {
            CUDA_TRY(cudaMemcpy(x_out, S->d_x, sizeof(double) * n, cudaMemcpyDeviceToHost));
            if (dy_out) CUDA_TRY(cudaMemcpy(dy_out, S->d_x + n, sizeof(double) * m, cudaMemcpyDeviceToHost));
        }
    } else if (a.x == S->d_x) {
        CUDA_TRY(cudaMemcpy(x_out, S->d_x, sizeof(double) * n, cudaMemcpyDeviceToDevice));
        if (dy_out) CUDA_TRY(cudaMemcpy(dy_out, S->d_x + n, sizeof(double) * m, cudaMemcpyDeviceToDevice));
    }
    if (hist && hist_cap > 0) {
        const int rows = solver == CPK_CPSYMMLQ ? 3 : 1;
        const int64_t len = std::min<int64_t>(std::min<int64_t>(st.hist_len, cap), hist_cap);
        for (int r = 0; r < rows; ++r)
            if (len > 0)
                CUDA_TRY(cudaMemcpy(hist + (size_t)r * hist_cap, S->d_hist + (size_t)r * cap, sizeof(double) * len, cudaMemcpyDeviceToHost));
    }
    fill_stats(stats, st, ms, 1);
    return status_to_rc(st, solver);
}

int cpk_solve(cpk_handle S, int solver, const double *b1, const cpk_opts *opts, double *dx, double *dy, cpk_mem mem,
              cpk_stats *stats, double *hist, int64_t hist_cap)
{
    return guarded([&] { return do_solve(S, solver, b1, opts, dx, dy, mem, stats, hist, hist_cap, 0); });
}

int cpk_reg_solve(cpk_handle S, int solver, const double *b, const cpk_opts *opts, double *x, cpk_mem mem,
                  cpk_stats *stats, double *hist, int64_t hist_cap)
{
    return guarded([&] { return do_solve(S, solver, b, opts, x, nullptr, mem, stats, hist, hist_cap, 1); });
}

// ---------------------------------------------------------------------------
// batch: one CTA per system, one launch
// ---------------------------------------------------------------------------
static int cpk_batch_reg_solve_impl(const cpk_handle *handles, int64_t count, int solver, const double *const *b, const cpk_opts *opts,
                        double *const *x, cpk_stats *stats, double *const *hist, int64_t hist_cap)
{
    CPK_LAUNCH_LOCK();
    if (!handles || count <= 0 || !b || !x || !opts) return fail(CPK_ERR_ARG, "cpk_batch_reg_solve: bad argument");
    if (count > kMaxBatch) return fail(CPK_ERR_UNSUPPORTED, "batch larger than %d systems: split it", kMaxBatch);
    Plan plan;
    int rc = make_plan(solver, opts, &plan);
    if (rc) return rc;
    std::vector<System *> sys(count);
    for (int64_t i = 0; i < count; ++i) {
        sys[i] = lookup<System>(handles[i], OBJ_SYSTEM);
        if (!sys[i]) return fail(CPK_ERR_ARG, "batch entry %lld is not a system handle", (long long)i);
        if (sys[i]->device != sys[0]->device) return fail(CPK_ERR_ARG, "all systems of a batch must live on one device");
        if (sys[i]->aop) return fail(CPK_ERR_UNSUPPORTED, "batch entry %lld has a matrix-free A: solve it with cpk_reg_solve", (long long)i);
        for (int64_t j = 0; j < i; ++j)
            if (sys[j] == sys[i]) return fail(CPK_ERR_ARG, "system %lld appears twice in the batch", (long long)i);
    }
    DeviceCtx *dc;
    rc = get_device_ctx(sys[0]->device, &dc);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(sys[0]->device));
    const int64_t cap = cpk_hist_capacity(solver, opts);
    std::vector<DevSystem> hsys(count);
    std::vector<SolveArgs> hargs(count);
    size_t dsm_total = plan.dsm;
    // one pinned staging area for all right-hand sides and solutions: the copies of the whole
    // batch are queued asynchronously around the single launch
    std::vector<size_t> off(count + 1, 0);
    for (int64_t i = 0; i < count; ++i) off[i + 1] = off[i] + (size_t)sys[i]->h.N;
    const size_t tot = off[count];
    const bool want_hist = hist && hist_cap > 0;
    const size_t hrows = solver == CPK_CPSYMMLQ ? 3 : 1;
    const size_t htot = want_hist ? (size_t)count * 3 * (size_t)cap : 0;    // device layout: 3 rows of `cap` per system
    const size_t need = 2 * tot + htot;
    if (need > dc->h_stage_cap) {
        if (dc->h_stage) cudaFreeHost(dc->h_stage);
        if (dc->d_stage) cudaFree(dc->d_stage);
        dc->h_stage = nullptr; dc->d_stage = nullptr; dc->h_stage_cap = 0;
        CUDA_TRY(cudaMallocHost(&dc->h_stage, sizeof(double) * need));
        CUDA_TRY(cudaMalloc(&dc->d_stage, sizeof(double) * need));
        dc->h_stage_cap = need;
    }
    if ((size_t)count > dc->d_batch_cap) {
        if (dc->d_bsys) cudaFree(dc->d_bsys);
        if (dc->d_bargs) cudaFree(dc->d_bargs);
        if (dc->d_bstatus) cudaFree(dc->d_bstatus);
        dc->d_bsys = nullptr; dc->d_bargs = nullptr; dc->d_bstatus = nullptr; dc->d_batch_cap = 0;
        CUDA_TRY(cudaMalloc(&dc->d_bsys, sizeof(DevSystem) * count));
        CUDA_TRY(cudaMalloc(&dc->d_bargs, sizeof(SolveArgs) * count));
        CUDA_TRY(cudaMalloc(&dc->d_bstatus, sizeof(DevStatus) * count));
        dc->d_batch_cap = (size_t)count;
    }
    double *hb = dc->h_stage, *hx = dc->h_stage + tot;
    double *db = dc->d_stage, *dx = dc->d_stage + tot;
    // launch order: the systems small enough for a one-CTA team first (ONE launch, CTA t on
    // system t), then the larger ones in waves of cooperative launches whose grid is cut into
    // sub-teams of `tc` CTAs, one system per sub-team
    std::vector<int64_t> order;
    order.reserve(count);
    for (int64_t i = 0; i < count; ++i) if (!use_grid(sys[i]->h.N)) order.push_back(i);
    const int64_t n_cta = (int64_t)order.size();
    for (int64_t i = 0; i < count; ++i) if (use_grid(sys[i]->h.N)) order.push_back(i);
    int tc = dc->grid_blocks;
    if (n_cta < count) {
        int maxN = 0;
        for (int64_t q = n_cta; q < count; ++q) maxN = std::max(maxN, sys[order[q]]->h.N);
        // sub-team size: at least ~8192 unknowns per CTA (smaller teams are bound by their barriers:
        // N = 80 000, 14 systems: 4 CTAs 11.9 ms, 10 CTAs 5.7 ms, whole grid one by one 11.7 ms);
        // when the whole batch fits in one wave the SMs left over widen the teams
        const int64_t n_grid = count - n_cta;
        const int tc_lo = std::min(dc->grid_blocks, std::max(2, (maxN + 8191) / 8192));
        const int tc_hi = std::min(dc->grid_blocks, std::max(tc_lo, (maxN + 2047) / 2048));
        tc = tc_lo;
        if (n_grid * tc_lo <= dc->grid_blocks) tc = std::min(tc_hi, (int)(dc->grid_blocks / n_grid));
        if (const char *e = getenv("CPK_TEAM_CTAS")) { const int v = atoi(e); if (v >= 1 && v <= dc->grid_blocks) tc = v; }
        if (plan.wide_cols) { rc = ensure_wide(dc, plan.wide_cols); if (rc) return rc; }
    }
    for (int64_t q = 0; q < count; ++q) {
        const int64_t i = order[q];
        System *S = sys[i];
        rc = ensure_buffers(S, plan, cap);
        if (rc) return rc;
        S->h.M = S->M->d;
        set_walk(S->h.M, S->M, q >= n_cta);
        hsys[q] = S->h;
        SolveArgs a{};
        a.solver = solver; a.reg_mode = 1;
        if (!b[i] || !x[i]) return fail(CPK_ERR_ARG, "batch entry %lld: null vector", (long long)i);
        memcpy(hb + off[i], b[i], sizeof(double) * S->h.N);
        a.b = db + off[i]; a.x = dx + off[i];
        a.atol = opts->atol; a.rtol = opts->rtol; a.btol = opts->btol; a.itmax = opts->itmax;
        a.restart = plan.restart; a.mem = plan.mem; a.profile = opts->profile;
        a.work = S->d_work; a.work_len = (long long)plan.nvec * S->h.N;
        a.hist = want_hist ? dc->d_stage + 2 * tot + (size_t)i * 3 * cap : S->d_hist;
        a.hist_cap = cap; a.gs = S->d_gs; a.status = dc->d_bstatus + i;
        a.cw_off = q < n_cta ? cw_place(dc, S->h.M, false, plan.dsm, &dsm_total) : -1;
        if (S->M->sweep_stale && a.cw_off < 0) return fail(CPK_ERR_UNSUPPORTED, "batch entry %lld: a refactorized operator only lives in the compact-walk stream, which does not fit this launch", (long long)i);
        hargs[q] = a;
    }
    CUDA_TRY(cudaMemcpyAsync(db, hb, sizeof(double) * tot, cudaMemcpyHostToDevice, dc->stream));
    CUDA_TRY(cudaMemsetAsync(dc->d_bstatus, 0, sizeof(DevStatus) * count, dc->stream));
    CUDA_TRY(cudaMemcpyAsync(dc->d_bsys, hsys.data(), sizeof(DevSystem) * count, cudaMemcpyHostToDevice, dc->stream));
    CUDA_TRY(cudaMemcpyAsync(dc->d_bargs, hargs.data(), sizeof(SolveArgs) * count, cudaMemcpyHostToDevice, dc->stream));
    CUDA_TRY(cudaMemsetAsync(dc->ctl, 0, sizeof(TeamCtl) * kMaxBatch, dc->stream));
    CUDA_TRY(cudaEventRecord(dc->ev0, dc->stream));
    int launches = 0;
    if (n_cta > 0) {
        const DevSystem *d_sys = dc->d_bsys; const SolveArgs *d_args = dc->d_bargs;
        double *nullp = nullptr; int zero = 0;
        void *params[] = {(void *)&d_sys, (void *)&d_args, (void *)&dc->ctl, (void *)&nullp, (void *)&nullp, (void *)&zero, (void *)&zero};
        // A batch wants throughput, not the latency of one system: with fewer threads per CTA two CTAs
        // share an SM (registers and shared memory allow it), and the level-bound LDL' walks of the two
        // systems overlap.  The walkers of the compact walk must all be there.
        // (256 systems of the cvxqp1 pattern: 23.4 ms with 896 threads per CTA, 17.7 ms with 448, 18.7 ms
        // with 384; batches that leave SMs idle keep the full CTA)
        static const int batch_env = [] { const char *e = getenv("CPK_BATCH_BLOCK"); return e ? atoi(e) : 0; }();
        int batch_block = batch_env > 0 ? batch_env : (n_cta > dc->num_sms ? kBatchBlock : kBlock);
        batch_block = std::max(32 * kCwWarps, std::min(kBlock, (batch_block + 31) & ~31));
        CUDA_TRY(cudaLaunchKernel(solver_kernel(solver, false), dim3((unsigned)n_cta), dim3(batch_block), params, dsm_total, dc->stream));
        ++launches;
    }
    const int per_wave = std::max(1, dc->grid_blocks / tc);
    for (int64_t q0 = n_cta; q0 < count; q0 += per_wave) {
        const int teams = (int)std::min<int64_t>(per_wave, count - q0);
        const DevSystem *d_sys = dc->d_bsys + q0; const SolveArgs *d_args = dc->d_bargs + q0;
        TeamCtl *ctl = dc->ctl + q0;
        double *wide = dc->wide; int wide_cols = dc->wide_cols;
        void *params[] = {(void *)&d_sys, (void *)&d_args, (void *)&ctl, (void *)&dc->partials, (void *)&wide, (void *)&wide_cols, (void *)&tc};
        CUDA_TRY(cudaLaunchCooperativeKernel(solver_kernel(solver, true), dim3((unsigned)(teams * tc)), dim3(kBlock), params, plan.dsm, dc->stream));
        ++launches;
    }
    g_launches += launches;
    CUDA_TRY(cudaEventRecord(dc->ev1, dc->stream));
    CUDA_TRY(cudaMemcpyAsync(dc->h_status, dc->d_bstatus, sizeof(DevStatus) * count, cudaMemcpyDeviceToHost, dc->stream));
    CUDA_TRY(cudaMemcpyAsync(hx, dx, sizeof(double) * (tot + htot), cudaMemcpyDeviceToHost, dc->stream));     // solutions (+ histories)
    CUDA_TRY(cudaStreamSynchronize(dc->stream));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, dc->ev0, dc->ev1));
    int first_rc = CPK_OK;
    std::string first_msg;
    for (int64_t i = 0; i < count; ++i) {
        System *S = sys[i];
        memcpy(x[i], hx + off[i], sizeof(double) * S->h.N);
        const DevStatus &st = dc->h_status[i];
        if (want_hist && hist[i]) {
            const int64_t len = std::min<int64_t>(std::min<int64_t>(st.hist_len, cap), hist_cap);
            const double *hh = dc->h_stage + 2 * tot + (size_t)i * 3 * cap;
            for (size_t r = 0; r < hrows; ++r)
                if (len > 0) memcpy(hist[i] + r * hist_cap, hh + r * cap, sizeof(double) * len);
        }
        if (stats) fill_stats(&stats[i], st, ms, i == 0 ? launches : 0);
        const int r = status_to_rc(st, solver);
        if (r && !first_rc) { first_rc = r; first_msg = "system " + std::to_string(i) + ": " + g_err; }
    }
    if (first_rc) { g_err = first_msg; return first_rc; }
    return CPK_OK;
}

// ---------------------------------------------------------------------------
// guarded public entry points
// ---------------------------------------------------------------------------
int cpk_ldl2_create(cpk_handle *out, const cpk_csc *A, const cpk_csc *B, const cpk_csc *C,
                    const cpk_csc *L, const cpk_csc *D, const int64_t *perm, int device)
{
    return guarded([&] { return cpk_ldl2_create_impl(out, A, B, C, L, D, perm, device); });
}

int cpk_ldl2_create_sqd(cpk_handle *out, const cpk_csc *A, const cpk_csc *B, const cpk_csc *C, const int64_t *perm, int device)
{
    return guarded([&] { return cpk_ldl2_create_sqd_impl(out, A, B, C, perm, device); });
}

int cpk_ldl2_refactor(cpk_handle h, const cpk_csc *A, const cpk_csc *B, const cpk_csc *C)
{
    return guarded([&] { return cpk_ldl2_refactor_impl(h, A, B, C); });
}

int cpk_system_update(cpk_handle h, const cpk_csc *A, const cpk_csc *C)
{
    return guarded([&] { return cpk_system_update_impl(h, A, C); });
}

int cpk_system_create(cpk_handle *out, const cpk_csc *A, const cpk_csc *C, cpk_handle Mh)
{
    return guarded([&] { return cpk_system_create_impl(out, A, C, Mh); });
}

int cpk_system_create_op(cpk_handle *out, int64_t n, cpk_matvec_fn Aop, void *ctx, const cpk_csc *C, cpk_handle Mh)
{
    return guarded([&] { return cpk_system_create_op_impl(out, n, Aop, ctx, C, Mh); });
}

int cpk_batch_reg_solve(const cpk_handle *handles, int64_t count, int solver, const double *const *b, const cpk_opts *opts,
                        double *const *x, cpk_stats *stats, double *const *hist, int64_t hist_cap)
{
    return guarded([&] { return cpk_batch_reg_solve_impl(handles, count, solver, b, opts, x, stats, hist, hist_cap); });
}

int cpk_ldl2_apply(cpk_handle h, const double *z, double *y, cpk_mem mem, cpk_stats *stats)
{
    return guarded([&] { return cpk_ldl2_apply_impl(h, z, y, mem, stats); });
}

int cpk_destroy(cpk_handle h)
{
    std::unique_ptr<Object> victim;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_obj.find(h);
        if (it == g_obj.end()) return fail(CPK_ERR_ARG, "unknown handle");
        if (it->second->kind == OBJ_LDL2 && static_cast<Ldl2 *>(it->second.get())->in_system)
            return fail(CPK_ERR_ARG, "opLDL2 handle is still owned by a system; destroy the system first");
        if (it->second->kind == OBJ_SYSTEM) static_cast<System *>(it->second.get())->M->in_system = false;
        victim = std::move(it->second);
        g_obj.erase(it);
    }
    return CPK_OK;
}

int cpk_destroy_all(void)
{
    std::lock_guard<std::mutex> lk(g_mu);
    // systems first (they reference their opLDL2)
    for (auto it = g_obj.begin(); it != g_obj.end();)
        if (it->second->kind == OBJ_SYSTEM) it = g_obj.erase(it); else ++it;
    g_obj.clear();
    return CPK_OK;
}

