/*
 * cpk_b200.h -- C ABI of libcpk_b200.so: the B200 (sm_100a) implementation of
 * cpkrylov's per-iteration hot path (constraint-preconditioned Krylov solvers
 * on [H B'; B -C] with the opLDL2 preconditioner apply).
 *
 * Every entry point is what a MEX gateway (MATLAB) or a ctypes stub (Python)
 * binds; nothing here mentions torch, numpy or mxArray.  Each declaration names
 * the reference interface it replaces (paths relative to the cpkrylov tree).
 *
 * Conventions
 *   - all reals are IEEE fp64; all indices are 0-based int64 (MATLAB mwIndex);
 *   - sparse inputs are CSC exactly as MATLAB stores them (mxGetJc/Ir/Pr);
 *   - every function returns 0 (CPK_OK) or a negative cpk_status; the message
 *     for the most recent failure on the calling thread is cpk_last_error();
 *   - the library copies its inputs to the device at create time; the caller
 *     may free them at once.  Outputs go to caller-allocated buffers;
 *   - vector arguments are host pointers unless the call takes a `mem`
 *     argument, in which case CPK_MEM_DEVICE means "device pointer on the
 *     handle's GPU" (used to keep rhs/solution resident in HBM);
 *   - calls are synchronous (return after the handle's stream has drained);
 *     the entry points that launch kernels serialise on one process-wide lock
 *     (a solve owns every SM of its device), so the library may be called from
 *     several host threads; a handle must still not be destroyed while another
 *     thread uses it.
 *   - `hist` buffers are always HOST pointers, whatever `mem` says.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point
 *     fails with CPK_ERR_CUDA.
 */
#ifndef CPK_B200_H
#define CPK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint64_t cpk_handle;

typedef enum cpk_status {
    CPK_OK              =  0,
    CPK_ERR_ARG         = -1,   /* bad argument / unknown handle                          */
    CPK_ERR_DIM         = -2,   /* opLDL2.m:68-75 'must be square' / 'Incompatible dimensions.' */
    CPK_ERR_ALLOC       = -3,
    CPK_ERR_CUDA        = -4,   /* no device, launch failure, ...                          */
    CPK_ERR_INDEFINITE  = -5,   /* cpminres.m:135-139,195-199; cpsymmlq.m:142-146,206-209,274-278;
                                   cpcglanczos.m:158-163,248-254 (P-inner product < -100*eps)      */
    CPK_ERR_BREAKDOWN   = -6,   /* negative P-inner product under a sqrt in cpgmres/cpdqgmres/cpcg
                                   (complex in MATLAB: cpgmres.m:173-176,219-222, cpcg.m:175)     */
    CPK_ERR_TIMEOUT     = -7,   /* device-side watchdog fired (never expected)             */
    CPK_ERR_UNSUPPORTED = -8
} cpk_status;

typedef enum cpk_solver {       /* the `method` handle of reg_cpkrylov.m:67-68             */
    CPK_CPCG        = 0,        /* kernels/cpcg.m        */
    CPK_CPCGLANCZOS = 1,        /* kernels/cpcglanczos.m */
    CPK_CPMINRES    = 2,        /* kernels/cpminres.m    */
    CPK_CPSYMMLQ    = 3,        /* kernels/cpsymmlq.m    */
    CPK_CPGMRES     = 4,        /* kernels/cpgmres.m     */
    CPK_CPDQGMRES   = 5         /* kernels/cpdqgmres.m   */
} cpk_solver;

typedef enum cpk_mem { CPK_MEM_HOST = 0, CPK_MEM_DEVICE = 1 } cpk_mem;

/* MATLAB sparse matrix, column-compressed (mxGetM/N/Jc/Ir/Pr). */
typedef struct cpk_csc {
    int64_t        nrows, ncols;
    const int64_t *colptr;      /* ncols+1 */
    const int64_t *rowind;      /* nnz, 0-based, any order inside a column */
    const double  *val;         /* nnz */
} cpk_csc;

/* The `opts` struct of the solvers (isfield logic: cpcg.m:101-115,
 * cpcglanczos.m:115-132, cpminres.m:97-111, cpsymmlq.m:104-118,
 * cpgmres.m:107-127, cpdqgmres.m:105-122).  Fill with cpk_opts_default()
 * (= "field absent"), then overwrite what the caller's struct has. */
typedef struct cpk_opts {
    double  atol;               /* 1e-6 */
    double  rtol;               /* 1e-6 */
    double  btol;               /* cpcglanczos only, 0 */
    int64_t itmax;              /* n (cg, cglanczos, minres, symmlq) or n+m (gmres, dqgmres) */
    int32_t restart;            /* cpgmres, 50 */
    int32_t mem;                /* cpdqgmres, 50 (clamped to [1,itmax], cpdqgmres.m:117,125) */
    int32_t profile;            /* 1: collect per-phase cycle counters (cpk_stats.phase_cycles) */
    int32_t reserved;
} cpk_opts;

#define CPK_NPHASE 8
/* phase_cycles[] slots (SM cycles seen by team thread 0; shares, not absolute time) */
#define CPK_PH_SPMV   0   /* H*v, C*q (+fused dots)                         */
#define CPK_PH_LDL    1   /* P L^-T D^-1 L^-1 P' solves                     */
#define CPK_PH_RESID  2   /* refinement residual x - K_P*y (+norms)          */
#define CPK_PH_VEC    3   /* Krylov vector algebra incl. reductions          */
#define CPK_PH_OTHER  4   /* prologue/epilogue (rhs shift, un-shift)         */

/* stats/flag structs of the solvers (e.g. cpminres.m:250-252) plus timings
 * of reg_cpkrylov.m:175-178. */
typedef struct cpk_stats {
    int64_t niters;             /* stats.niters                                          */
    int32_t solved;             /* flag.solved                                           */
    int32_t status;             /* cpcglanczos stats.status: 0 itmax, 1 residual small,
                                   2 backward error small (cpcglanczos.m:312-324)        */
    int32_t error_iter;         /* iteration index of an INDEFINITE/BREAKDOWN error      */
    int32_t error_second;       /* 1: raised while building the 2nd Lanczos vector (cpsymmlq.m:206) */
    double  error_value;        /* offending beta (before sqrt)                          */
    int64_t hist_len;           /* entries written per history row                       */
    int64_t napply;             /* preconditioner applies M*z                            */
    int64_t nldlsolve;          /* LDL' solves inside those applies (1 + refinement steps) */
    int64_t nresid;             /* K_P residual evaluations inside those applies         */
    int32_t shifted;            /* reg_cpkrylov.m:154: rhs tail was nonzero              */
    int32_t launches;           /* kernel launches made by this call                     */
    double  t_solve_ms;         /* CUDA-event time of the device work of this call (stats.stime analogue) */
    double  phase_cycles[CPK_NPHASE];
} cpk_stats;

/* ---- library / device ---------------------------------------------------- */
int  cpk_version(void);
/* number of CUDA devices visible (0 without a GPU; never fails) */
int  cpk_device_count(void);
/* copies the calling thread's last error text (NUL-terminated) into buf */
int  cpk_last_error(char *buf, int64_t buflen);
/* total kernel launches issued by the library in this process */
int64_t cpk_launch_count(void);

/* ---- opLDL2: ops/opLDL2.m ------------------------------------------------- */
/* Constructor opLDL2(A,B,C) (opLDL2.m:60-92) with the factorization of :82
 * ([L,D,P] = ldl([A B'; B C]), P'*K*P = L*D*L') supplied by the caller:
 *   A  nA x nA symmetric (the G of reg_cpkrylov.m:131), B  nC x nA, C  nC x nC
 *   L  N x N unit lower triangular (diagonal entries optional)
 *   D  N x N block diagonal, 1x1 and 2x2 blocks
 *   perm[k] = row index of the single 1 in column k of P, i.e. (P'x)[k] = x[perm[k]]
 * Errors: CPK_ERR_DIM with the reference's texts (opLDL2.m:68-75). */
int  cpk_ldl2_create(cpk_handle *M, const cpk_csc *A, const cpk_csc *B, const cpk_csc *C,
                     const cpk_csc *L, const cpk_csc *D, const int64_t *perm, int device);
/* public properties opLDL2.m:45-50, setters :97-115 (nitref := max(0,round(v))) */
int  cpk_ldl2_set_nitref(cpk_handle M, double v);
int  cpk_ldl2_set_itref_tol(cpk_handle M, double v);
int  cpk_ldl2_set_force_itref(cpk_handle M, int v);
int  cpk_ldl2_set_residual_update(cpk_handle M, int v);
/* Spot value-class (0, default: Aty/Cy assignments of opLDL2.m:170-171 are lost)
 * or handle-class (1: kept between applies) reading of residual_update. */
int  cpk_ldl2_set_ru_stateful(cpk_handle M, int v);
/* 1: also evaluate the residual after the last refinement step (only feeds
 * op.rNorm, opLDL2.m:183-186); 0 (default) skips that dead K_P product. */
int  cpk_ldl2_set_track_rnorm(cpk_handle M, int v);
int  cpk_ldl2_get_rnorm(cpk_handle M, double *rnorm);
/* size(M) = [N N] */
int  cpk_ldl2_size(cpk_handle M, int64_t *N, int64_t *nA, int64_t *nC);
/* y = M*z  (opLDL2.m:161-188: solve, residual update, iterative refinement) */
int  cpk_ldl2_apply(cpk_handle M, const double *z, double *y, cpk_mem mem, cpk_stats *stats);
/* y = M\b = K_P*b  (opLDL2.m:193-195) */
int  cpk_ldl2_matvec(cpk_handle M, const double *b, double *y, cpk_mem mem, cpk_stats *stats);
/* structure report: levels / items of the forward and backward sweeps */
int  cpk_ldl2_info(cpk_handle M, int64_t *nnz_L_off, int64_t *levels_fwd, int64_t *levels_bwd,
                   int64_t *n_2x2);

/* ---- device-side factorization for sequences with a fixed pattern ----
 * The reference re-assembles K_P and calls ldl for every system of an interior-point
 * run (opLDL2.m:81-82).  For symmetric quasi-definite K_P = [A B'; B C] (A positive,
 * C negative definite: every symmetric permutation has an LDL' factorization with a
 * diagonal D) the library factorizes on the device with a STATIC permutation:
 * perm[k] = original index of row k of the permuted matrix, chosen once for fill
 * (e.g. the column permutation of a first host factorization).  One-CTA systems only
 * (N <= ~22 000); A and C bring both triangles, as MATLAB stores them. */
int  cpk_ldl2_create_sqd(cpk_handle *M, const cpk_csc *A, const cpk_csc *B, const cpk_csc *C,
                         const int64_t *perm, int device);
/* next system of the sequence: new values, SAME sparsity patterns (entry order included).
 * Symbolic work, level schedule and device layouts are reused; the numeric LDL' runs on
 * the device and rewrites the operator in place. */
int  cpk_ldl2_refactor(cpk_handle M, const cpk_csc *A, const cpk_csc *B, const cpk_csc *C);
/* the factor the device computed: strict lower triangle of L (CSC, pattern with fill) and
 * diag(D).  Any output may be NULL; call once with colptr/rowind/Lval NULL to get nnz. */
int  cpk_ldl2_get_factor(cpk_handle M, int64_t *nnz, int64_t *colptr, int64_t *rowind,
                         double *Lval, double *d);

/* ---- system = the (A, C, M) arguments of method(b1, A, C, M, opts) -------- */
/* A n x n (may be nonsymmetric), C m x m symmetric, M from cpk_ldl2_create on
 * the same device.  B (m x n) is only used by cpk_reg_solve for the rhs shift
 * B'*y0 (reg_cpkrylov.m:157); it is taken from M. */
int  cpk_system_create(cpk_handle *S, const cpk_csc *A, const cpk_csc *C, cpk_handle M);
/* Matrix-free A: reg_cpkrylov.m:40 "the argument A may be a matrix or a linear operator"
 * (a Spot operator in the reference; the solvers only ever form A*v: cpcg.m:151,
 * cpcglanczos.m:228, cpminres.m:187, cpsymmlq.m:238, cpgmres.m:209, cpdqgmres.m:205, and the
 * rhs shift reg_cpkrylov.m:157).  `Aop(ctx, v, u, n)` must write u = A*v for HOST n-vectors and
 * return 0 (anything else aborts the solve with CPK_ERR_ARG).  The solve stays ONE persistent
 * kernel with every Krylov vector, the preconditioner apply and the recurrences on the device;
 * per product only v and A*v cross the bus: the kernel posts a request in mapped host memory,
 * the calling thread answers it from inside cpk_solve / cpk_reg_solve (the callback runs on
 * that thread) and the kernel resumes.  The callback must not launch kernels on the handle's
 * GPU (the resident solver owns every SM) nor call this library.  A CTA waits at most
 * CPK_HOSTOP_TIMEOUT_S seconds (environment, default 120) for an answer. */
typedef int (*cpk_matvec_fn)(void *ctx, const double *v, double *u, int64_t n);
int  cpk_system_create_op(cpk_handle *S, int64_t n, cpk_matvec_fn Aop, void *ctx,
                          const cpk_csc *C, cpk_handle M);
/* new values of A (= H) and C with the patterns the system was created with (the other
 * half of a sequence step next to cpk_ldl2_refactor; reg_cpkrylov.m:1 takes A and C anew
 * for every system) */
int  cpk_system_update(cpk_handle S, const cpk_csc *A, const cpk_csc *C);
/* y = A*x (which=0, n-vector) or y = C*x (which=1, m-vector): the sparse
 * mtimes call sites cpcg.m:151-152 etc. */
int  cpk_system_matvec(cpk_handle S, int which, const double *x, double *y, cpk_mem mem,
                       cpk_stats *stats);

void cpk_opts_default(cpk_opts *o, int solver, int64_t n, int64_t m);

/* [x, y, stats, flag] = method(b1, A, C, M, opts)   (e.g. cpminres.m:1)
 *   b1 n-vector; dx n-vector, dy m-vector out.
 *   hist: 3 rows of hist_cap doubles (row 0 = residHistory; cpsymmlq: rows
 *   0,1,2 = cg/lq/qr residHistory, cpsymmlq.m:363-366); may be NULL.
 * Returns CPK_ERR_INDEFINITE / CPK_ERR_BREAKDOWN for the reference's error()
 * paths; non-convergence is NOT an error (stats->solved = 0). */
int  cpk_solve(cpk_handle S, int solver, const double *b1, const cpk_opts *opts,
               double *dx, double *dy, cpk_mem mem,
               cpk_stats *stats, double *hist, int64_t hist_cap);
/* body of reg_cpkrylov.m:150-175: rhs shift if any(b(n+1:N)), solve, un-shift.
 *   b, x: N-vectors. */
int  cpk_reg_solve(cpk_handle S, int solver, const double *b, const cpk_opts *opts,
                   double *x, cpk_mem mem,
                   cpk_stats *stats, double *hist, int64_t hist_cap);
/* history length needed for (solver, opts): itmax+1 (+ a restart cycle for cpgmres) */
int64_t cpk_hist_capacity(int solver, const cpk_opts *opts);

/* Batch of independent systems on one GPU: the sharded unit of the multi-GPU path
 * (what a loop over reg_cpkrylov.m:1 calls does for the systems of an interior-point
 * run).  Systems small enough for a one-CTA team share ONE launch (CTA t solves system
 * t); larger ones share cooperative launches whose grid is cut into sub-teams of
 * consecutive CTAs, one system per sub-team, as many waves as needed
 * (stats[0].launches = launches of the whole call).  All systems must live on the same
 * device.  b/x: arrays of `count` host pointers (N_i each); hist may be NULL or
 * `count` pointers to 3*hist_cap doubles. */
int  cpk_batch_reg_solve(const cpk_handle *S, int64_t count, int solver, const double *const *b,
                         const cpk_opts *opts, double *const *x, cpk_stats *stats,
                         double *const *hist, int64_t hist_cap);

int  cpk_destroy(cpk_handle h);           /* any handle kind */
int  cpk_destroy_all(void);               /* mexAtExit hook   */

#ifdef __cplusplus
}
#endif
#endif /* CPK_B200_H */
