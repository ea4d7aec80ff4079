/*
 * oracle/kernels.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the three MATLAB builtins that cpkrylov's hot path
 * leans on and that are not under /root/reference (they live inside the closed
 * MATLAB runtime): sparse mtimes, sparse triangular "\" and the block-diagonal
 * D "\" of ldl().  Call sites that these stand in for:
 *   sparse mtimes     kernels/cpcg.m:151-152, kernels/cpminres.m:187-188,
 *                     ops/opLDL2.m:170-171,175,182, reg_cpkrylov.m:157
 *   L \ . , L' \ .    ops/opLDL2.m:86  (inv(op.L), inv(op.L'))
 *   D \ .             ops/opLDL2.m:86  (inv(op.D)), D from [L,D,P]=ldl(K) (:82)
 *
 * Single-threaded scalar loops, sequential left-to-right accumulation (the
 * order a CSC/CSR CPU kernel uses).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product (cpkrylov_b200/) never does.
 *
 * Parity status: UNPINNED -- the reference ships no golden vectors and neither
 * MATLAB nor Octave exists in the build container.
 */
#include <stdint.h>
#include <stddef.h>

/* y = A*x, A in CSR (rowptr[nrows+1], col, val). */
void orc_csr_matvec(int64_t nrows, const int64_t *rowptr, const int64_t *col,
                    const double *val, const double *x, double *y)
{
    for (int64_t i = 0; i < nrows; ++i) {
        double s = 0.0;
        for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k)
            s += val[k] * x[col[k]];
        y[i] = s;
    }
}

/* Solve L w = b in place (w holds b on entry).  L is unit lower triangular,
 * stored as CSR of its STRICT lower part (no diagonal entries). */
void orc_unit_lower_solve(int64_t n, const int64_t *rowptr, const int64_t *col,
                          const double *val, double *w)
{
    for (int64_t i = 0; i < n; ++i) {
        double s = w[i];
        for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k)
            s -= val[k] * w[col[k]];
        w[i] = s;
    }
}

/* Solve L' w = b in place.  Input is CSR of the strict lower part of L (the
 * same arrays as above): column sweep from the last row upward. */
void orc_unit_lower_transpose_solve(int64_t n, const int64_t *rowptr,
                                    const int64_t *col, const double *val,
                                    double *w)
{
    for (int64_t i = n - 1; i >= 0; --i) {
        const double wi = w[i];
        for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k)
            w[col[k]] -= val[k] * wi;
    }
}

/* Solve D w = b in place.  D is symmetric block diagonal with 1x1 and 2x2
 * blocks: d[i] = D(i,i), e[i] = D(i+1,i) (e[i] != 0 marks a 2x2 block that
 * starts at i; e has n entries, e[n-1] = 0).  2x2 blocks are solved by
 * Cramer's rule on the symmetric block [d_i e_i; e_i d_{i+1}]. */
void orc_block_diag_solve(int64_t n, const double *d, const double *e, double *w)
{
    int64_t i = 0;
    while (i < n) {
        if (i + 1 < n && e[i] != 0.0) {
            const double a = d[i], b = e[i], c = d[i + 1];
            const double det = a * c - b * b;
            const double w0 = w[i], w1 = w[i + 1];
            w[i]     = (c * w0 - b * w1) / det;
            w[i + 1] = (a * w1 - b * w0) / det;
            i += 2;
        } else {
            w[i] = w[i] / d[i];
            i += 1;
        }
    }
}

/* ---------------------------------------------------------------------------
 * Extended-precision twins (vectors and accumulation in `long double`; matrix and factor
 * entries stay the fp64 data the fp64 run uses).  They serve the ARBITER of the parity tests
 * (cpk_oracle.extended_precision): which of two fp64 runs is closer to the exactly-rounded
 * result of the same algorithm.  Same loops, same order.
 * ------------------------------------------------------------------------- */
void orc_csr_matvec_ld(int64_t nrows, const int64_t *rowptr, const int64_t *col,
                       const double *val, const long double *x, long double *y)
{
    for (int64_t i = 0; i < nrows; ++i) {
        long double s = 0.0L;
        for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k)
            s += (long double)val[k] * x[col[k]];
        y[i] = s;
    }
}

void orc_unit_lower_solve_ld(int64_t n, const int64_t *rowptr, const int64_t *col,
                             const double *val, long double *w)
{
    for (int64_t i = 0; i < n; ++i) {
        long double s = w[i];
        for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k)
            s -= (long double)val[k] * w[col[k]];
        w[i] = s;
    }
}

void orc_unit_lower_transpose_solve_ld(int64_t n, const int64_t *rowptr,
                                       const int64_t *col, const double *val,
                                       long double *w)
{
    for (int64_t i = n - 1; i >= 0; --i) {
        const long double wi = w[i];
        for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k)
            w[col[k]] -= (long double)val[k] * wi;
    }
}

void orc_block_diag_solve_ld(int64_t n, const double *d, const double *e, long double *w)
{
    int64_t i = 0;
    while (i < n) {
        if (i + 1 < n && e[i] != 0.0) {
            const long double a = d[i], b = e[i], c = d[i + 1];
            const long double det = a * c - b * b;
            const long double w0 = w[i], w1 = w[i + 1];
            w[i]     = (c * w0 - b * w1) / det;
            w[i + 1] = (a * w1 - b * w0) / det;
            i += 2;
        } else {
            w[i] = w[i] / (long double)d[i];
            i += 1;
        }
    }
}
