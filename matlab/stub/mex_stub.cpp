// matlab/stub/mex_stub.cpp -- TEST INFRASTRUCTURE: a toy MEX runtime (see mex.h next to it) plus a
// C interface through which a test (ctypes) builds mxArrays, calls the gateway's mexFunction and
// reads the results back.  error() inside the gateway (mexErrMsgIdAndTxt) unwinds to stub_call,
// which returns 1 and keeps identifier + message, like MATLAB's try/catch would.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "mex.h"

struct mxArray_tag {
    bool sparse = false, is_char = false, is_logical = false;
    size_t m = 0, n = 0;
    std::vector<double> pr;
    std::vector<mwIndex> jc, ir;
    std::string str;
};

struct StubError : std::runtime_error {
    std::string id;
    StubError(const char *i, const char *msg) : std::runtime_error(msg), id(i) {}
};
static std::string g_id, g_msg;
static void (*g_at_exit)(void) = nullptr;

extern "C" {
bool mxIsSparse(const mxArray *a) { return a->sparse; }
bool mxIsDouble(const mxArray *a) { return !a->is_char && !a->is_logical; }
bool mxIsComplex(const mxArray *) { return false; }
bool mxIsChar(const mxArray *a) { return a->is_char; }
bool mxIsNaN(double v) { return std::isnan(v); }
mwIndex *mxGetJc(const mxArray *a) { return const_cast<mwIndex *>(a->jc.data()); }
mwIndex *mxGetIr(const mxArray *a) { return const_cast<mwIndex *>(a->ir.data()); }
size_t mxGetM(const mxArray *a) { return a->m; }
size_t mxGetN(const mxArray *a) { return a->n; }
size_t mxGetNumberOfElements(const mxArray *a) { return a->is_char ? a->str.size() : a->m * a->n; }
double *mxGetDoubles(const mxArray *a) { return const_cast<double *>(a->pr.data()); }
double mxGetScalar(const mxArray *a) { return a->pr.empty() ? 0.0 : a->pr[0]; }
int mxGetString(const mxArray *a, char *buf, mwSize buflen)
{
    if (!a->is_char || buflen == 0) return 1;
    std::snprintf(buf, buflen, "%s", a->str.c_str());
    return a->str.size() + 1 > buflen;
}
mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity)
{
    mxArray *a = new mxArray_tag;
    a->m = m; a->n = n; a->pr.assign(m * n, 0.0);
    return a;
}
mxArray *mxCreateDoubleScalar(double v) { mxArray *a = mxCreateDoubleMatrix(1, 1, mxREAL); a->pr[0] = v; return a; }
mxArray *mxCreateLogicalScalar(bool v) { mxArray *a = mxCreateDoubleScalar(v ? 1.0 : 0.0); a->is_logical = true; return a; }
void mxDestroyArray(mxArray *a) { delete a; }
mxArray *mxDuplicateArray(const mxArray *a) { return new mxArray_tag(*a); }
void mexMakeArrayPersistent(mxArray *) {}
bool mxIsClass(const mxArray *a, const char *name) { return !std::strcmp(name, "double") && !a->is_char && !a->is_logical; }
static int g_callbacks = 0;
// the toy interpreter knows ONE function: mtimes(sparse matrix, dense vector)
int mexCallMATLAB(int nlhs, mxArray *plhs[], int nrhs, mxArray *prhs[], const char *fn)
{
    if (std::strcmp(fn, "mtimes") || nlhs != 1 || nrhs != 2 || !prhs[0]->sparse || prhs[1]->sparse || prhs[1]->m != prhs[0]->n) return 1;
    const mxArray *A = prhs[0], *v = prhs[1];
    mxArray *u = mxCreateDoubleMatrix(A->m, 1, mxREAL);
    for (size_t j = 0; j < A->n; ++j)
        for (mwIndex k = A->jc[j]; k < A->jc[j + 1]; ++k) u->pr[A->ir[k]] += A->pr[k] * v->pr[j];
    plhs[0] = u;
    ++g_callbacks;
    return 0;
}
int stub_callbacks(void) { return g_callbacks; }
void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...)
{
    char buf[2048];
    va_list ap;
    va_start(ap, fmt);
    std::vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw StubError(id, buf);
}
int mexAtExit(void (*fn)(void)) { g_at_exit = fn; return 0; }

// ---- the test's side ---------------------------------------------------------------------
mxArray *stub_dense(size_t m, size_t n, const double *v)
{
    mxArray *a = mxCreateDoubleMatrix(m, n, mxREAL);
    if (v) std::memcpy(a->pr.data(), v, sizeof(double) * m * n);
    return a;
}
mxArray *stub_sparse(size_t m, size_t n, const long long *jc, const long long *ir, const double *pr)
{
    mxArray *a = new mxArray_tag;
    a->sparse = true; a->m = m; a->n = n;
    a->jc.assign(jc, jc + n + 1);
    const size_t nnz = (size_t)jc[n];
    a->ir.assign(ir, ir + nnz);
    a->pr.assign(pr, pr + nnz);
    return a;
}
mxArray *stub_string(const char *s)
{
    mxArray *a = new mxArray_tag;
    a->is_char = true; a->str = s; a->m = 1; a->n = a->str.size();
    return a;
}
size_t stub_rows(const mxArray *a) { return a->m; }
size_t stub_cols(const mxArray *a) { return a->n; }
const double *stub_data(const mxArray *a) { return a->pr.data(); }
void stub_free(mxArray *a) { delete a; }
const char *stub_last_id(void) { return g_id.c_str(); }
const char *stub_last_msg(void) { return g_msg.c_str(); }
// [plhs...] = mexFunction(prhs...): 0 ok, 1 error() was raised (identifier / message kept)
int stub_call(int nlhs, mxArray **plhs, int nrhs, mxArray **prhs)
{
    try {
        for (int i = 0; i < nlhs; ++i) plhs[i] = nullptr;
        mexFunction(nlhs, plhs, nrhs, const_cast<const mxArray **>(prhs));
        return 0;
    } catch (const StubError &e) {
        g_id = e.id; g_msg = e.what();
        return 1;
    }
}
void stub_exit(void) { if (g_at_exit) g_at_exit(); }
}
