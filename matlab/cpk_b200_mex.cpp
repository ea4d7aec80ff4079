// cpk_b200_mex.cpp -- thin MEX gateway from MATLAB to libcpk_b200.so.
//
// The reference-side binding a maintainer builds with
//     mex -R2018a -I../include cpk_b200_mex.cpp -L../cpkrylov_b200 -lcpk_b200
// No MATLAB exists in this repository's build image: tests/test_mex_gateway.py compiles this file
// against the prototype header matlab/stub/mex.h, links it with a toy MEX runtime
// (matlab/stub/mex_stub.cpp) and drives mexFunction from there.
// Every branch is a mechanical translation mxArray <-> the plain-pointer C ABI of
// include/cpk_b200.h; no arithmetic happens here.
//
//   h  = cpk_b200_mex('ldl2_create', G, B, C22, L, D, P)     opLDL2(A,B,C) ctor, ops/opLDL2.m:60-92
//        cpk_b200_mex('ldl2_set', h, name, value)             public properties, opLDL2.m:45-50
//   y  = cpk_b200_mex('ldl2_apply', h, z)                     M*z, opLDL2.m:161-188
//   y  = cpk_b200_mex('ldl2_matvec', h, b)                    M\b, opLDL2.m:193-195
//   s  = cpk_b200_mex('system_create', A, C, h)               (A, C, M) of method(b1,A,C,M,opts)
//   s  = cpk_b200_mex('system_create_op', n, Aop, C, h)       A is an operator (Spot object, function handle):
//                                                             reg_cpkrylov.m:40; A*v is evaluated by mexCallMATLAB
//                                                             from inside reg_solve while the kernel stays resident
//   [x, niters, solved, status, hist, t_solve] = cpk_b200_mex('reg_solve', s, solver_id, b, optsvec, [n m])
//        hist: hist_len x 3 (columns = cg / lq / qr residHistory for cpsymmlq, column 1 otherwise)
//   h  = cpk_b200_mex('ldl2_create_sqd', G, B, C22, p)        device LDL' with the static permutation p (sequences)
//        cpk_b200_mex('ldl2_refactor', h, G, B, C22)           next system: new values, same patterns
//        cpk_b200_mex('system_update', s, A, C)                 ... and its A, C
//        cpk_b200_mex('destroy', h)
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "mex.h"
#include "cpk_b200.h"

static void fail_if(int rc)
{
    if (rc == CPK_OK) return;
    char msg[1024];
    cpk_last_error(msg, sizeof msg);
    const char *id = "cpk_b200:error";
    if (rc == CPK_ERR_INDEFINITE && std::strstr(msg, "second-order")) id = "CPCGLanczos:IndefiniteError";  // cpcglanczos.m:161
    mexErrMsgIdAndTxt(id, "%s", msg);
}

// MATLAB sparse (mwIndex == 64-bit with -R2018a / -largeArrayDims) -> cpk_csc
static cpk_csc as_csc(const mxArray *a, std::vector<int64_t> &jc, std::vector<int64_t> &ir)
{
    if (!mxIsSparse(a) || !mxIsDouble(a) || mxIsComplex(a))
        mexErrMsgIdAndTxt("cpk_b200:arg", "expected a real sparse double matrix");
    const mwIndex *Jc = mxGetJc(a), *Ir = mxGetIr(a);
    const size_t n = mxGetN(a), nnz = Jc[n];
    jc.assign(Jc, Jc + n + 1);
    ir.assign(Ir, Ir + nnz);
    cpk_csc c;
    c.nrows = (int64_t)mxGetM(a); c.ncols = (int64_t)n;
    c.colptr = jc.data(); c.rowind = ir.data(); c.val = mxGetDoubles(a);
    return c;
}

static cpk_handle as_handle(const mxArray *a) { return (cpk_handle)mxGetScalar(a); }

// Operators of matrix-free systems: persistent copies of the MATLAB objects, by system handle.
static std::map<cpk_handle, mxArray *> g_ops;

// cpk_matvec_fn: u = A*v through the interpreter.  Runs on the MATLAB thread (the library answers
// the kernel's requests from inside cpk_reg_solve, i.e. inside this MEX call).
static int op_matvec(void *ctx, const double *v, double *u, int64_t n)
{
    mxArray *op = static_cast<mxArray *>(ctx);
    mxArray *vin = mxCreateDoubleMatrix((mwSize)n, 1, mxREAL);
    std::memcpy(mxGetDoubles(vin), v, sizeof(double) * (size_t)n);
    mxArray *out = nullptr;
    mxArray *in[2] = {op, vin};
    // function handle: feval(A, v); anything else (Spot operator, matrix): mtimes(A, v)
    const int rc = mexCallMATLAB(1, &out, 2, in, mxIsClass(op, "function_handle") ? "feval" : "mtimes");
    mxDestroyArray(vin);
    if (rc != 0 || !out) return 1;
    const bool ok = mxIsDouble(out) && !mxIsSparse(out) && !mxIsComplex(out) && mxGetNumberOfElements(out) == (size_t)n;
    if (ok) std::memcpy(u, mxGetDoubles(out), sizeof(double) * (size_t)n);
    mxDestroyArray(out);
    return ok ? 0 : 2;
}

static void at_exit(void)
{
    cpk_destroy_all();
    for (auto &kv : g_ops) mxDestroyArray(kv.second);
    g_ops.clear();
}

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    if (nrhs < 1 || !mxIsChar(prhs[0])) mexErrMsgIdAndTxt("cpk_b200:arg", "first argument must be a command string");
    mexAtExit(at_exit);
    char cmdbuf[64];
    mxGetString(prhs[0], cmdbuf, sizeof cmdbuf);
    const std::string cmd(cmdbuf);

    if (cmd == "ldl2_create") {                                   // G, B, C22, L, D, P
        if (nrhs != 7) mexErrMsgIdAndTxt("cpk_b200:arg", "Invalid number of arguments.");   // opLDL2.m:61-63
        std::vector<int64_t> jc[5], ir[5];
        cpk_csc m[5];
        for (int i = 0; i < 5; ++i) m[i] = as_csc(prhs[1 + i], jc[i], ir[i]);
        // P is the sparse permutation MATRIX of [L,D,P] = ldl(K): column k has its 1 in row perm[k]
        const mwIndex *Pjc = mxGetJc(prhs[6]), *Pir = mxGetIr(prhs[6]);
        const size_t N = mxGetN(prhs[6]);
        std::vector<int64_t> perm(N);
        for (size_t k = 0; k < N; ++k) perm[k] = (int64_t)Pir[Pjc[k]];
        cpk_handle h = 0;
        fail_if(cpk_ldl2_create(&h, &m[0], &m[1], &m[2], &m[3], &m[4], perm.data(), 0));
        plhs[0] = mxCreateDoubleScalar((double)h);
    } else if (cmd == "ldl2_create_sqd") {                        // G, B, C22, p  (p: 1-based permutation vector, e.g. from amd/symamd)
        // symmetric quasi-definite K_P: numeric LDL' on the device with the static permutation p
        // (replaces the ldl call of opLDL2.m:82 for the systems of a sequence)
        if (nrhs != 5) mexErrMsgIdAndTxt("cpk_b200:arg", "Invalid number of arguments.");
        std::vector<int64_t> jc[3], ir[3];
        cpk_csc m[3];
        for (int i = 0; i < 3; ++i) m[i] = as_csc(prhs[1 + i], jc[i], ir[i]);
        const size_t N = mxGetNumberOfElements(prhs[4]);
        const double *pv = mxGetDoubles(prhs[4]);
        std::vector<int64_t> perm(N);
        for (size_t k = 0; k < N; ++k) perm[k] = (int64_t)pv[k] - 1;
        cpk_handle h = 0;
        fail_if(cpk_ldl2_create_sqd(&h, &m[0], &m[1], &m[2], perm.data(), 0));
        plhs[0] = mxCreateDoubleScalar((double)h);
    } else if (cmd == "ldl2_refactor") {                          // h, G, B, C22: new values, same patterns
        if (nrhs != 5) mexErrMsgIdAndTxt("cpk_b200:arg", "Invalid number of arguments.");
        std::vector<int64_t> jc[3], ir[3];
        cpk_csc m[3];
        for (int i = 0; i < 3; ++i) m[i] = as_csc(prhs[2 + i], jc[i], ir[i]);
        fail_if(cpk_ldl2_refactor(as_handle(prhs[1]), &m[0], &m[1], &m[2]));
    } else if (cmd == "system_update") {                          // s, A, C: new values, same patterns
        if (nrhs != 4) mexErrMsgIdAndTxt("cpk_b200:arg", "Invalid number of arguments.");
        std::vector<int64_t> jc[2], ir[2];
        cpk_csc a = as_csc(prhs[2], jc[0], ir[0]), c = as_csc(prhs[3], jc[1], ir[1]);
        fail_if(cpk_system_update(as_handle(prhs[1]), &a, &c));
    } else if (cmd == "ldl2_set") {                               // h, name, value
        char name[32];
        mxGetString(prhs[2], name, sizeof name);
        const double v = mxGetScalar(prhs[3]);
        const cpk_handle h = as_handle(prhs[1]);
        if (!std::strcmp(name, "nitref")) fail_if(cpk_ldl2_set_nitref(h, v));
        else if (!std::strcmp(name, "itref_tol")) fail_if(cpk_ldl2_set_itref_tol(h, v));
        else if (!std::strcmp(name, "force_itref")) fail_if(cpk_ldl2_set_force_itref(h, (int)v));
        else if (!std::strcmp(name, "residual_update")) fail_if(cpk_ldl2_set_residual_update(h, (int)v));
        else if (!std::strcmp(name, "ru_stateful")) fail_if(cpk_ldl2_set_ru_stateful(h, (int)v));
        else mexErrMsgIdAndTxt("cpk_b200:arg", "unknown opLDL2 property %s", name);
    } else if (cmd == "ldl2_apply" || cmd == "ldl2_matvec") {     // h, z
        const cpk_handle h = as_handle(prhs[1]);
        int64_t N = 0;
        fail_if(cpk_ldl2_size(h, &N, nullptr, nullptr));
        if ((int64_t)mxGetNumberOfElements(prhs[2]) != N) mexErrMsgIdAndTxt("cpk_b200:arg", "vector length must be %lld", (long long)N);
        plhs[0] = mxCreateDoubleMatrix((mwSize)N, 1, mxREAL);
        cpk_stats st;
        if (cmd == "ldl2_apply") fail_if(cpk_ldl2_apply(h, mxGetDoubles(prhs[2]), mxGetDoubles(plhs[0]), CPK_MEM_HOST, &st));
        else fail_if(cpk_ldl2_matvec(h, mxGetDoubles(prhs[2]), mxGetDoubles(plhs[0]), CPK_MEM_HOST, &st));
    } else if (cmd == "system_create") {                          // A, C, h
        std::vector<int64_t> jc[2], ir[2];
        cpk_csc a = as_csc(prhs[1], jc[0], ir[0]), c = as_csc(prhs[2], jc[1], ir[1]);
        cpk_handle s = 0;
        fail_if(cpk_system_create(&s, &a, &c, as_handle(prhs[3])));
        plhs[0] = mxCreateDoubleScalar((double)s);
    } else if (cmd == "system_create_op") {                       // n, Aop, C, h
        if (nrhs != 5) mexErrMsgIdAndTxt("cpk_b200:arg", "Invalid number of arguments.");
        std::vector<int64_t> jc, ir;
        cpk_csc c = as_csc(prhs[3], jc, ir);
        mxArray *op = mxDuplicateArray(prhs[2]);
        mexMakeArrayPersistent(op);
        cpk_handle s = 0;
        const int rc = cpk_system_create_op(&s, (int64_t)mxGetScalar(prhs[1]), op_matvec, op, &c, as_handle(prhs[4]));
        if (rc != CPK_OK) mxDestroyArray(op);
        fail_if(rc);
        g_ops[s] = op;
        plhs[0] = mxCreateDoubleScalar((double)s);
    } else if (cmd == "reg_solve") {                              // s, solver_id, b, [atol rtol btol itmax restart mem] (NaN = absent)
        const cpk_handle s = as_handle(prhs[1]);
        const int solver = (int)mxGetScalar(prhs[2]);
        const size_t N = mxGetNumberOfElements(prhs[3]);
        const double *ov = mxGetDoubles(prhs[4]);
        const double *nm = mxGetDoubles(prhs[5]);                 // [n m]
        cpk_opts o;
        cpk_opts_default(&o, solver, (int64_t)nm[0], (int64_t)nm[1]);
        if (!mxIsNaN(ov[0])) o.atol = ov[0];
        if (!mxIsNaN(ov[1])) o.rtol = ov[1];
        if (!mxIsNaN(ov[2])) o.btol = ov[2];
        if (!mxIsNaN(ov[3])) o.itmax = (int64_t)ov[3];
        if (!mxIsNaN(ov[4])) o.restart = (int32_t)ov[4];
        if (!mxIsNaN(ov[5])) o.mem = (int32_t)ov[5];
        const int64_t cap = cpk_hist_capacity(solver, &o);
        if (nrhs != 6) mexErrMsgIdAndTxt("cpk_b200:arg", "reg_cpkrylov: not enough inputs");       // reg_cpkrylov.m:122-125
        plhs[0] = mxCreateDoubleMatrix((mwSize)N, 1, mxREAL);
        std::vector<double> hbuf((size_t)3 * (size_t)cap);       // the ABI's 3 rows of `cap` entries
        cpk_stats st;
        const int rc = cpk_reg_solve(s, solver, mxGetDoubles(prhs[3]), &o, mxGetDoubles(plhs[0]), CPK_MEM_HOST,
                                     &st, hbuf.data(), cap);
        fail_if(rc);
        if (nlhs > 1) plhs[1] = mxCreateDoubleScalar((double)st.niters);
        if (nlhs > 2) plhs[2] = mxCreateLogicalScalar(st.solved != 0);
        if (nlhs > 3) plhs[3] = mxCreateDoubleScalar((double)st.status);
        if (nlhs > 4) {
            // hist_len x 3, column-major: column r = row r of the ABI buffer, cut to hist_len entries
            // (shrinking the first dimension of a cap x 3 array in place would leave columns 2 and 3
            // at their old offsets)
            const size_t len = (size_t)(st.hist_len < cap ? st.hist_len : cap);
            plhs[4] = mxCreateDoubleMatrix((mwSize)len, 3, mxREAL);
            double *h = mxGetDoubles(plhs[4]);
            for (size_t r = 0; r < 3; ++r) std::memcpy(h + r * len, hbuf.data() + r * (size_t)cap, sizeof(double) * len);
        }
        if (nlhs > 5) plhs[5] = mxCreateDoubleScalar(st.t_solve_ms * 1e-3);
    } else if (cmd == "destroy") {
        fail_if(cpk_destroy(as_handle(prhs[1])));
        auto it = g_ops.find(as_handle(prhs[1]));
        if (it != g_ops.end()) { mxDestroyArray(it->second); g_ops.erase(it); }
    } else {
        mexErrMsgIdAndTxt("cpk_b200:arg", "unknown command %s", cmd.c_str());
    }
}
