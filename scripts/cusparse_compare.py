"""Same-box comparison points (optional): cuSPARSE SpMV on H and K_P and cuSPARSE
triangular solves with L / L' of the cfg-3 system, reached through torch's sparse CSR tensors
(torch dispatches `A @ x` to cusparseSpMV and `torch.triangular_solve` to cusparseSpSM).
Library code, timed with CUDA events on torch's stream, L2 flushed before every call; printed
next to the library's own stand-alone kernels.  Comparison only: nothing here is on the product path."""
import json, os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); warnings.filterwarnings("ignore")
import numpy as np, scipy.sparse as sp, torch
from cpkrylov_b200 import synth
from cpkrylov_b200.ldl import ldl_superlu

g = int(sys.argv[1]) if len(sys.argv) > 1 else 100
w = synth.kkt_lap3d(g=g)
n, m = w["n"], w["m"]; N = n + m
KP = synth.kp_matrix(w).tocsr()
junk = torch.zeros(48 * 1024 * 1024, dtype=torch.float64, device="cuda")

def csr(A):
    A = sp.csr_matrix(A); A.sort_indices()
    return torch.sparse_csr_tensor(torch.from_numpy(A.indptr.astype(np.int32)), torch.from_numpy(A.indices.astype(np.int32)),
                                   torch.from_numpy(A.data), size=A.shape, device="cuda")

def timed(fn, reps=12):
    ts = []
    for _ in range(reps):
        junk.add_(1.0); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts[2:]))

spmv_bytes = lambda nnz, r, c: 12 * nnz + 4 * (r + 1) + 8 * c + 8 * r
out = {}
H = csr(w["H"]); x = torch.randn(n, dtype=torch.float64, device="cuda")
us = timed(lambda: H @ x); out["cusparse_spmv_H_csr"] = dict(us=us, GBs=spmv_bytes(w["H"].nnz, n, n) / us / 1e3)
K = csr(KP); z = torch.randn(N, dtype=torch.float64, device="cuda")
us = timed(lambda: K @ z); out["cusparse_spmv_KP_csr"] = dict(us=us, GBs=spmv_bytes(KP.nnz, N, N) / us / 1e3)
try:
    L, d, e, p = ldl_superlu(synth.kp_matrix(w))
    Lc = csr(sp.csr_matrix(L)); Lt = csr(sp.csr_matrix(L).T.tocsr())
    b = torch.randn(N, 1, dtype=torch.float64, device="cuda")
    nnz_off = L.nnz - N
    B_trsv = 12 * nnz_off + 4 * (N + 1) + 16 * N
    us = timed(lambda: torch.triangular_solve(b, Lc, upper=False, unitriangular=True), reps=6)
    out["cusparse_sptrsv_L"] = dict(us=us, GBs=B_trsv / us / 1e3)
    us = timed(lambda: torch.triangular_solve(b, Lt, upper=True, unitriangular=True), reps=6)
    out["cusparse_sptrsv_Lt"] = dict(us=us, GBs=B_trsv / us / 1e3)
except Exception as ex:                     # not every torch build routes sparse CSR triangular solves to cuSPARSE
    out["cusparse_sptrsv"] = dict(error=str(ex)[:200])
print(json.dumps(out))
