"""ctypes binding of libcpk_b200.so (C ABI in include/cpk_b200.h).

The library is the product; this module only marshals numpy/scipy data into the
plain-pointer ABI.  There is no CPU fallback: if the shared library has not been
built, importing a compute entry point raises ``CpkLibraryMissing``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CPK_LIB_PATH") or os.path.join(_HERE, "libcpk_b200.so")    # override: A/B runs of two builds
CSRC = os.path.join(_HERE, "csrc")

CPK_OK = 0
CPK_ERR_ARG, CPK_ERR_DIM, CPK_ERR_ALLOC, CPK_ERR_CUDA = -1, -2, -3, -4
CPK_ERR_INDEFINITE, CPK_ERR_BREAKDOWN, CPK_ERR_TIMEOUT, CPK_ERR_UNSUPPORTED = -5, -6, -7, -8
MEM_HOST, MEM_DEVICE = 0, 1
NPHASE = 8
SOLVER_IDS = {"cpcg": 0, "cpcglanczos": 1, "cpminres": 2, "cpsymmlq": 3, "cpgmres": 4, "cpdqgmres": 5}
PHASE_NAMES = ["spmv", "ldl", "resid", "vec", "other"]


class CpkLibraryMissing(RuntimeError):
    pass


class CpkError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


class CscStruct(C.Structure):
    _fields_ = [("nrows", C.c_int64), ("ncols", C.c_int64),
                ("colptr", C.POINTER(C.c_int64)), ("rowind", C.POINTER(C.c_int64)),
                ("val", C.POINTER(C.c_double))]


class OptsStruct(C.Structure):
    _fields_ = [("atol", C.c_double), ("rtol", C.c_double), ("btol", C.c_double),
                ("itmax", C.c_int64), ("restart", C.c_int32), ("mem", C.c_int32),
                ("profile", C.c_int32), ("reserved", C.c_int32)]


class StatsStruct(C.Structure):
    _fields_ = [("niters", C.c_int64), ("solved", C.c_int32), ("status", C.c_int32),
                ("error_iter", C.c_int32), ("error_second", C.c_int32), ("error_value", C.c_double),
                ("hist_len", C.c_int64), ("napply", C.c_int64), ("nldlsolve", C.c_int64),
                ("nresid", C.c_int64), ("shifted", C.c_int32), ("launches", C.c_int32),
                ("t_solve_ms", C.c_double), ("phase_cycles", C.c_double * NPHASE)]


# every symbol include/cpk_b200.h declares: (name, restype, argtypes)
_H = C.c_uint64
_PD = C.POINTER(C.c_double)
_PCSC = C.POINTER(CscStruct)
_PSTATS = C.POINTER(StatsStruct)
_POPTS = C.POINTER(OptsStruct)
# int (*cpk_matvec_fn)(void *ctx, const double *v, double *u, int64_t n)
MATVEC_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, _PD, _PD, C.c_int64)
API = [
    ("cpk_version", C.c_int, []),
    ("cpk_device_count", C.c_int, []),
    ("cpk_last_error", C.c_int, [C.c_char_p, C.c_int64]),
    ("cpk_launch_count", C.c_int64, []),
    ("cpk_ldl2_create", C.c_int, [C.POINTER(_H), _PCSC, _PCSC, _PCSC, _PCSC, _PCSC, C.POINTER(C.c_int64), C.c_int]),
    ("cpk_ldl2_set_nitref", C.c_int, [_H, C.c_double]),
    ("cpk_ldl2_set_itref_tol", C.c_int, [_H, C.c_double]),
    ("cpk_ldl2_set_force_itref", C.c_int, [_H, C.c_int]),
    ("cpk_ldl2_set_residual_update", C.c_int, [_H, C.c_int]),
    ("cpk_ldl2_set_ru_stateful", C.c_int, [_H, C.c_int]),
    ("cpk_ldl2_set_track_rnorm", C.c_int, [_H, C.c_int]),
    ("cpk_ldl2_get_rnorm", C.c_int, [_H, _PD]),
    ("cpk_ldl2_size", C.c_int, [_H, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    ("cpk_ldl2_apply", C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int, _PSTATS]),
    ("cpk_ldl2_matvec", C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int, _PSTATS]),
    ("cpk_ldl2_info", C.c_int, [_H] + [C.POINTER(C.c_int64)] * 4),
    ("cpk_ldl2_create_sqd", C.c_int, [C.POINTER(_H), _PCSC, _PCSC, _PCSC, C.POINTER(C.c_int64), C.c_int]),
    ("cpk_ldl2_refactor", C.c_int, [_H, _PCSC, _PCSC, _PCSC]),
    ("cpk_ldl2_get_factor", C.c_int, [_H, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64), _PD, _PD]),
    ("cpk_system_create", C.c_int, [C.POINTER(_H), _PCSC, _PCSC, _H]),
    ("cpk_system_create_op", C.c_int, [C.POINTER(_H), C.c_int64, MATVEC_FN, C.c_void_p, _PCSC, _H]),
    ("cpk_system_update", C.c_int, [_H, _PCSC, _PCSC]),
    ("cpk_system_matvec", C.c_int, [_H, C.c_int, C.c_void_p, C.c_void_p, C.c_int, _PSTATS]),
    ("cpk_opts_default", None, [_POPTS, C.c_int, C.c_int64, C.c_int64]),
    ("cpk_solve", C.c_int, [_H, C.c_int, C.c_void_p, _POPTS, C.c_void_p, C.c_void_p, C.c_int, _PSTATS, C.c_void_p, C.c_int64]),
    ("cpk_reg_solve", C.c_int, [_H, C.c_int, C.c_void_p, _POPTS, C.c_void_p, C.c_int, _PSTATS, C.c_void_p, C.c_int64]),
    ("cpk_hist_capacity", C.c_int64, [C.c_int, _POPTS]),
    ("cpk_batch_reg_solve", C.c_int, [C.POINTER(_H), C.c_int64, C.c_int, C.POINTER(C.c_void_p), _POPTS,
                                      C.POINTER(C.c_void_p), _PSTATS, C.POINTER(C.c_void_p), C.c_int64]),
    ("cpk_destroy", C.c_int, [_H]),
    ("cpk_destroy_all", C.c_int, []),
]

_lib = None


def build(verbose=False, jobs=8):
    """make -C cpkrylov_b200/csrc  (nvcc, sm_100a only)."""
    cmd = ["make", "-C", CSRC, "-j%d" % jobs]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or out.returncode:
        print(out.stdout)
    if out.returncode:
        raise RuntimeError("building libcpk_b200.so failed")
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CpkLibraryMissing(
                "%s not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the cpkrylov hot path)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, res, args in API:
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error():
    buf = C.create_string_buffer(1024)
    lib().cpk_last_error(buf, 1024)
    return buf.value.decode("utf-8", "replace")


def check(rc):
    if rc != CPK_OK:
        raise CpkError(rc, last_error())


class Csc:
    """Keeps the int64/float64 arrays of a MATLAB-style CSC matrix alive."""

    def __init__(self, A):
        A = sp.csc_matrix(A)
        A.sort_indices()
        self.shape = A.shape
        self.colptr = np.ascontiguousarray(A.indptr, dtype=np.int64)
        self.rowind = np.ascontiguousarray(A.indices, dtype=np.int64)
        self.val = np.ascontiguousarray(A.data, dtype=np.float64)
        if self.rowind.size == 0:       # keep pointers non-null
            self.rowind = np.zeros(1, dtype=np.int64)
            self.val = np.zeros(1, dtype=np.float64)
        self.struct = CscStruct(A.shape[0], A.shape[1],
                                self.colptr.ctypes.data_as(C.POINTER(C.c_int64)),
                                self.rowind.ctypes.data_as(C.POINTER(C.c_int64)),
                                self.val.ctypes.data_as(C.POINTER(C.c_double)))

    def ref(self):
        return C.byref(self.struct)


def stats_to_dict(s: StatsStruct):
    return {
        "niters": int(s.niters), "solved": bool(s.solved), "status": int(s.status),
        "hist_len": int(s.hist_len), "napply": int(s.napply), "nldlsolve": int(s.nldlsolve),
        "nresid": int(s.nresid), "shifted": bool(s.shifted), "launches": int(s.launches),
        "t_solve_ms": float(s.t_solve_ms),
        "phase_cycles": {PHASE_NAMES[i]: float(s.phase_cycles[i]) for i in range(len(PHASE_NAMES))},
    }
