"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every
symbol include/cpk_b200.h declares, fails loudly without a GPU (no fallback), and
the product never touches oracle/."""
import ctypes as ct
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "cpk_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cpk_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(cpk_lib):
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(cpk_lib, n), "libcpk_b200.so does not export %s" % n
    from cpkrylov_b200 import _lib
    assert sorted(n for n, _, _ in _lib.API) == names       # the ctypes table covers the whole header


def test_struct_layouts_match_header():
    from cpkrylov_b200 import _lib
    assert ct.sizeof(_lib.CscStruct) == 40
    assert ct.sizeof(_lib.OptsStruct) == 48
    assert ct.sizeof(_lib.StatsStruct) == 8 + 4 * 4 + 8 + 4 * 8 + 2 * 4 + 8 + 8 * _lib.NPHASE


def test_opts_defaults_follow_reference(cpk_lib):
    from cpkrylov_b200 import _lib
    o = _lib.OptsStruct()
    for name, sid in _lib.SOLVER_IDS.items():
        cpk_lib.cpk_opts_default(ct.byref(o), sid, 3000, 2500)
        assert (o.atol, o.rtol, o.btol) == (1e-6, 1e-6, 0.0)               # e.g. cpcg.m:97-98
        assert o.itmax == (5500 if name in ("cpgmres", "cpdqgmres") else 3000)   # cpgmres.m:105 / cpcg.m:99
        assert (o.restart, o.mem) == (50, 50)                              # cpgmres.m:103, cpdqgmres.m:103
    o.itmax = 105; o.restart = 50
    assert cpk_lib.cpk_hist_capacity(_lib.SOLVER_IDS["cpgmres"], ct.byref(o)) >= 151   # cpgmres.m:148 overshoot
    assert cpk_lib.cpk_hist_capacity(_lib.SOLVER_IDS["cpminres"], ct.byref(o)) >= 106


def test_dimension_errors_use_reference_texts(cpk_lib):
    import cpkrylov_b200 as cp
    with pytest.raises(ValueError, match="Incompatible dimensions"):       # opLDL2.m:73-75
        cp.opLDL2(sp.identity(3), sp.csr_matrix((2, 4)), -sp.identity(2))
    with pytest.raises(ValueError, match="must be square"):                # opLDL2.m:68-70
        cp.opLDL2(sp.csr_matrix((3, 4)), sp.csr_matrix((2, 3)), -sp.identity(2))
    with pytest.raises(ValueError, match="not enough inputs"):             # reg_cpkrylov.m:122-125
        cp.reg_cpkrylov(cp.cpminres, np.zeros(3), None, None, None, None)


def test_no_cpu_fallback_without_gpu(cpk_lib):
    import cpkrylov_b200 as cp
    if cpk_lib.cpk_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(cp.CpkError, match="no CUDA device"):
        cp.opLDL2(sp.identity(3), sp.csr_matrix((1, 3)), -sp.identity(1))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cpkrylov_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
                assert "liborc" not in txt and "cpk_oracle" not in txt, f


def test_method_dispatch_names():
    import cpkrylov_b200 as cp
    from cpkrylov_b200.solvers import _method_name
    assert _method_name(cp.cpminres) == "cpminres" and _method_name("@cpgmres") == "cpgmres"
    with pytest.raises(ValueError):
        _method_name("pcg")
