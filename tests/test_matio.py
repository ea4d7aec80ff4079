"""Loader for the reference's .mat example format (cpkrylov_b200/matio.py) and the
sequence driver.  CPU part: block extraction against the committed goldens, round trips
through scipy.io, and -- only in the build container, where /root/reference exists --
the shipped .mat files themselves.  GPU part: a short IPM-like sequence in one launch."""
import os

import numpy as np
import pytest
import scipy.io as sio
import scipy.sparse as sp

from cpkrylov_b200 import matio
from helpers import EX_OPTS, load_factors, load_system, relerr

REF = "/root/reference/examples"


def _same_blocks(s, g):
    for a, b in (("H", "Q"), ("B", "A"), ("C", "C"), ("G", "G")):
        assert abs(sp.csc_matrix(s[a]) - sp.csc_matrix(g[b])).max() == 0.0
    assert np.array_equal(s["rhs"], g["rhs"]) and (s["n"], s["m"], s["N"]) == (g["n"], g["m"], g["N"])


@pytest.mark.parametrize("name", ["cvxqp1_m", "cvxqp2_s"])
def test_system_from_K_matches_the_example_scripts(name):
    g = load_system(name)
    _same_blocks(matio.system_from_K(g["K"], g["rhs"], g["n"]), g)


@pytest.mark.parametrize("layout,nH,nZ", [("2x2", 40, 0), ("3x3", 25, 15)])
def test_mat_round_trip_both_layouts(tmp_path, layout, nH, nZ):
    rng = np.random.default_rng(1)
    n, m = (nH + nZ if layout == "3x3" else nH), 12
    K = sp.random(n + m, n + m, density=0.1, random_state=rng, format="csc") + sp.identity(n + m, format="csc")
    rhs = rng.standard_normal(n + m)
    path = os.path.join(tmp_path, "sys_%s_iter3.mat" % layout)
    sio.savemat(path, dict(K=K, rhs=rhs.reshape(-1, 1), nH=nH, nJ=m, nZ=nZ, n=n + m))
    s = matio.load_mat_system(path)                 # layout from the file name, like the shipped files
    assert s["params"]["layout"] == layout and (s["n"], s["m"]) == (n, m)
    assert abs(s["H"] - K[:n, :n]).max() == 0.0 and abs(s["B"] - K[n:, :n]).max() == 0.0
    assert abs(s["C"] + K[n:, n:]).max() == 0.0
    assert np.array_equal(s["G"].diagonal(), K[:n, :n].diagonal()) and s["G"].nnz <= n
    assert np.array_equal(s["rhs"], rhs)
    with pytest.raises(ValueError):
        matio.load_mat_system(path, layout="bogus")


def test_mat_loader_rejects_inconsistent_sizes(tmp_path):
    K = sp.identity(10, format="csc")
    path = os.path.join(tmp_path, "bad_2x2.mat")
    sio.savemat(path, dict(K=K, rhs=np.ones((10, 1)), nH=6, nJ=3, nZ=0))
    with pytest.raises(ValueError):
        matio.load_mat_system(path)
    sio.savemat(path, dict(K=K, rhs=np.ones((9, 1)), nH=6, nJ=4, nZ=0))
    with pytest.raises(ValueError):
        matio.load_mat_system(path)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("name,file", [("cvxqp1_m", "cvxqp1_m_2x2_symm_iter10.mat"),
                                       ("cvxqp2_s", "cvxqp2_s_3x3_nonsymm_perm_iter10.mat")])
def test_shipped_mat_files_load_like_the_goldens(name, file):
    _same_blocks(matio.load_mat_system(os.path.join(REF, file)), load_system(name))


@pytest.mark.gpu
def test_sequence_of_ipm_systems_one_launch():
    from cpkrylov_b200.synth import ipm_batch_system
    from cpkrylov_b200.ldl import ldl_superlu
    from cpkrylov_b200.synth import kp_matrix
    from oracle import cpk_oracle as orc
    from cpkrylov_b200 import _lib
    base = load_system("cvxqp1_m")
    systems = []
    for j in range(3):
        w = ipm_batch_system(base, j)
        s = matio.system_from_K(sp.bmat([[w["H"], w["B"].T], [w["B"], -w["C"]]], format="csc"), w["rhs"], w["n"])
        systems.append(s)
    facs = [ldl_superlu(kp_matrix(s)) for s in systems]
    l0 = _lib.lib().cpk_launch_count()
    xs, stats = matio.solve_sequence("cpminres", systems, dict(EX_OPTS), factors=facs)
    assert _lib.lib().cpk_launch_count() - l0 == 1
    for s, fac, x, st in zip(systems, facs, xs, stats):
        xo, so, fo = orc.reg_cpkrylov("cpminres", s["rhs"], s["H"], s["B"], s["C"], s["G"], dict(EX_OPTS), factor=lambda K, f=fac: f)
        assert st["solved"] == fo["solved"] and abs(st["niters"] - so["niters"]) <= 2
        assert relerr(x, xo) <= 1e-8


@pytest.mark.gpu
def test_ipm_sequence_in_place():
    """Generator-driven sequence: operator built once on the device, refreshed in place."""
    from cpkrylov_b200.synth import ipm_batch_system, kp_matrix
    from cpkrylov_b200.ldl import ldl_superlu
    from oracle import cpk_oracle as orc
    base = load_system("cvxqp1_m")
    seq = [ipm_batch_system(base, j) for j in range(3)]
    out = list(matio.solve_ipm_sequence("cpminres", (w for w in seq), dict(EX_OPTS)))
    assert len(out) == 3
    for j, (w, (x, st, fl)) in enumerate(zip(seq, out)):
        fac = ldl_superlu(kp_matrix(w))
        xo, so, fo = orc.reg_cpkrylov("cpminres", w["rhs"], w["H"], w["B"], w["C"], w["G"], dict(EX_OPTS), factor=lambda K, f=fac: f)
        assert fl["solved"] == fo["solved"] and abs(st["niters"] - so["niters"]) <= 2
        assert relerr(x, xo) <= 1e-7
        if j:
            assert st["t_setup"] < 0.05
    bad = dict(seq[1]); bad["B"] = sp.csc_matrix(seq[1]["B"].shape)
    with pytest.raises(ValueError):
        list(matio.solve_ipm_sequence("cpminres", [seq[0], bad], dict(EX_OPTS)))
