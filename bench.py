#!/usr/bin/env python
"""bench.py -- cpkrylov hot path on B200: Krylov iterations/s + achieved HBM GB/s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload kkt_lap3d|kkt_convdiff|ipm_batch] [--g G] [--solver NAME]

A *step* is one complete solve (reg_cpkrylov body: rhs shift, device-resident
Krylov loop, un-shift) of the workload's saddle-point system; the metric is
Krylov iterations per second over all ranks.  Default workload = BASELINE.json
configs[2], the ~10M-nonzero synthetic KKT system the north-star target is quoted
on: H = 3-D Laplacian n = 100^3, random sparse B (k=2) m = 250k, C = 1e-6 I,
G = diag(H), cpcg, reference-default options.

N > 1: launched by torch.distributed.run, one rank per GPU; every rank solves its
own independent system (same matrix, own right-hand side) -- weak scaling, no
collective inside the iteration; one NCCL all-reduce per step gathers the
convergence flags / iteration counts.

`--impl reference` times the CPU restatement of the reference (oracle/, the
reference itself is MATLAB and cannot run here) on the host cores of rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as ct
import json
import os
import subprocess
import sys
import threading
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

# V_s: N-vector passes of Krylov algebra per iteration under minimal fusion
V_S = {"cpcg": 11.4, "cpcglanczos": 11.0, "cpminres": 13.0, "cpsymmlq": 12.0}


def algorithmic_bytes(s, info, solver, gpu_stats, mem=0, restart=0):
    """Algorithmic HBM bytes of ONE solve (fp64 values, int32 indices, vectors
    counted once; formulas of SURVEY.md 8d / DESIGN.md)."""
    n, m = s["n"], s["m"]
    N = n + m
    spmv = lambda nnz, r, c: 12 * nnz + 4 * (r + 1) + 8 * c + 8 * r
    B_H = spmv(s["H"].nnz, n, n)
    B_C = spmv(s["C"].nnz, m, m)
    nnz_KP = s["G"].nnz + 2 * s["B"].nnz + s["C"].nnz
    B_resid = spmv(nnz_KP, N, N) + 8 * N
    B_ldl = 24 * info["nnz_L_off"] + 48 * N
    it = gpu_stats["niters"]
    # N-vector passes of the Krylov algebra over the whole solve
    if solver in V_S:
        vs_total = V_S[solver] * it
    elif solver == "cpdqgmres":             # 2p + p' + 10 with the window actually filled at iteration k
        vs_total = sum(2 * min(k, mem) + min(k - 1, mem) + 10 for k in range(1, it + 1))
    else:                                   # cpgmres: 2k+6 at inner iteration k, k+2 at the end of a cycle
        vs_total, k = 0, 0
        for _ in range(it):
            k += 1
            vs_total += 2 * k + 6
            if k == restart:
                vs_total += k + 2; k = 0
        if k:
            vs_total += k + 2
    vs = vs_total / max(it, 1)
    per_iter = B_H + B_C
    total = it * per_iter + 8 * N * vs_total + gpu_stats["nldlsolve"] * B_ldl + gpu_stats["nresid"] * B_resid
    parts = dict(B_spmv_H=B_H, B_spmv_C=B_C, B_ldl=B_ldl, B_resid=B_resid, B_vec=8 * N * vs)
    return total, parts


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the GPU works."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.samples = []
        self.index = index
        self.proc = None
        self.t = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.samples.append((time.perf_counter(), line.strip()))
        self.t = threading.Thread(target=pump, daemon=True)
        self.t.start()

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [l for (t, l) in self.samples if t0 <= t <= t1] or [l for (_, l) in self.samples]
        sm, smmax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in rows:
            f = [x.strip() for x in l.split(",")]
            try:
                sm.append(float(f[1])); smmax.append(float(f[2]))
                for nm, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(smmax) if smmax else None,
                "reasons": sorted(reasons), "samples": len(rows), "in_timed_region": bool([1 for (t, _) in self.samples if t0 <= t <= t1])}


def build_workload(args, rank):
    from cpkrylov_b200 import synth
    if args.workload == "kkt_lap3d":
        s = synth.kkt_lap3d(g=args.g or 100, k=args.k, window=args.window, seed_x=3 + rank)
        solver = args.solver or "cpcg"
        opts = dict(atol=1e-6, rtol=1e-6)                    # reference defaults: nitref=3, itref_tol=1e-8
    elif args.workload == "kkt_convdiff":
        s = synth.kkt_convdiff(g=args.g or 126, k=args.k, window=args.window, seed_x=3 + rank)
        solver = args.solver or "cpdqgmres"
        opts = dict(atol=1e-6, rtol=1e-6, itmax=500, mem=args.mem, nitref=1, force_itref=True)
    else:
        raise SystemExit("unknown workload " + args.workload)
    if args.nitref is not None:
        opts["nitref"] = args.nitref
    return s, solver, opts


def blas_threads():
    """Threads the CPU restatement can actually use: its sparse kernels (oracle/kernels.c)
    are single-threaded, NumPy's BLAS-1 (dot, norm, axpy) runs on the BLAS thread pool."""
    try:
        from threadpoolctl import threadpool_info
        return max([d.get("num_threads", 1) for d in threadpool_info()] or [1])
    except Exception:
        return 1


def peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def flush_l2(torch, buf):
    """writes a buffer larger than the 126 MB L2 so that the next launch starts cold"""
    buf.add_(1.0)
    torch.cuda.synchronize()


def roofline_parts(torch, L, S, M, s, info, peak, reps=12):
    """Per-kernel evidence (north star: SpMV and preconditioner apply EACH against the HBM
    roofline): every kernel is launched stand-alone through the C ABI on device-resident
    vectors, the L2 flushed before each launch, CUDA-event time of the launch (the library's
    own events on its stream), algorithmic bytes of SURVEY 8d."""
    from cpkrylov_b200 import _lib
    n, m = s["n"], s["m"]
    N = n + m
    spmv = lambda nnz, r, c: 12 * nnz + 4 * (r + 1) + 8 * c + 8 * r
    nnz_KP = s["G"].nnz + 2 * s["B"].nnz + s["C"].nnz
    B_H, B_KP = spmv(s["H"].nnz, n, n), spmv(nnz_KP, N, N)
    B_ldl = 24 * info["nnz_L_off"] + 48 * N
    x = torch.randn(N, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    junk = torch.zeros(48 * 1024 * 1024, dtype=torch.float64, device="cuda")     # 384 MB
    saved = (M.nitref, M.force_itref)

    def timed(call):
        # `us`: L2 flushed by WRITING a buffer larger than the L2 (the launch then also pays for the
        # write-back of the flush's dirty lines); `us_readflush`: the L2 filled with clean lines of a
        # buffer that was only read.  Both are CUDA events around the cooperative launch.
        res = []
        for writes in (True, False):
            ts = []
            for _ in range(reps):
                if writes:
                    flush_l2(torch, junk)
                else:
                    junk.sum(); torch.cuda.synchronize()
                st = _lib.StatsStruct()
                _lib.check(call(ct.byref(st)))
                ts.append(st.t_solve_ms * 1e3)
            res.append(float(np.median(ts[2:])))
        return res

    out = {}
    def put(name, us2, nbytes, what):
        us, us_rf = us2
        gbs = nbytes / us / 1e3
        out[name] = {"us": us, "bytes": nbytes, "GBs": gbs, "frac": gbs / peak, "us_readflush": us_rf, "frac_readflush": nbytes / us_rf / 1e3 / peak, "what": what}
    put("spmv_H", timed(lambda st: L.cpk_system_matvec(S.handle, 0, x.data_ptr(), y.data_ptr(), 1, st)), B_H,
        "H*v (cpk_system_matvec), 12 nnz + 4(r+1) + 8c + 8r")
    put("spmv_KP", timed(lambda st: L.cpk_ldl2_matvec(M.handle, x.data_ptr(), y.data_ptr(), 1, st)), B_KP,
        "K_P*v (cpk_ldl2_matvec: the product inside the refinement residual)")
    M.nitref = 0
    put("ldl_solve", timed(lambda st: L.cpk_ldl2_apply(M.handle, x.data_ptr(), y.data_ptr(), 1, st)), B_ldl,
        "M*z with nitref = 0: P L^-T D^-1 L^-1 P' alone, 24 nnz_off(L) + 48 N")
    M.nitref = 3; M.force_itref = False
    put("apply_default", timed(lambda st: L.cpk_ldl2_apply(M.handle, x.data_ptr(), y.data_ptr(), 1, st)), B_ldl + B_KP + 8 * N,
        "M*z with the reference defaults (nitref = 3, no step taken): solve + residual pass")
    M.nitref = 1; M.force_itref = True
    put("apply_forced1", timed(lambda st: L.cpk_ldl2_apply(M.handle, x.data_ptr(), y.data_ptr(), 1, st)), 2 * B_ldl + B_KP + 8 * N + 24 * N,
        "M*z with nitref = 1, force_itref (example options): 2 solves + 1 residual pass")
    M.nitref, M.force_itref = saved
    del junk
    return out


def parity_section(s, solver, opts, fac, x_gpu, stats_gpu):
    """Untimed: the CPU restatement run to completion on the same system with the same factors;
    backs the 'GPU = CPU restatement' columns of BASELINE.md with the driver's own run."""
    from oracle import cpk_oracle as orc
    orc.build_c_kernels()
    t = time.perf_counter()
    xo, so, fo = orc.reg_cpkrylov(solver, s["rhs"], s["H"], s["B"], s["C"], s["G"], dict(opts, print=False), factor=lambda K: fac)
    ho = np.asarray(so.get("residHistory", so.get("cgresidHistory")), dtype=float)
    return {"niters_oracle": int(so["niters"]), "niters_gpu": int(stats_gpu["niters"]), "solved_oracle": bool(fo["solved"]),
            "solved_gpu": bool(stats_gpu["solved"]),
            "relerr_vs_oracle": float(np.linalg.norm(x_gpu - xo) / np.linalg.norm(xo)), "oracle_s": time.perf_counter() - t,
            "oracle_final_resid": float(ho[-1]), "note": "oracle = oracle/cpk_oracle.py, same (L, D, p); see tests/test_gpu_parity_full.py for the arbitrated bar"}


def run_reference(args, rank, world):
    """CPU restatement of the reference on the host cores (rank 0 only)."""
    if rank != 0:
        return
    from oracle import cpk_oracle as orc
    from cpkrylov_b200 import synth
    from cpkrylov_b200.ldl import ldl_factor
    orc.build_c_kernels()
    s, solver, opts = build_workload(args, 0)
    fac = ldl_factor(synth.kp_matrix(s))
    sample_it = args.ref_iters
    o = dict(opts, print=False, itmax=sample_it)
    times, iters = [], 0
    for step in range(args.warmup + args.steps):
        t = time.perf_counter()
        x, st, fl = orc.reg_cpkrylov(solver, s["rhs"], s["H"], s["B"], s["C"], s["G"], o, factor=lambda K: fac)
        dt = st["stime"]
        if step >= args.warmup:
            times.append(dt); iters += st["niters"]
    total = sum(times)
    val = iters / total
    nthr = blas_threads()
    sample = ("first %d iterations of the %s solve per step (oracle/cpk_oracle.py + oracle/kernels.c; sparse "
              "kernels single-threaded, NumPy BLAS-1 on %d threads; host has %d cores)" % (sample_it, args.workload, nthr, os.cpu_count()))
    line = {
        "impl": "reference", "metric": "krylov_iterations_per_second", "value": val, "unit": "iterations/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, len(times)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(s["params"], solver=solver, opts={k: v for k, v in opts.items()}),
        "cpu_baseline": {"value": val, "unit": "iterations/s", "cores": nthr, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def batch_measure(args, rank, local_rank, world, steps, warmup, total=None, per_gpu=0, g=0):
    """BASELINE cfg 5: a batch of independent IPM-like KKT systems (pattern of the
    reference's cvxqp1 example), block-partitioned over the ranks; every rank solves
    its share in ONE launch (one CTA per system).  Timed end to end through the host-pointer
    batch ABI (H2D right-hand sides, D2H solutions inside the timed region), barrier on both
    sides, max over ranks.  The process group (world > 1) must exist.  Returns the line on rank 0."""
    import torch
    import torch.distributed as dist
    from cpkrylov_b200 import synth
    from cpkrylov_b200.batch import BatchSolver, partition
    from cpkrylov_b200.ldl import ldl_superlu
    total = per_gpu * world if per_gpu > 0 else (total or 256)
    lo, hi = partition(total, world, rank)
    opts = dict(atol=1e-6, rtol=1e-6, itmax=500, residual_update=True, nitref=1, force_itref=True)
    t0 = time.perf_counter()
    if g > 0:
        # larger variant: cfg-3 pattern on a g^3 grid; systems beyond the one-CTA
        # limit share cooperative launches as sub-teams of the grid
        systems = [synth.ipm_batch_lap3d(g, j) for j in range(lo, hi)]
        base = dict(n=systems[0]["n"], m=systems[0]["m"], N=systems[0]["n"] + systems[0]["m"])
    else:
        base = synth.load_cvxqp1()
        systems = [synth.ipm_batch_system(base, j) for j in range(lo, hi)]
    facs = [ldl_superlu(synth.kp_matrix(w)) for w in systems]
    bs = BatchSolver(systems, facs, opts, device=local_rank)
    t_setup = time.perf_counter() - t0
    rhs = [w["rhs"] for w in systems]
    for _ in range(max(warmup, 3)):
        bs.solve("cpminres", rhs, opts)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    iters = 0
    dev_ms = 0.0
    launches = 0
    solved = 1
    for _ in range(steps):
        xs, st = bs.solve("cpminres", rhs, opts)
        iters += sum(d["niters"] for d in st)
        solved = min(solved, min(int(d["solved"]) for d in st))
        dev_ms += bs.last_ms
        launches += bs.last_launches
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - t1
    tt = torch.tensor([wall, dev_ms, float(-solved)], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([float(iters)], dtype=torch.float64, device="cuda")
    if world > 1:
        # result gathering / global convergence reduction (the only collectives of the path)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    bs.close()
    if rank != 0:
        return None
    wall_max, dev_max, neg_solved = tt.tolist()
    return {
        "metric": "krylov_iterations_per_second", "value": cnt.item() / wall_max, "unit": "iterations/s",
        "n_gpus": world, "steps": steps, "warmup": max(warmup, 3), "ms_per_step": 1e3 * wall_max / steps,
        "higher_is_better": True, "scaling": "weak" if per_gpu > 0 else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "ipm_batch", "g": g, "systems": total, "n": base["n"], "m": base["m"], "solver": "cpminres",
                   "opts": opts, "systems_per_rank": hi - lo, "device_ms_per_step": dev_max / steps, "all_solved": bool(neg_solved < 0),
                   "setup_s": t_setup, "note": "end to end through the host-pointer batch ABI (H2D rhs, D2H solutions inside the timed region)"},
        "e2e": {"value": cnt.item() / wall_max, "unit": "iterations/s",
                "h2d_bytes_per_step": 8 * base["N"] * (hi - lo), "d2h_bytes_per_step": 8 * base["N"] * (hi - lo)},
        "gpu_launches": launches}


def run_batch(args, rank, local_rank, world):
    """`--workload ipm_batch`: BASELINE cfg 5 as a line of its own (also reported as the
    `cfg5_ipm_batch` section of the default line)."""
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    line = batch_measure(args, rank, local_rank, world, args.steps, args.warmup, total=args.batch, per_gpu=args.batch_per_gpu, g=args.g)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def stress_section(args, torch, peak):
    """SURVEY 8d stress variant of cfg 3 at full size (k = 6 entries per row of B inside a 64-wide
    window: nnz(K) ~ 10.1 M, a FILLED factor with hundreds of dependency levels -- the deep
    sparse triangular solves the k = 2 headline system does not exercise).  Rank 0, untimed set-up
    reported apart; the solve is timed like the headline (CUDA events of the one launch)."""
    from cpkrylov_b200 import _lib, synth
    from cpkrylov_b200.operators import opLDL2, KktSystem
    from cpkrylov_b200.solvers import _fill_opts, apply_opts_to_M
    L = _lib.lib()
    t0 = time.perf_counter()
    s = synth.kkt_lap3d(g=args.stress_g, k=6, window=64)
    solver, opts = "cpcg", dict(atol=1e-6, rtol=1e-6)
    n, m = s["n"], s["m"]
    N = n + m
    M = opLDL2(s["G"], s["B"], -s["C"])
    S = KktSystem(s["H"], s["C"], M)
    apply_opts_to_M(M, opts)
    info = M.info()
    t_setup = time.perf_counter() - t0
    sid, o = _fill_opts(solver, opts, n, m)
    cap = int(L.cpk_hist_capacity(sid, ct.byref(o)))
    hist = np.zeros((3, cap))
    b_dev = torch.from_numpy(s["rhs"]).to("cuda")
    x_dev = torch.empty(N, dtype=torch.float64, device="cuda")
    ms = []
    for _ in range(2 + args.stress_steps):
        st = _lib.StatsStruct()
        _lib.check(L.cpk_reg_solve(S.handle, sid, b_dev.data_ptr(), ct.byref(o), x_dev.data_ptr(), _lib.MEM_DEVICE,
                                   ct.byref(st), hist.ctypes.data, cap))
        ms.append(st.t_solve_ms)
    d = _lib.stats_to_dict(st)
    kernel_ms = float(np.mean(ms[2:]))
    bytes_solve, parts = algorithmic_bytes(s, info, solver, d)
    x_gpu = x_dev.cpu().numpy()
    out = {"config": dict(s["params"], solver=solver, opts=opts, iters_per_solve=d["niters"], solved=d["solved"],
                          relerr_vs_xstar=float(np.linalg.norm(x_gpu - s["xstar"]) / np.linalg.norm(s["xstar"])),
                          setup_s=t_setup, t_factor_s=M.t_factor, t_upload_s=M.t_upload, ldl=info),
           "steps": args.stress_steps, "ms_per_step": kernel_ms, "value": 1e3 * d["niters"] / kernel_ms, "unit": "iterations/s",
           "roofline": {"bound": "hbm", "achieved": bytes_solve / kernel_ms / 1e6, "peak": peak, "unit": "GB/s",
                        "frac": bytes_solve / kernel_ms / 1e6 / peak, "algorithmic_bytes_per_launch": bytes_solve, "bytes_parts": parts,
                        "ldl_solves": d["nldlsolve"]}}
    if not args.no_parity:
        out["parity"] = parity_section(s, solver, opts, M.factors, x_gpu, d)
    S.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="kkt_lap3d")
    ap.add_argument("--solver", default=None)
    ap.add_argument("--g", type=int, default=0)
    ap.add_argument("--k", type=int, default=2)
    ap.add_argument("--window", type=int, default=0)
    ap.add_argument("--mem", type=int, default=20)
    ap.add_argument("--nitref", type=int, default=None)
    ap.add_argument("--ref-iters", type=int, default=8)
    ap.add_argument("--cpu-baseline-iters", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true", help="collect per-phase cycle shares (slower)")
    ap.add_argument("--batch", type=int, default=256, help="systems in the ipm_batch workload")
    ap.add_argument("--batch-per-gpu", type=int, default=0, help="ipm_batch: systems PER GPU (weak scaling) instead of --batch in total")
    ap.add_argument("--no-parts", action="store_true", help="skip the stand-alone per-kernel roofline section")
    ap.add_argument("--no-parity", action="store_true", help="skip the untimed oracle run")
    ap.add_argument("--no-extras", action="store_true", help="skip the cfg 5 (ipm_batch) and stress (k=6 windowed) sections of the default line")
    ap.add_argument("--stress-g", type=int, default=100)
    ap.add_argument("--stress-steps", type=int, default=3)
    ap.add_argument("--extras-batch-steps", type=int, default=10)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.workload == "ipm_batch":
        run_batch(args, rank, local_rank, world)
        return

    import torch
    import torch.distributed as dist
    from cpkrylov_b200 import _lib, synth
    from cpkrylov_b200.operators import opLDL2, KktSystem
    from cpkrylov_b200.solvers import _fill_opts, apply_opts_to_M

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---------------- setup (untimed: generation, factorization, upload) -------------
    t_setup0 = time.perf_counter()
    s, solver, opts = build_workload(args, rank)
    n, m = s["n"], s["m"]
    N = n + m
    M = opLDL2(s["G"], s["B"], -s["C"], device=local_rank)
    S = KktSystem(s["H"], s["C"], M)
    apply_opts_to_M(M, opts)
    info = M.info()
    t_setup = time.perf_counter() - t_setup0
    sid, o = _fill_opts(solver, dict(opts, profile=args.profile), n, m)
    cap = int(_lib.lib().cpk_hist_capacity(sid, ct.byref(o)))
    hist = np.zeros((3, cap))
    L = _lib.lib()

    b_dev = torch.from_numpy(s["rhs"]).to("cuda")
    x_dev = torch.empty(N, dtype=torch.float64, device="cuda")
    b_pin = torch.from_numpy(s["rhs"]).pin_memory()
    x_pin = torch.empty(N, dtype=torch.float64).pin_memory()
    red = torch.zeros(2, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()

    def solve_dev():
        st = _lib.StatsStruct()
        _lib.check(L.cpk_reg_solve(S.handle, sid, b_dev.data_ptr(), ct.byref(o), x_dev.data_ptr(), _lib.MEM_DEVICE,
                                   ct.byref(st), hist.ctypes.data, cap))
        return st

    def solve_host():
        st = _lib.StatsStruct()
        _lib.check(L.cpk_reg_solve(S.handle, sid, b_pin.data_ptr(), ct.byref(o), x_pin.data_ptr(), _lib.MEM_HOST,
                                   ct.byref(st), hist.ctypes.data, cap))
        return st

    red_pin = torch.zeros(2, dtype=torch.int64).pin_memory()

    def step_collective(st):
        # global convergence reduction of the step (the only collective of the path).
        # It is completed before the next solve is launched: the persistent solver
        # kernel owns every SM (1 CTA/SM, all registers), so an NCCL kernel still
        # queued behind it would wait for the whole next solve.
        if world > 1:
            red_pin[0] = -int(st.solved); red_pin[1] = int(st.niters)      # MAX of -solved = -MIN(solved)
            red.copy_(red_pin, non_blocking=True)
            dist.all_reduce(red, op=dist.ReduceOp.MAX)
            torch.cuda.current_stream().synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        st = solve_dev(); step_collective(st)

    # ---------------- timed region 1: device-resident inputs -------------------------
    barrier()
    launches0 = L.cpk_launch_count()
    t0 = time.perf_counter()
    dev_ms, iters, stats_last = 0.0, 0, None
    phase = np.zeros(_lib.NPHASE)
    for _ in range(args.steps):
        st = solve_dev()
        step_collective(st)
        dev_ms += st.t_solve_ms
        iters += st.niters
        stats_last = _lib.stats_to_dict(st)
        phase += np.array([st.phase_cycles[i] for i in range(_lib.NPHASE)])
    barrier()
    t1 = time.perf_counter()
    launches = L.cpk_launch_count() - launches0
    wall_ms = 1e3 * (t1 - t0)

    # ---------------- timed region 2: end to end through the host-pointer ABI --------
    for _ in range(2):
        solve_host()
    barrier()
    t2 = time.perf_counter()
    e2e_iters = 0
    for _ in range(args.steps):
        st = solve_host(); step_collective(st)
        e2e_iters += st.niters
    barrier()
    t3 = time.perf_counter()
    clocks = sampler.stop(t0, t3) if rank == 0 else None

    # ---------------- max over ranks --------------------------------------------------
    tt = torch.tensor([wall_ms, dev_ms, 1e3 * (t3 - t2)], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([iters, e2e_iters], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    per_rank = [torch.zeros(3, dtype=torch.float64, device="cuda") for _ in range(world)]
    mine = torch.tensor([wall_ms, dev_ms, 1e3 * (t3 - t2)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_gather(per_rank, mine)
    else:
        per_rank = [mine]
    wall_ms_max, dev_ms_max, e2e_ms_max = [float(v) for v in tt.tolist()]
    tot_iters, tot_e2e_iters = [float(v) for v in cnt.tolist()]

    x_gpu = x_dev.cpu().numpy()
    err = float(np.linalg.norm(x_gpu - s["xstar"]) / np.linalg.norm(s["xstar"]))

    if rank == 0:
        peak, peak_src = peak_gbs()
        bytes_solve, parts = algorithmic_bytes(s, info, solver, stats_last, mem=args.mem, restart=50)
        kernel_ms = dev_ms / args.steps                     # CUDA events on the library's stream, one launch per step
        achieved = bytes_solve / (kernel_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic_%s.json" % args.workload)
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": "krylov_iterations_per_second", "value": tot_iters / (wall_ms_max * 1e-3), "unit": "iterations/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(s["params"], solver=solver, opts=opts, iters_per_solve=stats_last["niters"],
                           solved=stats_last["solved"], relerr_vs_xstar=err, parallelism="1 independent system per GPU",
                           l2_policy="working set (%.0f MB of matrices + vectors) exceeds the 126 MB L2; no explicit flush"
                                     % ((parts["B_spmv_H"] + parts["B_ldl"] + parts["B_resid"]) / 1e6 + 8 * N * 8 / 1e6),
                           setup_s=t_setup, t_factor_s=M.t_factor, t_upload_s=M.t_upload,
                           ldl=info, device_ms_per_step=dev_ms_max / args.steps,
                           device_ms_per_step_by_rank=[float(t[1]) / args.steps for t in per_rank]),
            "e2e": {"value": tot_e2e_iters / (e2e_ms_max * 1e-3), "unit": "iterations/s",
                    "h2d_bytes_per_step": 8 * N, "d2h_bytes_per_step": 8 * N + 8 * int(stats_last["hist_len"]) + 160},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "k_solve<%s,grid> (one persistent launch per solve)" % solver,
                         "algorithmic_bytes_per_launch": bytes_solve, "bytes_parts": parts, "peak_source": peak_src,
                         "kernel_ms": kernel_ms},
            "clocks": clocks,
        }
        if not args.no_parts:
            line["roofline_parts"] = roofline_parts(torch, L, S, M, s, info, peak)
        if not args.no_parity:
            line["parity"] = parity_section(s, solver, opts, M.factors, x_gpu, stats_last)
        if args.profile:
            tot = phase.sum() or 1.0
            line["phase_share"] = {nm: float(phase[i] / tot) for i, nm in enumerate(_lib.PHASE_NAMES)}
        if not args.no_cpu_baseline:
            from oracle import cpk_oracle as orc
            orc.build_c_kernels()
            kit = args.cpu_baseline_iters
            t = time.perf_counter()
            xo, so, fo = orc.reg_cpkrylov(solver, s["rhs"], s["H"], s["B"], s["C"], s["G"],
                                          dict(opts, print=False, itmax=kit), factor=lambda K: M.factors)
            line["cpu_baseline"] = {"value": so["niters"] / so["stime"], "unit": "iterations/s", "cores": blas_threads(), "kind": "port",
                                    "sample": "first %d iterations of the same solve (oracle/cpk_oracle.py + oracle/kernels.c: sparse kernels "
                                              "single-threaded, NumPy BLAS-1 on the BLAS thread pool; host has %d cores)" % (kit, os.cpu_count())}
    S.close()
    extras = not args.no_extras and args.workload == "kkt_lap3d" and args.k == 2 and not args.g
    if extras:
        # BASELINE cfg 5, sharded over the ranks of THIS run (strong scaling: 256 systems in total)
        b5 = batch_measure(args, rank, local_rank, world, args.extras_batch_steps, 3, total=args.batch)
        if rank == 0:
            line["cfg5_ipm_batch"] = {k: b5[k] for k in ("value", "unit", "n_gpus", "steps", "ms_per_step", "scaling", "config", "e2e", "gpu_launches")}
        if world > 1:
            dist.barrier()
    if rank == 0:
        if extras:
            line["stress_k6"] = stress_section(args, torch, peak_gbs()[0])
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
