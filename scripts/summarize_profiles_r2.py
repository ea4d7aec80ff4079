"""Round 2: turns the ncu exports a gpurun call brought back (gpurun_out/r2_*.raw.csv = `--page raw --csv`,
r2_*.source.csv.gz = `--page source --csv` of one `--set full` capture each) into the tracked summaries under
profiles/.  The captures themselves are taken by scripts/gpu_job.sh (commands quoted in every summary)."""
import csv, gzip, io, json, os, shutil, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
        'smsp__inst_executed.sum', 'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio']
MULT = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1, 'ms': 1e3, 'us': 1.0, 'ns': 1e-3, 'msecond': 1e3, 'usecond': 1.0, 'nsecond': 1e-3, 'second': 1e6}


def raw(tag):
    rows = list(csv.reader(open(os.path.join(G, tag + ".raw.csv"))))
    return dict(zip(rows[0], rows[1])), dict(zip(rows[0], rows[2]))


def stalls(tag):
    p = os.path.join(G, tag + ".source.csv.gz")
    if not os.path.exists(p):
        return None, 0, 0
    rows = list(csv.reader(io.TextIOWrapper(gzip.open(p), newline="")))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    names = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    agg = {s: 0 for s in names}
    for r in data:
        for s in names:
            try:
                agg[s] += int(r[ix[s]])
            except Exception:
                pass
    tot = sum(agg.values()) or 1
    return {k: round(100.0 * v / tot, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 100 >= tot}, tot, len(data)


def summary(tag, title, command, what, alg_bytes, alg_note, extra=""):
    units, d = raw(tag)
    rd = float(d['dram__bytes_read.sum']) * MULT[units['dram__bytes_read.sum']]
    wr = float(d['dram__bytes_write.sum']) * MULT[units['dram__bytes_write.sum']]
    us = float(d['gpu__time_duration.sum']) * MULT[units['gpu__time_duration.sum']]
    st, nsamp, ninst = stalls(tag)
    out = os.path.join(P, tag + "_ncu_summary.md")
    with open(out, "w") as f:
        f.write("# %s\n\nCommand (on the B200 box, directly after the same command exited 0 without ncu):\n\n```\n%s\n```\n\n%s\n\n" % (title, command, what))
        f.write("| metric | value | unit |\n|---|---|---|\n")
        for k in KEYS:
            if k in d:
                f.write("| `%s` | %s | %s |\n" % (k, d[k], units.get(k, '')))
        f.write("\nDuration under ncu (caches flushed before every replay pass, cold): %.1f us.  DRAM traffic of the launch = %.1f MB "
                "(read %.1f + write %.1f); algorithmic bytes = %.1f MB (%s) -> traffic / algorithmic = %.2f; algorithmic GB/s at this "
                "duration = %.0f (%.3f of the measured 6549 GB/s copy peak).\n" % (us, (rd + wr) / 1e6, rd / 1e6, wr / 1e6, alg_bytes / 1e6, alg_note,
                                                                               (rd + wr) / alg_bytes, alg_bytes / us / 1e3, alg_bytes / us / 1e3 / 6549.4))
        if st:
            f.write("\nWarp-state samples by reason (source page, %d samples over %d SASS instructions; reasons with at least 1 %%): %s.\n" % (nsamp, ninst, json.dumps(st)))
        if extra:
            f.write("\n" + extra + "\n")
    return dict(us=us, dram=rd + wr, read=rd, write=wr)


if __name__ == "__main__":
    bp = json.loads(open(os.path.join(G, "r2_bench_profile.json")).read().strip().splitlines()[-1])
    shutil.copy(os.path.join(G, "r2_bench_profile.json"), os.path.join(P, "r2_bench_profile.json"))
    shutil.copy(os.path.join(G, "r2_launches.csv"), os.path.join(P, "r2_launches.csv"))
    B = "python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-parts --no-parity"
    K = "python scripts/kernel_bench.py --reps 12"
    alg = bp['roofline']['algorithmic_bytes_per_launch']
    r = summary("r2_k_solve_cpcg", "Round 2 — ncu `--set full` of the dominant kernel (whole cpcg solve, BASELINE cfg 3)",
                "ncu --set full --clock-control none --import-source on -k regex:k_solve -s 3 -c 1 -o r2_k_solve_cpcg " + B,
                "Kernel: `cpk::k_solve<0, true>` = cpcg, the whole solve (%d iterations) in ONE cooperative launch, 148 CTAs x 896 threads."
                % bp['config']['iters_per_solve'], alg, "DESIGN.md section 5, counts reported by the device",
                "Share check against the launch list (`r2_launches.csv`, cold-cache, serialised): the timed region of `bench.py` holds only `k_solve` launches "
                "(one per step, `gpu_launches` = steps): the persistent kernel is 100 %% of the step on both views.  Phase shares of this build "
                "(`r2_bench_profile.json`, cycle counters of team thread 0): %s." % json.dumps({k: round(v, 3) for k, v in bp.get("phase_share", {}).items()}))
    json.dump({"dram_bytes_per_launch": r["dram"], "dram_read": r["read"], "dram_write": r["write"],
               "source": "ncu --set full, k_solve<0,true> (cpcg, grid team), launch 4 of `%s` (round 2 build)" % B},
              open(os.path.join(P, "traffic_kkt_lap3d.json"), "w"), indent=1)
    spmv = lambda nnz, rr, c: 12 * nnz + 4 * (rr + 1) + 8 * c + 8 * rr
    n, m = 1000000, 250000
    N = n + m
    summary("r2_k_matvec_H", "Round 2 — ncu `--set full` of the stand-alone `H*v` (cpk_system_matvec, cfg 3)",
            "ncu --set full --clock-control none --import-source on -k regex:k_matvec -s 6 -c 1 -o r2_k_matvec_H " + K,
            "Kernel: `k_matvec<true>`, the streaming SELL pass of the solvers' `spmv_sell` over H (n = 1 000 000, 6.94 M entries, packed 4-byte entries).",
            spmv(6940000, n, n), "12 nnz + 4(r+1) + 8c + 8r")
    B_ldl = 24 * 500000 + 48 * N
    B_kp = spmv(1250000 + 2 * 500000, N, N)
    summary("r2_k_apply_nitref0", "Round 2 — ncu `--set full` of the stand-alone `M*z`, nitref = 0 (the LDL' solve alone, cfg 3)",
            "ncu --set full --clock-control none --import-source on -k regex:k_apply -s 6 -c 1 -o r2_k_apply_nitref0 " + K,
            "Kernel: `k_apply<true>` on the grid team: forward level, team barrier, backward level (19+19 levels merged to 1+1).",
            B_ldl, "24 nnz_off(L) + 48 N")
    summary("r2_k_apply_default", "Round 2 — ncu `--set full` of the stand-alone `M*z`, reference defaults (solve + refinement residual, cfg 3)",
            "ncu --set full --clock-control none --import-source on -k regex:k_apply -s 18 -c 1 -o r2_k_apply_default " + K,
            "Kernel: `k_apply<true>`: the two sweep levels, then the residual pass x - K_P*y with its norms (no refinement step is taken).",
            B_ldl + B_kp + 8 * N, "B_ldl + B_spmv(K_P) + 8 N")
    if os.path.exists(os.path.join(G, "r2_k_solve_stress60.raw.csv")):
        bs = json.loads(open(os.path.join(G, "r2_bench_stress60.json")).read().strip().splitlines()[-1])
        shutil.copy(os.path.join(G, "r2_bench_stress60.json"), os.path.join(P, "r2_bench_stress60.json"))
        summary("r2_k_solve_stress60", "Round 2 — ncu `--set full` of a cpcg solve on the stress system (k = 6 windowed B, filled factor), g = 60",
                "ncu --set full --clock-control none --import-source on -k regex:k_solve -s 3 -c 1 -o r2_k_solve_stress60 "
                "python bench.py --g 60 --k 6 --window 64 --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-parts --no-parity",
                "Kernel: `cpk::k_solve<0, true>`: N = 270 000, %s; the sweeps run the sync-free tagged walk over the merged item list."
                % json.dumps(bs['config']['ldl']), bs['roofline']['algorithmic_bytes_per_launch'], "DESIGN.md section 5")
    for fn in sorted(os.listdir(P)):
        if fn.startswith("r2_k_") and fn.endswith("_ncu_summary.md"):
            print(open(os.path.join(P, fn)).read())
