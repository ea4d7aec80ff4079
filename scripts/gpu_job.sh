mkdir -p gpurun_out
# part 1: tests, default bench line, smoke, results table
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r2_pytest70.log
python bench.py > gpurun_out/r2_bench70.json 2> gpurun_out/r2_bench70.err; tail -2 gpurun_out/r2_bench70.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench70.json').read().strip().splitlines()[-1])
c=d['config']
print('cfg3 ms/solve %.4f it/s %d e2e %d frac %.3f setup_s %.2f'%(c['device_ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], c['setup_s']))
for k,v in d['roofline_parts'].items(): print(' ',k,'us %.1f frac %.3f'%(v['us'],v['frac']))
b=d['cfg5_ipm_batch']; print('cfg5', b['value'], b['ms_per_step'], b['config']['device_ms_per_step'])
s=d['stress_k6']; print('stress', s['ms_per_step'], s['roofline']['frac'], s['config']['setup_s'], s['parity']['relerr_vs_oracle'])
print('parity', d['parity'])
PY
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
CPK_RESULTS_TAG=r2 timeout 1500 python scripts/results_table.py 2>&1 | grep "^{" > gpurun_out/r2_results70.log
# part 2: launch list and ncu captures (each after its own command ran without ncu)
exp() { ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null; ncu -i gpurun_out/$1.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/$1.source.csv.gz; [ "$2" = keep ] || rm -f gpurun_out/$1.ncu-rep; }
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-parts --no-parity"
$B --profile > gpurun_out/r2_bench_profile.json 2> gpurun_out/r2_bench_profile.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_ncu1.log 2>&1
$B > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_solve -s 3 -c 1 -f -o gpurun_out/r2_k_solve_cpcg $B > gpurun_out/r2_ncu2.log 2>&1
exp r2_k_solve_cpcg
K="python scripts/kernel_bench.py --reps 12"
$K > gpurun_out/r2_kernel_bench.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_matvec -s 6 -c 1 -f -o gpurun_out/r2_k_matvec_H $K > gpurun_out/r2_ncu3.log 2>&1
exp r2_k_matvec_H
$K > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_apply -s 6 -c 1 -f -o gpurun_out/r2_k_apply_nitref0 $K > gpurun_out/r2_ncu4.log 2>&1
exp r2_k_apply_nitref0
$K > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_apply -s 18 -c 1 -f -o gpurun_out/r2_k_apply_default $K > gpurun_out/r2_ncu5.log 2>&1
exp r2_k_apply_default
du -sh gpurun_out; tail -2 gpurun_out/r2_kernel_bench.log | cut -c1-500
