mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r2_pytest71.log
python bench.py > gpurun_out/r2_bench71.json 2> gpurun_out/r2_bench71.err; tail -2 gpurun_out/r2_bench71.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench71.json').read().strip().splitlines()[-1])
c=d['config']
print('cfg3 ms/solve %.4f it/s %d e2e %d frac %.3f'%(c['device_ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac']))
for k,v in d['roofline_parts'].items(): print(' ',k,'us %.1f frac %.3f'%(v['us'],v['frac']))
PY
exp() { ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null; ncu -i gpurun_out/$1.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/$1.source.csv.gz; rm -f gpurun_out/$1.ncu-rep; }
K="python scripts/kernel_bench.py --reps 12"
$K > gpurun_out/r2_kernel_bench.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_matvec -s 6 -c 1 -f -o gpurun_out/r2_k_matvec_H $K > gpurun_out/r2_ncu3.log 2>&1
exp r2_k_matvec_H
tail -2 gpurun_out/r2_kernel_bench.log | cut -c1-300
