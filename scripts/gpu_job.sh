mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r2_pytest41.log
python bench.py > gpurun_out/r2_bench41.json 2> gpurun_out/r2_bench41.err; tail -2 gpurun_out/r2_bench41.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench41.json').read().strip().splitlines()[-1])
c=d['config']
print('cfg3 ms/solve %.4f it/s %d e2e %d frac %.3f setup_s %.2f'%(c['device_ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], c['setup_s']))
for k,v in d['roofline_parts'].items(): print(' ',k,'us %.1f frac %.3f'%(v['us'],v['frac']))
b=d['cfg5_ipm_batch']; print('cfg5', b['value'], b['ms_per_step'], b['config']['device_ms_per_step'])
s=d['stress_k6']; print('stress', s['ms_per_step'], s['roofline']['frac'], s['config']['setup_s'], s['parity']['relerr_vs_oracle'])
PY
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
