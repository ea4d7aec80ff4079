mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r2_pytest52.log
python bench.py > gpurun_out/r2_bench52.json 2> gpurun_out/r2_bench52.err; tail -2 gpurun_out/r2_bench52.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench52.json').read().strip().splitlines()[-1])
c=d['config']
print('cfg3 ms/solve %.4f it/s %d e2e %d frac %.3f setup_s %.2f'%(c['device_ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], c['setup_s']))
for k,v in d['roofline_parts'].items(): print(' ',k,'us %.1f frac %.3f'%(v['us'],v['frac']))
b=d['cfg5_ipm_batch']; print('cfg5', b['value'], b['ms_per_step'], b['config']['device_ms_per_step'])
s=d['stress_k6']; print('stress', s['ms_per_step'], s['roofline']['frac'], s['config']['setup_s'], s['parity']['relerr_vs_oracle'])
print('parity', d['parity'])
PY
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
CPK_RESULTS_TAG=r2 timeout 1500 python scripts/results_table.py 2>&1 | grep "^{" > gpurun_out/r2_results52.log; python - <<'PY'
import json
for l in open('gpurun_out/r2_results52.log'):
    r=json.loads(l); print(r['config'][:4], r['solver'], {k:v for k,v in r.get('opts',{}).items() if k not in ('atol','rtol','itmax','force_itref')}, 'iters', r.get('iters'), 'ms %.3f'%r.get('ms',0), 'it/s %d'%r.get('it_per_s',0), 'GBs %d'%r.get('GBs',0), 'frac %.3f'%r.get('frac',0))
PY
