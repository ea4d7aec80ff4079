mkdir -p gpurun_out
CPK_RESULTS_TAG=r2 timeout 1500 python scripts/results_table.py 2>&1 | grep "^{" > gpurun_out/r2_results46.log; python - <<'PY'
import json
for l in open('gpurun_out/r2_results46.log'):
    r=json.loads(l); print(r['config'][:4], r['solver'], {k:v for k,v in r.get('opts',{}).items() if k not in ('atol','rtol','itmax','force_itref')}, 'iters', r.get('iters'), 'ms %.3f'%r.get('ms',0), 'it/s %d'%r.get('it_per_s',0), 'GBs %d'%r.get('GBs',0), 'frac %.3f'%r.get('frac',0))
PY
