mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "gmres or dqgmres" 2>&1 | tail -3 | tee gpurun_out/r2_pytest29.log
CPK_RESULTS_ONLY=cfg4 CPK_RESULTS_TAG=r2_cfg4 timeout 1500 python scripts/results_table.py 2>&1 | grep "^{" > gpurun_out/r2_results29.log; python - <<'PY'
import json
for l in open('gpurun_out/r2_results29.log'):
    r=json.loads(l); print(r['config'], r['solver'], {k:v for k,v in r.get('opts',{}).items() if k not in ('atol','rtol')}, 'iters', r.get('iters'), 'ms %.3f'%r.get('ms',0), 'frac %.3f'%r.get('frac',0))
PY
