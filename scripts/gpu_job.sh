mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "gmres or dqgmres or cfg4 or convdiff or full" 2>&1 | tail -4 | tee gpurun_out/r2_pytest28.log
timeout 1500 python scripts/results_table.py 2>&1 | grep "^{" > gpurun_out/r2_results28.log; python - <<'PY'
import json
for l in open('gpurun_out/r2_results28.log'):
    r=json.loads(l); print(r['config'], r['solver'], {k:v for k,v in r.get('opts',{}).items() if k not in ('atol','rtol')}, 'iters', r.get('iters'), 'ms %.3f'%r.get('ms',0), 'frac %.3f'%r.get('frac',0))
PY
