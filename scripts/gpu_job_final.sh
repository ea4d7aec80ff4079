mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r2_pytest73.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
CPK_RESULTS_TAG=r2 timeout 900 python scripts/results_table.py 2>&1 | grep "^{" > gpurun_out/r2_results73.log
python - <<'PY'
import json
for l in open('gpurun_out/r2_results73.log'):
    r=json.loads(l); print(r['config'][:4], r['solver'], {k:v for k,v in r.get('opts',{}).items() if k not in ('atol','rtol','itmax','force_itref')}, 'iters', r.get('iters'), 'ms %.3f'%r.get('ms',0), 'it/s %d'%r.get('it_per_s',0), 'GBs %d'%r.get('GBs',0), 'frac %.3f'%r.get('frac',0))
PY
python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2_bench73.json 2> gpurun_out/r2_bench73.err; tail -1 gpurun_out/r2_bench73.json | cut -c1-300
