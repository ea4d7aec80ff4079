for lib in b200 T896; do
  export CPK_LIB_PATH=$PWD/cpkrylov_b200/libcpk_$lib.so
  echo "== $lib"
  python scripts/results_table.py 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(d['config'][:5], d['solver'], d['opts'].get('restart', d['opts'].get('mem','')), d['opts'].get('nitref',''), 'ms %.2f frac %.3f'%(d['ms'],d['frac']))
"
  python scripts/compact_probe.py --quick 2>&1 | grep -A1 "\"cvxqp\|apply_us" | grep "us_per_iter\|apply_us\|cvxqp" | tr "\n" " "; echo
  python bench.py --workload ipm_batch --batch 256 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ipm_batch 256', round(d['value']), d['config']['device_ms_per_step'])"
  python scripts/stress_probe.py 2>&1 | tail -2
done
