#!/bin/bash
# wider A/B of library builds: all solvers on cfg 3 / cfg 4 (results_table), cfg 5 and the stress system (bench extras)
for lib in ${LIBS:-b200}; do
  export CPK_LIB_PATH=$PWD/cpkrylov_b200/libcpk_$lib.so
  echo "== $lib"
  python scripts/results_table.py 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(d['config'][:5], d['solver'], d['opts'].get('restart', d['opts'].get('mem','')), d['opts'].get('nitref',''), 'iters', d['iters'], 'ms %.3f frac %.3f'%(d['ms'],d['frac']))
"
  python bench.py --no-cpu-baseline --no-parts --no-parity --steps 10 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
b=d['cfg5_ipm_batch']; s=d['stress_k6']
print('cfg5 it/s %d ms %.3f | stress ms %.3f frac %.3f' % (b['value'], b['config']['device_ms_per_step'], s['ms_per_step'], s['roofline']['frac']))
"
done
