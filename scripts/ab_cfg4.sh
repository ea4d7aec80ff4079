#!/bin/bash
# A/B of library builds on the GMRES-family lines of the results table (cfg 3 restart/mem 20, cfg 4)
for lib in ${LIBS:-b200}; do
  export CPK_LIB_PATH=$PWD/cpkrylov_b200/libcpk_$lib.so
  echo "== $lib"
  CPK_RESULTS_ONLY=cfg4 python scripts/results_table.py 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(d['config'][:5], d['solver'], d['opts'].get('restart', d['opts'].get('mem','')), 'iters', d['iters'], 'ms %.3f frac %.3f'%(d['ms'],d['frac']))
"
done
