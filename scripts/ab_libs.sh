#!/bin/bash
# A/B of library builds on one box (scripts/build_variant.sh): cfg 3 bench with and without phase clocks + stand-alone kernels
# usage: LIBS="b200 h1 ..." scripts/ab_libs.sh
for lib in ${LIBS:-b200}; do
  export CPK_LIB_PATH=$PWD/cpkrylov_b200/libcpk_$lib.so
  python bench.py --no-cpu-baseline --no-parts --no-parity --no-extras --steps 20 --profile 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); c=d['config']
print('$lib', 'cfg3 cpcg ms/solve %.4f iters %d frac %.3f e2e %d relerr %.2e' % (c['device_ms_per_step'], c['iters_per_solve'], d['roofline']['frac'], d['e2e']['value'], c['relerr_vs_xstar']), 'phases', {k: round(v,3) for k,v in d.get('phase_share',{}).items()})
" || echo "$lib bench failed"
  python bench.py --no-cpu-baseline --no-parts --no-parity --no-extras --steps 30 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); c=d['config']
print('$lib', 'cfg3 cpcg (no profile) ms/solve %.4f it/s %d frac %.3f' % (c['device_ms_per_step'], d['value'], d['roofline']['frac']))
" || echo "$lib bench failed"
  python scripts/kernel_bench.py 2>&1 | grep "^{" | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('$lib', 'kernels', {k: (round(v['us'],1), round(v['GBs'])) for k,v in d.items() if k!='info'})
" || echo "$lib kernel_bench failed"
done
