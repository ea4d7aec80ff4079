for lib in $LIBS; do
  export CPK_LIB_PATH=$PWD/cpkrylov_b200/libcpk_$lib.so
  python bench.py --no-cpu-baseline --no-parts --no-parity --no-extras --steps 30 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); c=d['config']
print('$lib', 'cfg3 cpcg ms/solve %.4f it/s %d frac %.3f' % (c['device_ms_per_step'], d['value'], d['roofline']['frac']))
"
done
