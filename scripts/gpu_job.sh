mkdir -p gpurun_out
: > gpurun_out/r2_ab44.log
for rep in 1 2; do
for lib in prev mixed; do
export CPK_LIB_PATH=$PWD/cpkrylov_b200/libcpk_$lib.so
python bench.py --no-cpu-baseline --no-parts --no-parity --no-extras --steps 30 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); c=d['config']
print('$lib cfg3 ms/solve %.4f frac %.3f' % (c['device_ms_per_step'], d['roofline']['frac']))" | tee -a gpurun_out/r2_ab44.log
python bench.py --workload kkt_convdiff --no-cpu-baseline --no-parts --no-parity --no-extras --steps 6 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); c=d['config']
print('$lib cfg4 ms/solve %.3f frac %.3f' % (c['device_ms_per_step'], d['roofline']['frac']))" | tee -a gpurun_out/r2_ab44.log
done
done
