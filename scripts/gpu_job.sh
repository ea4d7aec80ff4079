mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
: > gpurun_out/r2_ab51.log
for lib in prev queue; do
export CPK_LIB_PATH=$PWD/cpkrylov_b200/libcpk_$lib.so
python bench.py --no-cpu-baseline --no-parts --no-parity --no-extras --steps 30 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); c=d['config']
print('$lib cfg3 ms/solve %.4f frac %.3f' % (c['device_ms_per_step'], d['roofline']['frac']))" | tee -a gpurun_out/r2_ab51.log
timeout 600 python scripts/compact_probe.py --quick 2>&1 | grep "us_per_iter" | head -4 | tr '\n' ' ' | sed "s/^/$lib fixtures /" | tee -a gpurun_out/r2_ab51.log; echo
python scripts/subteam_probe.py 2>&1 | head -3 | sed "s/^/$lib /" | tee -a gpurun_out/r2_ab51.log
env A=1 timeout 400 python scripts/stress_env_probe.py 40 "$lib g40" 2>&1 | tail -1 | tee -a gpurun_out/r2_ab51.log
python bench.py --workload ipm_batch --steps 5 --warmup 3 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); c=d['config']
print('$lib cfg5 device ms/step %.3f' % c['device_ms_per_step'])" | tee -a gpurun_out/r2_ab51.log
done
