mkdir -p gpurun_out
out=gpurun_out/r2_stress23.log; : > $out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "stress or walk_variants or opldl2 or full_size" 2>&1 | tail -3 | tee -a $out
env A=1 timeout 300 python scripts/stress_env_probe.py 40 "g40" 2>&1 | tail -1 | tee -a $out
env A=1 timeout 300 python scripts/stress_env_probe.py 60 "g60" 2>&1 | tail -1 | tee -a $out
for W in 32 64; do
echo "== CPK_LDL_WIDE_ROW=$W" | tee -a $out
CPK_VERBOSE=1 CPK_LDL_WIDE_ROW=$W timeout 900 python bench.py --k 6 --window 64 --steps 5 --warmup 3 --no-cpu-baseline --no-parts --profile 2>gpurun_out/r2_stress23_W$W.err | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); c=d['config']
print('stress g100 ms/solve %.3f iters %d frac %.4f setup_s %.1f parity %s phases %s' % (c['device_ms_per_step'], c['iters_per_solve'], d['roofline']['frac'], c['setup_s'], d['parity']['relerr_vs_oracle'], {k: round(v,3) for k,v in d.get('phase_share',{}).items()}))" | tee -a $out
done
export LIBS=b200
timeout 600 scripts/ab_r2.sh 2>&1 | tee gpurun_out/r2_ab23.log
timeout 600 python scripts/compact_probe.py --quick 2>&1 | grep -A12 "^default" | grep "us_per_iter\|cvxqp" | tee -a $out
