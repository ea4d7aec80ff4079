mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r2_pytest38.log
python bench.py > gpurun_out/r2_bench38.json 2> gpurun_out/r2_bench38.err; tail -2 gpurun_out/r2_bench38.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench38.json').read().strip().splitlines()[-1])
c=d['config']
print('cfg3 ms/solve %.4f it/s %d e2e %d frac %.3f setup_s %.2f (factor %.2f upload %.2f)'%(c['device_ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], c['setup_s'], c['t_factor_s'], c['t_upload_s']))
b=d['cfg5_ipm_batch']; print('cfg5', b['value'], b['ms_per_step'], b['config']['device_ms_per_step'])
s=d['stress_k6']; print('stress', s['ms_per_step'], s['roofline']['frac'], s['config']['setup_s'], s['parity']['relerr_vs_oracle'], s['parity']['niters_gpu'], s['parity']['niters_oracle'])
PY
timeout 600 python scripts/compact_probe.py --quick 2>&1 | grep "us_per_iter\|cvxqp\|apply_us"
out=gpurun_out/r2_stress38.log; : > $out
for g in 40 60; do env A=1 timeout 400 python scripts/stress_env_probe.py $g "g$g" 2>&1 | tail -1 | tee -a $out; done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
