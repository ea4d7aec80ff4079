mkdir -p gpurun_out
( time python bench.py > gpurun_out/r2_bench24.json 2> gpurun_out/r2_bench24.err ) 2>&1 | tail -3
tail -3 gpurun_out/r2_bench24.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench24.json').read().strip().splitlines()[-1])
print({k:(v if not isinstance(v,dict) else '...') for k,v in d.items()})
print('cfg5', {k:v for k,v in d['cfg5_ipm_batch'].items() if k!='config'}, d['cfg5_ipm_batch']['config']['device_ms_per_step'])
s=d['stress_k6']; print('stress', s['ms_per_step'], s['value'], s['roofline']['frac'], s['config']['ldl'], s['config']['setup_s'], s.get('parity'))
PY
