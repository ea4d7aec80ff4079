mkdir -p gpurun_out
timeout 900 python scripts/compact_probe.py > gpurun_out/r2_compact30.log 2>&1
python - <<'PY'
import json
d=json.load(open('gpurun_out/compact_probe.json'))
for k,v in d.items():
    if v is None: print(k, None); continue
    print(k, {kk:(round(vv['us_per_iter'],1), vv['iters'], '%.1e'%vv['err_vs_oracle']) for kk,vv in v.items() if isinstance(vv,dict)}, 'apply_us', round(v['apply_us'],1), 'walk', v['walk_cycles(gather,ring wait,steps,scatter)'])
PY
for env in "CPK_CW_NO_CHAINS=1" "A=1" "CPK_CW_CHAIN_ITEMS=8"; do
env $env python bench.py --workload ipm_batch --steps 5 --warmup 3 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); c=d['config']
print('$env batch 256: device ms/step %.3f e2e ms/step %.3f it/s %d' % (c['device_ms_per_step'], d['ms_per_step'], d['value']))"
done
timeout 900 python -m pytest tests -m gpu -q -x -k "cvxqp or fixture or batch or cfg5 or device_factor or sequence or matio" 2>&1 | tail -3
