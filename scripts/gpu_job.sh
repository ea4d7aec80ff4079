mkdir -p gpurun_out
N=${NGPU:-4}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -2 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().splitlines()[-1])
print('n_gpus', d['n_gpus'], 'value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])
b=d['cfg5_ipm_batch']; print('cfg5', b['n_gpus'], b['value'], b['ms_per_step'], b['config']['systems_per_rank'], b['config']['device_ms_per_step'], b['scaling'])
PY
