mkdir -p gpurun_out
timeout 900 python scripts/compact_probe.py --quick 2>&1 | tee gpurun_out/r2_compact18.log | grep -v "^ " | cut -c1-1200
python bench.py --workload ipm_batch --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/r2_batch18.json
out=gpurun_out/r2_stress18.log; : > $out
run() { env "$@" timeout 300 python scripts/stress_env_probe.py 40 "$*" 2>&1 | tail -1 | tee -a $out; }
run A=0
run CPK_LDL_TAIL_MAXLEN=384
run CPK_LDL_TAIL_FILL=10
run CPK_LDL_TAIL_FILL=10 CPK_LDL_TAIL_MAXLEN=384
run CPK_LDL_TAIL_FILL=4
