# batch launches: threads per CTA (two CTAs per SM below ~450 threads)
run() { env "$@" python bench.py --workload ipm_batch --batch 256 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$*', 'e2e', round(d['value']), 'device ms', round(d['config']['device_ms_per_step'],2))"; }
run CPK_LIB_PATH=$PWD/cpkrylov_b200/libcpk_b200.so CPK_BATCH_BLOCK=896
run CPK_LIB_PATH=$PWD/cpkrylov_b200/libcpk_b200.so CPK_BATCH_BLOCK=512
run CPK_LIB_PATH=$PWD/cpkrylov_b200/libcpk_CW12.so CPK_BATCH_BLOCK=896
run CPK_LIB_PATH=$PWD/cpkrylov_b200/libcpk_CW12.so CPK_BATCH_BLOCK=448
run CPK_LIB_PATH=$PWD/cpkrylov_b200/libcpk_CW12.so CPK_BATCH_BLOCK=384
