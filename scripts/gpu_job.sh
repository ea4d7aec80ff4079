set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_matio.py tests/test_device_factorization.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest5.log; cat gpurun_out/r2_pytest5.log
LIBS="b200 pf ch8 b4422 b2211 t768 r1" timeout 1500 scripts/ab_r2.sh 2>&1 | tee gpurun_out/r2_ab5.log
