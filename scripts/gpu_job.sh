set -x
mkdir -p gpurun_out
CPK_HOSTOP_TIMEOUT_S=10 timeout 600 python -m pytest tests/test_matrix_free.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r2_pytest11.log; cat gpurun_out/r2_pytest11.log
