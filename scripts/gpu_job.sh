mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2_pytest26.log; cat gpurun_out/r2_pytest26.log
export LIBS=b200
timeout 600 scripts/ab_r2.sh 2>&1 | tee gpurun_out/r2_ab26.log
