set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "walk_variants or opldl2 or full_size_properties or stress" 2>&1 | tail -15 > gpurun_out/r2_pytest12.log; cat gpurun_out/r2_pytest12.log
export LIBS=b200
timeout 600 scripts/ab_r2.sh 2>&1 | sed 's/^b200/items/' | tee gpurun_out/r2_ab12.log
CPK_LDL_RC=1 timeout 600 scripts/ab_r2.sh 2>&1 | sed 's/^b200/rc-staged/' | tee -a gpurun_out/r2_ab12.log
CPK_LDL_RC=1 CPK_RESID_RC=1 timeout 600 scripts/ab_r2.sh 2>&1 | sed 's/^b200/rc+resid-staged/' | tee -a gpurun_out/r2_ab12.log
CPK_LDL_RC=1 CPK_RESID_RC=1 CPK_RC_STAGE=0 timeout 600 scripts/ab_r2.sh 2>&1 | sed 's/^b200/rc+resid-unstaged/' | tee -a gpurun_out/r2_ab12.log
