set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_table.jsonl
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest9.log; cat gpurun_out/r2_pytest9.log
LIBS="b200 up4 up6 r1" timeout 1500 scripts/ab_r2.sh 2>&1 | tee gpurun_out/r2_ab9.log
CPK_SELL_PACK=0 LIBS="b200" timeout 600 scripts/ab_r2.sh 2>&1 | sed 's/^b200/b200-nopack/' | tee -a gpurun_out/r2_ab9.log
CPK_LDL_RC=1 CPK_RESID_RC=1 LIBS="b200" timeout 600 scripts/ab_r2.sh 2>&1 | sed 's/^b200/b200-rc/' | tee -a gpurun_out/r2_ab9.log
