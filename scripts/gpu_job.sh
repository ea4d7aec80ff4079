set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_table.jsonl
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2_pytest10.log; cat gpurun_out/r2_pytest10.log
timeout 600 python bench.py > gpurun_out/r2_bench10.json 2> gpurun_out/r2_bench10.err; tail -c 3000 gpurun_out/r2_bench10.json; tail -5 gpurun_out/r2_bench10.err
timeout 600 python scripts/cusparse_compare.py > gpurun_out/r2_cusparse.json 2> gpurun_out/r2_cusparse.err; cat gpurun_out/r2_cusparse.json; tail -3 gpurun_out/r2_cusparse.err
