mkdir -p gpurun_out
exp() { ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null; ncu -i gpurun_out/$1.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/$1.source.csv.gz; [ "$2" = keep ] || rm -f gpurun_out/$1.ncu-rep; }
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-parts --no-parity"
$B --profile > gpurun_out/r2_bench_profile.json 2> gpurun_out/r2_bench_profile.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_ncu1.log 2>&1
$B > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_solve -s 3 -c 1 -f -o gpurun_out/r2_k_solve_cpcg $B > gpurun_out/r2_ncu2.log 2>&1
exp r2_k_solve_cpcg
K="python scripts/kernel_bench.py --reps 12"
$K > gpurun_out/r2_kernel_bench.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_matvec -s 6 -c 1 -f -o gpurun_out/r2_k_matvec_H $K > gpurun_out/r2_ncu3.log 2>&1
exp r2_k_matvec_H
$K > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_apply -s 6 -c 1 -f -o gpurun_out/r2_k_apply_nitref0 $K > gpurun_out/r2_ncu4.log 2>&1
exp r2_k_apply_nitref0
$K > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_apply -s 18 -c 1 -f -o gpurun_out/r2_k_apply_default $K > gpurun_out/r2_ncu5.log 2>&1
exp r2_k_apply_default
du -sh gpurun_out; tail -2 gpurun_out/r2_kernel_bench.log | cut -c1-500
