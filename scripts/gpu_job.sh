mkdir -p gpurun_out
timeout 1200 python scripts/team_crossover.py 2>&1 | tee gpurun_out/r2_team45.log
