# A/B runs on ONE box (boxes differ by ~2 %): alternative builds of the library through CPK_LIB_PATH,
# e.g.  make -C cpkrylov_b200/csrc clean && make -C cpkrylov_b200/csrc -j8 EXTRA="-DCPK_TEAM_MAP_B=4" OUT=../libcpk_TM4.so
# compile-time knobs: CPK_SPMV_U (8), CPK_SWEEP_B (4), CPK_TEAM_MAP_B (1), CPK_TEAM_MAP_SCALAR, CPK_BLOCK (512)
for lib in ${LIBS:-b200}; do
  CPK_LIB_PATH=$PWD/cpkrylov_b200/libcpk_$lib.so python bench.py --no-cpu-baseline --steps 30 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', round(d['config']['device_ms_per_step'],4), round(d['value']))"
done
