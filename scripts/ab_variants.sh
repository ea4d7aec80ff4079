# A/B runs on one box: environment switches / alternative builds of the library (CPK_LIB_PATH)
for sig in 4096 256 65536 32; do CPK_SELL_SIGMA=$sig python bench.py --no-cpu-baseline --steps 30 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('sigma $sig', round(d['config']['device_ms_per_step'],4), round(d['value']))"; done
