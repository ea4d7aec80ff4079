run() { env "$@" python bench.py --no-cpu-baseline --steps 30 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$*', round(d['config']['device_ms_per_step'],4), round(d['value']), d['config']['ldl'])"; }
run X=1
run CPK_LDL_CUT0=0
run CPK_LDL_TAIL_FILL=3
run CPK_LDL_TAIL_FILL=12
run CPK_LDL_TAIL_MAXLEN=32
run CPK_SELL_SIGMA=1024
run X=1
