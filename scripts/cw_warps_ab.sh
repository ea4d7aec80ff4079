for lib in b200 CW8 CW12 CW24 b200; do
  export CPK_LIB_PATH=$PWD/cpkrylov_b200/libcpk_$lib.so
  python scripts/compact_probe.py --quick 2>&1 | python -c "
import sys,json
t=sys.stdin.read(); i=t.index('{'); d=json.loads(t[i:t.rindex('}')+1]) if False else None
" 2>/dev/null
  python - <<'PY'
import json,subprocess,sys,os
p=subprocess.run([sys.executable,'scripts/compact_probe.py','--quick'],stdout=subprocess.PIPE,stderr=subprocess.STDOUT,text=True)
d=json.load(open('gpurun_out/compact_probe.json'))['compact']
print(os.environ['CPK_LIB_PATH'].split('_')[-1], 'cvxqp1 minres %.1f cg %.1f | cvxqp2 gmres %.1f dqgmres %.1f | apply %.1f us | steps %d' % (d['cvxqp1_m:cpminres']['us_per_iter'], d['cvxqp1_m:cpcg']['us_per_iter'], d['cvxqp2_s:cpgmres']['us_per_iter'], d['cvxqp2_s:cpdqgmres']['us_per_iter'], d['apply_us'], d['walk_cycles(gather,ring wait,steps,scatter)'][2]))
PY
done
