mkdir -p gpurun_out
CPK_VERBOSE=1 python bench.py --no-cpu-baseline --no-parts --no-parity --no-extras --steps 3 2> gpurun_out/r2_setup33.err | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); c=d['config']; print('setup_s %.2f factor %.2f upload %.2f'%(c['setup_s'], c['t_factor_s'], c['t_upload_s']))"
grep "set-up ms\|build_sweeps ms" gpurun_out/r2_setup33.err
