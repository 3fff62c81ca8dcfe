# quick C4 bench line (no tests): step time and the residual kernel's time per launch
mkdir -p gpurun_out
timeout 300 python bench.py --workload c4 --no-cpu-baseline --steps 10 --warmup 3 --sustain 0.5 2> gpurun_out/quick_c4.err | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=j['roofline']
print('step ms %.3f' % j['ms_per_step'], 'kernel ms/launch %.3f' % r['kernel_ms_per_launch'], 'frac %.3f' % r['frac'], 'e2e ms %.3f' % j['e2e']['ms_per_step'])"
