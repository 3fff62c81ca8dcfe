# round 2ae: tile sort with one thread per comparator and right-sized shared memory; accumulate pass size sweep (C4)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_regions.py tests/test_gpu_cv.py tests/test_gpu_full_size.py -m gpu -q > gpurun_out/r2ae_tests.log 2>&1; tail -4 gpurun_out/r2ae_tests.log
for ap in 32 16 8 64; do
VB200_CVT_ACCPASS=$ap timeout 300 python bench.py --workload c4 --no-cpu-baseline --steps 10 --warmup 3 --sustain 0.5 2> gpurun_out/r2ae.err | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=j['roofline']
print('accpass $ap', 'step ms %.3f' % j['ms_per_step'], 'kernel ms/launch %.3f' % r['kernel_ms_per_launch'], 'e2e ms %.3f' % j['e2e']['ms_per_step'])" | tee -a gpurun_out/r2ae_sweep.txt
done
