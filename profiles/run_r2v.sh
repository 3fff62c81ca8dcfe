# round 2v: CTA-per-split rounds in the batched refinement; first residual candidates of four samples from one Philox block
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_regions.py tests/test_gpu_tolerance.py tests/test_gpu_cv.py tests/test_gpu_full_size.py -m gpu -q > gpurun_out/r2v_tests.log 2>&1; tail -5 gpurun_out/r2v_tests.log
for m in 0 296 1184 4736 100000000; do
VB200_SPLIT_CTA_MAX=$m timeout 600 python bench.py --workload c4 --no-cpu-baseline --steps 10 --warmup 3 --sustain 0.5 2> gpurun_out/r2v_bench_c4.err | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=j['roofline']
print('cta_max $m', 'step ms %.3f' % j['ms_per_step'], 'kernel ms/launch %.3f' % r['kernel_ms_per_launch'], 'frac %.3f' % r['frac'], 'e2e ms %.3f' % j['e2e']['ms_per_step'])" | tee -a gpurun_out/r2v_sweep.txt
done
