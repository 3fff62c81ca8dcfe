# round 2ak: histograms zeroed by the write kernel (6 launches per big round)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_regions.py tests/test_gpu_tolerance.py -m gpu -q > gpurun_out/r2ak_tests.log 2>&1; tail -4 gpurun_out/r2ak_tests.log
for wl in c4 c3; do
timeout 300 python bench.py --workload $wl --no-cpu-baseline --steps 10 --warmup 3 --sustain 0.5 2> gpurun_out/r2ak.err | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl', 'step ms %.3f' % j['ms_per_step'], 'e2e ms %.3f' % j['e2e']['ms_per_step'])" | tee -a gpurun_out/r2ak_sweep.txt
done
