# round 2aa: tile lists by brute-force ordered compaction (no sort) against region-major binning + sort, C4
set -x
mkdir -p gpurun_out
for lim in 268435456 1073741824 4294967296; do
VB200_TILE_PAIR_LIMIT=$lim timeout 300 python bench.py --workload c4 --no-cpu-baseline --steps 10 --warmup 3 --sustain 0.5 2> gpurun_out/r2aa.err | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=j['roofline']
print('pair_limit $lim', 'step ms %.3f' % j['ms_per_step'], 'e2e ms %.3f' % j['e2e']['ms_per_step'], 'launches %.0f' % j['gpu_launches_per_step'])" | tee -a gpurun_out/r2aa_sweep.txt
done
VB200_TILE_PAIR_LIMIT=4294967296 timeout 300 python bench.py --workload c3 --no-cpu-baseline --steps 10 --warmup 3 --sustain 0.5 2> gpurun_out/r2aa.err | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c3 pair_limit 2^32', 'step ms %.3f' % j['ms_per_step'])" | tee -a gpurun_out/r2aa_sweep.txt
timeout 300 python bench.py --workload c3 --no-cpu-baseline --steps 10 --warmup 3 --sustain 0.5 2> gpurun_out/r2aa.err | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c3 default', 'step ms %.3f' % j['ms_per_step'])" | tee -a gpurun_out/r2aa_sweep.txt
timeout 600 python -m pytest tests/test_gpu_cv.py -m gpu -q -k "kernel_timer" > gpurun_out/r2aa_tests.log 2>&1; tail -3 gpurun_out/r2aa_tests.log
