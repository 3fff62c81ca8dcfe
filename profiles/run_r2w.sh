# round 2w: separable fast bin walk (row coefficients x column power differences), templated split_points
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_regions.py tests/test_gpu_tolerance.py tests/test_gpu_cv.py tests/test_gpu_full_size.py tests/test_gpu_fubini.py -m gpu -q > gpurun_out/r2w_tests.log 2>&1; tail -5 gpurun_out/r2w_tests.log
timeout 600 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/r2w_bench_c4.json 2> gpurun_out/r2w_bench_c4.err; python -c "
import json
j=json.loads(open('gpurun_out/r2w_bench_c4.json').read().strip().splitlines()[-1]); r=j['roofline']
print('step ms %.3f' % j['ms_per_step'], 'kernel ms/launch %.3f' % r['kernel_ms_per_launch'], 'frac %.3f' % r['frac'], 'e2e ms %.3f' % j['e2e']['ms_per_step'], 'exact ms %.1f' % j['exact_mode']['ms_per_step'])"
BATCH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2w_launches_c4.csv python profiles/run_full.py c4 > gpurun_out/r2w_c4_run.log 2>&1; tail -3 gpurun_out/r2w_c4_run.log
python profiles/summarize_launches.py gpurun_out/r2w_launches_c4.csv
