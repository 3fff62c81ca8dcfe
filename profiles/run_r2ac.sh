# round 2ac: inverse extents in the staged record, closed-form S = 3 basis, hoisted slot positions, constant tile size
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cv.py tests/test_gpu_full_size.py tests/test_gpu_fubini.py -m gpu -q > gpurun_out/r2ac_tests.log 2>&1; tail -5 gpurun_out/r2ac_tests.log
timeout 600 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/r2ac_bench_c4.json 2> gpurun_out/r2ac_bench_c4.err; python -c "
import json
j=json.loads(open('gpurun_out/r2ac_bench_c4.json').read().strip().splitlines()[-1]); r=j['roofline']
print('step ms %.3f' % j['ms_per_step'], 'kernel ms/launch %.3f' % r['kernel_ms_per_launch'], 'frac %.3f' % r['frac'], 'e2e ms %.3f' % j['e2e']['ms_per_step'], 'exact ms %.1f' % j['exact_mode']['ms_per_step'])"
