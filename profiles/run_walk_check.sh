python -m pytest tests/test_gpu_walk.py tests/test_gpu_fubini.py tests/test_gpu_examples.py tests/test_gpu_full_size.py -x -q -m gpu 2>&1 | tail -4
python bench.py --workload c5 --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2_bench_c5b.json 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_c5b.json').read().strip().splitlines()[-1])
print('c5 value %.1f G paths/s'%(d['value']/1e9), 'ms', d['ms_per_step'], 'e2e %.1f'%(d['e2e']['value']/1e9))
PY
