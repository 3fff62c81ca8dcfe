# round 2, last check of the committed build: the whole GPU suite and smoke()
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2K_tests.log 2>&1; tail -4 gpurun_out/r2K_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2K_smoke.log 2>&1; tail -2 gpurun_out/r2K_smoke.log
timeout 200 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r2K_bench_default.json 2> gpurun_out/r2K_bench_default.err; python -c "
import json
j=json.loads(open('gpurun_out/r2K_bench_default.json').read().strip().splitlines()[-1]); print('c2 %.4g frac %.3f e2e %.4g | cv ms %.3f' % (j['value'], j['roofline']['frac'], j['e2e']['value'], j['cv']['ms_per_step']))"
