# round 1d: two-tile window walk kernel (C5) — parity, timing against the tile-at-a-time kernel, ncu, then the whole GPU suite and the C2 bench
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_walk.py -x -q -m gpu > gpurun_out/r1d_walk_tests.log 2>&1; tail -5 gpurun_out/r1d_walk_tests.log
timeout 200 python bench.py --workload c5 --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r1d_bench_c5_window.json 2> gpurun_out/r1d_bench_c5_window.err
VB200_WALK_WINDOW=0 timeout 200 python bench.py --workload c5 --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r1d_bench_c5_tile.json 2> gpurun_out/r1d_bench_c5_tile.err
python - <<'PY'
import json
for n in ("window", "tile"):
    try:
        d = json.loads(open(f'gpurun_out/r1d_bench_c5_{n}.json').read().strip().splitlines()[-1])
        print(n, 'c5 value %.1f G paths/s' % (d['value'] / 1e9), 'ms', d['ms_per_step'], 'e2e %.1f' % (d['e2e']['value'] / 1e9))
    except Exception as e:
        print(n, "failed", e)
PY
timeout 200 ncu --set full --clock-control none --import-source on -k regex:walk_block -s 1 -c 1 -o gpurun_out/r1d_c5_window -f python profiles/run_c2.py c5 3 > gpurun_out/r1d_ncu.log 2>&1; tail -3 gpurun_out/r1d_ncu.log
timeout 600 python -m pytest tests -m gpu -x -q --durations=10 > gpurun_out/r1d_tests.log 2>&1; tail -20 gpurun_out/r1d_tests.log
timeout 200 python bench.py > gpurun_out/r1d_bench_c2.json 2> gpurun_out/r1d_bench_c2.err; tail -c 1500 gpurun_out/r1d_bench_c2.json
timeout 100 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1d_bench_ref.json 2>&1; tail -c 600 gpurun_out/r1d_bench_ref.json
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1d_smoke.log 2>&1; tail -2 gpurun_out/r1d_smoke.log
