# round 2j: generic greedy kernel (float / double tables, 8- / 16-byte heap entries), error_heuristic_mixed, fp64 adaptive generation
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2j_tests.log 2>&1; tail -14 gpurun_out/r2j_tests.log
timeout 300 python bench.py --workload c3 --exact --no-cpu-baseline > gpurun_out/r2j_bench_c3.json 2> gpurun_out/r2j_bench_c3.err; tail -c 300 gpurun_out/r2j_bench_c3.json
