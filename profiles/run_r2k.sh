# round 2k: importance / MIS / roulette region sampling, stratified allocation (Optimized integrator), faster batched refinement kernels
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2k_tests.log 2>&1; tail -14 gpurun_out/r2k_tests.log
timeout 300 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/r2k_bench_c4.json 2> gpurun_out/r2k_bench_c4.err; tail -c 300 gpurun_out/r2k_bench_c4.json; tail -3 gpurun_out/r2k_bench_c4.err
timeout 300 python bench.py --workload c3 --no-cpu-baseline > gpurun_out/r2k_bench_c3.json 2> gpurun_out/r2k_bench_c3.err; tail -c 300 gpurun_out/r2k_bench_c3.json
BATCH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2k_launches_c4.csv python profiles/run_full.py c4 > gpurun_out/r2k_c4_run.log 2>&1; tail -3 gpurun_out/r2k_c4_run.log
python profiles/summarize_launches.py gpurun_out/r2k_launches_c4.csv
