# round 2i: tile kernel with rejection-sampled regions, batched loads in the accumulate kernel
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cv.py tests/test_gpu_full_size.py -m gpu -q --durations=5 > gpurun_out/r2i_tests.log 2>&1; tail -12 gpurun_out/r2i_tests.log
timeout 300 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/r2i_bench_c4.json 2> gpurun_out/r2i_bench_c4.err; tail -c 400 gpurun_out/r2i_bench_c4.json; tail -5 gpurun_out/r2i_bench_c4.err
BATCH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2i_launches_c4.csv python profiles/run_full.py c4 > gpurun_out/r2i_c4_run.log 2>&1; tail -3 gpurun_out/r2i_c4_run.log
python profiles/summarize_launches.py gpurun_out/r2i_launches_c4.csv
