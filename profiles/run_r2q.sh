# round 2q: 128-bit shared loads in the tile-major residual kernel, one Philox block per 5-D point, C4 roofline from the kernel's own event time
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cv.py tests/test_gpu_full_size.py -m gpu -q > gpurun_out/r2q_tests.log 2>&1; tail -8 gpurun_out/r2q_tests.log
timeout 600 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/r2q_bench_c4.json 2> gpurun_out/r2q_bench_c4.err; cat gpurun_out/r2q_bench_c4.json; tail -3 gpurun_out/r2q_bench_c4.err
BATCH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2q_launches_c4.csv python profiles/run_full.py c4 > gpurun_out/r2q_c4_run.log 2>&1; tail -3 gpurun_out/r2q_c4_run.log
python profiles/summarize_launches.py gpurun_out/r2q_launches_c4.csv
