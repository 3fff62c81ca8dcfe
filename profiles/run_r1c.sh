set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt
python bench.py > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err; tail -c 600 gpurun_out/r2_bench_c2.json
python bench.py --workload c2k16 --no-cpu-baseline > gpurun_out/r2_bench_c2k16.json 2>&1
python bench.py --workload c5 --no-cpu-baseline > gpurun_out/r2_bench_c5.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_c2.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mc_per_bin_kernel -s 2 -c 1 -o gpurun_out/r2_c2_k1 -f python profiles/run_c2.py c2 4 > gpurun_out/r2_ncu_full.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r2_tests.log 2>&1; tail -25 gpurun_out/r2_tests.log
