set -x
mkdir -p gpurun_out
profiles/exp/bin/greedy_phases 200000 > gpurun_out/r2o_greedy_phases.txt 2>&1; cat gpurun_out/r2o_greedy_phases.txt
profiles/exp/bin/greedy_phases 1000000 >> gpurun_out/r2o_greedy_phases.txt 2>&1; tail -3 gpurun_out/r2o_greedy_phases.txt
timeout 900 python -m pytest tests/test_gpu_regions.py tests/test_gpu_f64.py tests/test_gpu_examples.py -m gpu -q -x > gpurun_out/r2o_tests.log 2>&1; tail -8 gpurun_out/r2o_tests.log
timeout 300 python bench.py --workload c3 --exact --steps 3 --warmup 1 > gpurun_out/r2o_c3_exact.json 2> gpurun_out/r2o_c3_exact.err; cat gpurun_out/r2o_c3_exact.json; tail -3 gpurun_out/r2o_c3_exact.err
