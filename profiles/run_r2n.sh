set -x
mkdir -p gpurun_out
profiles/exp/bin/greedy_phases 200000 > gpurun_out/r2n_greedy_phases.txt 2>&1; cat gpurun_out/r2n_greedy_phases.txt
profiles/exp/bin/greedy_phases 1000000 >> gpurun_out/r2n_greedy_phases.txt 2>&1; tail -2 gpurun_out/r2n_greedy_phases.txt
timeout 600 python -m pytest tests/test_gpu_cv.py -m gpu -q -k "sampling or optimized" > gpurun_out/r2n_tests.log 2>&1; tail -5 gpurun_out/r2n_tests.log
