# round 2r: exact greedy kernel with a cooperative heap warp (subtree prefetch in pop, path-parallel pushes, next-parent prefetch) and a one-warp worker side for <= 2-D regions
set -x
mkdir -p gpurun_out
timeout 120 profiles/exp/bin/greedy_phases 200000 > gpurun_out/r2r_greedy_phases.txt 2>&1; cat gpurun_out/r2r_greedy_phases.txt
timeout 120 profiles/exp/bin/greedy_phases 1000000 >> gpurun_out/r2r_greedy_phases.txt 2>&1; tail -2 gpurun_out/r2r_greedy_phases.txt
timeout 900 python -m pytest tests/test_gpu_regions.py tests/test_gpu_f64.py tests/test_gpu_examples.py tests/test_gpu_tolerance.py -m gpu -q -x > gpurun_out/r2r_tests.log 2>&1; tail -12 gpurun_out/r2r_tests.log
