set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_cv.py tests/test_gpu_examples.py tests/test_gpu_regions.py -m gpu -q --durations=5 > gpurun_out/r2l_tests.log 2>&1; tail -14 gpurun_out/r2l_tests.log
