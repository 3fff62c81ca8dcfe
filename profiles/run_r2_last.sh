# round 2: the timer events are released with the context; quick check (capi, cv and mc tests, smoke)
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_cv.py tests/test_gpu_mc.py -m gpu -q -k "kernel_timer or golden or replay or shard" > gpurun_out/r2L_tests.log 2>&1; tail -3 gpurun_out/r2L_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
