# round 2: compute-sanitizer memcheck over the kernels that changed this round (profiles/memcheck_r2.py)
mkdir -p gpurun_out
timeout 170 compute-sanitizer --tool memcheck --error-exitcode 3 python profiles/memcheck_r2.py > gpurun_out/r2_memcheck.log 2>&1; echo "rc $?" >> gpurun_out/r2_memcheck.log
tail -12 gpurun_out/r2_memcheck.log
