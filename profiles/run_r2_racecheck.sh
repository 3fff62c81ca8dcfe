# round 2: compute-sanitizer racecheck (shared-memory hazards) over the same pass as memcheck_r2.py, after the select_small fix; regions tests first
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_regions.py -m gpu -q -x > gpurun_out/r2_race_tests.log 2>&1; tail -1 gpurun_out/r2_race_tests.log
timeout 120 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/memcheck_r2.py > gpurun_out/r2_racecheck.log 2>&1; echo "rc $?" >> gpurun_out/r2_racecheck.log
grep -c "Race reported" gpurun_out/r2_racecheck.log; tail -4 gpurun_out/r2_racecheck.log
