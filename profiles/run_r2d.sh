# round 2d: K1 with the self-resetting scheduler only (ticket prefetch reverted, 3 CTAs/SM); Fubini failure hunt
set -x
mkdir -p gpurun_out; out=gpurun_out/k1_mix_r2e.txt; : > $out
for rep in 1 2; do for b in profiles/exp/bin/k1_mix_xoshiro_minb* profiles/exp/bin/k1_mix_p5t0_r12; do $b 30 >> $out; done; done
cat $out
timeout 1200 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2d_tests.log 2>&1; tail -12 gpurun_out/r2d_tests.log
if grep -q "FAILED tests/test_gpu_fubini" gpurun_out/r2d_tests.log; then
  timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_fubini.py -x -q -m gpu -k "poly3-1" > gpurun_out/r2d_sanitizer.log 2>&1; grep -m 30 -E "Invalid|Error|at |by |=========" gpurun_out/r2d_sanitizer.log | head -60
fi
timeout 600 python bench.py > gpurun_out/r2d_bench_default.json 2> gpurun_out/r2d_bench_default.err; tail -c 300 gpurun_out/r2d_bench_default.json; tail -5 gpurun_out/r2d_bench_default.err
