# round 2c: self-resetting scheduler + ticket prefetch + 4 CTAs/SM in K1, NCCL entry points, fixed full-size gates
set -x
mkdir -p gpurun_out; out=gpurun_out/k1_mix_r2d.txt; : > $out
for rep in 1 2; do for b in profiles/exp/bin/k1_mix_xoshiro_minb* profiles/exp/bin/k1_mix_p5t0_r12; do $b 30 >> $out; done; done
cat $out
timeout 1200 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/r2c_tests.log 2>&1; tail -25 gpurun_out/r2c_tests.log
timeout 600 python bench.py > gpurun_out/r2c_bench_default.json 2> gpurun_out/r2c_bench_default.err; tail -c 600 gpurun_out/r2c_bench_default.json; tail -5 gpurun_out/r2c_bench_default.err
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c_smoke.log 2>&1; tail -2 gpurun_out/r2c_smoke.log
