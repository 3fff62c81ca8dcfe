# round 2b: K1 pipeline / chain variants (harness), GPU suite with the fixed generator mapping, default bench, c5 reference leg
set -x
mkdir -p gpurun_out; out=gpurun_out/k1_mix_r2c.txt; : > $out
for rep in 1 2; do for b in profiles/exp/bin/k1_mix_*; do $b 30 >> $out; done; done
cat $out
timeout 900 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/r2b_tests.log 2>&1; tail -25 gpurun_out/r2b_tests.log
timeout 600 python bench.py > gpurun_out/r2b_bench_default.json 2> gpurun_out/r2b_bench_default.err; tail -c 600 gpurun_out/r2b_bench_default.json; tail -5 gpurun_out/r2b_bench_default.err
timeout 300 python bench.py --workload c5 > gpurun_out/r2b_bench_c5.json 2> gpurun_out/r2b_bench_c5.err; tail -c 900 gpurun_out/r2b_bench_c5.json; tail -5 gpurun_out/r2b_bench_c5.err
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.log 2>&1; tail -2 gpurun_out/r2b_smoke.log
