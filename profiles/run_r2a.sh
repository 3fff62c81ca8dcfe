# round 2a: xoshiro-stream K1 in the product, new bench.py (c2 + cv record, c3, c5, reference arm), full GPU suite
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r2a_tests.log 2>&1; tail -25 gpurun_out/r2a_tests.log
timeout 600 python bench.py > gpurun_out/r2a_bench_default.json 2> gpurun_out/r2a_bench_default.err; tail -c 3000 gpurun_out/r2a_bench_default.json; tail -5 gpurun_out/r2a_bench_default.err
timeout 300 python bench.py --workload c3 --exact > gpurun_out/r2a_bench_c3.json 2> gpurun_out/r2a_bench_c3.err; tail -c 1500 gpurun_out/r2a_bench_c3.json; tail -5 gpurun_out/r2a_bench_c3.err
timeout 300 python bench.py --workload c5 > gpurun_out/r2a_bench_c5.json 2> gpurun_out/r2a_bench_c5.err; tail -c 1500 gpurun_out/r2a_bench_c5.json; tail -5 gpurun_out/r2a_bench_c5.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_bench_ref_c2.json 2>&1; tail -c 1200 gpurun_out/r2a_bench_ref_c2.json
timeout 300 python bench.py --impl reference --workload c4 --steps 2 --warmup 0 > gpurun_out/r2a_bench_ref_c4.json 2>&1; tail -c 1200 gpurun_out/r2a_bench_ref_c4.json
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; tail -2 gpurun_out/r2a_smoke.log
nproc; free -g | head -2; lscpu | grep "Model name"
