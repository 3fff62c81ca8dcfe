set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2p_tests.log 2>&1; tail -15 gpurun_out/r2p_tests.log
timeout 600 python bench.py > gpurun_out/r2p_bench_default.json 2> gpurun_out/r2p_bench_default.err; cat gpurun_out/r2p_bench_default.json
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p_smoke.log 2>&1; tail -3 gpurun_out/r2p_smoke.log
