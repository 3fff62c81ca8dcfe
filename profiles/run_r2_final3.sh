# round 2 final, one GPU: the whole GPU suite, smoke, every bench line (default = C2 + C4 sub-record, C3 with the exact mode, C5) and the reference arm
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2I_tests.log 2>&1; tail -6 gpurun_out/r2I_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2I_smoke.log 2>&1; tail -2 gpurun_out/r2I_smoke.log
timeout 600 python bench.py > gpurun_out/r2I_bench_default.json 2> gpurun_out/r2I_bench_default.err; tail -c 600 gpurun_out/r2I_bench_default.json
timeout 600 python bench.py --workload c4 > gpurun_out/r2I_bench_c4.json 2> gpurun_out/r2I_bench_c4.err; tail -c 300 gpurun_out/r2I_bench_c4.json
timeout 600 python bench.py --workload c3 --exact > gpurun_out/r2I_bench_c3.json 2> gpurun_out/r2I_bench_c3.err; tail -c 300 gpurun_out/r2I_bench_c3.json
timeout 600 python bench.py --workload c5 > gpurun_out/r2I_bench_c5.json 2> gpurun_out/r2I_bench_c5.err; tail -c 300 gpurun_out/r2I_bench_c5.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2I_bench_ref_c2.json 2> gpurun_out/r2I_bench_ref_c2.err; tail -c 300 gpurun_out/r2I_bench_ref_c2.json
BATCH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2I_launches_c4.csv python profiles/run_full.py c4 > gpurun_out/r2I_c4_run.log 2>&1
python profiles/summarize_launches.py gpurun_out/r2I_launches_c4.csv
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2I_launches_default.csv python bench.py --steps 2 --warmup 1 --sustain 0 --no-cpu-baseline > gpurun_out/r2I_default_under_ncu.log 2>&1
