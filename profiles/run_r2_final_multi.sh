# round 2 final: multi-GPU runs (one process per GPU under torchrun): NCCL paths of the product, C2 weak + strong, C4 replicated + broadcast, C3, C5
N=${1:-2}
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/multigpu_check.py > gpurun_out/r2G_check_n$N.log 2>&1; tail -3 gpurun_out/r2G_check_n$N.log
timeout 300 $TR --master-port 29512 profiles/time_allreduce.py > gpurun_out/r2G_allreduce_n$N.log 2>&1; grep monte_carlo gpurun_out/r2G_allreduce_n$N.log
timeout 600 $TR --master-port 29513 bench.py --gpus $N > gpurun_out/r2G_bench_default_n$N.json 2> gpurun_out/r2G_bench_default_n$N.err; tail -c 300 gpurun_out/r2G_bench_default_n$N.json; tail -3 gpurun_out/r2G_bench_default_n$N.err
timeout 300 $TR --master-port 29514 bench.py --gpus $N --workload c2 --scaling strong > gpurun_out/r2G_bench_c2_strong_n$N.json 2> gpurun_out/r2G_bench_c2_strong_n$N.err; tail -c 200 gpurun_out/r2G_bench_c2_strong_n$N.json
timeout 300 $TR --master-port 29515 bench.py --gpus $N --workload c4 --table broadcast > gpurun_out/r2G_bench_c4_bcast_n$N.json 2> gpurun_out/r2G_bench_c4_bcast_n$N.err; tail -c 200 gpurun_out/r2G_bench_c4_bcast_n$N.json; tail -3 gpurun_out/r2G_bench_c4_bcast_n$N.err
timeout 300 $TR --master-port 29516 bench.py --gpus $N --workload c3 > gpurun_out/r2G_bench_c3_n$N.json 2> gpurun_out/r2G_bench_c3_n$N.err; tail -c 200 gpurun_out/r2G_bench_c3_n$N.json
timeout 300 $TR --master-port 29517 bench.py --gpus $N --workload c5 > gpurun_out/r2G_bench_c5_n$N.json 2> gpurun_out/r2G_bench_c5_n$N.err; tail -c 200 gpurun_out/r2G_bench_c5_n$N.json
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/r2G_bench_*_n$N.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'value %.4g %s' % (d['value'], d['unit']), 'ms %.4f' % d['ms_per_step'], 'e2e %.4g' % d['e2e']['value'], d['scaling'], ('cv %.4g ms %.3f' % (d['cv']['value'], d['cv']['ms_per_step'])) if 'cv' in d else '')
    except Exception as e:
        print(f, 'FAILED', e)
PY
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r2G_tests_n$N.log 2>&1; tail -4 gpurun_out/r2G_tests_n$N.log
