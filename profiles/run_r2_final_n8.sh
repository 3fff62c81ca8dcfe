# round 2 final: the default bench at 8 GPUs (the driver's scaling run takes N = 1, 2, 4, 8)
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 > gpurun_out/r2J_bench_default_n8.json 2> gpurun_out/r2J_bench_default_n8.err; tail -c 300 gpurun_out/r2J_bench_default_n8.json; tail -2 gpurun_out/r2J_bench_default_n8.err
