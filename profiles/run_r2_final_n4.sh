# round 2 final: the default bench at 4 GPUs (the driver's scaling run takes N = 1, 2, 4, 8)
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 > gpurun_out/r2J_bench_default_n4.json 2> gpurun_out/r2J_bench_default_n4.err; tail -c 300 gpurun_out/r2J_bench_default_n4.json; tail -2 gpurun_out/r2J_bench_default_n4.err
