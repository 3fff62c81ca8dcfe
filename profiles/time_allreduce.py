"""Split-sample monte_carlo (VB200_MC_ALLREDUCE) on the C1 shape under torchrun: N ranks each draw 1/N of the samples, ncclAllReduce sums the
partial grids.  Prints (rank 0) the time per call next to the single-GPU call over all samples.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 profiles/time_allreduce.py"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from viltrum_b200 import Context, Range
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = Context(local)
if world > 1:
    ctx.comm_init_from_torch()
rng = Range([0.0, 0.0], [1.0, 1.0])
for samples, res in ((8192, [10]), (1 << 24, [10]), (1 << 28, [10]), (1 << 28, [32, 32])):
    nb = int(np.prod(res))
    dev = torch.zeros(nb, dtype=torch.float32, device="cuda")
    def run(allreduce):
        ctx.monte_carlo("x2y2", dev, res, rng, samples, 7, allreduce=allreduce)
    out = {}
    for mode in ((False, True) if world > 1 else (False,)):
        for _ in range(3): run(mode)
        ctx.synchronize()
        if world > 1: dist.barrier()
        t0 = time.perf_counter()
        for _ in range(20): run(mode)
        ctx.synchronize()
        if world > 1: dist.barrier()
        out[mode] = (time.perf_counter() - t0) / 20
    if rank == 0:
        line = f"monte_carlo x2y2 {samples} samples into {res}: one GPU {out[False]*1e6:9.1f} us ({samples/out[False]/1e9:7.2f} G samples/s)"
        if world > 1: line += f" | {world} GPUs + ncclAllReduce {out[True]*1e6:9.1f} us ({samples/out[True]/1e9:7.2f} G samples/s, x{out[False]/out[True]:.2f})"
        print(line, flush=True)
if world > 1:
    ctx.comm_destroy(); dist.destroy_process_group()
