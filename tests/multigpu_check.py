"""Multi-GPU checks of the product's NCCL paths, one process per GPU (not collected by pytest; tests/test_gpu_multi.py launches it under
torchrun when the box has >= 2 GPUs, and profiles/run_r2_multi.sh runs it on the 2- and 8-GPU boxes):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py

  1. vb200_comm_init over a token handed out by torch.distributed;
  2. split-sample monte_carlo (VB200_MC_ALLREDUCE, C1 shape: x^2+y^2 into 10 bins): every rank ends with the SAME bins, equal to the
     single-GPU call over the whole sample range up to the summation order of the partial grids;
  3. vb200_regions_broadcast: the table rank 0 generated arrives bit-identical on every rank;
  4. C4 shape sharded over the ranks (bins slabbed, table replicated or broadcast): the gathered image is bit-identical to one GPU's.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from viltrum_b200 import Context, Range, shard_for_rank
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = Context(local)
    ctx.comm_init_from_torch()
    assert ctx.comm_size == world and ctx.comm_rank == rank
    # -- 2. split-sample allreduce ---------------------------------------------------------------------------------------------
    rng2 = Range([0.0, 0.0], [1.0, 1.0])
    for samples, res in ((8192, [10]), (1 << 22, [10]), (1 << 20, [32, 32]), (1000003, [7])):
        nb = int(np.prod(res))
        whole = np.zeros(nb, np.float32)
        ctx.monte_carlo("x2y2", whole, res, rng2, samples, 5)                       # one GPU, every sample
        part = np.full(nb, 1.0, np.float32)
        ctx.monte_carlo("x2y2", part, res, rng2, samples, 5, allreduce=True)        # '+=' of the total onto bins that held 1.0
        assert np.allclose(part - 1.0, whole, rtol=2e-5, atol=1e-6), (rank, samples, res, np.abs(part - 1.0 - whole).max())
        t = torch.from_numpy(part.copy()).cuda(); lo = t.clone(); hi = t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), "ranks disagree on the reduced bins"
        dev = torch.ones(nb, dtype=torch.float32, device="cuda")                    # device-resident bins
        ctx.monte_carlo("x2y2", dev, res, rng2, samples, 5, allreduce=True); ctx.synchronize()
        assert np.allclose(dev.cpu().numpy(), part, rtol=2e-6, atol=1e-7)        # same samples; the float atomics of the scatter add in a different order every run
    # -- 3. region-table broadcast ---------------------------------------------------------------------------------------------
    rng5 = Range([0.0] * 5, [1.0] * 5)
    mine = ctx.regions_generate_adaptive("shade5_16", rng5, "simpson_trapezoidal", "size", "relative", 3000, 1e-5, batch=0, exact=True)
    got = ctx.regions_broadcast(mine if rank == 0 else None, 0)
    a, b = mine.download(), got.download()
    for k in a:
        assert np.array_equal(a[k], b[k]), f"broadcast table differs in {k} on rank {rank}"
    # -- 4. C4 shape over the ranks ---------------------------------------------------------------------------------------------
    res, spp = [128, 96], 16
    nb = res[0] * res[1]
    one = np.zeros(nb, np.float32)
    mine.cv_integrate("shade5_16", one, res, rng5, spp, 3)
    slab = torch.zeros(nb, dtype=torch.float32, device="cuda")
    got.cv_integrate("shade5_16", slab, res, rng5, spp, 3, shard=shard_for_rank(res, rank, world)); ctx.synchronize()
    dist.all_reduce(slab, op=dist.ReduceOp.SUM)                                      # slabs are disjoint and zero elsewhere: the sum IS the gather
    assert np.array_equal(slab.cpu().numpy().view(np.uint32), one.view(np.uint32)), "sharded C4 image differs from one GPU's"
    if rank != 0:
        got.free()
    mine.free()
    ctx.comm_destroy()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print(f"multigpu_check ok: world {world}, NCCL {ctx._L.vb200_nccl_version()}")


if __name__ == "__main__":
    main()
