"""CPU, world_size 2 over gloo: the host-side logic of the N>1 path — bin-grid slabs per rank, slab gather,
max-over-ranks timing reduction, and the split-sample mode's allreduce (sample counters keyed globally, partial grids summed).
The per-rank 'compute' is a host emulation of the scatter sampler built on the library's own Philox routine, so no GPU is used."""
import ctypes
import os
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def scatter_emulation(seed, s0, s1, res):
    """host twin of mc_scatter_kernel for f = x^2+y^2 on [0,1]^2 (include/viltrum_b200/device/mc_scatter.cuh): counter = (sample, 0xffffffff, block)"""
    from viltrum_b200 import _capi
    L = _capi.lib()
    out = np.zeros(res, np.float64)
    c = (ctypes.c_uint32 * 4)(); k = (ctypes.c_uint32 * 2)(seed & 0xffffffff, seed >> 32); o = (ctypes.c_uint32 * 4)()
    for s in range(s0, s1):
        c[0], c[1], c[2], c[3] = s & 0xffffffff, s >> 32, 0xffffffff, 0
        L.vb200_philox4x32_10(c, k, o)
        x, y = (o[0] >> 8) * 2.0 ** -24, (o[1] >> 8) * 2.0 ** -24
        out[min(int(res * x), res - 1)] += x * x + y * y
    return out


def worker(rank, world, port, res, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, ROOT)
        from viltrum_b200 import shard_for_rank, sample_shard_for_rank
        nb = int(np.prod(res))
        # (1) slabs: contiguous, disjoint, whole rows of the last dimension, cover the grid
        b, e = shard_for_rank(res, rank, world)
        row = nb // res[-1]
        assert b % row == 0 and e % row == 0
        spans = [None] * world
        dist.all_gather_object(spans, (b, e))
        assert spans[0][0] == 0 and spans[-1][1] == nb and all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        # (2) slab gather: every rank fills its slab, the concatenation is the full grid
        full = torch.arange(nb, dtype=torch.float32) * 0.5 + 1.0
        mine = torch.zeros(nb, dtype=torch.float32); mine[b:e] = full[b:e]
        dist.all_reduce(mine, op=dist.ReduceOp.SUM)
        assert torch.equal(mine, full)
        # (3) timing reduction used by bench.py: max over ranks
        t = torch.tensor([1.0 + rank], dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert float(t) == float(world)
        # (4) split-sample mode: partial grids from disjoint sample ranges sum to the single-process estimate
        n, seed, bins = 4000, 0x1234ABCD5678, 10
        s0, s1 = sample_shard_for_rank(n, rank, world)
        part = torch.from_numpy(scatter_emulation(seed, s0, s1, bins))
        dist.all_reduce(part, op=dist.ReduceOp.SUM)
        if rank == 0:
            whole = scatter_emulation(seed, 0, n, bins)
            assert np.allclose(part.numpy(), whole, rtol=1e-12)
            est = whole * bins / n
            analytic = np.array([(3 * k * k + 3 * k + 1) / 300 + 1 / 3 for k in range(bins)])
            assert np.max(np.abs(est - analytic)) < 0.2
        q.put((rank, "ok"))
    except Exception as ex:      # surface the failure in the parent
        q.put((rank, repr(ex)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("res", [[1024, 1024], [37, 5], [100], [8, 6, 3]])
def test_world_size_two_gloo(res):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() + len(res) * 7 + res[0]) % 300
    procs = [ctx.Process(target=worker, args=(r, 2, port, res, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results


def test_shards_edge_cases():
    from viltrum_b200 import shard_for_rank
    assert shard_for_rank([10, 3], 0, 8) == (0, 0) or shard_for_rank([10, 3], 0, 8)[0] == 0        # fewer rows than ranks: some slabs are empty
    spans = [shard_for_rank([10, 3], r, 8) for r in range(8)]
    assert spans[-1][1] == 30 and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert sum(e - b for b, e in spans) == 30
