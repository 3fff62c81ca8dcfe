"""-m gpu: the product's NCCL paths (split-sample allreduce of vb200_monte_carlo, region-table broadcast, C4 sharded over ranks) under
torchrun, one process per GPU — needs >= 2 GPUs on the box (skipped on the single-GPU test box; profiles/run_r2_multi.sh runs the same
script on the 2- and 8-GPU boxes).  The CPU-side logic of the N > 1 path is covered by tests/test_distributed_cpu.py (gloo)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_nccl_paths_under_torchrun():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 2 if n < 4 else 4
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(ROOT, "tests", "multigpu_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "multigpu_check ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_comm_entry_points_without_a_communicator(ctx=None):
    """single GPU: the collective entry points fail cleanly without a communicator; NCCL binds at run time"""
    import numpy as np
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from viltrum_b200 import Context, Range, Vb200Error
    c = Context(0)
    assert c.comm_size == 1 and c.comm_rank == 0
    with pytest.raises(Vb200Error):
        c.monte_carlo("x2y2", np.zeros(10, np.float32), [10], Range([0, 0], [1, 1]), 1000, 1, allreduce=True)
    with pytest.raises(Vb200Error):
        c.regions_broadcast(None, 0)
    token = c.comm_unique_id()
    assert len(token) == 128 and c._L.vb200_nccl_version() >= 20000
    c.comm_init(token, 0, 1)                      # a world of one: allreduce and broadcast are identities
    b = np.zeros(10, np.float32); w = np.zeros(10, np.float32)
    c.monte_carlo("x2y2", b, [10], Range([0, 0], [1, 1]), 8192, 0, allreduce=True)
    c.monte_carlo("x2y2", w, [10], Range([0, 0], [1, 1]), 8192, 0)
    assert np.allclose(b, w, rtol=1e-6)          # same samples; the order of the float atomics differs run to run
    c.comm_destroy()
    c.close()
