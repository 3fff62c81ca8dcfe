"""Times the global scatter Monte Carlo (monte_carlo(N), reference monte-carlo.h:39-63) for few bins / many samples."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from viltrum_b200 import Context, Range
ctx = Context(0)
for integ, d, res, n in (("x2y2", 2, [10], 1 << 28), ("x2y2", 2, [32, 32], 1 << 28), ("shade4_64", 4, [64, 64], 1 << 28), ("shade4_64", 4, [1024, 1024], 1 << 28)):
    nb = int(np.prod(res))
    bins = torch.zeros(nb, dtype=torch.float32, device="cuda")
    rng = Range([0.0] * d, [1.0] * d)
    for rep in range(3):
        bins.zero_(); ctx.synchronize(); t0 = time.perf_counter()
        ctx.monte_carlo(integ, bins, res, rng, n, rep)
        ctx.synchronize(); dt = time.perf_counter() - t0
    print(f"monte_carlo({n}) {integ} into {res}: {dt*1e3:.3f} ms = {n/dt/1e9:.1f} G samples/s, mean {float(bins.mean()):.5f}", flush=True)
