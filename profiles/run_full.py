"""Times the region-based configs at BASELINE sizes (C3, C4) on the GPU; prints phase times (wall clock with synchronisation:
these are multi-kernel pipelines, not single launches)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from viltrum_b200 import Context, Range
which = sys.argv[1] if len(sys.argv) > 1 else "c3"
batch = int(os.environ.get("BATCH", "1"))
ctx = Context(0)
def tic(): ctx.synchronize(); return time.perf_counter()
if which == "c3":
    it = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
    w = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    rng = Range([0, 0], [1, 1])
    for rep in range(2):
        t0 = tic()
        regs = ctx.regions_generate_adaptive("smooth_edge2", rng, "boole_simpson", "size", "relative", it, 1e-5, batch=batch, exact=True)
        t1 = tic()
        bins = torch.zeros(w * w, dtype=torch.float32, device="cuda")
        regs.integrate_bins(bins, [w, w], rng)
        t2 = tic()
        print(f"C3 rep{rep}: {it} iterations: generation {t1-t0:.3f} s ({(t1-t0)/it*1e6:.2f} us/iter, {it/(t1-t0)/1e6:.3f} M regions/s), region->bin ({w}x{w}) {t2-t1:.4f} s, mean {float(bins.mean()):.6f}", flush=True)
        regs.free()
elif which == "c4":
    it = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
    w = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    spp = int(sys.argv[4]) if len(sys.argv) > 4 else 64
    rng = Range([0] * 5, [1] * 5)
    for rep in range(2):
        t0 = tic()
        regs = ctx.regions_generate_adaptive("shade5_64", rng, "simpson_trapezoidal", "size", "relative", it, 1e-5, batch=batch, exact=True)
        t1 = tic()
        bins = torch.zeros(w * w, dtype=torch.float32, device="cuda")
        nreg = torch.zeros(w * w, dtype=torch.int32, device="cuda")
        regs.cv_integrate("shade5_64", bins, [w, w], rng, spp, rep, nregions=nreg)
        t2 = tic()
        ev = w * w * spp
        print(f"C4 rep{rep}: generation {it} iterations {t1-t0:.3f} s ({(t1-t0)/it*1e6:.2f} us/iter); CV+residual {w}x{w}x{spp}spp {t2-t1:.3f} s = {ev/(t2-t1)/1e6:.1f} M evals/s; "
              f"total {ev/(t2-t0)/1e6:.1f} M evals/s; regions/bin {float(nreg.float().mean()):.1f} pairs {float(nreg.double().sum()):.3e}; mean {float(bins.mean()):.5f}", flush=True)
        regs.free()
