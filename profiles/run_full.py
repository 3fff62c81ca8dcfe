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
elif which == "tol":
    # integrator_adaptive_tolerance on the C3 integrand: every failing region of a round is split at once, leaves sorted into DFS order
    tol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-10
    w = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    rng = Range([0, 0], [1, 1])
    for rep in range(2):
        t0 = tic()
        regs = ctx.regions_generate_tolerance("smooth_edge2", rng, "boole_simpson", "default", "absolute", tol, 1e-5, exact=True)
        t1 = tic()
        bins = torch.zeros(w * w, dtype=torch.float32, device="cuda")
        regs.integrate_bins(bins, [w, w], rng)
        t2 = tic()
        print(f"TOL rep{rep}: tolerance {tol:g}: {len(regs)} leaves in {t1-t0:.4f} s ({len(regs)/(t1-t0)/1e6:.2f} M leaves/s), region->bin ({w}x{w}) {t2-t1:.4f} s, mean {float(bins.mean()):.6f}", flush=True)
        regs.free()
elif which == "fub":
    # integrator_crespo2021_infinite<2> on the C5 integrand (random walk): region table over the first two dimensions from a noisy g
    from viltrum_b200 import integrate, integrator_crespo2021_infinite, RangeInfinite
    it = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
    w = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    mc = int(sys.argv[4]) if len(sys.argv) > 4 else 16
    spp = int(sys.argv[5]) if len(sys.argv) > 5 else 64
    for rep in range(2):
        bins = torch.zeros(w * w, dtype=torch.float32, device="cuda")
        t0 = tic()
        integrate(integrator_crespo2021_infinite(2, it, mc, spp, seed=rep, batch=batch), bins, [w, w], "walk", RangeInfinite(), ctx=ctx)
        t1 = tic()
        print(f"FUB rep{rep}: crespo2021_infinite<2>({it},{mc},{spp}) {w}x{w} bins, batch={batch}: {t1-t0:.4f} s = {w*w*spp/(t1-t0)/1e6:.1f} M residual paths/s, mean {float(bins.mean()):.5f}", flush=True)
