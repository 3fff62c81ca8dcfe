"""Small pass over the kernels that changed in round 2, meant to run under compute-sanitizer --tool memcheck:
per-bin sampler (both generators), batched refinement with both split kernels and both selection paths, exact greedy (float / mixed keys),
tile lists + sort, separable fast walk, tile-major residual pass, importance sampling, tolerance-driven refinement."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from viltrum_b200 import Context, Range, _capi
ctx = Context(0)
r4, r5, r2 = Range([0.0] * 4, [1.0] * 4), Range([0.0] * 5, [1.0] * 5), Range([0.0] * 2, [1.0] * 2)
b = np.zeros(96 * 80, np.float32)
for gen in ("xoshiro", "philox"):
    ctx.mc_per_bin("shade4_64", b, [96, 80], r4, 16, 3, _capi.MC_PER_BIN, generator=gen)
print("mc", float(b.mean()))
for knob in (None, "0"):
    if knob is None: os.environ.pop("VB200_SPLIT_CTA_MAX", None); os.environ.pop("VB200_SELECT_SMALL_MAX", None)
    else: os.environ["VB200_SPLIT_CTA_MAX"] = knob; os.environ["VB200_SELECT_SMALL_MAX"] = "64"
    regs = ctx.regions_generate_adaptive("shade5_16", r5, "simpson_trapezoidal", "size", "relative", 3000, 1e-5, batch=0, exact=True)
    bins = np.zeros(48 * 40, np.float32)
    regs.cv_integrate("shade5_16", bins, [48, 40], r5, 16, 5)
    regs.cv_integrate("shade5_16", bins, [48, 40], r5, 8, 6, rs="importance")
    print("cv", knob, float(bins.mean()))
    regs.free()
os.environ.pop("VB200_SPLIT_CTA_MAX", None); os.environ.pop("VB200_SELECT_SMALL_MAX", None)
regs = ctx.regions_generate_adaptive("smooth_edge2", r2, "boole_simpson", "size", "relative", 20000, 1e-5, batch=0, exact=True)
bins = np.zeros(64 * 64, np.float32); regs.integrate_bins(bins, [64, 64], r2); print("nc batched", float(bins.mean())); regs.free()
regs = ctx.regions_generate_adaptive("smooth_edge2", r2, "boole_simpson", "size", "relative", 3000, 1e-5, batch=1, exact=True)
bins = np.zeros(64 * 64, np.float32); regs.integrate_bins(bins, [64, 64], r2); print("nc greedy", float(bins.mean())); regs.free()
ctx.close()
print("memcheck pass done")
