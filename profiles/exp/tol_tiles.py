"""experiment: tile list lengths of the tolerance-generated region table (smooth_edge2, 3e-12) at 512x512 bins"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from viltrum_b200 import Context, Range
ctx = Context(0)
rng = Range([0, 0], [1, 1])
regs = ctx.regions_generate_tolerance("smooth_edge2", rng, "boole_simpson", "default", "absolute", 3e-12, 1e-5, exact=True)
t = regs.download()
mn, mx = t["min"], t["max"]
print("regions", len(regs), mn.shape)
w = 512
mn = mn.reshape(-1, 2) if mn.ndim == 1 else mn
mx = mx.reshape(-1, 2) if mx.ndim == 1 else mx
if mn.shape[0] == 2 and mn.shape[1] != 2: mn, mx = mn.T, mx.T
ps = np.floor(w * mn).astype(np.int64); pe = np.maximum(ps + 1, np.minimum(w, (0.99 + w * mx).astype(np.int64)))
ts = ps // 16; te = (pe - 1) // 16
cnt = np.zeros((32, 32), np.int64)
ntile = (te[:, 0] - ts[:, 0] + 1) * (te[:, 1] - ts[:, 1] + 1)
print("tile entries", int(ntile.sum()), "max tiles per region", int(ntile.max()))
for dx in range(int((te[:, 0] - ts[:, 0]).max()) + 1):
    for dy in range(int((te[:, 1] - ts[:, 1]).max()) + 1):
        m = (ts[:, 0] + dx <= te[:, 0]) & (ts[:, 1] + dy <= te[:, 1])
        np.add.at(cnt, (ts[m, 0] + dx, ts[m, 1] + dy), 1)
print("tiles:", (cnt > 0).sum(), "max list", int(cnt.max()), "mean", float(cnt.mean()), "top10", np.sort(cnt.ravel())[-10:])
area = (pe[:, 0] - ps[:, 0]) * (pe[:, 1] - ps[:, 1])
print("bins per region: mean", float(area.mean()), "max", int(area.max()), "pairs", int(area.sum()))
ext = mx - mn
print("smallest extents", ext.min(axis=0), "aspect max", float((ext.max(axis=1) / ext.min(axis=1)).max()))
