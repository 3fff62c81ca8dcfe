"""Minimal driver for ncu captures: runs the C2 workload (device-resident bins) a few times."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from viltrum_b200 import Context, Range, RangeInfinite, _capi
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ctx = Context(0)
if wl == "c2":
    d = torch.zeros(1 << 20, dtype=torch.float32, device="cuda")
    for i in range(n):
        ctx.mc_per_bin("shade4_64", d, [1024, 1024], Range([0.0] * 4, [1.0] * 4), 64, i)
elif wl == "c5":
    d = torch.zeros(1 << 22, dtype=torch.float32, device="cuda")
    for i in range(n):
        ctx.mc_per_bin_inf("walk", d, [2048, 2048], RangeInfinite(), 256, i)
ctx.synchronize()
print("done", float(d.mean()))
