import sys,time; sys.path.insert(0,".")
import torch
from viltrum_b200 import Context, RangeInfinite
ctx=Context(0)
d=torch.zeros(1<<22,dtype=torch.float32,device="cuda")
for name in ("walk","walk_plain"):
    for rep in range(3):
        ctx.synchronize(); t=time.perf_counter()
        ctx.mc_per_bin_inf(name, d, [2048,2048], RangeInfinite(), 256, rep)
        ctx.synchronize(); dt=time.perf_counter()-t
    print(name, round(dt*1e3,2),"ms", round((1<<22)*256/dt/1e9,1),"G paths/s")
