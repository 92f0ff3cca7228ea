import sys, os
sys.path.insert(0, "/root/repo")
import torch, xsdba_b200 as xs
N=11520
tt = xs.TimeAxis.daily(1981, 30, "noleap")
T=len(tt)
gen = torch.Generator(device="cuda").manual_seed(1)
sim = torch.empty((T, N), device="cuda").normal_(280, 5, generator=gen)
for _ in range(2):
    tr = xs.loess_trend(sim, time=tt, f=0.2, niter=1, d=0, kind="+")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); tr = xs.loess_trend(sim, time=tt, f=0.2, niter=1, d=0, kind="+"); e1.record(); e1.synchronize()
print("loess_trend ms", e0.elapsed_time(e1))
