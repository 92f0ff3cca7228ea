"""EQM train on the day-of-year x 31 windows of 11 520 points x 30 years, with and without the fused jitter option."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import xsdba_b200 as xs

N = 11520
tt = xs.TimeAxis.daily(1981, 30, "noleap")
T = len(tt)
gen = torch.Generator(device="cuda").manual_seed(1)
def pr():
    x = torch.empty((T, N), device="cuda").exponential_(0.2, generator=gen)
    x[torch.rand((T, N), device="cuda", generator=gen) < 0.4] = 0.0
    return x
ref, hist = pr(), pr()
g = xs.Grouper("time.dayofyear", 31)
def run(**kw):
    return xs.EmpiricalQuantileMapping.train(ref, hist, time=tt, nquantiles=100, group=g, kind="*", **kw)
for name, kw in (("plain", {}), ("jitter_under_thresh", {"jitter_under_thresh_value": "0.01 mm/d"})):
    run(**kw); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(**kw); e1.record(); e1.synchronize()
    print(f"{name}: {e0.elapsed_time(e1):.1f} ms")
