"""Device-resident throughput of the other BASELINE.json configs (parity-test cases, not bench lines) on one
slab, for DESIGN.md.  Run on the GPU box:  python profiles/bench_configs.py [n_lat_rows]"""
import json, os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import xsdba_b200 as xs

warnings.simplefilter("ignore")
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N = rows * 1440
tt = xs.TimeAxis.daily(1981, 30, "noleap"); ts = xs.TimeAxis.daily(2041, 30, "noleap")
T = len(tt)
gen = torch.Generator(device="cuda").manual_seed(1)
def tas(off): return torch.empty((T, N), device="cuda").normal_(280 + off, 5, generator=gen)
def pr():
    x = torch.empty((T, N), device="cuda").exponential_(0.2, generator=gen)
    dry = torch.rand((T, N), device="cuda", generator=gen) < 0.4
    x[dry] = torch.rand(int(dry.sum()), device="cuda", generator=gen) * 0.01 + 1e-6   # pre-jittered dry days
    return x
def timeit(fn, n=2):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n
res = {"points": N, "days": T}
# cfg3: QDM kind='*' pr, Grouper('time.dayofyear', 31), nq=100
ref, hist, sim = pr(), pr(), pr()
g = xs.Grouper("time.dayofyear", 31)
def cfg3(rw):
    obj = xs.QuantileDeltaMapping.train(ref, hist, time=tt, nquantiles=100, group=g, kind="*")
    return obj.adjust(sim, time=ts, interp="nearest", extrapolation="constant", rank_window=rw)
for rw in (False, True):
    ms = timeit(lambda: cfg3(rw), 1)
    res[f"cfg3_qdm_doy31_rank_window_{rw}"] = {"ms": ms, "gp_days_per_s": N * T / ms * 1e3}
# cfg4: DQM tas '+', doy x 31 QM, detrend = PolyDetrend(1) and LoessDetrend(f=0.2, niter=1, d=0, group="time")
ref, hist, sim = tas(0), tas(1.5), tas(3.5)
dq = xs.DetrendedQuantileMapping.train(ref, hist, time=tt, nquantiles=50, group=g, kind="+")
torch.cuda.synchronize()
ms = timeit(lambda: xs.DetrendedQuantileMapping.train(ref, hist, time=tt, nquantiles=50, group=g, kind="+"), 1)
res["cfg4_dqm_train_doy31"] = {"ms": ms, "gp_days_per_s": N * T / ms * 1e3}
ms = timeit(lambda: dq.adjust(sim, time=ts, detrend=1), 1)
res["cfg4_dqm_adjust_poly1"] = {"ms": ms, "gp_days_per_s": N * T / ms * 1e3}
ms = timeit(lambda: dq.adjust(sim, time=ts, detrend=xs.LoessDetrend(group="time", f=0.2, niter=1, d=0)), 1)
res["cfg4_dqm_adjust_loess_f0.2"] = {"ms": ms, "gp_days_per_s": N * T / ms * 1e3}
del ref, hist, sim, dq
# cfg5: MBCn 5 variables, n_iter=20, group="time" (the reference refuses time.month), on fewer points
Nm = min(N, 1440)
mk = lambda off: torch.stack([tas(off)[:, :Nm] for _ in range(5)])
ref5, hist5, sim5 = mk(0), mk(1), mk(2)
t0 = time.perf_counter()
obj = xs.MBCn.train(ref5, hist5, time=tt, base_kws={"nquantiles": 20, "group": "time"}, n_iter=20, seed=1)
torch.cuda.synchronize(); t1 = time.perf_counter()
out = obj.adjust(sim5, ref5, hist5, time=tt)
torch.cuda.synchronize(); t2 = time.perf_counter()
res["cfg5_mbcn_time_5var_20iter"] = {"points": Nm, "train_s": t1 - t0, "adjust_s": t2 - t1,
                                     "gp_days_per_s": Nm * T / (t2 - t0)}
print(json.dumps(res, indent=1))
