"""MBCn (5 variables, 20 iterations, group="time") on 1440 points x 30 years: wall time of train and adjust."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import xsdba_b200 as xs

N = 1440
tt = xs.TimeAxis.daily(1981, 30, "noleap")
T = len(tt)
gen = torch.Generator(device="cuda").manual_seed(1)
mk = lambda off: torch.stack([torch.empty((T, N), device="cuda").normal_(280 + off, 5, generator=gen) for _ in range(5)])
ref5, hist5, sim5 = mk(0), mk(1), mk(2)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    obj = xs.MBCn.train(ref5, hist5, time=tt, base_kws={"nquantiles": 20, "group": "time"}, n_iter=20, seed=1)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    out = obj.adjust(sim5, ref5, hist5, time=tt)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"rep {rep}: train {t1 - t0:.3f} s, adjust {t2 - t1:.3f} s")
