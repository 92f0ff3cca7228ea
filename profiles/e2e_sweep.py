"""Sweep the slab size of the host end-to-end entry point (pinned host buffers, PCIe copies inside the
timed region).  Run on the GPU box: python profiles/e2e_sweep.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import xsdba_b200 as xs
T = 10950; n = 16 * 1440
tt = xs.TimeAxis.daily(1981, 30, "noleap"); ts = xs.TimeAxis.daily(2041, 30, "noleap")
bufs = [torch.empty((T, n), dtype=torch.float32).normal_(280, 5).pin_memory() for _ in range(3)]
out = torch.empty((T, n), dtype=torch.float32).pin_memory()
res = {}
for slab in (1024, 2048, 4096, 8192, 11520, 23040):
    ts_ = []
    for i in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        xs.train_adjust_host(bufs[0].numpy(), bufs[1].numpy(), bufs[2].numpy(), time=tt, sim_time=ts, nquantiles=50,
                             group="time.month", kind="+", slab_points=slab, out=out.numpy())
        ts_.append(time.perf_counter() - t0)
    best = min(ts_[1:])
    res[slab] = {"ms": 1e3 * best, "gp_days_per_s": n * T / best, "GBps_pcie": 4 * n * T * 4 / best / 1e9}
print(json.dumps(res, indent=1))
