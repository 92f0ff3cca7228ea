import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import numpy as np, xsdba_b200 as xs, qm_oracle as o
g = np.load("/root/repo/tests/golden/loess_gaussian.npz")
x, y = g["loess_x"], g["loess_y"]; n = x.size
t = xs.TimeAxis.daily(2001, 1, "noleap")[:n]
series = np.stack([y, y[::-1].copy()], axis=1)
for k in range(3):
    d, f, niter, dx = g[f"case{k}_params"]
    got = xs.loess_trend(series, time=t, f=float(f), niter=int(niter), d=int(d), weights="gaussian").cpu().numpy()
    want = g[f"case{k}_out"]
    bad = np.where(~np.isclose(got[:, 0], want, rtol=1e-9, atol=1e-10, equal_nan=True))[0]
    print(k, d, f, niter, "bad", bad.tolist()[:40], got[bad[:6], 0], want[bad[:6]])
    got_t = xs.loess_trend(series, time=t, f=float(f), niter=int(niter), d=int(d)).cpu().numpy()
    want_t = o.loess_nb(x, y, f=float(f), niter=int(niter), d=int(d), dx=float(dx))
    print("  tricube maxdiff", np.nanmax(np.abs(got_t[:, 0] - want_t)))
    # time coordinate used by the library vs golden x
o_ = np.asarray(t.ordinal, np.float64); xn = (o_ - o_[0]) / (o_[-1] - o_[0])
print("xn vs x maxdiff", np.max(np.abs(xn - x)), "dx", xn[1]-xn[0], x[1]-x[0])
