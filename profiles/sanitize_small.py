"""Small end-to-end run of the round-2 kernels for compute-sanitizer (memcheck / racecheck / synccheck):
K1b (+ sorter fallback columns), K2p/K2t/K2x, K1w, K3w on 96 points x 6 years."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np, torch
import xsdba_b200 as xs
import qm_oracle as o, synth
rng = np.random.default_rng(5)
years = 30
to = o.daily_time_axis(1981, years, "noleap"); tx = xs.TimeAxis.daily(1981, years, "noleap")
P = 70
ref = synth.tas(rng, to, P, "ref", np.float32); hist = synth.tas(rng, to, P, "hist", np.float32); sim = synth.tas(rng, to, P, "sim", np.float32)
hist[:, 1] = 281.0; ref[:, 2] = np.round(ref[:, 2]); hist[::7, 3] = np.inf; hist[:, 7] = np.nan; sim[::50, 9] = np.nan
q = o.equally_spaced_nodes(50).astype(np.float32)
g = xs.Grouper("time.month")
tr = xs.eqm_train(xs.Dataset({"ref": ref, "hist": hist}, time=tx), group=g, kind="+", quantiles=q)
out = xs.qm_adjust(xs.Dataset({"sim": sim, "af": tr.af, "hist_q": tr.hist_q}, time=tx), group=g, interp="nearest", extrapolation="constant", kind="+")
torch.cuda.synchronize()
gw = xs.Grouper("time.dayofyear", 31)
pr = [synth.pr(rng, to, 10, w) for w in ("ref", "hist", "sim")]
trw = xs.eqm_train(xs.Dataset({"ref": pr[0], "hist": pr[1]}, time=tx), group=gw, kind="*", quantiles=o.equally_spaced_nodes(100).astype(np.float32))
outw = xs.qdm_adjust(xs.Dataset({"sim": pr[2], "af": trw.af, "quantiles": o.equally_spaced_nodes(100).astype(np.float32)}, time=tx), group=gw,
                     interp="nearest", extrapolation="constant", kind="*", rank_window=True)
torch.cuda.synchronize()
print("ok", float(torch.nansum(out.scen)), float(torch.nansum(outw.scen)))
