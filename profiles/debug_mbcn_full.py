import sys
sys.path.insert(0, "."); sys.path.insert(0, "oracle"); sys.path.insert(0, "tests")
import numpy as np, torch
import qm_oracle as o, synth
import xsdba_b200 as xs
rng = np.random.default_rng(51)
to = o.daily_time_axis(1981, 30, "noleap"); tx = xs.TimeAxis.daily(1981, 30, "noleap")
T = len(to); N = 4
def mk(which):
    tas = synth.tas(rng, to, N, which, nan_frac=0)
    d = np.abs(rng.normal(4, 1, size=(2, T, N))).astype(np.float32)
    pr = synth.pr(rng, to, N, which, nan_frac=0)
    hurs = np.clip(100 * rng.beta(5, 2, size=(T, N)), 0, 100).astype(np.float32)
    return np.stack([hurs, pr, tas, tas + d[0], tas - d[1]])
ref, hist, sim = mk("ref"), mk("hist"), mk("sim")
rots = o.rand_rot_matrices(5, 20, 20260117); q = o.equally_spaced_nodes(20)
kinds = ["+", "*", "+", "+", "+"]
blocks = o.mbcn_blocks(to, "time", 1)
tr_ = lambda a: np.ascontiguousarray(a.transpose(0, 2, 1))
afq_o = o.mbcn_train(tr_(ref), tr_(hist), rots, q, blocks)
scen_o = o.mbcn_adjust(tr_(ref), tr_(hist), tr_(sim), afq_o, rots, q, blocks, kinds)
obj = xs.MBCn.train(ref, hist, time=tx, base_kws={"nquantiles": q, "group": "time"}, n_iter=20, rot_matrices=rots)
afq = obj.ds["af_q"].cpu().numpy()
print("af_q bit mismatch", (afq.view(np.int32) != afq_o.view(np.int32)).mean())
scen = tr_(obj.adjust(sim, ref, hist, time=tx, kinds=kinds).cpu().numpy())
neq = scen.view(np.int32) != scen_o.view(np.int32)
print("scen bit mismatch per (var, point):\n", neq.mean(axis=-1))
