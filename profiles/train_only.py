"""One slab of the bench workload through train (and optionally adjust) only -- the target of the ncu captures.
    python profiles/train_only.py [lat_rows] [reps] [adjust]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import xsdba_b200 as xs
from xsdba_b200 import _lib

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 48
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
with_adjust = len(sys.argv) > 3 and sys.argv[3] in ("adjust", "fresh")
fresh = len(sys.argv) > 3 and sys.argv[3] == "fresh"   # regenerate the inputs before every repetition, like bench.py
dev = torch.device("cuda", 0)
lib = _lib.load()
tt = xs.TimeAxis.daily(1981, 30, "noleap"); ts = xs.TimeAxis.daily(2041, 30, "noleap")
T = len(tt); n = rows * bench.NLON
g = xs.Grouper("time.month")
ht, hs = g.handle(tt), g.handle(ts, with_window=False)
q = torch.from_numpy(xs.equally_spaced_nodes(50).astype(np.float32)).to(dev)
doy = torch.from_numpy(tt.dayofyear.astype(np.float32)).to(dev)
year = torch.from_numpy((tt.year - tt.year[0]).astype(np.float32)).to(dev)
gen = torch.Generator(device=dev); gen.manual_seed(1)
ref, hist, sim = (bench.synth_slab(torch, gen, T, rows, 300, w, doy, year, dev) for w in ("ref", "hist", "sim"))
af = torch.empty((n, 12, 50), device=dev); hq = torch.empty_like(af); scen = torch.empty_like(sim)
s = torch.cuda.current_stream().cuda_stream
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
for i in range(reps):
    if fresh and i:
        del ref, hist, sim
        ref, hist, sim = (bench.synth_slab(torch, gen, T, rows, 300 + i, w, doy, year, dev) for w in ("ref", "hist", "sim"))
    e0.record()
    _lib.check(lib.xsdba_qm_train_f32(ref.data_ptr(), hist.data_ptr(), n, 1, n, ht.ptr, q.data_ptr(), 50, 43, 0,
                                      af.data_ptr(), hq.data_ptr(), None, s))
    e1.record()
    if with_adjust:
        _lib.check(lib.xsdba_qm_adjust_f32(sim.data_ptr(), n, 1, n, hs.ptr, af.data_ptr(), hq.data_ptr(), 50, 0, 0, 43,
                                           scen.data_ptr(), s))
    e2.record(); e2.synchronize()
    print(f"train {e0.elapsed_time(e1):.3f} ms  adjust {e1.elapsed_time(e2):.3f} ms  ({n} points)")
