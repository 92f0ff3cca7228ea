"""Config 3's train step alone (QDM / EQM train, Grouper("time.dayofyear", 31), nq = 100, pre-jittered pr) on 8 lat rows,
optionally followed by the rank_window=True adjust (K3w).
    python profiles/cfg3_train_only.py [lat_rows] [reps] [adjust]"""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import xsdba_b200 as xs
warnings.simplefilter("ignore")
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
tt = xs.TimeAxis.daily(1981, 30, "noleap")
T = len(tt); n = rows * bench.NLON
gen = torch.Generator(device=dev); gen.manual_seed(5); torch.manual_seed(6)
ref, hist = (xs.jitter_under_thresh(bench.synth_pr(torch, gen, T, n, w, dev), "0.01 mm/d", seed=i) for i, w in enumerate(("ref", "hist")))
g = xs.Grouper("time.dayofyear", 31)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(reps):
    e0.record()
    obj = xs.QuantileDeltaMapping.train(ref, hist, time=tt, nquantiles=100, group=g, kind="*")
    e1.record(); e1.synchronize()
    print(f"cfg3 train {e0.elapsed_time(e1):.3f} ms ({n} points)")
    if len(sys.argv) > 3 and sys.argv[3] == "adjust":
        ts = xs.TimeAxis.daily(2041, 30, "noleap")
        e0.record()
        scen = obj.adjust(hist, time=ts, interp="nearest", extrapolation="constant", rank_window=True)
        e1.record(); e1.synchronize()
        print(f"cfg3 adjust rank_window=True {e0.elapsed_time(e1):.3f} ms")
