import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
import xsdba_b200 as xs
from xsdba_b200 import _lib
rows=48; dev=torch.device("cuda",0); lib=_lib.load()
tt = xs.TimeAxis.daily(1981, 30, "noleap"); ts = xs.TimeAxis.daily(2041, 30, "noleap")
T=len(tt); n=rows*bench.NLON
g=xs.Grouper("time.month"); ht,hs=g.handle(tt),g.handle(ts,with_window=False)
q=torch.from_numpy(xs.equally_spaced_nodes(50).astype(np.float32)).to(dev)
doy=torch.from_numpy(tt.dayofyear.astype(np.float32)).to(dev); year=torch.from_numpy((tt.year-tt.year[0]).astype(np.float32)).to(dev)
gen=torch.Generator(device=dev); gen.manual_seed(1)
ref,hist,sim=(bench.synth_slab(torch,gen,T,rows,300,w,doy,year,dev) for w in ("ref","hist","sim"))
af=torch.empty((n,12,50),device=dev); hq=torch.empty_like(af); scen=torch.empty_like(sim)
s=torch.cuda.current_stream().cuda_stream
e=[torch.cuda.Event(enable_timing=True) for _ in range(3)]
def train(): _lib.check(lib.xsdba_qm_train_f32(ref.data_ptr(),hist.data_ptr(),n,1,n,ht.ptr,q.data_ptr(),50,43,0,af.data_ptr(),hq.data_ptr(),None,s))
def adjust(): _lib.check(lib.xsdba_qm_adjust_f32(sim.data_ptr(),n,1,n,hs.ptr,af.data_ptr(),hq.data_ptr(),50,0,0,43,scen.data_ptr(),s))
train(); adjust(); torch.cuda.synchronize()
for mode in ("back-to-back","sync+sleep","adjust x3"):
    for i in range(3):
        e[0].record(); train(); e[1].record()
        if mode=="sync+sleep": torch.cuda.synchronize(); time.sleep(0.05); e[1].record()
        adjust(); e[2].record()
        if mode=="adjust x3":
            adjust(); e3=torch.cuda.Event(enable_timing=True); e3.record(); adjust(); e4=torch.cuda.Event(enable_timing=True); e4.record()
        torch.cuda.synchronize()
        extra = f" 2nd {e[2].elapsed_time(e3):.3f} 3rd {e3.elapsed_time(e4):.3f}" if mode=="adjust x3" else ""
        print(mode, f"train {e[0].elapsed_time(e[1]):.3f} adjust {e[1].elapsed_time(e[2]):.3f}"+extra)
