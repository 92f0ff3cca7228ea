"""What does HBM give the (point tile x group) row-piece access pattern of the adjust/train kernels?
Copies a time-major (T, N) float32 array group by group with 128 / 256 / 512-byte row pieces and
compares with a flat torch copy.  Run on the GPU box:  python profiles/microbench_rows.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import xsdba_b200 as xs
from xsdba_b200 import _lib

lib = _lib.load()
T, N = 10950, 48 * 1440
t = xs.TimeAxis.daily(1981, 30, "noleap")
x = torch.randn((T, N), device="cuda")
y = torch.empty_like(x)
res = {}
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n
ms = timeit(lambda: y.copy_(x)); res["torch_copy_GBps"] = 2 * x.numel() * 4 / ms / 1e6
for group in ("time.month", "time.dayofyear"):
    h = xs.Grouper(group).handle(t)
    for v in (1, 2, 4):
        ms = timeit(lambda: _lib.check(lib.xsdba_debug_copy_rows_f32(x.data_ptr(), N, N, h.ptr, y.data_ptr(), v, torch.cuda.current_stream().cuda_stream)))
        res[f"{group}_v{v}_{128*v}B_GBps"] = 2 * x.numel() * 4 / ms / 1e6
        assert torch.equal(x, y)
print(json.dumps(res, indent=1))
