import sys
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import numpy as np, torch
import qm_oracle as o
import xsdba_b200 as xs
from xsdba_b200 import mbcn as M
rng = np.random.default_rng(3)
for T, N in ((10950, 4), (10950, 1), (3000, 4), (900, 40), (10950, 37)):
    sim = rng.standard_normal((T, N)).astype(np.float32)
    ref = rng.standard_normal((T, N)).astype(np.float32)
    blk = M._Block(T, N, torch.float32)
    got = blk.reorder(torch.from_numpy(sim).cuda(), torch.from_numpy(ref).cuda()).cpu().numpy()
    want = np.stack([o.reordering_1d(sim[:, i], ref[:, i]) for i in range(N)], 1)
    print(T, N, "mismatch per column", (got != want).mean(axis=0)[:8])
