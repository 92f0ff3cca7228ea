import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "oracle"); sys.path.insert(0, "tests")
import numpy as np, torch
import qm_oracle as o, synth
import xsdba_b200 as xs
from xsdba_b200 import mbcn as M

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
def neq(a, b): return float((a.view(np.int32) != b.view(np.int32)).mean())
blk = M._Block(T, N, torch.float32)
xd = torch.from_numpy(sim).cuda()
# standardize
sg = blk.standardize(xd).cpu().numpy()
so = np.stack([o._standardize(sim[:, :, i]) for i in range(N)], axis=2)   # [V, T, N]
print("standardize mismatch frac", neq(sg, so))
# rotations
x = torch.from_numpy(so).cuda()
rot = rots[3] @ rots[2].T
yf = blk.rotate(x, rot, fused=True).cpu().numpy(); yu = blk.rotate(x, rot, fused=False).cpu().numpy()
mm = np.stack([rot @ so[:, :, i] for i in range(N)], axis=2)
es = np.stack([np.einsum("ij,j...->i...", rot, so[:, None, :, i])[:, 0] for i in range(N)], axis=2)
print("fused vs matmul", neq(yf, mm), " unfused vs einsum", neq(yu, es), " (fused vs einsum", neq(yf, es), ")")
est = np.stack([np.einsum("ij,j...->i...", rots[-1].T, so[:, None, :, i])[:, 0] for i in range(N)], axis=2)
yt = blk.rotate(x, rots[-1].T, fused=False).cpu().numpy()
print("unfused vs einsum (transposed view)", neq(yt, est))
# one npdft adjust iteration on variable 0: rank_bn lookup + add
afq = rng.normal(0, 0.1, size=(N, 1, 20)).astype(np.float32)
q64 = torch.from_numpy(np.asarray(q, np.float64)).cuda()
h0 = torch.from_numpy(np.ascontiguousarray(es[0])).cuda()
g = blk.add_factor_at_rank(h0, torch.from_numpy(afq).cuda(), q64, "nearest", "constant").cpu().numpy()
want = np.empty_like(g)
for i in range(N):
    a = o.interp_on_quantiles_1d(o.rank_bn(es[0, :, i]), np.asarray(q, np.float64), afq[i, 0].astype(np.float64), "nearest", "constant")
    want[:, i] = (es[0, :, i] + a).astype(np.float32)
print("rank lookup + add mismatch", neq(g, want))
# full npdft adjust of point 0 with random af_q, compare step by step
afq_full = rng.normal(0, 0.05, size=(20, 5, 20)).astype(np.float32)
so0 = so[:, :, 0].copy()
ora = so0.copy()[:, None, :]
gx = torch.from_numpy(np.ascontiguousarray(so[:, :, :1])).cuda()
blk1 = M._Block(T, 1, torch.float32)
for ii in range(20):
    rot = rots[ii] if ii == 0 else rots[ii] @ rots[ii - 1].T
    ora = np.einsum("ij,j...->i...", rot, ora)
    gx = blk1.rotate(gx, rot, fused=False)
    m0 = neq(gx.cpu().numpy()[:, :, 0], ora[:, 0])
    for iv in range(5):
        a = o.interp_on_quantiles_1d(o.rank_bn(ora[iv, 0]), np.asarray(q, np.float64), afq_full[ii, iv].astype(np.float64), "nearest", "constant")
        ora[iv, 0] = ora[iv, 0] + a
        gx[iv] = blk1.add_factor_at_rank(gx[iv], torch.from_numpy(afq_full[ii, iv].reshape(1, 1, -1)).cuda(), q64, "nearest", "constant")
    m1 = neq(gx.cpu().numpy()[:, :, 0], ora[:, 0])
    print(ii, "after rotate", m0, "after lookups", m1)

# ---- train: af_q per iteration against the oracle -----------------------------------------------------------------
tr_ = lambda a: np.ascontiguousarray(a.transpose(0, 2, 1))
blocks = o.mbcn_blocks(to, "time", 1)
afq_o = o.mbcn_train(tr_(ref), tr_(hist), rots, q, blocks)            # (1, N, n_iter, V, nq)
afq_g = xs.mbcn_train(ref, hist, time=tx, rot_matrices=rots, quantiles=q, group="time").cpu().numpy()
bad = afq_o.view(np.int32) != afq_g.view(np.int32)
print("af_q mismatch frac", bad.mean(), "per iteration:", bad.mean(axis=(0, 1, 3, 4)))
ix = np.argwhere(bad)
print("first mismatches (blk, pt, iter, var, node):", ix[:8].tolist())
for b_, p_, i_, v_, k_ in ix[:5]:
    print(afq_o[b_, p_, i_, v_, k_], afq_g[b_, p_, i_, v_, k_])

# ---- adjust: univariate QDM, N-pdf block and reorder against the oracle, point by point ---------------------------
from xsdba_b200 import _adjustment as L4
kinds = ["+", "*", "+", "+", "+"]
dt = torch.float32
refd, histd, simd = (torch.from_numpy(a).cuda() for a in (ref, hist, sim))
q_dt = np.asarray(q).astype(np.float32)
blkN = M._Block(T, N, dt)
fake_time = M._BlockTime(T)
tgrp = xs.Grouper("time")
scen_block = torch.empty_like(simd)
for v in range(5):
    tr = L4.eqm_train(L4.Dataset({"ref": refd[v], "hist": histd[v]}, time=fake_time), group=tgrp, kind=kinds[v], quantiles=q_dt)
    out = L4.qdm_adjust(L4.Dataset({"sim": simd[v], "af": tr["af"], "quantiles": tr["quantiles"]}, time=fake_time),
                        group=tgrp, interp="nearest", extrapolation="constant", kind=kinds[v])
    scen_block[v] = out["scen"]
x = blkN.standardize(simd)
afq_d = torch.from_numpy(afq_o).cuda()
for ii in range(20):
    x = blkN.rotate(x, M._iter_rot(rots, ii), fused=False)
    for iv in range(5):
        x[iv] = blkN.add_factor_at_rank(x[iv], afq_d[0, :, ii, iv, :].reshape(N, 1, -1).contiguous(), q64, "nearest", "constant")
x = blkN.rotate(x, rots[-1].T, fused=False)
xg = x.cpu().numpy(); sbg = scen_block.cpu().numpy()
for i in range(N):
    sb_o = np.empty((5, T), np.float32)
    for v in range(5):
        r_, h_, s_ = ref[v, :, i][None], hist[v, :, i][None], sim[v, :, i][None]
        af, _ = o.eqm_train(r_, h_, np.zeros(T, np.int32), 1, 1, q_dt, kinds[v])
        sq = o.rank_pct(s_)
        afi = o.interp_on_quantiles_1d(sq[0], q_dt, af[0, 0], "nearest", "constant")
        sb_o[v] = o.apply_correction(s_[0], afi.astype(np.float32), kinds[v])
    nb_o = o.npdft_adjust(o._standardize(sim[:, :, i]), afq_o[0, i].astype(np.float64), rots, q)
    print("pt", i, "scen_block mismatch", [neq(sbg[v, :, i], sb_o[v]) for v in range(5)],
          "npdft mismatch", [neq(xg[v, :, i], nb_o[v].astype(np.float32)) for v in range(5)])
    for v in range(5):
        want = o.reordering_1d(sb_o[v], nb_o[v])
        got = blkN.reorder(torch.from_numpy(np.ascontiguousarray(np.stack([sb_o[v]] * N, 1))).cuda(),
                           torch.from_numpy(np.ascontiguousarray(np.stack([nb_o[v].astype(np.float32)] * N, 1))).cuda()).cpu().numpy()[:, 0]
        print("   reorder var", v, "mismatch", neq(got, want), "npdft dtype", nb_o.dtype, "ties in npdft", T - np.unique(nb_o[v]).size)
