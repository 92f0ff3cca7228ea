"""MBCn (multivariate bias correction, N-pdf transform) on the GPU: host mirror of
``xsdba.adjustment.MBCn`` (adjustment.py:1718-1973) / ``mbcn_train`` / ``mbcn_adjust``
(_adjustment.py:331-591).

The iteration loop (n_iter random rotations x variables) is host code; every array operation inside it
is a CUDA kernel of the library: rotation (``xsdba_rotate``), standardisation (``xsdba_standardize``),
quantiles + factors (``xsdba_qm_train_q64``), rank + factor lookup + add (``xsdba_rank_lookup`` with the
``_rank_bn`` normalisation), the per-variable QDM (``eqm_train`` / ``qdm_adjust``) and the final
Schaake shuffle (``xsdba_reorder``).  Arrays are ``(multivar, time, *points)``.

As in the reference, ``group`` is "time" or ``Grouper("time.dayofyear", window)``; monthly grouping
raises NotImplementedError (adjustment.py:1851-1852).  Time blocks follow ``grouped_time_indexes``
(processing.py:829-918): each block is gathered to a dense ``(V, T_block, N)`` array (window slots
outside the series dropped) and treated as one "time" group.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from . import _adjustment as L4
from .base import Grouper, grouping_handle, parse_group
from .utils import equally_spaced_nodes


def rand_rot_matrix(n_var: int, num: int = 1, seed=None) -> np.ndarray:
    """``utils.rand_rot_matrix`` (utils.py:924-974; Mezzadri 2007): float32 [num, N, N]."""
    rng = np.random.default_rng(seed)
    out = np.empty((num, n_var, n_var), np.float32)
    for i in range(num):
        Z = rng.standard_normal((n_var, n_var))
        Q, R = np.linalg.qr(Z)
        d = np.diag(R)
        out[i] = (Q @ np.diag(d / np.abs(d))).astype(np.float32)
    return out


def grouped_time_indexes(time, group: Grouper):
    """processing.py:829-918 for "time" and "time.dayofyear": list of (windowed indices, exact indices)."""
    T = len(time)
    if group.prop == "group":
        return [(np.arange(T), np.arange(T))]
    if group.prop != "dayofyear":
        raise NotImplementedError(f"Grouping {group.name} not implemented.")  # processing.py:913
    gidx = group.zero_based_index(time)
    half = group.window // 2
    out = []
    for g in range(group.n_groups(time)):
        sel = np.nonzero(gidx == g)[0]
        if sel.size == 0:
            continue
        w = (sel[:, None] - half + np.arange(group.window)[None, :]).ravel()
        out.append((w[(w >= 0) & (w < T)], sel))
    return out


def _sfx(dt):
    return "f32" if dt == torch.float32 else "f64"


class _Block:
    """Kernel plumbing for one dense (V, Tb, N) block treated as a single "time" group."""

    def __init__(self, n_time: int, n_pts: int, dt):
        self.lib = _lib.load()
        self.h = grouping_handle(np.zeros(n_time, np.int32), 1, 1)
        self.Tb, self.N, self.dt = n_time, n_pts, dt
        self.stream = torch.cuda.current_stream().cuda_stream

    def standardize(self, x):
        V = x.shape[0]
        y = torch.empty_like(x)
        fn = getattr(self.lib, f"xsdba_standardize_{_sfx(self.dt)}")
        _lib.check(fn(x.data_ptr(), self.N, 1, self.N, self.Tb, V, self.Tb * self.N, y.data_ptr(), self.stream), "standardize")
        return y

    def rotate(self, x, rot: np.ndarray, fused: bool = True):
        """``rot @ x`` (fused multiply-add chain, _adjustment.py:311) or ``einsum("ij,j...->i...", rot, x)`` (separately
        rounded products and sums, _adjustment.py:449, 462)."""
        V = x.shape[0]
        y = torch.empty_like(x)
        rot = np.ascontiguousarray(rot, np.float32)
        fn = getattr(self.lib, f"xsdba_rotate_{'' if fused else 'unfused_'}{_sfx(self.dt)}")
        _lib.check(fn(x.data_ptr(), self.Tb * self.N, V, rot.ctypes.data_as(_lib.c_f32p), y.data_ptr(), self.stream), "rotate")
        return y

    def factors(self, ref_v, hist_v, q64):
        """af = quantile(ref_v) - quantile(hist_v) at the float64 nodes (_adjustment.py:315-316) -> [N, nq]."""
        nq = q64.numel()
        af = torch.empty((self.N, 1, nq), dtype=self.dt, device=ref_v.device)
        hq = torch.empty_like(af)
        if self.dt == torch.float32:
            st = self.lib.xsdba_qm_train_q64_f32(ref_v.data_ptr(), hist_v.data_ptr(), self.N, 1, self.N, self.h.ptr,
                                                 q64.data_ptr(), nq, 43, af.data_ptr(), hq.data_ptr(), self.stream)
        else:
            st = self.lib.xsdba_qm_train_f64(ref_v.data_ptr(), hist_v.data_ptr(), self.N, 1, self.N, self.h.ptr,
                                             q64.data_ptr(), nq, 43, 0, af.data_ptr(), hq.data_ptr(), None, self.stream)
        _lib.check(st, "npdft factors")
        return af

    def add_factor_at_rank(self, x_v, af, q64, interp, extrap):
        """x + interp(rank_bn(x) -> af) (_adjustment.py:317-324, 453-460).  The lookup runs on float64 nodes and
        ranks like the reference (af_q lives in a float64 array there), the sum is rounded to the data dtype."""
        x64 = x_v.to(torch.float64).contiguous()
        af64 = af.to(torch.float64).contiguous()
        out = torch.empty_like(x64)
        st = self.lib.xsdba_rank_lookup_f64(x64.data_ptr(), self.N, 1, self.N, self.h.ptr, af64.data_ptr(), q64.data_ptr(),
                                            q64.numel(), _lib.INTERP[interp], _lib.EXTRAP[extrap], 43, 0, 1, out.data_ptr(),
                                            None, self.stream)
        _lib.check(st, "npdft rank lookup")
        return out.to(self.dt)

    def step(self, ref_v, x_v, q64, interp, extrap, af=None):
        """One variable of one N-pdf iteration in ONE launch (float32): quantiles of ref_v / x_v at the float64 nodes,
        af = ref_q - x_q (returned; or ``af`` given when ref_v is None), x_v += af looked up at rank_bn(x_v), in place."""
        nq = q64.numel()
        if af is None:
            af = torch.empty((self.N, nq), dtype=self.dt, device=x_v.device)
        st = self.lib.xsdba_npdft_step_f32(None if ref_v is None else ref_v.data_ptr(), x_v.data_ptr(), self.N, 1, self.N,
                                           self.h.ptr, q64.data_ptr(), nq, _lib.INTERP[interp], _lib.EXTRAP[extrap],
                                           af.data_ptr(), self.stream)
        _lib.check(st, "npdft step")
        return af

    def reorder(self, sim_v, ref_v):
        out = torch.empty_like(sim_v)
        fn = getattr(self.lib, f"xsdba_reorder_{_sfx(self.dt)}")
        _lib.check(fn(sim_v.data_ptr(), ref_v.data_ptr(), self.N, 1, self.N, self.h.ptr, out.data_ptr(), self.stream), "reorder")
        return out


def _prep(x, dt=None):
    x = L4._as_device(x)
    if x.dtype not in (torch.float32, torch.float64):
        x = x.to(torch.float32)
    if dt is not None and x.dtype != dt:
        x = x.to(dt)
    pshape = tuple(x.shape[2:])
    return x.reshape(x.shape[0], x.shape[1], -1).contiguous(), pshape


def _iter_rot(rots, ii):
    return rots[ii] if ii == 0 else rots[ii] @ rots[ii - 1].T  # _adjustment.py:310, 448


def mbcn_train(ref, hist, *, time, rot_matrices, quantiles, group, interp="nearest", extrapolation="constant",
               n_escore=-1):
    """``mbcn_train`` (_adjustment.py:331-423): ref, hist (V, time, *points) -> af_q (n_blocks, *points, n_iter, V, nq)
    and, when ``n_escore > 0``, the energy score after every iteration (n_blocks, *points, n_iter)."""
    group = parse_group(group)
    ref, pshape = _prep(ref)
    hist, _ = _prep(hist, ref.dtype)
    dt = ref.dtype
    V, T, N = ref.shape
    rots = np.asarray(rot_matrices, np.float32)
    q64 = L4._as_device(np.asarray(quantiles, np.float64)).contiguous()
    blocks = grouped_time_indexes(time, group)
    af_q = torch.empty((len(blocks), N, len(rots), V, q64.numel()), dtype=dt, device=ref.device)
    escores = torch.full((len(blocks), N, len(rots)), float("nan"), dtype=dt, device=ref.device)
    for ib, (gw, _) in enumerate(blocks):
        idx = torch.as_tensor(gw, device=ref.device)
        blk = _Block(len(gw), N, dt)
        r = blk.standardize(ref.index_select(1, idx).contiguous())   # _adjustment.py:303-305
        h = blk.standardize(hist.index_select(1, idx).contiguous())
        for ii in range(len(rots)):
            rot = _iter_rot(rots, ii)
            r, h = blk.rotate(r, rot), blk.rotate(h, rot)
            for iv in range(V):
                if dt == torch.float32:   # fused: two sorts and one launch instead of three and three
                    af_q[ib, :, ii, iv, :] = blk.step(r[iv], h[iv], q64, interp, extrapolation)
                    continue
                af = blk.factors(r[iv], h[iv], q64)
                af_q[ib, :, ii, iv, :] = af[:, 0, :]
                h[iv] = blk.add_factor_at_rank(h[iv], af, q64, interp, extrapolation)
            if n_escore > 0:  # _adjustment.py:307-308, 325-326
                from .processing import escore
                escores[ib, :, ii] = escore(r, h, N=n_escore)
    af_q = af_q.reshape(len(blocks), *pshape, len(rots), V, q64.numel())
    if n_escore > 0:
        return af_q, escores.reshape(len(blocks), *pshape, len(rots))
    return af_q


def mbcn_adjust(ref, hist, sim, *, time, af_q, rot_matrices, quantiles, group, kinds, interp="nearest",
                extrapolation="constant"):
    """``mbcn_adjust`` (_adjustment.py:467-591) with ``base = QuantileDeltaMapping``: (V, time, *points) -> scen."""
    group = parse_group(group)
    ref, pshape = _prep(ref)
    hist, _ = _prep(hist, ref.dtype)
    sim, _ = _prep(sim, ref.dtype)
    dt = ref.dtype
    V, T, N = sim.shape
    rots = np.asarray(rot_matrices, np.float32)
    q64 = L4._as_device(np.asarray(quantiles, np.float64)).contiguous()
    q_dt = np.asarray(quantiles).astype(np.float32 if dt == torch.float32 else np.float64)  # adjustment.py:480-483
    af_q = L4._as_device(af_q, dt).reshape(-1, N, len(rots), V, q64.numel())
    blocks = grouped_time_indexes(time, group)
    scen = torch.zeros_like(sim)
    tgrp = Grouper("time")
    for ib, (gw, g) in enumerate(blocks):
        idx = torch.as_tensor(gw, device=sim.device)
        blk = _Block(len(gw), N, dt)
        rb, hb, sb = (a.index_select(1, idx).contiguous() for a in (ref, hist, sim))
        # 1. univariate QDM of every variable on the block (_adjustment.py:548-559)
        scen_block = torch.empty_like(sb)
        fake_time = _BlockTime(len(gw))
        for v in range(V):
            tr = L4.eqm_train(L4.Dataset({"ref": rb[v], "hist": hb[v]}, time=fake_time), group=tgrp, kind=kinds[v],
                              quantiles=q_dt)
            out = L4.qdm_adjust(L4.Dataset({"sim": sb[v], "af": tr["af"], "quantiles": tr["quantiles"]}, time=fake_time),
                                group=tgrp, interp=interp, extrapolation=extrapolation, kind=kinds[v])
            scen_block[v] = out["scen"]
        # 2. N-pdf transform of the standardised block (_adjustment.py:561-586)
        x = blk.standardize(sb)
        for ii in range(len(rots)):
            x = blk.rotate(x, _iter_rot(rots, ii), fused=False)
            for iv in range(V):
                if dt == torch.float32:
                    blk.step(None, x[iv], q64, interp, extrapolation, af=af_q[ib, :, ii, iv, :].contiguous())
                    continue
                x[iv] = blk.add_factor_at_rank(x[iv], af_q[ib, :, ii, iv, :].reshape(N, 1, -1).contiguous(), q64, interp,
                                               extrapolation)
        x = blk.rotate(x, rots[-1].T, fused=False)
        # 3. reorder the univariate scenario by the ranks of the transformed block, keep the exact-group days
        keep = torch.as_tensor(np.nonzero(np.isin(gw, g))[0], device=sim.device)
        gi = torch.as_tensor(g, device=sim.device)
        for v in range(V):
            reordered = blk.reorder(scen_block[v].contiguous(), x[v].contiguous())
            scen[v].index_copy_(0, gi, reordered.index_select(0, keep))
    return scen.reshape(V, T, *pshape)


class _BlockTime:
    """Stand-in time coordinate for a gathered block: only its length matters (group = "time")."""

    def __init__(self, n):
        self.n = n
        self.calendar = "noleap"

    def __len__(self):
        return self.n


class MBCn:
    """``xsdba.adjustment.MBCn`` (adjustment.py:1718-1973): ``MBCn.train(ref, hist, ...)`` then
    ``obj.adjust(sim, ref, hist, ...)``; arrays are (multivar, time, *points)."""

    def __init__(self, ds, group, interp, extrapolation):
        self.ds, self.group, self.interp, self.extrapolation = ds, group, interp, extrapolation

    @classmethod
    def train(cls, ref, hist, *, time, base_kws=None, adj_kws=None, n_escore=-1, n_iter=20, rot_matrices=None, seed=None):
        base_kws = dict(base_kws or {})
        adj_kws = dict(adj_kws or {})
        base_kws.setdefault("nquantiles", 20)
        base_kws.setdefault("group", Grouper("time", 1))
        adj_kws.setdefault("interp", "nearest")
        adj_kws.setdefault("extrapolation", "constant")
        if np.isscalar(base_kws["nquantiles"]):
            base_kws["nquantiles"] = equally_spaced_nodes(base_kws["nquantiles"])
        group = parse_group(base_kws["group"])
        if group.name == "time.month":
            raise NotImplementedError("Received `group==time.month` in `base_kws`. Monthly grouping is not currently "
                                      "supported in the MBCn class.")  # adjustment.py:1851-1852
        if n_escore == 0:
            raise NotImplementedError("n_escore=0 is a no-op in the reference (_adjustment.py:307, 325); use n_escore > 0")
        n_var = ref.shape[0]
        rots = rand_rot_matrix(n_var, n_iter, seed) if rot_matrices is None else np.asarray(rot_matrices, np.float32)
        res = mbcn_train(ref, hist, time=time, rot_matrices=rots, quantiles=base_kws["nquantiles"], group=group,
                         interp=adj_kws["interp"], extrapolation=adj_kws["extrapolation"], n_escore=n_escore)
        af_q, esc = res if n_escore > 0 else (res, None)
        ds = {"af_q": af_q, "escores": esc, "rot_matrices": rots, "quantiles": np.asarray(base_kws["nquantiles"], np.float64)}
        return cls(ds, group, adj_kws["interp"], adj_kws["extrapolation"])

    def adjust(self, sim, ref, hist, *, time, kinds=None):
        kinds = kinds or ["+"] * sim.shape[0]
        return mbcn_adjust(ref, hist, sim, time=time, af_q=self.ds["af_q"], rot_matrices=self.ds["rot_matrices"],
                           quantiles=self.ds["quantiles"], group=self.group, kinds=kinds, interp=self.interp,
                           extrapolation=self.extrapolation)


def npdf_transform(ref, hist, sim, *, time, sim_time, rot_matrices, base_kws=None, adj_kws=None, n_escore=-1, base="qdm"):
    """``xsdba._adjustment.npdf_transform`` (_adjustment.py:977-1057): iterative univariate adjustment in randomly
    rotated spaces.  ref / hist (V, time, *points), sim (V, sim_time, *points) -> dict(scenh, scen, escores).

    Every iteration rotates the three clouds (``x @ R`` along the variable dimension, i.e. ``R.T`` applied: xarray's
    dot / einsum arithmetic, products and sums rounded separately), trains the univariate ``base`` adjustment
    (QuantileDeltaMapping, kind "+", any grouping -- this is the route by which monthly grouping has a meaning for the
    multivariate path, SURVEY.md section 8 note) on every rotated variable, adjusts hist and sim with it and rotates
    back (``x' @ R``: ``R`` applied).  ``n_escore >= 0`` also returns the energy score of (ref, hist) after every
    iteration (scale=True; 0 = all time steps)."""
    if base not in ("qdm", "eqm"):
        raise NotImplementedError("npdf_transform is built for base = QuantileDeltaMapping / EmpiricalQuantileMapping")
    base_kws = dict(base_kws or {})
    adj_kws = dict(adj_kws or {})
    if "kind" in base_kws and base_kws["kind"] != "+":
        import warnings
        warnings.warn('The adjustment kind cannot be controlled when using NpdfTransform, it defaults to "+".', stacklevel=2)
    base_kws.pop("kind", None)   # always "+" (adjustment.py:1331-1337)
    group = parse_group(base_kws.pop("group", "time"), base_kws.pop("window", 1))
    nquantiles = base_kws.pop("nquantiles", 20)
    adj_kws.setdefault("interp", "nearest")
    adj_kws.setdefault("extrapolation", "constant")
    ref, pshape = _prep(ref)
    hist, _ = _prep(hist, ref.dtype)
    sim, _ = _prep(sim, ref.dtype)
    dt = ref.dtype
    V, T, N = ref.shape
    Ts = sim.shape[1]
    rots = np.asarray(rot_matrices, np.float32)
    npdt = np.float32 if dt == torch.float32 else np.float64
    q = equally_spaced_nodes(int(nquantiles)).astype(npdt) if np.isscalar(nquantiles) else np.asarray(nquantiles, npdt)
    blk_t, blk_s = _Block(T, N, dt), _Block(Ts, N, dt)
    flat = lambda a: a.permute(1, 0, 2).reshape(a.shape[1], V * N).contiguous()            # (V, t, N) -> (t, V*N)
    unflat = lambda a, t: a.reshape(t, V, N).permute(1, 0, 2).contiguous()                   # noqa: E731
    escores = []
    for R in rots:
        refp, histp, simp = blk_t.rotate(ref, R.T, fused=False), blk_t.rotate(hist, R.T, fused=False), blk_s.rotate(sim, R.T, fused=False)
        tr = L4.eqm_train(L4.Dataset({"ref": flat(refp), "hist": flat(histp)}, time=time), group=group, kind="+",
                          quantiles=q, **base_kws)
        outs = []
        for x, tx in ((histp, time), (simp, sim_time)):
            if base == "qdm":
                o = L4.qdm_adjust(L4.Dataset({"sim": flat(x), "af": tr["af"], "quantiles": tr["quantiles"]}, time=tx),
                                  group=group, kind="+", **adj_kws)
            else:
                o = L4.qm_adjust(L4.Dataset({"sim": flat(x), "af": tr["af"], "hist_q": tr["hist_q"]}, time=tx),
                                 group=group, kind="+", **adj_kws)
            outs.append(unflat(o["scen"], len(tx)))
        hist = blk_t.rotate(outs[0], R, fused=False)
        sim = blk_s.rotate(outs[1], R, fused=False)
        if n_escore >= 0:
            from .processing import escore
            escores.append(escore(ref, hist, N=n_escore, scale=True))
    esc = torch.stack(escores) if escores else torch.full((len(rots), N), float("nan"), dtype=dt, device=ref.device)
    return {"scenh": hist.reshape(V, T, *pshape), "scen": sim.reshape(V, Ts, *pshape),
            "escores": esc.reshape(len(rots), *pshape), "rotation_matrices": rots}


class NpdfTransform:
    """``xsdba.adjustment.NpdfTransform`` (adjustment.py:1239-1391): ``NpdfTransform.adjust(ref, hist, sim, ...)``
    trains and adjusts in one call; arrays are (multivar, time, *points)."""

    @classmethod
    def adjust(cls, ref, hist, sim, *, time, sim_time=None, base="qdm", base_kws=None, n_escore=0, n_iter=20,
               adj_kws=None, rot_matrices=None, seed=None):
        rots = rand_rot_matrix(ref.shape[0], n_iter, seed) if rot_matrices is None else np.asarray(rot_matrices, np.float32)
        return npdf_transform(ref, hist, sim, time=time, sim_time=time if sim_time is None else sim_time, rot_matrices=rots,
                              base_kws=base_kws, adj_kws=adj_kws, n_escore=n_escore, base=base)
