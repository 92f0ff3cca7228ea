"""Class-level mirror of ``xsdba.adjustment`` for the quantile-mapping family (adjustment.py:414-742).

``Cls.train(ref, hist, ...) -> obj`` and ``obj.adjust(sim, ...)`` keep the reference's argument names
and defaults; ``obj.ds`` holds the trained ``af`` / ``hist_q`` (/ ``scaling``) like the reference's
trained Dataset, so an object can also be rebuilt from saved tables with ``from_dataset``.
Arrays are ``(time, *points)`` (time_axis=0) or ``(*points, time)`` (time_axis=-1); a
:class:`~xsdba_b200.calendar.TimeAxis` plays the role of the xarray time coordinate.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _adjustment as L4
from . import _lib
from .base import parse_group
from .utils import equally_spaced_nodes


class _TrainAdjust:
    _train_fn = None

    def __init__(self, ds, group, kind, adapt_freq_thresh=None, max_tail_factor=None):
        self.ds, self.group, self.kind = ds, group, kind
        self.adapt_freq_thresh, self.max_tail_factor = adapt_freq_thresh, max_tail_factor

    @classmethod
    def from_dataset(cls, ds, *, group, kind):
        """Rebuild a trained object from tables it did not train (base.py:80-91)."""
        return cls(ds, parse_group(group), kind)

    @classmethod
    def train(cls, ref, hist, *, time, nquantiles=20, kind="+", group="time", window=1, time_axis=0, **kw):
        group = parse_group(group, window)
        if np.isscalar(nquantiles):
            dtype = getattr(ref, "dtype", np.float32)
            npdt = np.float64 if "64" in str(dtype) else np.float32
            quantiles = equally_spaced_nodes(int(nquantiles)).astype(npdt)  # adjustment.py:480-483 (no end points)
        else:
            quantiles = np.asarray(nquantiles)
        ds = cls._train_fn(L4.Dataset({"ref": ref, "hist": hist}, time=time, time_axis=time_axis), group=group,
                           kind=kind, quantiles=quantiles, **kw)
        return cls(ds, group, kind, kw.get("adapt_freq_thresh"), kw.get("max_tail_factor"))

    # -- trained object <-> file (the reference keeps a trained adjustment as an xr.Dataset with jsonpickled
    #    parameters in attrs and round-trips it through netCDF + from_dataset, base.py:75-100; without xarray the
    #    same content goes to a .npz: the tables as arrays, the parameters as a JSON string) -------------------
    def save(self, path):
        import json
        arrays = {}
        for k, v in self.ds.items():
            if v is None:
                continue
            arrays[k] = v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
        params = {"cls": type(self).__name__, "group": self.group.name, "window": self.group.window, "kind": self.kind,
                  "adapt_freq_thresh": self.adapt_freq_thresh, "max_tail_factor": self.max_tail_factor}
        np.savez(path, _xsdba_adjustment=np.array(json.dumps(params)), **arrays)

    @classmethod
    def load(cls, path):
        import json
        with np.load(path, allow_pickle=False) as f:
            params = json.loads(str(f["_xsdba_adjustment"]))
            ds = L4.Dataset({k: f[k] for k in f.files if k != "_xsdba_adjustment"})
        if params["cls"] != cls.__name__:
            raise ValueError(f"{path} holds a {params['cls']}, not a {cls.__name__}")
        from .base import Grouper
        return cls(ds, Grouper(params["group"], params["window"]), params["kind"], params["adapt_freq_thresh"],
                   params["max_tail_factor"])

    def _extra(self):
        keys = [k for k in ("P0_ref", "P0_hist", "pth", "hist_q_raw") if self.ds.get(k) is not None]
        return {k: self.ds[k] for k in keys}


class EmpiricalQuantileMapping(_TrainAdjust):
    """adjustment.py:414-528."""
    _train_fn = staticmethod(L4.eqm_train)

    def adjust(self, sim, *, time, interp="nearest", extrapolation="constant", time_axis=0):
        ds = L4.Dataset({"sim": sim, "af": self.ds["af"], "hist_q": self.ds["hist_q"], **self._extra()}, time=time,
                        time_axis=time_axis)
        return L4.qm_adjust(ds, group=self.group, interp=interp, extrapolation=extrapolation, kind=self.kind,
                            adapt_freq_thresh=self.adapt_freq_thresh, max_tail_factor=self.max_tail_factor)["scen"]


class QuantileDeltaMapping(EmpiricalQuantileMapping):
    """adjustment.py:674-742 (train is EQM's)."""

    def adjust(self, sim, *, time, interp="nearest", extrapolation="constant", rank_window=None, time_axis=0,
               extra_output=False):
        ds = L4.Dataset({"sim": sim, "af": self.ds["af"], "quantiles": self.ds["quantiles"], **self._extra()}, time=time,
                        time_axis=time_axis)
        out = L4.qdm_adjust(ds, group=self.group, interp=interp, extrapolation=extrapolation, kind=self.kind,
                            rank_window=rank_window, adapt_freq_thresh=self.adapt_freq_thresh,
                            max_tail_factor=self.max_tail_factor)
        return out if extra_output else out["scen"]  # OPTIONS[EXTRA_OUTPUT] (adjustment.py:738)


class DetrendedQuantileMapping(_TrainAdjust):
    """adjustment.py:531-671."""
    _train_fn = staticmethod(L4.dqm_train)

    def adjust(self, sim, *, time, interp="nearest", extrapolation="constant", detrend=1, time_axis=0):
        ds = L4.Dataset({"sim": sim, "af": self.ds["af"], "hist_q": self.ds["hist_q"], "scaling": self.ds["scaling"],
                         **self._extra()}, time=time, time_axis=time_axis)
        return L4.dqm_adjust(ds, group=self.group, interp=interp, extrapolation=extrapolation, detrend=detrend,
                             kind=self.kind, adapt_freq_thresh=self.adapt_freq_thresh,
                             max_tail_factor=self.max_tail_factor)["scen"]


_WORKSPACES = __import__("threading").local()   # caller-owned device workspaces, one per (host thread, device)


def _host_workspace(nbytes: int):
    import torch
    dev = torch.cuda.current_device()
    cache = getattr(_WORKSPACES, "cache", None)
    if cache is None:
        cache = _WORKSPACES.cache = {}
    ws = cache.get(dev)
    if ws is None or ws.numel() < nbytes:
        cache[dev] = ws = None            # release before growing
        cache[dev] = ws = torch.empty(int(nbytes), dtype=torch.uint8, device=torch.device("cuda", dev))
    return ws


def train_adjust_host(ref: np.ndarray, hist: np.ndarray, sim: np.ndarray, *, time, sim_time, nquantiles=50,
                      group="time.month", window=1, kind="+", method="eqm", interp="nearest",
                      extrapolation="constant", detrend=1, rank_window=False, slab_points=16384, out=None,
                      return_tables=False):
    """EQM / QDM / DQM train + adjust on HOST (numpy, time-major ``(time, *points)``, float32 or float64) arrays
    through the C ABI's end-to-end entry point: slabs of points are streamed H2D -> kernels -> D2H on three streams.
    This is the call the xarray-facing Adjustment classes make for in-memory data.  The device staging workspace is
    owned by this (host thread, device) and reused between calls; ``detrend`` is the PolyDetrend degree of DQM."""
    import torch
    if not torch.cuda.is_available():
        raise _lib.XsdbaB200Error("xsdba_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    lib = _lib.load()
    group = parse_group(group, window)
    dt = ref.dtype
    if dt not in (np.float32, np.float64):
        raise ValueError("train_adjust_host takes float32 or float64 arrays")
    for a in (ref, hist, sim):
        if a.dtype != dt or not a.flags.c_contiguous:
            raise ValueError("train_adjust_host takes C-contiguous arrays of one dtype")
    pshape = ref.shape[1:]
    n_pts = int(np.prod(pshape)) if pshape else 1
    if ref.shape[0] != len(time) or hist.shape != ref.shape or sim.shape[0] != len(sim_time) or sim.shape[1:] != pshape:
        raise ValueError("shape mismatch between ref / hist / sim and their time coordinates")
    m = {"eqm": 0, "qdm": 1, "dqm": 2}[method]
    ht = group.handle(time)
    hs = group.handle(sim_time, with_window=(m == 2) or (m == 1 and bool(rank_window)))
    q = equally_spaced_nodes(nquantiles).astype(dt) if np.isscalar(nquantiles) else np.asarray(nquantiles, dt)
    scen = np.empty_like(sim) if out is None else out
    af = hq = sc = None
    if return_tables:
        af = np.empty(pshape + (ht.n_groups, q.size), dt)
        hq = np.empty_like(af)
        sc = np.empty(pshape + (ht.n_groups,), dt) if m == 2 else None
    tc = np.ascontiguousarray(sim_time.ordinal, np.float64) if m == 2 else None
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)  # noqa: E731
    es = 4 if dt == np.float32 else 8
    nbytes = lib.xsdba_qm_train_adjust_host_workspace_bytes(n_pts, ht.ptr, hs.ptr, q.size, es, m, int(slab_points))
    if nbytes < 0:
        _lib.check(int(nbytes), "train_adjust_host workspace")
    ws = _host_workspace(nbytes)
    fn = lib.xsdba_qm_train_adjust_host_ws_f32 if dt == np.float32 else lib.xsdba_qm_train_adjust_host_ws_f64
    st = fn(p(ref), p(hist), p(sim), n_pts, ht.ptr, hs.ptr, p(q), q.size, _lib.KIND[kind], m, _lib.INTERP[interp],
            _lib.EXTRAP[extrapolation], int(detrend), None if tc is None else tc.ctypes.data_as(_lib.c_f64p), p(scen),
            p(af), p(hq), p(sc), int(slab_points), ws.data_ptr(), ws.numel())
    _lib.check(st, "train_adjust_host")
    if return_tables:
        return (scen, af, hq, sc) if m == 2 else (scen, af, hq)
    return scen
