"""Host mirror of the reference's L4 compute functions for the quantile-mapping path.

``eqm_train`` / ``dqm_train`` / ``qm_adjust`` / ``qdm_adjust`` have the argument names and meaning of
``xsdba._adjustment`` (_adjustment.py:86-286, 594-886) but take a :class:`Dataset` of arrays instead
of an ``xarray.Dataset`` (xarray is not a dependency): series are ``(time, *points)`` (time_axis=0,
the reference's natural netCDF order, fastest on the GPU) or ``(*points, time)`` (time_axis=-1).
Inputs may be numpy arrays or torch CUDA tensors; outputs are torch CUDA tensors.  All arithmetic runs
in the CUDA library through the C ABI -- there is no CPU fallback.
"""
from __future__ import annotations

import warnings

import numpy as np
import torch

from . import _lib
from .base import Grouper, parse_group
from .calendar import TimeAxis
from .detrending import LoessDetrend, PolyDetrend


class Dataset(dict):
    """Minimal stand-in for the ``xr.Dataset`` the reference passes around: a dict of arrays with
    attribute access plus the time coordinate(s)."""

    def __init__(self, data=None, *, time: TimeAxis | None = None, time_axis: int = 0, **kw):
        super().__init__(data or {}, **kw)
        self.time = time
        self.time_axis = time_axis

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def assign(self, **kw):
        out = Dataset(self, time=self.time, time_axis=self.time_axis)
        out.update(kw)
        return out


def _device():
    if not torch.cuda.is_available():
        raise _lib.XsdbaB200Error("xsdba_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _as_device(x, dtype=None):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(np.asarray(x))
    if dtype is not None and x.dtype != dtype:
        x = x.to(dtype)
    if not x.is_cuda:
        x = x.to(_device(), non_blocking=True)
    return x


class _Series(tuple):
    """(tensor, n_pts, stride_pt, stride_time, points_shape) plus ``.axis``: the resolved time axis of the tensor (0 or
    -1), so that outputs are labelled by what was done rather than by what the strides suggest -- a (T, 1) array has
    stride_time == n_pts == 1 and is still time-major."""
    axis = 0


def _series(x, time_axis, n_time, dtype=None):
    """-> (tensor, n_pts, stride_pt, stride_time, points_shape), with ``.axis``."""
    x = _as_device(x, dtype)
    if x.dtype not in (torch.float32, torch.float64):
        x = x.to(torch.float32)
    ax = time_axis % x.ndim
    if ax not in (0, x.ndim - 1):
        x = x.movedim(ax, 0)
        ax = 0
    x = x.contiguous()
    if x.shape[ax] != n_time:
        raise ValueError(f"time axis has {x.shape[ax]} steps, the time coordinate has {n_time}")
    if ax == 0:
        pts_shape = tuple(x.shape[1:])
        n_pts = int(np.prod(pts_shape)) if pts_shape else 1
        out = _Series((x, n_pts, 1, n_pts, pts_shape))
        out.axis = 0
        return out
    pts_shape = tuple(x.shape[:-1])
    n_pts = int(np.prod(pts_shape)) if pts_shape else 1
    out = _Series((x, n_pts, n_time, 1, pts_shape))
    out.axis = -1
    return out


def _widest(*xs):
    dt = torch.float32
    for x in xs:
        d = x.dtype if isinstance(x, torch.Tensor) else torch.from_numpy(np.empty(0, np.asarray(x).dtype)).dtype
        if d == torch.float64:
            dt = torch.float64
    return dt  # base.py:681-685: output dtype = widest input dtype


def _sfx(dt):
    return "f32" if dt == torch.float32 else "f64"


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _reject_unsupported(**kw):
    for k, v in kw.items():
        if v is not None:
            raise NotImplementedError(f"{k} is not built in xsdba_b200 yet (SURVEY.md 8f rank 1)")


def _train(ds, *, group, kind, quantiles, normalize, adapt_freq_thresh=None, jitter_under_thresh_value=None,
           jitter_over_thresh_value=None, jitter_over_thresh_upper_bnd=None, max_tail_factor=None, seed=0):
    if (jitter_over_thresh_value is None) ^ (jitter_over_thresh_upper_bnd is None):
        raise ValueError("`jitter_over_thresh_value` and `jitter_over_thresh_upper_bnd` must both be specified or both "
                         "be `None` (default)")  # _adjustment.py:64-65
    if kind not in _lib.KIND:
        raise ValueError("kind must be + or *.")  # utils.py:139
    group = parse_group(group)
    time = ds.time
    if time is None:
        raise ValueError("the Dataset needs a time coordinate (ds.time)")
    lib = _lib.load()
    dt = _widest(ds["ref"], ds["hist"])
    ref, n_pts, sp, st, pshape = _series(ds["ref"], ds.time_axis, len(time), dt)
    hist, n_pts_h, sp_h, st_h, _ = _series(ds["hist"], ds.time_axis, len(time), dt)
    if n_pts_h != n_pts:
        raise ValueError("ref and hist must have the same points")
    h = group.handle(time)
    G = h.n_groups
    q = _as_device(np.asarray(quantiles), dt).contiguous()
    nq = q.numel()
    af = torch.empty((n_pts, G, nq), dtype=dt, device=ref.device)
    hq = torch.empty_like(af)
    sc = torch.empty((n_pts, G), dtype=dt, device=ref.device) if normalize else None
    import ctypes as C
    from .processing import _quantity, jitter_params
    p0r = p0h = pth = None
    hq_raw = None
    if max_tail_factor is not None:  # quantiles of hist before any pre-processing (_adjustment.py:254-256)
        hq_raw = torch.empty_like(hq)
        fnq = getattr(lib, f"xsdba_group_quantile_{_sfx(dt)}")
        _lib.check(fnq(hist.data_ptr(), n_pts, sp, st, h.ptr, q.data_ptr(), nq, hq_raw.data_ptr(), _stream()), "hist_q_raw")
    if adapt_freq_thresh is not None:
        j4 = None
        if jitter_under_thresh_value is not None or jitter_over_thresh_value is not None:
            j4 = jitter_params(dt, lower=jitter_under_thresh_value, upper=jitter_over_thresh_value,
                               maximum=jitter_over_thresh_upper_bnd)
        p0r = torch.empty((n_pts, G), dtype=torch.float64, device=ref.device)
        p0h = torch.empty_like(p0r)
        pth = torch.empty((n_pts, G), dtype=dt, device=ref.device)
        fn = getattr(lib, f"xsdba_{'dqm' if normalize else 'qm'}_train_adapt_{_sfx(dt)}")
        outs = (af.data_ptr(), hq.data_ptr()) + ((sc.data_ptr(),) if normalize else ())
        status = fn(ref.data_ptr(), hist.data_ptr(), n_pts, sp, st, h.ptr, q.data_ptr(), nq, _lib.KIND[kind],
                    None if j4 is None else j4.ctypes.data_as(_lib.c_f64p), C.c_double(_quantity(adapt_freq_thresh)),
                    C.c_uint64(seed), *outs, p0r.data_ptr(), p0h.data_ptr(), pth.data_ptr(), _stream())
    elif jitter_under_thresh_value is not None or jitter_over_thresh_value is not None:
        j4 = jitter_params(dt, lower=jitter_under_thresh_value, upper=jitter_over_thresh_value,
                           maximum=jitter_over_thresh_upper_bnd)
        fn = getattr(lib, f"xsdba_qm_train_jitter_{_sfx(dt)}")
        status = fn(ref.data_ptr(), hist.data_ptr(), n_pts, sp, st, h.ptr, q.data_ptr(), nq, _lib.KIND[kind],
                    1 if normalize else 0, j4.ctypes.data_as(_lib.c_f64p), C.c_uint64(seed), af.data_ptr(), hq.data_ptr(),
                    sc.data_ptr() if normalize else None, _stream())
    else:
        fn = getattr(lib, f"xsdba_qm_train_{_sfx(dt)}")
        status = fn(ref.data_ptr(), hist.data_ptr(), n_pts, sp, st, h.ptr, q.data_ptr(), nq, _lib.KIND[kind],
                    1 if normalize else 0, af.data_ptr(), hq.data_ptr(), sc.data_ptr() if normalize else None, _stream())
    _lib.check(status, "dqm_train" if normalize else "eqm_train")
    nan_like = lambda t: torch.full_like(t, float("nan"))  # noqa: E731
    out = Dataset(time=None)
    out["af"] = af.reshape(*pshape, G, nq)
    out["hist_q"] = hq.reshape(*pshape, G, nq)
    out["hist_q_raw"] = None if hq_raw is None else hq_raw.reshape(*pshape, G, nq)  # NaN dummy in the reference otherwise
    if p0r is not None:
        out["P0_ref"], out["P0_hist"], out["pth"] = (t.reshape(*pshape, G) for t in (p0r, p0h, pth))
    if normalize:
        out["scaling"] = sc.reshape(*pshape, G)
    out["quantiles"] = q
    out[group.prop] = group.get_coordinate(time)
    out.group = group
    out.kind = kind
    return out


def eqm_train(ds, *, group, kind, quantiles, **kw):
    """EQM train on every group: ``xsdba._adjustment.eqm_train`` (_adjustment.py:193-286)."""
    return _train(ds, group=group, kind=kind, quantiles=quantiles, normalize=False, **kw)


def dqm_train(ds, *, group, kind, quantiles, **kw):
    """DQM train on every group: ``xsdba._adjustment.dqm_train`` (_adjustment.py:86-190)."""
    return _train(ds, group=group, kind=kind, quantiles=quantiles, normalize=True, **kw)


def group_quantile(x, *, time, group, quantiles, time_axis=0):
    """``Grouper.apply(nbutils.quantile, x, q=quantiles)`` (nbutils.py:224-271) -> (*points, G, nq)."""
    group = parse_group(group)
    lib = _lib.load()
    dt = _widest(x)
    xs, n_pts, sp, st, pshape = _series(x, time_axis, len(time), dt)
    h = group.handle(time)
    q = _as_device(np.asarray(quantiles), dt).contiguous()
    out = torch.empty((n_pts, h.n_groups, q.numel()), dtype=dt, device=xs.device)
    fn = getattr(lib, f"xsdba_group_quantile_{_sfx(dt)}")
    _lib.check(fn(xs.data_ptr(), n_pts, sp, st, h.ptr, q.data_ptr(), q.numel(), out.data_ptr(), _stream()),
               "group_quantile")
    return out.reshape(*pshape, h.n_groups, q.numel())


def _tables(ds, n_pts, G, dt, names):
    out = []
    for name in names:
        t = _as_device(ds[name], dt).contiguous()
        nq = t.shape[-1]
        if t.numel() != n_pts * G * nq:
            raise ValueError(f"{name} has shape {tuple(t.shape)}, expected (*points={n_pts}, {G}, nq)")
        out.append(t)
    return out


def _adapt_freq_preprocess(ds, sim, n_pts, sp, st, group, time, dt, adapt_freq_thresh, seed=0):
    """``_adapt_freq_preprocess`` (_adjustment.py:32-45): frequency-adapt ``sim`` per exact group with the stored
    P0_ref / P0_hist / pth."""
    import ctypes as C
    from .processing import _quantity
    lib = _lib.load()
    h = Grouper(group.name).handle(time)
    G = h.n_groups
    for k in ("P0_ref", "P0_hist", "pth"):
        if ds.get(k) is None:
            raise ValueError("`P0_ref`, `P0_hist`, `pth` must all be given for adapt_freq_thresh")  # _processing.py:66-67
    p0r = _as_device(ds["P0_ref"], torch.float64).contiguous()
    p0h = _as_device(ds["P0_hist"], torch.float64).contiguous()
    pth = _as_device(ds["pth"], dt).contiguous()
    if p0r.numel() != n_pts * G:
        raise ValueError("P0_ref must be (*points, n_groups)")
    out = torch.empty_like(sim)
    fn = getattr(lib, f"xsdba_adapt_freq_apply_{_sfx(dt)}")
    _lib.check(fn(sim.data_ptr(), n_pts, sp, st, h.ptr, C.c_double(_quantity(adapt_freq_thresh)), p0r.data_ptr(),
                  p0h.data_ptr(), pth.data_ptr(), C.c_uint64(seed), out.data_ptr(), _stream()), "adapt_freq")
    return out


def _apply_tail_mask(ds, adapted, scen, n_pts, sp, st, group, time, dt, max_tail_factor):
    """max_tail_factor (_adjustment.py:647-658, 672-673): keep the (pre-processed) sim where it exceeds
    ``max_tail_factor`` times the last node of ``hist_q_raw``."""
    import ctypes as C
    lib = _lib.load()
    if ds.get("hist_q_raw") is None:
        raise ValueError("max_tail_factor needs `hist_q_raw` (train with max_tail_factor set)")
    h = Grouper(group.name).handle(time)
    (hq_raw,) = _tables(ds, n_pts, h.n_groups, dt, ("hist_q_raw",))
    fn = getattr(lib, f"xsdba_tail_mask_{_sfx(dt)}")
    _lib.check(fn(adapted.data_ptr(), n_pts, sp, st, h.ptr, hq_raw.data_ptr(), hq_raw.shape[-1], C.c_double(float(max_tail_factor)),
                  scen.data_ptr(), _stream()), "max_tail_factor")
    return scen


def qm_adjust(ds, *, group, interp, extrapolation, kind, adapt_freq_thresh=None, max_tail_factor=None, seed=0):
    """``xsdba._adjustment.qm_adjust`` (_adjustment.py:594-676): ds holds af, hist_q, sim (+ P0_ref, P0_hist, pth
    for ``adapt_freq_thresh``; + hist_q_raw for ``max_tail_factor``)."""
    if interp not in ("nearest", "linear", "cubic") or extrapolation not in _lib.EXTRAP:
        raise ValueError("interp must be nearest/linear/cubic and extrapolation constant/nan")
    group = parse_group(group)
    if interp == "cubic" and group.prop != "group":
        raise NotImplementedError("cubic interpolation over time.<prop> groups (2-D Clough-Tocher) is not built in xsdba_b200")
    lib = _lib.load()
    time = ds.time
    dt = _widest(ds["sim"], ds["af"])
    ser = _series(ds["sim"], ds.time_axis, len(time), dt)
    sim, n_pts, sp, st, pshape = ser
    h = group.handle(time, with_window=False)
    af, hq = _tables(ds, n_pts, h.n_groups, dt, ("af", "hist_q"))
    nq = af.shape[-1]
    if adapt_freq_thresh is not None:
        sim = _adapt_freq_preprocess(ds, sim, n_pts, sp, st, group, time, dt, adapt_freq_thresh, seed)
    scen = torch.empty_like(sim)
    fn = getattr(lib, f"xsdba_qm_adjust_{_sfx(dt)}")
    status = fn(sim.data_ptr(), n_pts, sp, st, h.ptr, af.data_ptr(), hq.data_ptr(), nq, _lib.INTERP[interp],
                _lib.EXTRAP[extrapolation], _lib.KIND[kind], scen.data_ptr(), _stream())
    _lib.check(status, "qm_adjust")
    if max_tail_factor is not None:
        scen = _apply_tail_mask(ds, sim, scen, n_pts, sp, st, group, time, dt, max_tail_factor)
    return Dataset({"scen": scen}, time=time, time_axis=ser.axis)


_GEOM_CACHE: dict = {}


def _qdm_linear_geometry(group, time, q, n_groups):
    """Host part of grouped ``interp="linear"`` for QDM: the fractional group coordinate of every time step
    (``Grouper.get_index(interp=True)``, base.py:306-320) and, from ONE Qhull triangulation of the lattice
    (quantile node, padded group coordinate) -- the point set SciPy's ``griddata(method="linear")`` receives for
    every gridpoint (utils.py:391-396) -- which diagonal splits each lattice cell."""
    if group.prop not in ("month", "dayofyear"):
        raise NotImplementedError(f"linear interpolation over time.{group.prop} groups is not built yet")
    qn = q.detach().cpu().numpy().astype(np.float64)
    key = (group.prop, n_groups, qn.tobytes(), torch.cuda.current_device())
    if key not in _GEOM_CACHE:
        from scipy.spatial import Delaunay  # SciPy is the reference's own dependency for this step
        gg = np.arange(n_groups + 2, dtype=np.float64)
        X, Y = np.meshgrid(qn, gg)
        tri = Delaunay(np.stack([X.ravel(), Y.ravel()], axis=1))
        nq = qn.size
        diag = np.full((n_groups + 1, nq - 1), 255, np.uint8)
        for smp in tri.simplices:
            r, k = smp // nq, smp % nq
            r0, k0 = int(r.min()), int(k.min())
            if r.max() - r0 != 1 or k.max() - k0 != 1:
                raise RuntimeError("unexpected Qhull simplex on the quantile lattice")
            corners = set(zip((r - r0).tolist(), (k - k0).tolist()))
            diag[r0, k0] = 0 if {(0, 0), (1, 1)} <= corners else 1
        if (diag == 255).any():
            raise RuntimeError("incomplete Qhull triangulation of the quantile lattice")
        _GEOM_CACHE[key] = _as_device(diag).contiguous()
    gcoord = _as_device(np.asarray(group.get_index(time, interp=True), np.float64)).contiguous()
    return gcoord, _GEOM_CACHE[key]


def qdm_adjust(ds, *, group, interp, extrapolation, kind, adapt_freq_thresh=None, rank_window=None,
               max_tail_factor=None, seed=0):
    """``xsdba._adjustment.qdm_adjust`` (_adjustment.py:783-886): ds holds af, quantiles, sim."""
    group = parse_group(group)
    if interp == "cubic" and group.prop != "group":
        raise NotImplementedError("cubic interpolation over time.<prop> groups (2-D Clough-Tocher) is not built in xsdba_b200")
    if rank_window is None:
        rank_window = False
        if group.window > 1:
            warnings.warn("QDM ranks within exact groups (rank_window=False); xsdba>=0.8 will honour the grouping "
                          "window (rank_window=True).", category=DeprecationWarning, stacklevel=2)  # _adjustment.py:858-871
    lib = _lib.load()
    time = ds.time
    dt = _widest(ds["sim"], ds["af"])
    ser = _series(ds["sim"], ds.time_axis, len(time), dt)
    sim, n_pts, sp, st, pshape = ser
    if max_tail_factor is not None and interp != "nearest" and group.prop in ("month", "season"):
        # the reference interpolates the last raw quantile between months for the mask (_adjustment.py:847-858,
        # u.broadcast(..., interp=interp)); only the nearest-group broadcast is built here
        raise NotImplementedError("max_tail_factor with a month-interpolated last quantile is not built yet")
    h = group.handle(time, with_window=bool(rank_window))
    (af,) = _tables(ds, n_pts, h.n_groups, dt, ("af",))
    q = _as_device(ds["quantiles"], dt).contiguous()
    nq = af.shape[-1]
    if adapt_freq_thresh is not None:
        sim = _adapt_freq_preprocess(ds, sim, n_pts, sp, st, group, time, dt, adapt_freq_thresh, seed)
    scen = torch.empty_like(sim)
    sim_q = torch.empty(sim.shape, dtype=torch.float64, device=sim.device)
    if interp == "linear" and h.n_groups > 1:
        gcoord, diag = _qdm_linear_geometry(group, time, q, h.n_groups)
        if bool(torch.isnan(af).any()):
            raise NotImplementedError("grouped linear interpolation with NaN adjustment factors is not built yet")
        fn = getattr(lib, f"xsdba_qdm_adjust_linear_{_sfx(dt)}")
        status = fn(sim.data_ptr(), n_pts, sp, st, h.ptr, af.data_ptr(), q.data_ptr(), nq, _lib.EXTRAP[extrapolation],
                    _lib.KIND[kind], 1 if rank_window else 0, gcoord.data_ptr(), diag.data_ptr(), scen.data_ptr(),
                    sim_q.data_ptr(), _stream())
    else:
        fn = getattr(lib, f"xsdba_qdm_adjust_{_sfx(dt)}")
        status = fn(sim.data_ptr(), n_pts, sp, st, h.ptr, af.data_ptr(), q.data_ptr(), nq, _lib.INTERP[interp],
                    _lib.EXTRAP[extrapolation], _lib.KIND[kind], 1 if rank_window else 0, scen.data_ptr(),
                    sim_q.data_ptr(), _stream())
    _lib.check(status, "qdm_adjust")
    if max_tail_factor is not None:
        scen = _apply_tail_mask(ds, sim, scen, n_pts, sp, st, group, time, dt, max_tail_factor)
    return Dataset({"scen": scen, "sim_q": sim_q}, time=time, time_axis=ser.axis)


def group_rank(x, *, time, group, rank_window=False, time_axis=0):
    """``Grouper.apply(utils.rank, x, main_only=not rank_window, pct=True)`` (utils.py:573-638) -> float64."""
    group = parse_group(group)
    lib = _lib.load()
    dt = _widest(x)
    xs, n_pts, sp, st, _ = _series(x, time_axis, len(time), dt)
    h = group.handle(time, with_window=bool(rank_window))
    out = torch.empty(xs.shape, dtype=torch.float64, device=xs.device)
    fn = getattr(lib, f"xsdba_group_rank_{_sfx(dt)}")
    _lib.check(fn(xs.data_ptr(), n_pts, sp, st, h.ptr, 1 if rank_window else 0, out.data_ptr(), _stream()), "rank")
    return out


def poly_trend(x, *, time, group, degree, kind="+", scaling=None, time_axis=0, preserve_mean=False):
    """Trend of ``x`` (optionally of ``x (+|*) scaling``) as ``PolyDetrend(degree, group).fit`` computes it
    (detrending.py:189-208): float64 tensor shaped like ``x``."""
    group = parse_group(group)
    lib = _lib.load()
    dt = _widest(x)
    xs_ = _series(x, time_axis, len(time), dt)
    xs, n_pts, sp, st, _ = xs_
    h = group.handle(time)
    sc = None
    if scaling is not None:
        sc = _as_device(scaling, dt).contiguous()
        if sc.numel() != n_pts * h.n_groups:
            raise ValueError("scaling must be (*points, n_groups)")
    tc = _as_device(np.asarray(time.ordinal, np.float64)).contiguous()
    trend = torch.empty(xs.shape, dtype=torch.float64, device=xs.device)
    fn = getattr(lib, f"xsdba_poly_trend_{_sfx(dt)}")
    _lib.check(fn(xs.data_ptr(), n_pts, sp, st, h.ptr, sc.data_ptr() if sc is not None else None, _lib.KIND[kind],
                  int(degree), tc.data_ptr(), trend.data_ptr(), _stream()), "poly_trend")
    if preserve_mean:
        # detrending.py:205: trend (+|*) invert(trend.mean(time)) inside every group -- a reduction over 30-10 950
        # values per (point, group) on a tensor that is already on the device (torch ops, no kernel of its own)
        tm = trend if xs_.axis == 0 else trend.movedim(-1, 0)
        tm2 = tm.reshape(tm.shape[0], -1)
        gi = torch.as_tensor(group.zero_based_index(time), device=trend.device, dtype=torch.long)
        ok = ~torch.isnan(tm2)
        sums = torch.zeros((h.n_groups, tm2.shape[1]), dtype=trend.dtype, device=trend.device).index_add_(0, gi, torch.nan_to_num(tm2))
        cnts = torch.zeros_like(sums).index_add_(0, gi, ok.to(trend.dtype))
        mean = (sums / cnts)[gi]
        tm2.copy_(tm2 - mean if kind == "+" else tm2 * (1.0 / mean))
    return trend


def _loess_spacing(o, equal_spacing):
    """loess.py:250-260: the dx > 0 branch for an equally spaced coordinate unless it is switched off, the dx == 0
    branch otherwise."""
    diffs = np.diff(o)
    if diffs.size and np.all(diffs == diffs[0]):
        return True if equal_spacing is None else bool(equal_spacing)
    if equal_spacing:
        raise NotImplementedError("equal_spacing=True on an unequally spaced coordinate (the reference warns of 'strange "
                                  "results', loess.py:255) is not built")
    return False


def loess_trend(x, *, time, f=0.2, niter=1, d=0, kind="+", scaling=None, scaling_group=None, loess_group="time",
                time_axis=0, weights="tricube", equal_spacing=None):
    """``LoessDetrend(group, f, niter, d, weights, equal_spacing).fit(x (+|*) scaling).ds.trend`` (detrending.py:211-296 ->
    loess.loess_smoothing, loess.py:182-279): float64 tensor shaped like ``x``.  ``loess_group="time"`` smooths the
    whole series; a grouped LoessDetrend (``"time.month"``, ``"time.season"``, ``"time.dayofyear"``, window 1) smooths
    the members of every group on their own, normalised time coordinate -- never equally spaced, hence the dx == 0
    branch."""
    loess_group = parse_group(loess_group)
    if weights not in ("tricube", "gaussian"):
        raise NotImplementedError("LOESS weights: 'tricube' or 'gaussian' (loess.py:247)")
    wflag = 1 if weights == "gaussian" else 0
    lib = _lib.load()
    dt = _widest(x)
    ser = _series(x, time_axis, len(time), dt)
    xs, n_pts, sp, st, _ = ser
    o = np.asarray(time.ordinal, np.float64)
    fn = getattr(lib, f"xsdba_loess_trend_w_{_sfx(dt)}")
    if loess_group.prop == "group":
        if len(o) < 3:
            raise NotImplementedError("LOESS needs at least 3 time steps")
        eq = _loess_spacing(o, equal_spacing)
        xn = _as_device((o - o[0]) / (o[-1] - o[0])).contiguous()   # loess.py:244-245
        g = parse_group(scaling_group) if scaling is not None else parse_group("time")
        h = g.handle(time, with_window=False)
        sc = None
        if scaling is not None:
            sc = _as_device(scaling, dt).contiguous()
            if sc.numel() != n_pts * h.n_groups:
                raise ValueError("scaling must be (*points, n_groups)")
        trend = torch.empty(xs.shape, dtype=torch.float64, device=xs.device)
        _lib.check(fn(xs.data_ptr(), n_pts, sp, st, h.ptr, sc.data_ptr() if sc is not None else None, _lib.KIND[kind],
                      float(f), int(niter), int(d), wflag, 1 if eq else 0, xn.data_ptr(), trend.data_ptr(), _stream()),
                   "loess_trend")
        return trend
    if loess_group.window != 1:
        raise NotImplementedError("a grouped LoessDetrend with a window is not built in xsdba_b200")
    # grouped: (time, points) time-major copy, scaling applied first, one smoothing per group on the gathered members
    from .base import grouping_handle
    x2 = (xs if ser.axis == 0 else xs.movedim(-1, 0)).reshape(len(time), n_pts).contiguous()
    if scaling is not None:
        sg = parse_group(scaling_group)
        sc = _as_device(scaling, dt).reshape(n_pts, -1)
        gi_s = torch.from_numpy(sg.zero_based_index(time).astype(np.int64)).to(x2.device)
        sct = sc.t()[gi_s]                                         # (time, points)
        x2 = x2 + sct if kind == "+" else x2 * sct
    gi = loess_group.zero_based_index(time)
    trend2 = torch.full((len(time), n_pts), float("nan"), dtype=torch.float64, device=x2.device)
    for g in range(loess_group.n_groups(time)):
        rows = np.nonzero(gi == g)[0]
        if rows.size < 3:
            continue
        og = o[rows]
        eq = _loess_spacing(og, equal_spacing)
        xn = _as_device((og - og[0]) / (og[-1] - og[0])).contiguous()
        rows_t = torch.from_numpy(rows.astype(np.int64)).to(x2.device)
        sub = x2[rows_t].contiguous()
        tr = torch.empty(sub.shape, dtype=torch.float64, device=x2.device)
        h = grouping_handle(np.zeros(rows.size, np.int32), 1, 1)
        _lib.check(fn(sub.data_ptr(), n_pts, 1, n_pts, h.ptr, None, _lib.KIND[kind], float(f), int(niter), int(d), wflag,
                      1 if eq else 0, xn.data_ptr(), tr.data_ptr(), _stream()), "loess_trend")
        trend2[rows_t] = tr
    trend2 = trend2.reshape((len(time),) + tuple(ser[4]))
    return trend2 if ser.axis == 0 else trend2.movedim(0, -1).contiguous()


def dqm_adjust(ds, *, group, interp, kind, extrapolation, detrend=1, adapt_freq_thresh=None, max_tail_factor=None,
               seed=0):
    """``xsdba._adjustment.dqm_adjust`` (_adjustment.py:679-780): ds holds scaling, af, hist_q, sim (+ P0_ref,
    P0_hist, pth for ``adapt_freq_thresh``; + hist_q_raw for ``max_tail_factor``).
    ``detrend`` is an int (PolyDetrend degree on the adjust group) or a PolyDetrend / LoessDetrend."""
    group = parse_group(group)
    if interp == "cubic" and group.prop != "group":
        raise NotImplementedError("cubic interpolation over time.<prop> groups (2-D Clough-Tocher) is not built in xsdba_b200")
    if group.prop not in ("group", "dayofyear") and interp != "nearest":
        raise NotImplementedError("broadcasting `scaling` with linear interpolation over months is not built yet")
    lib = _lib.load()
    time = ds.time
    dt = _widest(ds["sim"], ds["af"])
    ser = _series(ds["sim"], ds.time_axis, len(time), dt)
    sim, n_pts, sp, st, pshape = ser
    ta = ser.axis
    h = group.handle(time)
    af, hq = _tables(ds, n_pts, h.n_groups, dt, ("af", "hist_q"))
    scaling = _as_device(ds["scaling"], dt).contiguous()
    if scaling.numel() != n_pts * h.n_groups:
        raise ValueError("scaling must be (*points, n_groups)")
    if adapt_freq_thresh is not None:   # _adjustment.py:727-733: on the exact groups, before anything else
        sim = _adapt_freq_preprocess(ds, sim, n_pts, sp, st, group, time, dt, adapt_freq_thresh, seed)
    adapted = sim
    if max_tail_factor is not None and group.prop not in ("group", "dayofyear") and interp != "nearest":
        raise NotImplementedError("max_tail_factor with a month-interpolated last quantile is not built yet")
    if isinstance(detrend, (int, np.integer)):
        detrend = PolyDetrend(degree=int(detrend), kind=kind, group=group)   # _adjustment.py:759-762
    if isinstance(detrend, PolyDetrend):
        trend = poly_trend(sim, time=time, group=detrend.group, degree=detrend.degree, kind=kind,
                           preserve_mean=getattr(detrend, "preserve_mean", False),
                           scaling=scaling if detrend.group.name == group.name and detrend.group.window == group.window else None,
                           time_axis=ta)
        if not (detrend.group.name == group.name and detrend.group.window == group.window):
            raise NotImplementedError("a PolyDetrend with a group different from the adjustment group is not built yet")
    elif isinstance(detrend, LoessDetrend):
        trend = loess_trend(sim, time=time, f=detrend.f, niter=detrend.niter, d=detrend.d, kind=kind, scaling=scaling,
                            scaling_group=group, loess_group=detrend.group, time_axis=ta,
                            weights=getattr(detrend, "weights", "tricube"),
                            equal_spacing=getattr(detrend, "equal_spacing", None))
    else:
        raise TypeError("detrend must be an int, a PolyDetrend or a LoessDetrend")
    nq = af.shape[-1]
    scen = torch.empty_like(sim)
    fn = getattr(lib, f"xsdba_dqm_adjust_{_sfx(dt)}")
    status = fn(sim.data_ptr(), n_pts, sp, st, h.ptr, af.data_ptr(), hq.data_ptr(), scaling.data_ptr(),
                trend.data_ptr(), nq, _lib.INTERP[interp], _lib.EXTRAP[extrapolation], _lib.KIND[kind],
                scen.data_ptr(), _stream())
    _lib.check(status, "dqm_adjust")
    if max_tail_factor is not None:     # _adjustment.py:734-746, 776-777
        scen = _apply_tail_mask(ds, adapted, scen, n_pts, sp, st, group, time, dt, max_tail_factor)
    return Dataset({"scen": scen, "trend": trend}, time=time, time_axis=ta)


def vecquantiles(x, rnk, *, time, group="time", time_axis=0):
    """``nbutils.vecquantiles`` per group (nbutils.py:164-195): the quantile of every (point, group) segment at its
    own rank ``rnk`` (*points, n_groups) -> (*points, n_groups)."""
    group = parse_group(group)
    lib = _lib.load()
    dt = _widest(x)
    xs, n_pts, sp, st, pshape = _series(x, time_axis, len(time), dt)
    h = group.handle(time)
    r = _as_device(rnk, dt).contiguous()
    if r.numel() != n_pts * h.n_groups:
        raise ValueError("rnk must be (*points, n_groups)")
    out = torch.empty((n_pts, h.n_groups), dtype=dt, device=xs.device)
    fn = getattr(lib, f"xsdba_group_vecquantile_{_sfx(dt)}")
    _lib.check(fn(xs.data_ptr(), n_pts, sp, st, h.ptr, r.data_ptr(), out.data_ptr(), _stream()), "vecquantiles")
    return out.reshape(*pshape, h.n_groups)


def map_cdf(ds, *, y_value, group="time"):
    """``utils.map_cdf`` under ``Grouper.apply`` (utils.py:47-84): ds holds x and y; returns the value of x with the
    same empirical CDF as ``y_value`` in y: (*points, n_groups, len(y_value))."""
    group = parse_group(group)
    lib = _lib.load()
    time = ds.time
    dt = _widest(ds["x"], ds["y"])
    xs, n_pts, sp, st, pshape = _series(ds["x"], ds.time_axis, len(time), dt)
    ys, n_pts_y, _, _, _ = _series(ds["y"], ds.time_axis, len(time), dt)
    if n_pts_y != n_pts:
        raise ValueError("x and y must have the same points")
    h = group.handle(time)
    yv = _as_device(np.atleast_1d(np.asarray(y_value, np.float64))).contiguous()
    out = torch.empty((n_pts, h.n_groups, yv.numel()), dtype=dt, device=xs.device)
    fn = getattr(lib, f"xsdba_map_cdf_{_sfx(dt)}")
    _lib.check(fn(xs.data_ptr(), ys.data_ptr(), n_pts, sp, st, h.ptr, yv.data_ptr(), yv.numel(), out.data_ptr(), _stream()),
               "map_cdf")
    return out.reshape(*pshape, h.n_groups, yv.numel())
