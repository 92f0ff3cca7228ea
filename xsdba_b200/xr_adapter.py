"""Drop-in seam: GPU-backed replacements of the five L4 functions that ``xsdba.adjustment`` imports from
``xsdba._adjustment`` (adjustment.py:18-35) -- ``eqm_train``, ``dqm_train``, ``qm_adjust``, ``dqm_adjust``,
``qdm_adjust`` -- with the reference's ``(ds, *, group, ...) -> Dataset`` contract (_adjustment.py:95-106, 201-212,
595-604, 680-690, 784-794), so that the unmodified ``EmpiricalQuantileMapping`` / ``DetrendedQuantileMapping`` /
``QuantileDeltaMapping`` classes run on the CUDA library.

    import xsdba_b200.xr_adapter as b200
    b200.patch()            # rebinds xsdba.adjustment.<name> and xsdba._adjustment.<name>; b200.unpatch() restores
    EQM = xsdba.EmpiricalQuantileMapping.train(ref, hist, nquantiles=50, group="time.month", kind="+")
    scen = EQM.adjust(sim)

What the replacements do that the L4 originals do (and the array-level mirror in ``_adjustment.py`` does not):
  * take an ``xr.Dataset`` (anything with its duck type: ``ds[name]`` -> arrays with ``.dims``, ``.values``,
    ``.transpose``, ``.attrs``; ``ds.indexes["time"]``), put ``time`` first and flatten every other dimension into
    gridpoints (time-major, the layout the kernels are tuned for), widen to the largest input dtype (base.py:681-685);
  * accept ``group`` as a reference ``Grouper`` (``.name``, ``.window``, ``.add_dims``), a str, or this package's;
  * convert unit-string thresholds (``"0.01 mm/d"``) to the data's units with ``xsdba.units.convert_units_to`` when
    xsdba is importable (_adjustment.py:37, 61-67), else require the units to be spelled identically;
  * return a Dataset of the caller's own type with dims ``(<points...>, group.prop, "quantiles")`` for the trained
    tables -- a size-1 ``group`` dimension for ``group="time"`` -- coords ``quantiles`` and ``group.prop`` (base.py:
    661-667), the NaN dummies ``hist_q_raw`` / ``P0_ref`` / ``P0_hist`` / ``pth`` the classes drop afterwards
    (_adjustment.py:277-286), and keep the undecorated callable as ``.func`` (base.py:723, 775; ``dqm_adjust`` calls
    ``qm_adjust.func``, _adjustment.py:766).
There is no CPU fallback: without the CUDA library every call raises.  xarray itself is only needed by ``patch``.
"""
from __future__ import annotations

import numpy as np

from . import _adjustment as L4
from .base import Grouper
from .calendar import TimeAxis
from .detrending import LoessDetrend, PolyDetrend

_PATCHED: dict = {}
NAMES = ("eqm_train", "dqm_train", "qm_adjust", "dqm_adjust", "qdm_adjust")


# ------------------------------------------------------------------------------------------------------------
# marshalling
# ------------------------------------------------------------------------------------------------------------
def _group(group, window=1):
    """Reference Grouper / str / local Grouper -> local Grouper (base.py:128-175, 504-538)."""
    if isinstance(group, Grouper):
        return group
    if isinstance(group, str):
        return Grouper(group, window=window)
    add_dims = list(getattr(group, "add_dims", []) or [])
    return Grouper(group.name, window=int(getattr(group, "window", 1)), add_dims=add_dims or None)


def time_axis_of(ds, dim="time") -> TimeAxis:
    """``ds.indexes[dim]`` (pandas DatetimeIndex, or a cftime index with ``.calendar``) -> TimeAxis."""
    idx = ds.indexes[dim] if hasattr(ds, "indexes") else ds[dim].values
    cal = getattr(idx, "calendar", None)
    if cal is None and hasattr(idx, "year") and hasattr(idx, "dayofyear") and not hasattr(idx[0], "calendar"):
        return TimeAxis.from_datetime64(np.asarray(idx.values if hasattr(idx, "values") else idx))
    if cal is None:
        cal = getattr(idx[0], "calendar", "standard")
    cal = {"gregorian": "standard", "proleptic_gregorian": "standard", "365_day": "noleap", "366_day": "all_leap"}.get(cal, cal)
    get = lambda f: np.asarray([getattr(t, f) for t in idx], np.int64)  # noqa: E731
    return TimeAxis.from_fields(get("year"), get("month"), get("day"), cal)


class _PooledGrouper(Grouper):
    """``Grouper(..., add_dims=[...])`` after pooling (base.py:410-415): the extra dimensions are laid end to end along the
    time axis -- block a holds the series of pooled coordinate a, blocks are separated by ``gap`` NaN time steps that
    belong to no group, so that a window around the first days of one block cannot reach the last days of the block
    before it (NaN samples are ignored by every group-wise reduction).  Group index, coordinate and number of groups
    are those of the ORIGINAL time axis, repeated per block."""

    def __init__(self, base: Grouper, time: TimeAxis, n_rep: int, gap: int):
        super().__init__(base.name, window=base.window)
        self._base_time, self._n_rep, self._gap = time, int(n_rep), int(gap)

    def _tile(self, idx, fill):
        T = len(self._base_time)
        out = np.full((self._n_rep, T + self._gap), fill, dtype=np.asarray(idx).dtype)
        out[:, :T] = idx
        return out.reshape(-1)

    def zero_based_index(self, time=None):
        return self._tile(super().zero_based_index(self._base_time), -1).astype(np.int32)

    def n_groups(self, time=None):
        return super().n_groups(self._base_time)

    def get_coordinate(self, time=None):
        return super().get_coordinate(self._base_time)

    def pooled_time(self) -> TimeAxis:
        t = self._base_time
        tile = lambda a: self._tile(np.asarray(a), np.asarray(a)[0])  # noqa: E731  (spacer steps: any valid date)
        return TimeAxis(tile(t.year), tile(t.month), tile(t.day), tile(t.dayofyear), tile(t.days_in_month), t.calendar)


def _pool(da, add_dims, gap, dtype):
    """DataArray -> ((n_rep * (T + gap), n_pts) array with NaN spacer steps, other dims, their sizes, n_rep)."""
    missing = [d for d in add_dims if d not in da.dims]
    if missing:
        raise ValueError(f"add_dims {missing} are not dimensions of the data")
    other = [d for d in da.dims if d != "time" and d not in add_dims]
    a = np.asarray(da.transpose(*add_dims, "time", *other).values)
    n_rep = int(np.prod(a.shape[:len(add_dims)]))
    T = a.shape[len(add_dims)]
    sizes = a.shape[len(add_dims) + 1:]
    a = a.reshape(n_rep, T, -1)
    out = np.full((n_rep, T + gap, a.shape[2]), np.nan, dtype=dtype)
    out[:, :T] = a
    return out.reshape(n_rep * (T + gap), -1), other, tuple(sizes), n_rep


def _series(da, dtype):
    """DataArray (time among its dims) -> (numpy (time, n_pts) C-contiguous, other dims, their sizes)."""
    other = [d for d in da.dims if d != "time"]
    a = np.asarray(da.transpose("time", *other).values)
    sizes = a.shape[1:]
    return np.ascontiguousarray(a.reshape(a.shape[0], -1), dtype=dtype), other, tuple(sizes)


def _table(da, lead_dims, tail_dims, dtype):
    """Trained table with dims (<points...>, *tail_dims) in any order -> numpy (n_pts, *tail sizes)."""
    a = np.asarray(da.transpose(*lead_dims, *tail_dims).values)
    n_tail = len(tail_dims)
    return np.ascontiguousarray(a.reshape((-1,) + a.shape[a.ndim - n_tail:]), dtype=dtype)


def _widest(ds, names):
    return np.float64 if any(np.asarray(ds[n].values).dtype == np.float64 for n in names if n in ds) else np.float32


def _thresh(value, like):
    """Unit-string threshold -> float in the units of ``like`` (_adjustment.py:37, 61-67)."""
    if value is None or not isinstance(value, str):
        return value
    try:
        from xsdba.units import convert_units_to   # the reference's own conversion (pint)
        return float(convert_units_to(value, like))
    except ImportError:
        num, _, unit = value.partition(" ")
        have = str(getattr(like, "attrs", {}).get("units", unit)).strip()
        if unit.strip() and have and unit.strip() != have:
            raise ValueError(f"cannot convert '{value}' to '{have}' without xsdba.units (pint)")
        return float(num)


def _np(t):
    return t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)


def _types(ds):
    """(Dataset type, DataArray type) of the caller's objects, so that results are of the caller's own classes."""
    first = next(iter(ds.data_vars.values())) if hasattr(ds, "data_vars") else ds[next(iter(ds))]
    return type(ds), type(first)


def _dataset(ds, variables: dict, coords: dict):
    DS, DA = _types(ds)
    out = {}
    for name, (dims, data) in variables.items():
        out[name] = DA(data, dims=tuple(dims), coords={d: coords[d] for d in dims if d in coords}, name=name)
    return DS(out)


def _point_coords(ds, dims):
    return {d: ds[d].values for d in dims if hasattr(ds, "coords") and d in ds.coords}


# ------------------------------------------------------------------------------------------------------------
# the five L4 replacements
# ------------------------------------------------------------------------------------------------------------
def _train(ds, *, group, kind, quantiles, normalize, adapt_freq_thresh=None, jitter_under_thresh_value=None,
           jitter_over_thresh_value=None, jitter_over_thresh_upper_bnd=None, max_tail_factor=None):
    grp = _group(group)
    dt = _widest(ds, ("ref", "hist"))
    time = time_axis_of(ds)
    l4_time = time
    if grp.add_dims:
        # base.py:410-415: the group-wise reductions of train (quantiles, means, frequencies) pool the extra dimensions
        if adapt_freq_thresh is not None:
            raise NotImplementedError("adapt_freq_thresh together with add_dims is not built in xsdba_b200")
        gap = grp.window // 2
        ref, pdims, psizes, n_rep = _pool(ds["ref"], grp.add_dims, gap, dt)
        hist, pdims_h, psizes_h, n_rep_h = _pool(ds["hist"], grp.add_dims, gap, dt)
        if n_rep_h != n_rep:
            raise ValueError("ref and hist must share their pooled dimensions")
        grp = _PooledGrouper(grp, time, n_rep, gap)
        l4_time = grp.pooled_time()
    else:
        ref, pdims, psizes = _series(ds["ref"], dt)
        hist, pdims_h, psizes_h = _series(ds["hist"].transpose(*ds["ref"].dims), dt)
    if (pdims_h, psizes_h) != (pdims, psizes):
        raise ValueError("ref and hist must share their non-time dimensions")
    kw = dict(adapt_freq_thresh=_thresh(adapt_freq_thresh, ds["hist"]),
              jitter_under_thresh_value=_thresh(jitter_under_thresh_value, ds["hist"]),
              jitter_over_thresh_value=_thresh(jitter_over_thresh_value, ds["hist"]),
              jitter_over_thresh_upper_bnd=_thresh(jitter_over_thresh_upper_bnd, ds["hist"]),
              max_tail_factor=max_tail_factor)
    fn = L4.dqm_train if normalize else L4.eqm_train
    res = fn(L4.Dataset({"ref": ref, "hist": hist}, time=l4_time, time_axis=0), group=grp, kind=kind,
             quantiles=np.asarray(quantiles), **kw)
    G, nq = res["af"].shape[-2], res["af"].shape[-1]
    shp = psizes + (G, nq)
    tdims = pdims + [grp.prop, "quantiles"]
    gdims = pdims + [grp.prop]
    nan_t = np.full(shp, np.nan, dt)
    nan_g = np.full(psizes + (G,), np.nan, dt)
    coords = {**_point_coords(ds, pdims), "quantiles": np.asarray(quantiles), grp.prop: grp.get_coordinate(time)}
    out = {
        "af": (tdims, _np(res["af"]).reshape(shp)),
        "hist_q": (tdims, _np(res["hist_q"]).reshape(shp)),
        "hist_q_raw": (tdims, nan_t if res.get("hist_q_raw") is None else _np(res["hist_q_raw"]).reshape(shp)),
        "P0_ref": (gdims, nan_g if res.get("P0_ref") is None else _np(res["P0_ref"]).reshape(psizes + (G,)).astype(dt)),
        "P0_hist": (gdims, nan_g if res.get("P0_hist") is None else _np(res["P0_hist"]).reshape(psizes + (G,)).astype(dt)),
        "pth": (gdims, nan_g if res.get("pth") is None else _np(res["pth"]).reshape(psizes + (G,))),
    }
    if normalize:
        out["scaling"] = (gdims, _np(res["scaling"]).reshape(psizes + (G,)))
    return _dataset(ds, out, coords)


def eqm_train(ds, *, group, kind, quantiles, adapt_freq_thresh=None, jitter_under_thresh_value=None,
              jitter_over_thresh_value=None, jitter_over_thresh_upper_bnd=None, max_tail_factor=None):
    """``xsdba._adjustment.eqm_train`` (_adjustment.py:193-286)."""
    return _train(ds, group=group, kind=kind, quantiles=quantiles, normalize=False, adapt_freq_thresh=adapt_freq_thresh,
                  jitter_under_thresh_value=jitter_under_thresh_value, jitter_over_thresh_value=jitter_over_thresh_value,
                  jitter_over_thresh_upper_bnd=jitter_over_thresh_upper_bnd, max_tail_factor=max_tail_factor)


def dqm_train(ds, *, group, kind, quantiles, adapt_freq_thresh=None, jitter_under_thresh_value=None,
              jitter_over_thresh_value=None, jitter_over_thresh_upper_bnd=None, max_tail_factor=None):
    """``xsdba._adjustment.dqm_train`` (_adjustment.py:86-190)."""
    return _train(ds, group=group, kind=kind, quantiles=quantiles, normalize=True, adapt_freq_thresh=adapt_freq_thresh,
                  jitter_under_thresh_value=jitter_under_thresh_value, jitter_over_thresh_value=jitter_over_thresh_value,
                  jitter_over_thresh_upper_bnd=jitter_over_thresh_upper_bnd, max_tail_factor=max_tail_factor)


def _adjust_group(group, pooled_ranks=False):
    """The grouper of an adjust call: add_dims only matter to group-wise reductions, and adjust has one -- the ranks of
    qdm_adjust with rank_window=True (Grouper.apply(main_only=False), base.py:410-415), not built for pooled dims."""
    grp = _group(group)
    if grp.add_dims:
        if pooled_ranks:
            raise NotImplementedError("qdm_adjust(rank_window=True) with add_dims (ranks pooled over the extra "
                                      "dimensions) is not built in xsdba_b200")
        grp = Grouper(grp.name, window=grp.window)
    return grp


def _adjust_inputs(ds, grp, names, extra=()):
    """sim as (time, points); trained tables as (points, ...).  Dimensions of sim that the tables do not have (the
    add_dims pooled at training time, base.py:410-415) are adjusted with the same table: the tables are repeated along
    them."""
    dt = _widest(ds, ("sim", "af"))
    sim, pdims, psizes = _series(ds["sim"], dt)
    tdims_of = lambda n: [d for d in pdims if d in ds[n].dims]  # noqa: E731
    rep_dims = [d for d in pdims if d not in ds[names[0]].dims]
    if rep_dims and list(pdims[:len(rep_dims)]) != rep_dims:   # pooled dims first, so that a repeat is a plain tile
        order = rep_dims + [d for d in pdims if d not in rep_dims]
        sim, pdims, psizes = _series(ds["sim"].transpose("time", *order), dt)
    n_rep = int(np.prod([psizes[pdims.index(d)] for d in rep_dims])) if rep_dims else 1

    def table(n, tail, tdt):
        t = _table(ds[n], tdims_of(n), tail, tdt)
        return np.ascontiguousarray(np.tile(t, (n_rep,) + (1,) * (t.ndim - 1))) if n_rep > 1 else t
    tables = {}
    for n in names:
        tables[n] = table(n, [grp.prop, "quantiles"], dt)
    for n in extra:
        if n in ds and not np.isnan(np.asarray(ds[n].values)).all():
            tail = [grp.prop, "quantiles"] if n == "hist_q_raw" else [grp.prop]
            tables[n] = table(n, tail, np.float64 if n.startswith("P0") else dt)
    return dt, sim, pdims, psizes, tables


def _series_out(ds, name, data, pdims, psizes, like="sim"):
    """(time, n_pts) result -> DataArray with the dims of ``ds[like]`` in their original order."""
    a = np.asarray(data).reshape((data.shape[0],) + psizes)
    order = ["time"] + pdims
    want = list(ds[like].dims)
    a = np.transpose(a, [order.index(d) for d in want])
    coords = {**_point_coords(ds, pdims), "time": ds["time"].values if "time" in getattr(ds, "coords", {}) else None}
    return want, a, {k: v for k, v in coords.items() if v is not None}


def qm_adjust(ds, *, group, interp, extrapolation, kind, adapt_freq_thresh=None, max_tail_factor=None):
    """``xsdba._adjustment.qm_adjust`` (_adjustment.py:594-676)."""
    grp = _adjust_group(group)
    dt, sim, pdims, psizes, tb = _adjust_inputs(ds, grp, ("af", "hist_q"), ("hist_q_raw", "P0_ref", "P0_hist", "pth"))
    res = L4.qm_adjust(L4.Dataset({"sim": sim, **tb}, time=time_axis_of(ds), time_axis=0), group=grp, interp=interp,
                       extrapolation=extrapolation, kind=kind, adapt_freq_thresh=_thresh(adapt_freq_thresh, ds["sim"]),
                       max_tail_factor=max_tail_factor)
    dims, a, coords = _series_out(ds, "scen", _np(res["scen"]), pdims, psizes)
    return _dataset(ds, {"scen": (dims, a)}, coords)


def qdm_adjust(ds, *, group, interp, extrapolation, kind, adapt_freq_thresh=None, rank_window=None,
               max_tail_factor=None):
    """``xsdba._adjustment.qdm_adjust`` (_adjustment.py:783-886)."""
    grp = _adjust_group(group, pooled_ranks=bool(rank_window))
    dt, sim, pdims, psizes, tb = _adjust_inputs(ds, grp, ("af",), ("hist_q_raw", "P0_ref", "P0_hist", "pth"))
    q = np.asarray(ds["quantiles"].values)
    res = L4.qdm_adjust(L4.Dataset({"sim": sim, "quantiles": q, **tb}, time=time_axis_of(ds), time_axis=0), group=grp,
                        interp=interp, extrapolation=extrapolation, kind=kind, rank_window=rank_window,
                        adapt_freq_thresh=_thresh(adapt_freq_thresh, ds["sim"]), max_tail_factor=max_tail_factor)
    dims, a, coords = _series_out(ds, "scen", _np(res["scen"]), pdims, psizes)
    _, b, _ = _series_out(ds, "sim_q", _np(res["sim_q"]), pdims, psizes)
    return _dataset(ds, {"scen": (dims, a), "sim_q": (dims, b)}, coords)


def _detrend(detrend, kind, grp):
    """int / reference ``BaseDetrend`` object (parameters read like base.py:26-58) -> local detrend object."""
    if isinstance(detrend, (int, np.integer, PolyDetrend, LoessDetrend)):
        return detrend
    name = type(detrend).__name__
    params = dict(getattr(detrend, "parameters", {}) or {})
    g = _group(params.get("group", getattr(detrend, "group", "time")))
    k = params.get("kind", getattr(detrend, "kind", kind))
    if name == "PolyDetrend":
        return PolyDetrend(degree=int(params.get("degree", getattr(detrend, "degree", 4))), kind=k, group=g,
                           preserve_mean=bool(params.get("preserve_mean", False)))
    if name == "LoessDetrend":
        return LoessDetrend(group=g, kind=k, f=params.get("f", 0.2), niter=params.get("niter", 1), d=params.get("d", 0),
                            weights=params.get("weights", "tricube"), equal_spacing=params.get("equal_spacing"),
                            skipna=params.get("skipna", True))
    raise NotImplementedError(f"detrending with {name} is not built in xsdba_b200")


def dqm_adjust(ds, *, group, interp, extrapolation, kind, detrend=1, adapt_freq_thresh=None, max_tail_factor=None):
    """``xsdba._adjustment.dqm_adjust`` (_adjustment.py:679-780)."""
    grp = _group(group)
    if grp.add_dims:
        raise NotImplementedError("dqm_adjust with add_dims is not built in xsdba_b200")
    dt, sim, pdims, psizes, tb = _adjust_inputs(ds, grp, ("af", "hist_q"), ("hist_q_raw", "P0_ref", "P0_hist", "pth"))
    tb["scaling"] = _table(ds["scaling"], pdims, [grp.prop], dt)
    res = L4.dqm_adjust(L4.Dataset({"sim": sim, **tb}, time=time_axis_of(ds), time_axis=0), group=grp, interp=interp,
                        extrapolation=extrapolation, kind=kind, detrend=_detrend(detrend, kind, grp),
                        adapt_freq_thresh=_thresh(adapt_freq_thresh, ds["sim"]), max_tail_factor=max_tail_factor)
    dims, a, coords = _series_out(ds, "scen", _np(res["scen"]), pdims, psizes)
    _, b, _ = _series_out(ds, "trend", _np(res["trend"]), pdims, psizes)
    return _dataset(ds, {"scen": (dims, a), "trend": (dims, b)}, coords)


for _f in (eqm_train, dqm_train, qm_adjust, dqm_adjust, qdm_adjust):
    _f.func = _f   # base.py:723, 775: the undecorated callable (nothing wraps these: they take whole arrays)


# ------------------------------------------------------------------------------------------------------------
# patch / unpatch
# ------------------------------------------------------------------------------------------------------------
def patch(modules=None):
    """Rebind the five names in ``xsdba.adjustment`` (where the Adjustment classes look them up, adjustment.py:18-35)
    and in ``xsdba._adjustment``.  ``modules`` (for tests) replaces the list of module objects to patch."""
    if modules is None:
        import importlib
        modules = [importlib.import_module("xsdba.adjustment"), importlib.import_module("xsdba._adjustment")]
    g = globals()
    for mod in modules:
        saved = _PATCHED.setdefault(mod, {})
        for name in NAMES:
            if name not in saved:
                saved[name] = getattr(mod, name, None)
            setattr(mod, name, g[name])
    return list(modules)


def unpatch():
    for mod, saved in list(_PATCHED.items()):
        for name, orig in saved.items():
            if orig is None:
                if hasattr(mod, name):
                    delattr(mod, name)
            else:
                setattr(mod, name, orig)
        del _PATCHED[mod]
