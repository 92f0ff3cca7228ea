"""CPU oracle for the quantile-mapping hot path of xsdba  --  TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this file.  The product (``xsdba_b200``) never does; it fails loudly when its CUDA
library is missing.

What this is
------------
A numpy / SciPy restatement of the reference's algorithm for the path named by BASELINE.json
(EQM / DQM / QDM train + adjust).  Every function cites the reference file:line it follows
(paths relative to ``/root/reference/src/xsdba``).  Where the reference delegates arithmetic to a
third-party package that exists in this image (SciPy 1.18.1 ``interp1d`` / ``griddata``; numpy
``sort`` / ``interp``) the oracle calls that same package, so that the third-party semantics
(cKDTree nearest, interp1d mid-point rule ...) are the real ones and not a re-guess.
Where the reference delegates to packages that are *absent* here (xarray 2023.11+, bottleneck 1.3+,
numba-compiled kernels on the GPU box are available but the reference sources are not), the
published algorithm is restated:

* numba ``_nan_quantile_1d``  (nbutils.py:108-148)  -> :func:`nan_quantile`
* bottleneck ``nanrankdata`` through ``xarray.DataArray.rank`` (utils.py:573-638) -> :func:`rank_pct`
* xarray ``rolling(center=True).construct`` (base.py:261-265) -> :func:`window_gather`
* xarray ``groupby("time.month")`` etc. (base.py:267-345) -> :func:`group_index`
* xarray ``polyfit`` / ``polyval`` (detrending.py:196-208) -> :func:`poly_trend`
* numba ``_loess_nb`` (loess.py:49-179) -> :func:`loess_nb`

Parity pinning
--------------
``oracle/gen_golden.py`` loads the reference's own ``nbutils.py`` / ``utils.py`` / ``loess.py`` from
``/root/reference`` (stubbed imports, see ``oracle/ref_loader.py``), runs its numba / SciPy kernels
on seeded inputs and commits the results under ``tests/golden/``.  ``tests/test_oracle_golden.py``
checks this oracle against those fixtures (bit-exact for the float32 quantiles, ranks and nearest
lookups) and against the known-answer tests of the reference's own test-suite
(tests/test_utils.py:68-113,149-194; tests/test_nbutils.py:23-34; tests/test_loess.py:18-38;
tests/test_base.py:46-65; tests/test_processing.py:248-281).

Pinned arithmetic facts (measured against the reference's numba build in the build container,
numba 0.65 / LLVM, x86-64 with FMA):

* the virtual index is ``(n-1)*q`` rounded ONCE in float64 (LLVM folds
  ``n*q + (1 + q*(1-1-1)) - 1`` under ``reassoc``/``contract``), with ``q`` the data-dtype node
  promoted to float64;
* gamma is cast to the data dtype, ``1-gamma`` is evaluated in the data dtype;
* both lerp branches are fused multiply-adds (``fma(diff, gamma, left)`` /
  ``fma(-diff, 1-gamma, right)``).
"""
from __future__ import annotations

import warnings
from dataclasses import dataclass

import numpy as np

ADDITIVE = "+"
MULTIPLICATIVE = "*"

# ----------------------------------------------------------------------------------------------
# time axis and groups  (base.py:105-115, 207-230, 274-345)
# ----------------------------------------------------------------------------------------------

MAX_DOY = {
    "standard": 366, "gregorian": 366, "proleptic_gregorian": 366, "julian": 366,
    "noleap": 365, "365_day": 365, "all_leap": 366, "366_day": 366, "360_day": 360,
}  # base.py:105-115

_DIM_NOLEAP = np.array([31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31])


def _is_leap(year, calendar):
    if calendar in ("noleap", "365_day", "360_day"):
        return np.zeros_like(year, dtype=bool)
    if calendar in ("all_leap", "366_day"):
        return np.ones_like(year, dtype=bool)
    if calendar == "julian":
        return year % 4 == 0
    return (year % 4 == 0) & ((year % 100 != 0) | (year % 400 == 0))


@dataclass
class TimeAxis:
    """Daily time coordinate: what ``ds.time.dt`` / ``ds.indexes['time']`` give the reference."""

    year: np.ndarray
    month: np.ndarray
    day: np.ndarray
    dayofyear: np.ndarray
    days_in_month: np.ndarray
    calendar: str

    def __len__(self):
        return self.year.shape[0]

    def __getitem__(self, sl):
        return TimeAxis(self.year[sl], self.month[sl], self.day[sl], self.dayofyear[sl],
                        self.days_in_month[sl], self.calendar)


def daily_time_axis(start_year: int, n_years: int, calendar: str = "noleap") -> TimeAxis:
    """All days of ``n_years`` consecutive years starting 1 January ``start_year``."""
    ys, ms, ds, doys, dims = [], [], [], [], []
    for y in range(start_year, start_year + n_years):
        if calendar == "360_day":
            mlen = np.full(12, 30)
        else:
            mlen = _DIM_NOLEAP.copy()
            if _is_leap(np.array([y]), calendar)[0]:
                mlen[1] = 29
        doy = 1
        for m in range(12):
            n = int(mlen[m])
            ys.append(np.full(n, y)); ms.append(np.full(n, m + 1)); ds.append(np.arange(1, n + 1))
            doys.append(np.arange(doy, doy + n)); dims.append(np.full(n, n))
            doy += n
    cat = lambda l: np.concatenate(l).astype(np.int64)
    return TimeAxis(cat(ys), cat(ms), cat(ds), cat(doys), cat(dims), calendar)


def group_index(time: TimeAxis, group: str):
    """0-based group index of each time step, number of groups and the group coordinate.

    Follows ``Grouper.get_index(interp=False)`` (base.py:321-329) and ``Grouper.get_coordinate``
    (base.py:207-230).  ``group`` is ``"time"``, ``"time.month"``, ``"time.dayofyear"`` or
    ``"time.season"``.
    """
    if group == "time":
        return np.zeros(len(time), np.int32), 1, np.array([1])
    prop = group.split(".", 1)[1]
    if prop == "month":
        return (time.month - 1).astype(np.int32), 12, np.arange(1, 13)
    if prop == "dayofyear":
        mdoy = MAX_DOY[time.calendar]
        return (time.dayofyear - 1).astype(np.int32), mdoy, np.arange(1, mdoy + 1)
    if prop == "season":
        return (time.month % 12 // 3).astype(np.int32), 4, np.arange(4)
    raise NotImplementedError(group)


def group_index_interp(time: TimeAxis, group: str) -> np.ndarray:
    """Fractional group coordinate used when ``interp != 'nearest'`` (base.py:306-320)."""
    prop = group.split(".", 1)[1]
    if prop == "month":
        return time.month - 0.5 + time.day / time.days_in_month
    if prop == "dayofyear":
        return time.dayofyear.astype(np.float64)
    if prop == "season":
        if time.calendar == "360_day":
            length_year = 360
        else:
            length_year = 365 + (0 if time.calendar == "noleap" else _is_leap(time.year, time.calendar))
        return time.dayofyear / length_year * 4 - 1 / 6
    raise ValueError(prop)


def window_gather(x: np.ndarray, window: int) -> np.ndarray:
    """``x.rolling(time=W, center=True).construct('window')`` on the last axis (base.py:261-265).

    Positional, NaN padded: ``out[..., t, j] = x[..., t - W//2 + j]``.
    """
    if window == 1:
        return x[..., None]
    T = x.shape[-1]
    half = window // 2
    pad = [(0, 0)] * (x.ndim - 1) + [(half, window - 1 - half)]
    xp = np.pad(x, pad, constant_values=np.nan)
    idx = np.arange(T)[:, None] + np.arange(window)[None, :]
    return xp[..., idx]


def group_segment(x: np.ndarray, gidx: np.ndarray, g: int, window: int) -> np.ndarray:
    """All samples the reference's reducing functions see for group ``g``: dims [time_g, window]
    flattened (``Grouper.apply`` with ``main_only=False``, base.py:410-420)."""
    xw = window_gather(x, window)  # [..., T, W]
    sel = np.nonzero(gidx == g)[0]
    seg = xw[..., sel, :]
    return seg.reshape(*x.shape[:-1], -1)


# ----------------------------------------------------------------------------------------------
# quantiles  (nbutils.py:24-148, 198-221; utils.py:251-281)
# ----------------------------------------------------------------------------------------------

def equally_spaced_nodes(n: int, eps=None) -> np.ndarray:
    """utils.py:251-281."""
    dq = 1 / n / 2
    q = np.linspace(dq, 1 - dq, n)
    if eps is None:
        return q
    return np.insert(np.append(q, 1 - eps), 0, eps)


def _split(a):
    c = 134217729.0 * a  # 2**27 + 1 (Veltkamp)
    hi = c - (c - a)
    return hi, a - hi


def _fma64(a, b, c):
    """float64 fused multiply-add emulated with error-free transformations (Dekker/Knuth).
    Faithful to ~1 ulp of the exact fma; the C oracle uses the hardware ``fma`` instead."""
    a, b, c = np.broadcast_arrays(np.asarray(a, np.float64), np.asarray(b, np.float64), np.asarray(c, np.float64))
    with np.errstate(all="ignore"):
        p = a * b
        ah, al = _split(a)
        bh, bl = _split(b)
        e = ((ah * bh - p) + ah * bl + al * bh) + al * bl
        s = p + c
        bb = s - p
        t = (p - (s - bb)) + (c - bb)
        r = s + (t + e)
    bad = ~np.isfinite(r) | ~np.isfinite(e)
    return np.where(bad, p + c, r)


def _fma(a, b, c, dtype):
    if np.dtype(dtype) == np.float32:
        # product of two float32 is exact in float64; one extra rounding at 2^-29 probability
        return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)
    return _fma64(a, b, c)


def nan_quantile(arr: np.ndarray, q: np.ndarray, cast_q: bool = True) -> np.ndarray:
    """Type-7 NaN-aware quantiles of each row of ``arr`` [rows, S] -> [rows, nq].

    Restates ``_nan_quantile_1d`` + ``_get_indexes`` + ``_linear_interpolation`` +
    ``_wrapper_quantile1d`` (nbutils.py:24-148, 198-203) with the arithmetic pinned in the module
    docstring.  ``q`` is cast to the data dtype first (nbutils.py:253).
    """
    arr = np.asarray(arr)
    dt = arr.dtype
    # nbutils.quantile casts q to the data dtype (nbutils.py:253); direct _quantile callers
    # (_npdft_train, _adjustment.py:315) pass float64 nodes un-cast
    q = np.asarray(q, dtype=dt) if cast_q else np.asarray(q, dtype=np.float64)
    rows, S = arr.shape
    if S == 0:
        return np.full((rows, q.size), np.nan, dt)
    s = np.sort(arr, axis=1)  # NaNs last, like numba's ndarray.sort
    n = (~np.isnan(arr)).sum(axis=1).astype(np.int64)  # nbutils.py:128
    q64 = q.astype(np.float64)
    vi = (n[:, None] - 1).astype(np.float64) * q64[None, :]  # nbutils.py:131 (LLVM-folded)
    prev = np.floor(vi)
    nxt = prev + 1
    above = vi >= (n[:, None] - 1)  # nbutils.py:47-51
    prev[above] = -1
    nxt[above] = -1
    below = vi < 0  # nbutils.py:53-56
    prev[below] = 0
    nxt[below] = 0
    prev_i = prev.astype(np.intp)
    nxt_i = nxt.astype(np.intp)
    left = np.take_along_axis(s, prev_i % S, 1)
    right = np.take_along_axis(s, nxt_i % S, 1)
    gamma = (vi - prev_i).astype(dt)  # nbutils.py:142
    with np.errstate(invalid="ignore", over="ignore"):
        diff = (right - left).astype(dt)
        lo = _fma(diff, gamma, left, dt)  # nbutils.py:101-102
        hi = _fma(-diff, (dt.type(1) - gamma).astype(dt), right, dt)  # nbutils.py:103-104
    res = np.where(gamma >= 0.5, hi, lo).astype(dt)
    mx = np.take_along_axis(s, ((n - 1) % S)[:, None], 1)  # nbutils.py:146
    return np.where(np.isnan(res), mx, res).astype(dt)


def vecquantiles(arr: np.ndarray, rnk: np.ndarray) -> np.ndarray:
    """nbutils.py:151-161: one ``np.nanquantile(row, rnk[row])`` per row, NaN rank -> NaN."""
    out = np.full(arr.shape[0], np.nan, arr.dtype)
    for i in range(arr.shape[0]):
        if not np.isnan(rnk[i]):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                out[i] = np.nanquantile(arr[i], rnk[i])
    return out


# ----------------------------------------------------------------------------------------------
# corrections  (utils.py:130-177)
# ----------------------------------------------------------------------------------------------

def get_correction(x, y, kind):
    """utils.py:130-143: ``y - x`` or ``y / x``."""
    with np.errstate(all="ignore"):
        return y - x if kind == ADDITIVE else y / x


def apply_correction(x, factor, kind):
    """utils.py:146-162: ``x + factor`` or ``x * factor``."""
    with np.errstate(all="ignore"):
        return x + factor if kind == ADDITIVE else x * factor


def invert(x, kind):
    """utils.py:165-177: ``-x`` or ``1 / x``."""
    with np.errstate(all="ignore"):
        return -x if kind == ADDITIVE else 1 / x


# ----------------------------------------------------------------------------------------------
# train  (_adjustment.py:86-286)
# ----------------------------------------------------------------------------------------------

def eqm_train(ref, hist, gidx, n_groups, window, q, kind):
    """``eqm_train`` over all groups (_adjustment.py:253-286) for point-major inputs [N, T].

    Returns ``af``, ``hist_q`` of shape [N, G, nq] in the data dtype.
    """
    dt = ref.dtype
    N = ref.shape[0]
    q = np.asarray(q, dt)
    af = np.full((N, n_groups, q.size), np.nan, dt)
    hq = np.full((N, n_groups, q.size), np.nan, dt)
    for g in range(n_groups):
        if not np.any(gidx == g):
            continue
        ref_q = nan_quantile(group_segment(ref, gidx, g, window), q)    # _adjustment.py:271
        hist_q = nan_quantile(group_segment(hist, gidx, g, window), q)  # _adjustment.py:272
        af[:, g] = get_correction(hist_q, ref_q, kind)                  # _adjustment.py:276
        hq[:, g] = hist_q
    return af, hq


def _nanmean_rows(seg):
    """xarray ``.mean(dim)`` (skipna) of each row; accumulation order of the reference is not
    pinned -- float64 accumulation, cast back to the data dtype."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return np.nanmean(seg.astype(np.float64), axis=1).astype(seg.dtype)


def dqm_train(ref, hist, gidx, n_groups, window, q, kind):
    """``dqm_train`` over all groups (_adjustment.py:150-190). Returns af, hist_q [N,G,nq], scaling [N,G]."""
    dt = ref.dtype
    N = ref.shape[0]
    q = np.asarray(q, dt)
    af = np.full((N, n_groups, q.size), np.nan, dt)
    hq = np.full((N, n_groups, q.size), np.nan, dt)
    sc = np.full((N, n_groups), np.nan, dt)
    for g in range(n_groups):
        if not np.any(gidx == g):
            continue
        rseg = group_segment(ref, gidx, g, window)
        hseg = group_segment(hist, gidx, g, window)
        mu_ref = _nanmean_rows(rseg)
        mu_hist = _nanmean_rows(hseg)
        refn = apply_correction(rseg, invert(mu_ref, kind)[:, None].astype(dt), kind)    # :167
        histn = apply_correction(hseg, invert(mu_hist, kind)[:, None].astype(dt), kind)  # :168
        ref_q = nan_quantile(refn.astype(dt), q)
        hist_q = nan_quantile(histn.astype(dt), q)
        af[:, g] = get_correction(hist_q, ref_q, kind)
        hq[:, g] = hist_q
        sc[:, g] = get_correction(mu_hist, mu_ref, kind)  # :177-179
    return af, hq, sc


# ----------------------------------------------------------------------------------------------
# adjust: factor lookup  (utils.py:284-513, nbutils.py:375-416)
# ----------------------------------------------------------------------------------------------

def interp_on_quantiles_1d(newx, oldx, oldy, method, extrap):
    """utils.py:350-377 -- the third-party arithmetic is SciPy's own ``interp1d``."""
    from scipy.interpolate import interp1d

    mask_new = np.isnan(newx)
    mask_old = np.isnan(oldy) | np.isnan(oldx)
    out = np.full_like(newx, np.nan, dtype=f"float{oldy.dtype.itemsize * 8}")
    if np.all(mask_new) or np.all(mask_old):
        return out
    if extrap == "constant":
        fill_value = (oldy[~np.isnan(oldy)][0], oldy[~np.isnan(oldy)][-1])
    else:
        fill_value = np.nan
    with np.errstate(all="ignore"):
        out[~mask_new] = interp1d(oldx[~mask_old], oldy[~mask_old], kind=method, bounds_error=False,
                                  fill_value=fill_value)(newx[~mask_new])
    return out


def first_and_last_nonnull(arr):
    """nbutils.py:375-389."""
    out = np.empty((arr.shape[0], 2))
    for i in range(arr.shape[0]):
        idxs = np.where(~np.isnan(arr[i]))[0]
        if idxs.size > 0:
            out[i] = arr[i][idxs[np.array([0, -1])]]
        else:
            out[i] = np.nan
    return out


def extrapolate_on_quantiles(interp, oldx, oldg, oldy, newx, newg, method="constant"):
    """nbutils.py:392-416 (``np.interp`` is numpy's own)."""
    bnds = first_and_last_nonnull(oldx)
    xp = oldg[:, 0]
    with np.errstate(all="ignore"):
        toolow = newx < np.interp(newg, xp, bnds[:, 0])
        toohigh = newx > np.interp(newg, xp, bnds[:, 1])
        if method == "constant":
            constants = first_and_last_nonnull(oldy)
            cnstlow = np.interp(newg, xp, constants[:, 0])
            cnsthigh = np.interp(newg, xp, constants[:, 1])
            interp[toolow] = cnstlow[toolow]
            interp[toohigh] = cnsthigh[toohigh]
        else:
            interp[toolow] = np.nan
            interp[toohigh] = np.nan
    return interp


def interp_on_quantiles_2d(newx, newg, oldx, oldy, oldg, method, extrap):
    """utils.py:380-400 -- ``griddata`` is SciPy's own (cKDTree / Qhull)."""
    from scipy.interpolate import griddata

    mask_new = np.isnan(newx) | np.isnan(newg)
    mask_old = np.isnan(oldy) | np.isnan(oldx) | np.isnan(oldg)
    out = np.full_like(newx, np.nan, dtype=f"float{oldy.dtype.itemsize * 8}")
    if np.all(mask_new) or np.all(mask_old):
        return out
    out[~mask_new] = griddata((oldx[~mask_old], oldg[~mask_old]), oldy[~mask_old],
                              (newx[~mask_new], newg[~mask_new]), method=method)
    if method == "nearest" or extrap != "nan":
        out = extrapolate_on_quantiles(out, oldx, oldg, oldy, newx, newg, extrap)
    return out


def add_cyclic_bounds(tab, coords):
    """utils.py:284-314 with ``cyclic_coords=False``: wrap-pad one row each side of the group axis
    (axis -2 of ``tab`` [..., G, nq]); the new coordinates continue the neighbour step."""
    padded = np.concatenate([tab[..., -1:, :], tab, tab[..., :1, :]], axis=-2)
    c = np.asarray(coords, np.float64)
    if c.size > 1:
        cc = np.concatenate([[c[0] - (c[1] - c[0])], c, [c[-1] + (c[-1] - c[-2])]])
    else:
        cc = np.concatenate([[c[0] - 1], c, [c[0] + 1]])
    return padded, cc


def interp_on_quantiles(newx, xq, yq, *, group, time, method, extrapolation):
    """utils.py:408-513 for point-major arrays: newx [N,T]; xq,yq [N,G,nq] (or xq [nq] shared)."""
    N, T = newx.shape
    odt = np.dtype(f"float{yq.dtype.itemsize * 8}")
    out = np.empty((N, T), odt)
    if group == "time":
        for i in range(N):
            x_i = xq if xq.ndim == 1 else xq[i, 0]
            out[i] = interp_on_quantiles_1d(newx[i], x_i, yq[i, 0], method, extrapolation)
        return out
    _, G, coords = group_index(time, group)
    if method == "nearest":
        newg = group_index(time, group)[0].astype(np.int64) + (0 if group.endswith("season") else 1)
    else:
        newg = group_index_interp(time, group)
    for i in range(N):
        x_i = np.broadcast_to(xq, (G, yq.shape[-1])) if xq.ndim == 1 else xq[i]
        oldx, cc = add_cyclic_bounds(x_i, coords)
        oldy, _ = add_cyclic_bounds(yq[i], coords)
        oldg = np.broadcast_to(cc[:, None], oldx.shape)
        out[i] = interp_on_quantiles_2d(newx[i], newg, oldx, oldy, oldg, method, extrapolation)
    return out


# ----------------------------------------------------------------------------------------------
# ranks  (utils.py:573-646)
# ----------------------------------------------------------------------------------------------

def nanrankdata(a):
    """bottleneck.nanrankdata (>=1.3.1) along the last axis: average ties, 1-based, NaN -> NaN."""
    from scipy.stats import rankdata

    a = np.asarray(a)
    if a.shape[-1] == 0:
        return np.empty(a.shape, np.float64)
    return rankdata(a, method="average", axis=-1, nan_policy="omit").astype(np.float64)


def rank_pct(seg):
    """``utils.rank(da, dim, pct=True)`` on the last axis (utils.py:612-638)."""
    with np.errstate(all="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = nanrankdata(seg)
        cnt = (~np.isnan(seg)).sum(axis=-1, keepdims=True)
        r = r / cnt
        mn = np.nanmin(r, axis=-1, keepdims=True)
        mx = np.nanmax(r, axis=-1, keepdims=True)
        return mx * (r - mn) / (mx - mn)


def rank_bn(arr):
    """utils.py:641-646 (MBCn flavour) along the last axis."""
    with np.errstate(all="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        rnk = nanrankdata(arr)
        rnk = rnk / np.nanmax(rnk, axis=-1, keepdims=True)
        mn = np.nanmin(rnk, axis=-1, keepdims=True)
        return 1 * (rnk - mn) / (1 - mn)


def grouped_rank_pct(sim, gidx, n_groups, window, rank_window):
    """``group.apply(u.rank, sim, main_only=not rank_window, pct=True)`` (_adjustment.py:872):
    ranks inside each exact group, or over the window with the centre column kept
    (base.py:438-439).  Returns float64 [N, T]."""
    N, T = sim.shape
    out = np.full((N, T), np.nan, np.float64)
    for g in range(n_groups):
        sel = np.nonzero(gidx == g)[0]
        if sel.size == 0:
            continue
        if rank_window and window > 1:
            xw = window_gather(sim, window)[:, sel, :]  # [N, n_g, W]
            r = rank_pct(xw.reshape(N, -1)).reshape(N, sel.size, window)
            out[:, sel] = r[:, :, window // 2]
        else:
            out[:, sel] = rank_pct(sim[:, sel])
    return out


# ----------------------------------------------------------------------------------------------
# adjust  (_adjustment.py:594-886)
# ----------------------------------------------------------------------------------------------

def qm_adjust(sim, af, hist_q, *, group, time, interp, extrapolation, kind):
    """``qm_adjust.func`` without the optional adapt_freq / max_tail_factor steps
    (_adjustment.py:660-669)."""
    afi = interp_on_quantiles(sim, hist_q, af, group=group, time=time, method=interp,
                              extrapolation=extrapolation)
    return apply_correction(sim, afi.astype(af.dtype), kind).astype(sim.dtype)


def qdm_adjust(sim, af, quantiles, *, group, time, window, interp, extrapolation, kind, rank_window=False):
    """``qdm_adjust.func`` (_adjustment.py:872-881). Returns scen [N,T] and sim_q float64 [N,T]."""
    gidx, G, _ = group_index(time, group)
    sim_q = grouped_rank_pct(sim, gidx, G, window, rank_window)
    afi = interp_on_quantiles(sim_q, np.asarray(quantiles), af, group=group, time=time, method=interp,
                              extrapolation=extrapolation)
    return apply_correction(sim, afi.astype(af.dtype), kind).astype(sim.dtype), sim_q


def broadcast_nearest(grouped, gidx):
    """``utils.broadcast(grouped, x, group, interp='nearest')`` (utils.py:209-219): pick the value of
    each time step's own group. grouped [N, G] -> [N, T]."""
    return grouped[:, gidx]


def broadcast_month_linear(grouped, time):
    """``utils.broadcast`` with ``interp='linear'`` for month groups (utils.py:220-236): cyclic pad
    then xarray ``interp`` (scipy ``interp1d`` linear on float64 coords), cast to the input dtype."""
    g = np.concatenate([grouped[:, -1:], grouped, grouped[:, :1]], axis=1).astype(np.float64)
    xp = np.arange(0, 14, dtype=np.float64)
    newg = group_index_interp(time, "time.month")
    out = np.empty((grouped.shape[0], len(time)), np.float64)
    for i in range(grouped.shape[0]):
        out[i] = np.interp(newg, xp, g[i])
    return out.astype(grouped.dtype)


# ----------------------------------------------------------------------------------------------
# detrending  (detrending.py:165-296, loess.py:16-279)
# ----------------------------------------------------------------------------------------------

def poly_trend(y, x, degree):
    """``PolyDetrend`` fit+evaluate on one series (detrending.py:196-208 -> xarray polyfit/polyval).

    xarray.polyfit: Vandermonde of the float64 coordinate, columns scaled by their 2-norm, numpy
    ``lstsq`` (rows with NaN in y dropped when skipna), then polyval by Horner.  ``x`` is the time
    coordinate as float64 (xarray uses ns since 1970 for datetime64, days-based for cftime; only
    the fitted *values* matter, not the scale, up to conditioning).
    """
    x = np.asarray(x, np.float64)
    y64 = np.asarray(y, np.float64)
    ok = ~np.isnan(y64)
    if ok.sum() <= degree:
        return np.full_like(y64, np.nan)
    V = np.vander(x, degree + 1)  # highest power first
    scale = np.sqrt((V[ok] * V[ok]).sum(axis=0))
    coef, *_ = np.linalg.lstsq(V[ok] / scale, y64[ok], rcond=None)
    coef = coef / scale
    out = np.zeros_like(x)
    for c in coef:  # Horner, highest degree first
        out = out * x + c
    return out


def _tricube(x):
    w = (1 - x ** 3) ** 3
    w[x >= 1] = 0
    return w


def _gaussian(x):
    w = np.exp(-(x ** 2) / (2 * (1 / 1.96) ** 2))
    w[x >= 1] = 0
    return w


def loess_nb(x, y, f=0.5, niter=2, weights="tricube", d=1, dx=0.0, skipna=True):
    """``_loess_nb`` (loess.py:49-179) in float64 -- Python loop over points (small cases only).

    ``d`` selects ``_constant_regression`` (0, loess.py:38-39) or ``_linear_regression``
    (1, loess.py:42-46); ``dx > 0`` is the equal-spacing branch (loess.py:112-119, 136-150).
    """
    wf = _tricube if weights == "tricube" else _gaussian
    x = np.asarray(x, np.float64)
    y = np.asarray(y, np.float64)
    if skipna:
        nan = np.isnan(y)
        out = np.full(x.size, np.nan)
        y = y[~nan]
        x = x[~nan]
        if x.size == 0:
            return out
    n = x.size
    yest = np.zeros(n)
    delta = np.ones(n)
    if dx == 0:
        r = int(np.round(f * n))
        HW = min(r + 2, n)
        R = min(2 * HW, n)
    else:
        r = int(2 * (f * n // 2) + 1)
        hw = int((r - 1) / 2)
        R = min(r + 4, n)
        HW = hw + 2
    wi = None
    for iteration in range(niter):
        for i in range(n):
            if i < HW:
                sl = slice(0, R)
            elif i >= n - HW - 1:
                sl = slice(n - R, n)
            else:
                sl = slice(i - HW, i + HW + 1)
            xi, yi, di = x[sl], y[sl], delta[sl]
            if dx > 0:
                if i <= HW or i >= n - HW:
                    diffs = np.abs(xi - x[i])
                    if i < hw:
                        h = (r - i) * dx
                    elif i >= n - hw:
                        h = (i - (n - r) + 1) * dx
                    else:
                        h = (hw + 1) * dx
                    wi = wf(diffs / h)
                w = di * wi
            else:
                diffs = np.abs(xi - x[i])
                h = np.sort(diffs)[r]
                w = di * wf(diffs / h)
            if d == 0:
                yest[i] = (w * yi).sum() / w.sum()
            else:
                b = np.array([np.sum(w * yi), np.sum(w * yi * xi)])
                A = np.array([[np.sum(w), np.sum(w * xi)], [np.sum(w * xi), np.sum(w * xi * xi)]])
                beta = np.linalg.solve(A, b)
                yest[i] = beta[0] + beta[1] * x[i]
        if iteration < niter - 1:
            residuals = y - yest
            s = np.median(np.abs(residuals))
            if s == 0:
                xres = (residuals != 0) * 1.0
            else:
                xres = residuals / (6.0 * s)
            delta = (1 - xres ** 2) ** 2
            delta[np.abs(xres) >= 1] = 0
    if skipna:
        out[~nan] = yest
        return out
    return yest


def loess_smoothing(y, time_coord, d=1, f=0.5, niter=2, weights="tricube", equal_spacing=None, skipna=True):
    """``loess_smoothing`` (loess.py:244-278) on one series: x rescaled to [0, 1]; equal spacing is
    decided from the coordinate (loess.py:251-260)."""
    t = np.asarray(time_coord, np.float64)
    x = (t - t[0]) / (t[-1] - t[0])
    diffx = np.diff(t)
    if np.all(diffx == diffx[0]) and equal_spacing is None:
        equal_spacing = True
    dx = float(x[1] - x[0]) if equal_spacing else 0
    return loess_nb(x, y, f=f, niter=niter, weights=weights, d=d, dx=dx, skipna=skipna)


# ----------------------------------------------------------------------------------------------
# tie detection for the 2-D nearest rule (test helper)
# ----------------------------------------------------------------------------------------------

def nearest_2d_candidates(newx, newg, oldx, oldy, oldg, chunk=256):
    """Brute-force version of the 2-D nearest rule: for every query return (ymin, ymax, n_tied) over
    ALL nodes at the minimal squared distance ``dx*dx + dg*dg`` (float64, SciPy's arithmetic).

    SciPy's cKDTree breaks exact distance ties by tree-traversal order, which the reference does not
    control ("parity unpinned on ties").  Tests use this to require bit-equality
    wherever the nearest node is unique and "one of the tied nodes" elsewhere.
    """
    m = ~(np.isnan(oldx) | np.isnan(oldy) | np.isnan(oldg))
    px = oldx[m].astype(np.float64)
    pg = oldg[m].astype(np.float64)
    py = oldy[m]
    n = newx.shape[0]
    ymin = np.full(n, np.nan)
    ymax = np.full(n, np.nan)
    ntied = np.zeros(n, np.int64)
    if px.size == 0:
        return ymin, ymax, ntied
    for s in range(0, n, chunk):
        x = newx[s:s + chunk].astype(np.float64)[:, None]
        g = np.asarray(newg[s:s + chunk], np.float64)[:, None]
        dx = x - px[None, :]
        dg = g - pg[None, :]
        d2 = dx * dx + dg * dg
        best = np.nanmin(np.where(np.isnan(d2), np.inf, d2), axis=1, keepdims=True)
        tied = d2 == best
        yy = np.where(tied, py[None, :].astype(np.float64), np.nan)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ymin[s:s + chunk] = np.nanmin(yy, axis=1)
            ymax[s:s + chunk] = np.nanmax(yy, axis=1)
        ntied[s:s + chunk] = tied.sum(axis=1)
    bad = np.isnan(newx) | np.isnan(np.asarray(newg, np.float64))
    ymin[bad] = np.nan
    ymax[bad] = np.nan
    return ymin, ymax, ntied


def qm_adjust_factor_bounds(newx, af, xq, *, group, time, extrapolation):
    """For grouped nearest lookups: (lo, hi) factor each value may legally take -- equal wherever the
    nearest node is unique, the min/max over the tied nodes otherwise; extrapolation applied.
    ``xq`` is hist_q [N,G,nq] (EQM/DQM) or the shared quantile axis [nq] (QDM, newx = sim_q)."""
    N, T = newx.shape
    _, G, coords = group_index(time, group)
    newg = group_index(time, group)[0].astype(np.int64) + (0 if group.endswith("season") else 1)
    lo = np.empty((N, T)); hi = np.empty((N, T))
    for i in range(N):
        x_i = np.broadcast_to(xq, (G, af.shape[-1])) if xq.ndim == 1 else xq[i]
        oldx, cc = add_cyclic_bounds(x_i, coords)
        oldy, _ = add_cyclic_bounds(af[i], coords)
        oldg = np.broadcast_to(cc[:, None], oldx.shape)
        a, b, _ = nearest_2d_candidates(newx[i], newg, oldx, oldy, oldg)
        lo[i] = extrapolate_on_quantiles(a, oldx, oldg, oldy, newx[i], newg, extrapolation)
        hi[i] = extrapolate_on_quantiles(b, oldx, oldg, oldy, newx[i], newg, extrapolation)
    return lo, hi


# ----------------------------------------------------------------------------------------------
# DQM adjust  (_adjustment.py:679-780; detrending.py:59-120, 165-296)
# ----------------------------------------------------------------------------------------------

def time_ordinal(time: TimeAxis) -> np.ndarray:
    """Days since the first step as float64 (x-axis of the trend fits; xarray uses ns since 1970,
    which is the same polynomial up to conditioning)."""
    y0 = int(time.year.min())
    years = np.arange(y0, int(time.year.max()) + 1)
    if time.calendar == "360_day":
        ylen = np.full(years.shape, 360)
    else:
        ylen = 365 + _is_leap(years, time.calendar).astype(np.int64)
    start = np.concatenate([[0], np.cumsum(ylen)[:-1]])
    o = start[time.year - y0] + time.dayofyear - 1
    return (o - o[0]).astype(np.float64)


def group_trend_poly(x, gidx, n_groups, window, tcoord, degree, preserve_mean=False, kind="+"):
    """``PolyDetrend(degree, group).fit(x).ds.trend`` (detrending.py:189-208 through map_groups /
    Grouper.apply, base.py:410-420): per group, window dims are averaged first (NaN-skipping), then a
    polynomial is fitted on the group's time steps and evaluated there.  Returns float64 [N, T]."""
    N, T = x.shape
    trend = np.full((N, T), np.nan)
    xw = window_gather(x, window) if window > 1 else None
    for g in range(n_groups):
        sel = np.nonzero(gidx == g)[0]
        if sel.size == 0:
            continue
        if window > 1:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                series = np.nanmean(xw[:, sel, :].astype(np.float64), axis=2)  # detrending.py:199-200
        else:
            series = x[:, sel].astype(np.float64)
        for i in range(N):
            trend[i, sel] = poly_trend(series[i], tcoord[sel], degree)
        if preserve_mean:   # detrending.py:205: trend (+|*) invert(mean of the group's trend)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                m = np.nanmean(trend[:, sel], axis=1, keepdims=True)
            trend[:, sel] = apply_correction(trend[:, sel], invert(m, kind), kind)
    return trend


def dqm_adjust(sim, af, hist_q, scaling, *, group, window, time, interp, extrapolation, kind, detrend=1,
               loess=None):
    """``dqm_adjust.func`` without adapt_freq / max_tail_factor (_adjustment.py:748-780).
    ``detrend`` is the PolyDetrend degree; ``loess`` (dict f, niter, d, weights) selects a
    ``LoessDetrend(group="time")`` instead.  Returns (scen float64 [N,T], trend float64 [N,T])."""
    gidx, G, _ = group_index(time, group)
    if group == "time":
        sc_b = scaling[:, :1]
    elif group.endswith("dayofyear") or interp == "nearest":
        sc_b = broadcast_nearest(scaling, gidx)                      # _adjustment.py:750-756
    else:
        sc_b = broadcast_month_linear(scaling, time)
    scaled = apply_correction(sim, sc_b, kind)                       # _adjustment.py:748-757
    tcoord = time_ordinal(time)
    if loess is not None:
        trend = np.stack([loess_smoothing(scaled[i], tcoord, **loess) for i in range(sim.shape[0])])
    else:
        trend = group_trend_poly(scaled, gidx, G, window, tcoord, detrend)  # _adjustment.py:759-765
    detr = apply_correction(scaled, invert(trend, kind), kind)       # detrending.py:99
    afi = interp_on_quantiles(detr, hist_q, af, group=group, time=time, method=interp, extrapolation=extrapolation)
    scen = apply_correction(detr, afi, kind)                         # qm_adjust.func, _adjustment.py:669
    scen = apply_correction(scen, trend, kind)                       # detrending.py:120
    return scen, trend


# ----------------------------------------------------------------------------------------------
# MBCn / N-pdf transform  (_adjustment.py:289-591; processing.py:323-350, 829-918; _processing.py:184-247)
# ----------------------------------------------------------------------------------------------

def rand_rot_matrices(n_var: int, n_iter: int, seed: int) -> np.ndarray:
    """``utils.rand_rot_matrix`` recipe (utils.py:961-973; Mezzadri 2007) with a seeded generator:
    float32 [n_iter, V, V]."""
    rng = np.random.default_rng(seed)
    out = np.empty((n_iter, n_var, n_var), np.float32)
    for i in range(n_iter):
        Z = rng.standard_normal((n_var, n_var))
        Q, R = np.linalg.qr(Z)
        num = np.diag(R)
        out[i] = (Q @ np.diag(num / np.abs(num))).astype(np.float32)
    return out


def mbcn_blocks(time: TimeAxis, group: str, window: int):
    """``grouped_time_indexes`` (processing.py:829-918) for "time" and "time.dayofyear": per block the
    windowed time indices (sorted, out-of-range dropped) and the exact-group indices."""
    T = len(time)
    if group == "time":
        return [(np.arange(T), np.arange(T))]
    gidx, G, _ = group_index(time, group)
    half = window // 2
    out = []
    for g in range(G):
        sel = np.nonzero(gidx == g)[0]
        if sel.size == 0:
            continue
        w = (sel[:, None] - half + np.arange(window)[None, :]).ravel()
        w = w[(w >= 0) & (w < T)]
        out.append((w, sel))
    return out


def _standardize(x):
    """``(x - nanmean) / nanstd`` along the last axis as ``_npdft_train`` computes it (_adjustment.py:303-305).  The
    rows are made C-contiguous first: numpy sums a contiguous axis pairwise and any other layout sequentially (the
    float32 mean of a 30-year series then differs by ~1e-5 relative), and ``x[:, i, idx]`` -- basic and advanced
    indexing mixed -- silently returns a Fortran-ordered array.  The N-pdf iteration amplifies such differences by
    orders of magnitude, so the layout is pinned here.  (``mbcn_adjust`` standardises ``sim`` through xarray's
    mean / std, i.e. bottleneck when installed -- absent from this image, unpinned; this is the stand-in.)"""
    x = np.ascontiguousarray(x)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return ((x - np.nanmean(x, axis=-1, keepdims=True)) / np.nanstd(x, axis=-1, keepdims=True)).astype(x.dtype)


def npdft_train(ref, hist, rots, quantiles, method="nearest", extrap="constant"):
    """``_npdft_train`` (_adjustment.py:289-328) without escores: ref, hist [V, T] -> af_q [n_iter, V, nq]."""
    ref = _standardize(ref)
    hist = _standardize(hist)
    quantiles = np.asarray(quantiles, np.float64)
    af_q = np.zeros((len(rots), ref.shape[0], len(quantiles)))
    for ii, _rot in enumerate(rots):
        rot = _rot if ii == 0 else _rot @ rots[ii - 1].T
        ref, hist = rot @ ref, rot @ hist
        for iv in range(ref.shape[0]):
            ref_q = nan_quantile(ref[iv][None, :], quantiles, cast_q=False)[0]
            hist_q = nan_quantile(hist[iv][None, :], quantiles, cast_q=False)[0]
            af_q[ii, iv] = ref_q - hist_q
            af = interp_on_quantiles_1d(rank_bn(hist[iv]), quantiles, af_q[ii, iv], method, extrap)
            hist[iv] = hist[iv] + af
    return af_q


def npdft_adjust(sim, af_q, rots, quantiles, method="nearest", extrap="constant"):
    """``_npdft_adjust`` (_adjustment.py:426-464): sim [V, T] (already standardized) -> [V, T]."""
    sim = sim.copy()[:, np.newaxis, :]  # the reference adds a dummy period dim (_adjustment.py:441-442)
    quantiles = np.asarray(quantiles, np.float64)
    for ii, _rot in enumerate(rots):
        rot = _rot if ii == 0 else _rot @ rots[ii - 1].T
        sim = np.einsum("ij,j...->i...", rot, sim)  # (einsum, not matmul: the float32 summation order differs)
        for iv in range(sim.shape[0]):
            af = interp_on_quantiles_1d(rank_bn(sim[iv, 0]), quantiles, af_q[ii, iv], method, extrap)
            sim[iv, 0] = sim[iv, 0] + af
    return np.einsum("ij,j...->i...", rots[-1].T, sim)[:, 0, :]


def reordering_1d(data, ordr):
    """_processing.py:204-205."""
    return np.sort(data)[np.argsort(np.argsort(ordr, kind="stable"), kind="stable")]


def mbcn_train(ref, hist, rots, quantiles, blocks, method="nearest", extrap="constant"):
    """``mbcn_train`` (_adjustment.py:385-421): ref, hist [V, N, T] -> af_q [n_blocks, N, n_iter, V, nq] (data dtype)."""
    V, N, T = ref.shape
    out = np.empty((len(blocks), N, len(rots), V, len(quantiles)), ref.dtype)
    for ib, (gw, _) in enumerate(blocks):
        for i in range(N):
            out[ib, i] = npdft_train(ref[:, i, gw].copy(), hist[:, i, gw].copy(), rots, quantiles, method, extrap)
    return out


def mbcn_adjust(ref, hist, sim, af_q, rots, quantiles, blocks, kinds, method="nearest", extrap="constant"):
    """``mbcn_adjust`` (_adjustment.py:528-591) with base = QuantileDeltaMapping(group="time"): [V, N, T]."""
    V, N, T = sim.shape
    dt = sim.dtype
    q_dt = np.asarray(quantiles).astype(dt)   # QDM._train casts the nodes (adjustment.py:480-483)
    scen = np.zeros_like(sim)
    for ib, (gw, g) in enumerate(blocks):
        keep = np.isin(gw, g)
        for i in range(N):
            scen_block = np.empty((V, gw.size), dt)
            for v in range(V):
                r, h, s_ = ref[v, i, gw][None], hist[v, i, gw][None], sim[v, i, gw][None]
                af, _ = eqm_train(r, h, np.zeros(gw.size, np.int32), 1, 1, q_dt, kinds[v])
                sq = rank_pct(s_)
                afi = interp_on_quantiles_1d(sq[0], q_dt, af[0, 0], method, extrap)
                scen_block[v] = apply_correction(s_[0], afi.astype(dt), kinds[v])
            npdft_block = npdft_adjust(_standardize(sim[:, i, gw]), af_q[ib, i].astype(np.float64), rots, quantiles,
                                       method, extrap)
            for v in range(V):
                scen[v, i, g] = reordering_1d(scen_block[v], npdft_block[v])[keep]
    return scen


def npdf_transform(ref, hist, sim, rots, quantiles, *, group, time, sim_time, window=1, interp="nearest",
                   extrap="constant"):
    """``npdf_transform`` (_adjustment.py:977-1057) with base = QuantileDeltaMapping: ref, hist [V, N, T], sim
    [V, N, Ts] -> (scenh, scens).  ``x @ R`` of xarray is a dot product over the variable dimension -- restated with
    numpy's einsum (xarray is absent from this image, its dot calls einsum too; parity unpinned beyond that)."""
    V, N, T = ref.shape
    dt = ref.dtype
    q = np.asarray(quantiles, dt)
    gidx, G, _ = group_index(time, group)
    hist, sim = hist.copy(), sim.copy()
    for R in rots:
        refp, histp, simp = (np.einsum("xnt,xy->ynt", a, R).astype(dt) for a in (ref, hist, sim))
        sh, ss = np.empty_like(histp), np.empty_like(simp)
        for v in range(V):
            af, _ = eqm_train(refp[v], histp[v], gidx, G, window, q, "+")
            sh[v], _ = qdm_adjust(histp[v], af, q, group=group, time=time, window=window, interp=interp,
                                  extrapolation=extrap, kind="+")
            ss[v], _ = qdm_adjust(simp[v], af, q, group=group, time=sim_time, window=window, interp=interp,
                                  extrapolation=extrap, kind="+")
        hist = np.einsum("ynt,xy->xnt", sh, R).astype(dt)
        sim = np.einsum("ynt,xy->xnt", ss, R).astype(dt)
    return hist, sim


# ----------------------------------------------------------------------------------------------
# vecquantiles (numba flavour) and map_cdf  (nbutils.py:151-195; utils.py:35-84)
# ----------------------------------------------------------------------------------------------

def numba_nanquantile(row, q):
    """numba's np.nanquantile as compiled into ``_vecquantiles`` (numba/np/arraymath.py
    ``_collect_percentiles_inner``, numba 0.65): float64, rank = 1 + (n-1)*((q*100)/100)."""
    a = np.sort(np.asarray(row, np.float64)[~np.isnan(row)])
    n = a.size
    if n == 0:
        return np.nan
    pct = np.float64(q) * 100.0
    if n == 1:
        return a[0]
    if pct == 100:
        return a[-1]
    if pct == 0:
        return a[0]
    rank = 1 + (n - 1) * (pct / 100.0)
    f = np.floor(rank)
    m = rank - f
    k = int(f - 1)
    return a[k] * (1 - m) + a[k + 1] * m


def vecquantiles_numba(arr, rnk):
    """nbutils.py:151-161 with numba's nanquantile; output in the data dtype."""
    out = np.full(arr.shape[0], np.nan, arr.dtype)
    for i in range(arr.shape[0]):
        if not np.isnan(rnk[i]):
            out[i] = numba_nanquantile(arr[i], rnk[i])
    return out


def ecdf_1d(x, value):
    """utils.py:35-37."""
    sx = np.r_[-np.inf, np.sort(x, axis=None)]
    return np.searchsorted(sx, value, side="right") / np.sum(~np.isnan(sx))


def map_cdf_1d(x, y, y_value):
    """utils.py:40-44 (numpy's own nanquantile)."""
    q = ecdf_1d(y, np.atleast_1d(y_value))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return np.nanquantile(x, q=q)


# ----------------------------------------------------------------------------------------------
# frequency adaptation  (_processing.py:20-142; _adjustment.py:32-45, 69-70, 639-646)
# ----------------------------------------------------------------------------------------------

def ecdf(x, value):
    """utils.ecdf (utils.py:87-106) along the last axis: (x <= value).sum / notnull.sum (float64)."""
    with np.errstate(all="ignore"):
        return (x <= value).sum(axis=-1) / (~np.isnan(x)).sum(axis=-1)


def rank_pct_tiebreak(seg, rng):
    """utils.rank(pct=True, use_random_tiebreak=True) (utils.py:618-634) along the last axis."""
    r = nanrankdata(seg)
    r = r + rng.uniform(0.1, 0.25, size=seg.shape)
    r = nanrankdata(r)
    cnt = (~np.isnan(seg)).sum(axis=-1, keepdims=True)
    with np.errstate(all="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = r / cnt
        mn = np.nanmin(r, axis=-1, keepdims=True)
        mx = np.nanmax(r, axis=-1, keepdims=True)
        return mx * (r - mn) / (mx - mn)


def adapt_freq_segment(sim, thresh, rng, ref=None, P0_ref=None, P0_hist=None, pth=None):
    """``_adapt_freq.func`` on one group (_processing.py:75-131): sim [N, S] (and ref [N, S_ref], or the stored
    triplet) -> sim_ad, pth, dP0, P0_ref, P0_hist.  The random parts use ``rng`` (the reference uses numpy's
    global RNG)."""
    P0_sim = ecdf(sim, thresh)
    P0_hist = P0_sim if P0_hist is None else P0_hist
    P0_ref = ecdf(ref, thresh) if P0_ref is None else P0_ref
    with np.errstate(all="ignore"):
        dP0 = np.where(P0_hist == 0, np.nan, (P0_hist - P0_ref) / P0_hist)
    if ((dP0 <= 0) | np.isnan(dP0)).all():
        return sim.copy(), np.nan * dP0, dP0, P0_ref, P0_hist
    if pth is None:
        pth = vecquantiles_numba(ref.astype(np.float64), P0_hist).astype(ref.dtype)
        pth = np.where(dP0 > 0, pth, np.nan).astype(ref.dtype)
    rnk = rank_pct_tiebreak(sim, rng)
    no_adapt = ((dP0 <= 0) | np.isnan(dP0))[:, None]
    with np.errstate(all="ignore"):
        keep = (rnk < ((P0_ref / P0_hist) * P0_sim)[:, None]) | (rnk > P0_sim[:, None]) | np.isnan(sim)
        fill = (pth[:, None] - thresh) * rng.random(sim.shape).astype(sim.dtype) + thresh
    sim_ad = np.where(no_adapt, sim, np.where(keep, sim, fill)).astype(sim.dtype)
    return sim_ad, pth, dP0, P0_ref, P0_hist


def eqm_train_adapt_freq(ref, hist, gidx, n_groups, window, q, kind, thresh, rng):
    """eqm_train with ``adapt_freq_thresh`` (_adjustment.py:259-286 -> _preprocess_dataset:69-70): returns
    af, hist_q [N,G,nq] and P0_ref, P0_hist, pth [N,G]."""
    dt = ref.dtype
    N = ref.shape[0]
    q = np.asarray(q, dt)
    af = np.full((N, n_groups, q.size), np.nan, dt); hq = af.copy()
    P0r = np.full((N, n_groups), np.nan); P0h = P0r.copy(); pth = np.full((N, n_groups), np.nan, dt)
    for g in range(n_groups):
        if not np.any(gidx == g):
            continue
        rseg = group_segment(ref, gidx, g, window)
        hseg = group_segment(hist, gidx, g, window)
        h_ad, pth[:, g], _, P0r[:, g], P0h[:, g] = adapt_freq_segment(hseg, thresh, rng, ref=rseg)
        ref_q = nan_quantile(rseg, q)
        hist_q = nan_quantile(h_ad, q)
        af[:, g] = get_correction(hist_q, ref_q, kind)
        hq[:, g] = hist_q
    return af, hq, P0r, P0h, pth


def dqm_train_adapt_freq(ref, hist, gidx, n_groups, window, q, kind, thresh, rng):
    """dqm_train with ``adapt_freq_thresh`` (_adjustment.py:155-190): hist is frequency-adapted by
    ``_preprocess_dataset`` first, then both series are normalised by their means.  Returns af, hist_q [N,G,nq],
    scaling [N,G] and P0_ref, P0_hist, pth [N,G]."""
    dt = ref.dtype
    N = ref.shape[0]
    q = np.asarray(q, dt)
    af = np.full((N, n_groups, q.size), np.nan, dt); hq = af.copy()
    sc = np.full((N, n_groups), np.nan, dt)
    P0r = np.full((N, n_groups), np.nan); P0h = P0r.copy(); pth = np.full((N, n_groups), np.nan, dt)
    for g in range(n_groups):
        if not np.any(gidx == g):
            continue
        rseg = group_segment(ref, gidx, g, window)
        hseg = group_segment(hist, gidx, g, window)
        h_ad, pth[:, g], _, P0r[:, g], P0h[:, g] = adapt_freq_segment(hseg, thresh, rng, ref=rseg)
        h_ad = h_ad.astype(dt)
        mu_ref = _nanmean_rows(rseg)
        mu_hist = _nanmean_rows(h_ad)
        refn = apply_correction(rseg, invert(mu_ref, kind)[:, None].astype(dt), kind)
        histn = apply_correction(h_ad, invert(mu_hist, kind)[:, None].astype(dt), kind)
        ref_q = nan_quantile(refn.astype(dt), q)
        hist_q = nan_quantile(histn.astype(dt), q)
        af[:, g] = get_correction(hist_q, ref_q, kind)
        hq[:, g] = hist_q
        sc[:, g] = get_correction(mu_hist, mu_ref, kind)
    return af, hq, sc, P0r, P0h, pth


def escore(tgt, sim):
    """``_escore`` (nbutils.py:347-372): tgt [K, N], sim [K, M] (float64)."""
    sim = sim[:, ~np.isnan(sim).any(axis=0)]
    tgt = tgt[:, ~np.isnan(tgt).any(axis=0)]
    n1, n2 = sim.shape[1], tgt.shape[1]
    if 0 in (n1, n2):
        return np.nan
    dist = lambda A, B: np.sqrt(((A[:, :, None] - B[:, None, :]) ** 2).sum(axis=0))  # noqa: E731
    sXY = dist(tgt, sim).mean()
    sXX = dist(tgt, tgt).sum() / tgt.shape[1] ** 2
    sYY = dist(sim, sim).sum() / sim.shape[1] ** 2
    w = n1 * n2 / (n1 + n2)
    return w * (sXY + sXY - sXX - sYY) / 2
