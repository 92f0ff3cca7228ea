"""GPU parity: the CUDA path (through the C ABI / host mirror) against the CPU oracle on the same
seeded inputs.  Bit-exact for quantiles, ranks, nearest lookups; <= 1e-6 (f32) / 1e-12 (f64) relative
for the rest (BASELINE.json north_star)."""
import numpy as np
import pytest

import qm_oracle as o
import synth
from conftest import bits_equal

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _xs():
    import xsdba_b200 as xs
    return xs


def _time(calendar, n_years, start=1981):
    xs = _xs()
    return xs.TimeAxis.daily(start, n_years, calendar), o.daily_time_axis(start, n_years, calendar)


def _np(t):
    return t.detach().cpu().numpy()


CASES = [
    # group, window, calendar, years, nq, kind, var, dtype
    ("time", 1, "noleap", 3, 50, "+", "tas", np.float32),
    ("time.month", 1, "noleap", 6, 50, "+", "tas", np.float32),
    ("time.month", 1, "standard", 4, 20, "*", "pr", np.float32),
    ("time.month", 1, "noleap", 4, 50, "+", "tas", np.float64),
    ("time.dayofyear", 31, "noleap", 4, 100, "*", "pr", np.float32),
    ("time.dayofyear", 5, "standard", 5, 30, "+", "tas", np.float32),
    ("time.dayofyear", 31, "360_day", 3, 50, "+", "tas", np.float64),
    ("time.season", 1, "noleap", 3, 40, "+", "tas", np.float32),
]


def _make(case, n_pts=37, seed=0):
    group, window, cal, years, nq, kind, var, dt = case
    rng = np.random.default_rng(seed)
    tx, to = _time(cal, years)
    gen = getattr(synth, var)
    ref, hist, sim = (gen(rng, to, n_pts, w, dt) for w in ("ref", "hist", "sim"))
    ref[:, 3] = np.nan            # an all-NaN point (mask must be preserved bit-exactly)
    hist[:, 3] = np.nan
    hist[5:, 4] = np.nan          # a point with only 5 valid hist samples
    return tx, to, ref, hist, sim


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}-w{c[1]}-{c[2]}-{c[5]}-{np.dtype(c[7]).name}")
def test_eqm_train_matches_oracle(case):
    xs = _xs()
    group, window, cal, years, nq, kind, var, dt = case
    tx, to, ref, hist, sim = _make(case)
    q = o.equally_spaced_nodes(nq).astype(dt)
    gidx, G, _ = o.group_index(to, group)
    af_o, hq_o = o.eqm_train(ref.T.copy(), hist.T.copy(), gidx, G, window, q, kind)
    ds = xs.eqm_train(xs.Dataset({"ref": ref, "hist": hist}, time=tx), group=xs.Grouper(group, window), kind=kind,
                      quantiles=q)
    af, hq = _np(ds.af), _np(ds.hist_q)
    assert af.shape == (ref.shape[1], G, nq)
    if dt == np.float32:
        assert bits_equal(hq, hq_o)
        assert bits_equal(af, af_o)
    else:
        np.testing.assert_allclose(hq, hq_o, rtol=1e-12, atol=0, equal_nan=True)
        np.testing.assert_allclose(af, af_o, rtol=1e-9, atol=1e-12 * np.nanmax(np.abs(hq_o)), equal_nan=True)
    # point-major input gives the same tables
    ds2 = xs.eqm_train(xs.Dataset({"ref": ref.T.copy(), "hist": hist.T.copy()}, time=tx, time_axis=-1),
                       group=xs.Grouper(group, window), kind=kind, quantiles=q)
    assert bits_equal(_np(ds2.af), af) and bits_equal(_np(ds2.hist_q), hq)


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}-w{c[1]}-{c[2]}-{c[5]}-{np.dtype(c[7]).name}")
@pytest.mark.parametrize("extrap", ["constant", "nan"])
def test_qm_adjust_nearest_matches_oracle(case, extrap):
    xs = _xs()
    group, window, cal, years, nq, kind, var, dt = case
    tx, to, ref, hist, sim = _make(case, n_pts=13)
    q = o.equally_spaced_nodes(nq).astype(dt)
    gidx, G, _ = o.group_index(to, group)
    af_o, hq_o = o.eqm_train(ref.T.copy(), hist.T.copy(), gidx, G, window, q, kind)
    scen_o = o.qm_adjust(sim.T.copy(), af_o, hq_o, group=group, time=to, interp="nearest", extrapolation=extrap,
                         kind=kind)
    out = xs.qm_adjust(xs.Dataset({"sim": sim, "af": af_o, "hist_q": hq_o}, time=tx), group=xs.Grouper(group, window),
                       interp="nearest", extrapolation=extrap, kind=kind)
    scen = _np(out.scen).T
    if group == "time":
        assert bits_equal(scen, scen_o)  # interp1d nearest is an exact, tie-defined table pick
        return
    # 2-D nearest: bit-exact wherever the nearest node is unique; where two nodes are exactly
    # equidistant SciPy's cKDTree picks by traversal order ("parity unpinned on ties"):
    # there the CUDA value must be one of the tied candidates.
    lo, hi = o.qm_adjust_factor_bounds(sim.T.copy(), af_o, hq_o, group=group, time=to, extrapolation=extrap)
    simT = sim.T
    s_lo = o.apply_correction(simT, lo.astype(dt), kind).astype(dt)
    s_hi = o.apply_correction(simT, hi.astype(dt), kind).astype(dt)
    unique = (s_lo == s_hi) | (np.isnan(s_lo) & np.isnan(s_hi))
    assert unique.mean() > 0.9  # (the 5-valid-sample point has many duplicated nodes = ties)
    assert bits_equal(np.where(unique, scen, 0), np.where(unique, scen_o, 0))
    tie = ~unique
    mn, mx = np.minimum(s_lo, s_hi)[tie], np.maximum(s_lo, s_hi)[tie]  # >2 nodes may tie: any of them is legal
    assert ((scen[tie] >= mn) & (scen[tie] <= mx)).all()
    assert ((scen_o[tie] >= mn) & (scen_o[tie] <= mx)).all()


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("extrap", ["constant", "nan"])
def test_qm_adjust_time_linear_matches_oracle(dt, extrap):
    xs = _xs()
    case = ("time", 1, "noleap", 3, 50, "+", "tas", dt)
    tx, to, ref, hist, sim = _make(case, n_pts=11)
    q = o.equally_spaced_nodes(50).astype(dt)
    gidx, G, _ = o.group_index(to, "time")
    af_o, hq_o = o.eqm_train(ref.T.copy(), hist.T.copy(), gidx, G, 1, q, "+")
    scen_o = o.qm_adjust(sim.T.copy(), af_o, hq_o, group="time", time=to, interp="linear", extrapolation=extrap, kind="+")
    out = xs.qm_adjust(xs.Dataset({"sim": sim, "af": af_o, "hist_q": hq_o}, time=tx), group="time", interp="linear",
                       extrapolation=extrap, kind="+")
    np.testing.assert_allclose(_np(out.scen).T, scen_o, rtol=1e-6 if dt == np.float32 else 1e-12, atol=0, equal_nan=True)


QDM_CASES = [
    ("time", 1, "noleap", 3, 50, "*", "pr", np.float32, False),
    ("time.month", 1, "noleap", 4, 50, "*", "pr", np.float32, False),
    ("time.dayofyear", 31, "noleap", 4, 100, "*", "pr", np.float32, False),
    ("time.dayofyear", 31, "noleap", 3, 100, "*", "pr", np.float32, True),
    ("time.dayofyear", 7, "standard", 5, 20, "+", "tas", np.float64, True),
    ("time.month", 1, "noleap", 3, 30, "+", "tas", np.float64, False),
]


@pytest.mark.parametrize("case", QDM_CASES, ids=lambda c: f"{c[0]}-w{c[1]}-{c[2]}-{c[5]}-{np.dtype(c[7]).name}-rw{int(c[8])}")
def test_qdm_adjust_matches_oracle(case):
    xs = _xs()
    group, window, cal, years, nq, kind, var, dt, rank_window = case
    tx, to, ref, hist, sim = _make(case[:8], n_pts=9)
    sim[:, 2] = np.round(sim[:, 2])  # ties -> average ranks
    q = o.equally_spaced_nodes(nq).astype(dt)
    gidx, G, _ = o.group_index(to, group)
    af_o, hq_o = o.eqm_train(ref.T.copy(), hist.T.copy(), gidx, G, window, q, kind)
    scen_o, simq_o = o.qdm_adjust(sim.T.copy(), af_o, q, group=group, time=to, window=window, interp="nearest",
                                  extrapolation="constant", kind=kind, rank_window=rank_window)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", DeprecationWarning)
        out = xs.qdm_adjust(xs.Dataset({"sim": sim, "af": af_o, "quantiles": q}, time=tx),
                            group=xs.Grouper(group, window), interp="nearest", extrapolation="constant", kind=kind,
                            rank_window=rank_window)
    assert bits_equal(_np(out.sim_q).T, simq_o)      # ranks are exact rationals in float64
    scen = _np(out.scen).T
    if group == "time":
        assert bits_equal(scen, scen_o)
        return
    # exact ties between two quantile nodes are resolved by SciPy's KD-tree traversal order (unpinned)
    lo, hi = o.qm_adjust_factor_bounds(simq_o, af_o, q, group=group, time=to, extrapolation="constant")
    s_lo = o.apply_correction(sim.T, lo.astype(dt), kind).astype(dt)
    s_hi = o.apply_correction(sim.T, hi.astype(dt), kind).astype(dt)
    unique = (s_lo == s_hi) | (np.isnan(s_lo) & np.isnan(s_hi))
    assert unique.mean() > 0.95
    assert bits_equal(np.where(unique, scen, 0), np.where(unique, scen_o, 0))
    tie = ~unique
    mn, mx = np.minimum(s_lo, s_hi)[tie], np.maximum(s_lo, s_hi)[tie]
    assert ((scen[tie] >= mn) & (scen[tie] <= mx)).all()


def test_group_quantile_reference_golden(golden):
    """The CUDA quantiles against outputs of the reference's own numba kernel (tests/golden)."""
    xs = _xs()
    for tag, dt in (("f32", np.float32), ("f64", np.float64)):
        for nq in (50, 100):
            a, q, ref = (golden[f"quant_{tag}_{nq}_{k}"] for k in ("in", "q", "out"))
            t = xs.TimeAxis.daily(2000, 3, "noleap")[: a.shape[1]]
            got = _np(xs.group_quantile(a, time=t, group="time", quantiles=q, time_axis=-1))[:, 0, :]
            assert bits_equal(got, ref), (tag, nq)
    a, q, ref = golden["quant_long_in"], golden["quant_long_q"], golden["quant_long_out"]
    t = xs.TimeAxis.daily(1981, 30, "noleap")
    assert bits_equal(_np(xs.group_quantile(a, time=t, group="time", quantiles=q, time_axis=-1))[:, 0, :], ref)


@pytest.mark.parametrize("group,window,nq,kind,var", [("time.month", 1, 50, "+", "tas"), ("time.dayofyear", 31, 100, "*", "pr")])
def test_train_full_size_segments(group, window, nq, kind, var):
    """30-year daily series: 840-930-slot segments, the shape the register-blocked sorter is built for."""
    xs = _xs()
    case = (group, window, "noleap", 30, nq, kind, var, np.float32)
    tx, to, ref, hist, sim = _make(case, n_pts=33, seed=3)
    q = o.equally_spaced_nodes(nq).astype(np.float32)
    gidx, G, _ = o.group_index(to, group)
    sel = np.arange(G) if G <= 12 else np.array([0, 1, 14, 15, 16, 100, 200, 349, 350, 363, 364])
    ds = xs.eqm_train(xs.Dataset({"ref": ref, "hist": hist}, time=tx), group=xs.Grouper(group, window), kind=kind,
                      quantiles=q)
    af, hq = _np(ds.af), _np(ds.hist_q)
    for g in sel:  # the oracle's window gather is slow: check a subset of the day-of-year groups
        ref_q = o.nan_quantile(o.group_segment(ref.T.copy(), gidx, g, window), q)
        hist_q = o.nan_quantile(o.group_segment(hist.T.copy(), gidx, g, window), q)
        assert bits_equal(hq[:, g], hist_q), g
        assert bits_equal(af[:, g], o.get_correction(hist_q, ref_q, kind).astype(np.float32)), g


def test_dqm_train_matches_oracle():
    xs = _xs()
    for case in [("time.month", 1, "noleap", 5, 50, "+", "tas", np.float32), ("time.dayofyear", 31, "noleap", 4, 50, "*", "pr", np.float32),
                 ("time", 1, "noleap", 3, 20, "*", "pr", np.float64)]:
        group, window, cal, years, nq, kind, var, dt = case
        tx, to, ref, hist, sim = _make(case, n_pts=35)
        q = o.equally_spaced_nodes(nq).astype(dt)
        gidx, G, _ = o.group_index(to, group)
        af_o, hq_o, sc_o = o.dqm_train(ref.T.copy(), hist.T.copy(), gidx, G, window, q, kind)
        ds = xs.dqm_train(xs.Dataset({"ref": ref, "hist": hist}, time=tx), group=xs.Grouper(group, window), kind=kind,
                          quantiles=q)
        # group means are accumulated in a different order than the oracle's: tolerance, not bits
        rtol = 2e-6 if dt == np.float32 else 1e-12
        scale = np.nanmax(np.abs(hq_o[np.isfinite(hq_o)]))
        np.testing.assert_allclose(_np(ds.scaling), sc_o, rtol=rtol, atol=rtol * scale, equal_nan=True)
        np.testing.assert_allclose(_np(ds.hist_q), hq_o, rtol=rtol, atol=rtol * scale, equal_nan=True)
        fin = np.isfinite(af_o)
        assert (np.isfinite(_np(ds.af)) == fin).all()
        np.testing.assert_allclose(_np(ds.af)[fin], af_o[fin], rtol=50 * rtol, atol=50 * rtol * scale)


DQM_CASES = [
    ("time", 1, "noleap", 4, 30, "+", "tas", np.float32, 1),
    ("time.month", 1, "noleap", 5, 50, "+", "tas", np.float32, 1),
    ("time.dayofyear", 31, "noleap", 4, 20, "*", "pr", np.float32, 1),
    ("time.dayofyear", 15, "standard", 4, 20, "+", "tas", np.float64, 2),
    ("time", 1, "noleap", 3, 20, "*", "pr", np.float64, 0),
]


@pytest.mark.parametrize("case", DQM_CASES, ids=lambda c: f"{c[0]}-w{c[1]}-{c[2]}-{c[5]}-{np.dtype(c[7]).name}-deg{c[8]}")
def test_dqm_adjust_matches_oracle(case):
    """dqm_adjust with PolyDetrend: trend within 1e-9, scen within 1e-6 (f32) / 1e-9 (f64) relative; samples
    whose nearest node changes under that tolerance (ties) are excluded like in the EQM test."""
    xs = _xs()
    group, window, cal, years, nq, kind, var, dt, degree = case
    tx, to, ref, hist, sim = _make(case[:8], n_pts=9)
    q = o.equally_spaced_nodes(nq).astype(dt)
    gidx, G, _ = o.group_index(to, group)
    af_o, hq_o, sc_o = o.dqm_train(ref.T.copy(), hist.T.copy(), gidx, G, window, q, kind)
    scen_o, trend_o = o.dqm_adjust(sim.T.copy(), af_o, hq_o, sc_o, group=group, window=window, time=to,
                                   interp="nearest", extrapolation="constant", kind=kind, detrend=degree)
    out = xs.dqm_adjust(xs.Dataset({"sim": sim, "af": af_o, "hist_q": hq_o, "scaling": sc_o}, time=tx),
                        group=xs.Grouper(group, window), interp="nearest", extrapolation="constant", kind=kind,
                        detrend=degree)
    trend = _np(out.trend).T
    scen = _np(out.scen).T
    assert scen.dtype == dt
    np.testing.assert_allclose(trend, trend_o, rtol=1e-9, atol=1e-9, equal_nan=True)
    rtol = 2e-6 if dt == np.float32 else 1e-9
    close = np.isclose(scen, scen_o, rtol=rtol, atol=0, equal_nan=True)
    # a detrended value within ~1e-9 (relative) of a mid-point between two nodes may pick the other node
    assert close.mean() > 0.999, close.mean()
    assert (np.isnan(scen) == np.isnan(scen_o)).all()


@pytest.mark.parametrize("d,f", [(0, 0.2), (1, 0.3)])
def test_loess_trend_reference_golden(golden, d, f):
    """CUDA LOESS against the reference's own numba _loess_nb on the golden series (equal spacing, NaNs)."""
    xs = _xs()
    x, y = golden["loess_x"], golden["loess_y"]
    n = x.size
    t = xs.TimeAxis.daily(2001, 1, "noleap")[:n]
    if d == 0:  # golden case 0 is exactly (d=0, f=0.2, niter=1, equal spacing)
        assert tuple(golden["loess_case0_params"][:3]) == (0, 0.2, 1)
        want = golden["loess_case0_out"]
    else:
        want = o.loess_nb(x, y, f=f, niter=1, d=1, dx=float(x[1] - x[0]))
    series = np.stack([y, y[::-1].copy(), np.where(np.arange(n) < 30, np.nan, y)], axis=1)  # (time, 3 points)
    got = _np(xs.loess_trend(series, time=t, f=f, niter=1, d=d))
    np.testing.assert_allclose(got[:, 0], want, rtol=1e-11, atol=1e-12, equal_nan=True)
    for j in (1, 2):
        wj = o.loess_nb(x, series[:, j], f=f, niter=1, d=d, dx=float(x[1] - x[0]))
        np.testing.assert_allclose(got[:, j], wj, rtol=1e-11, atol=1e-12, equal_nan=True)


def test_loess_robust_iterations_reference_golden(golden):
    """niter > 1 (loess.py:166-176): golden case 1 of the reference's numba _loess_nb is (d=1, f=0.5, niter=2, equal
    spacing); d=0 / niter=3 against the float64 restatement."""
    xs = _xs()
    x, y = golden["loess_x"], golden["loess_y"]
    n = x.size
    t = xs.TimeAxis.daily(2001, 1, "noleap")[:n]
    assert tuple(golden["loess_case1_params"][:3]) == (1, 0.5, 2)
    series = np.stack([y, y[::-1].copy(), np.where(np.arange(n) % 17 == 3, np.nan, y)], axis=1)
    got = _np(xs.loess_trend(series, time=t, f=0.5, niter=2, d=1))
    np.testing.assert_allclose(got[:, 0], golden["loess_case1_out"], rtol=1e-9, atol=1e-10, equal_nan=True)
    dx = float(x[1] - x[0])
    for d, f, niter in ((0, 0.2, 3), (1, 0.3, 2)):
        got = _np(xs.loess_trend(series, time=t, f=f, niter=niter, d=d))
        for j in range(3):
            want = o.loess_nb(x, series[:, j], f=f, niter=niter, d=d, dx=dx)
            np.testing.assert_allclose(got[:, j], want, rtol=1e-9, atol=1e-10, equal_nan=True)


def test_loess_complete_series_shared_tables():
    """Complete (NaN-free) series take the shared-table paths of K6 (broadcast interior weights for a warp of 32
    complete points, L2-resident edge-weight table); one column with a gap keeps its warp on the per-point path.
    Both must reproduce the float64 restatement of _loess_nb to 1e-11."""
    xs = _xs()
    rng = np.random.default_rng(8)
    t = xs.TimeAxis.daily(2001, 3, "noleap")
    n = len(t)
    x = np.arange(n, dtype=np.float64)
    y = (280 + 5 * np.sin(2 * np.pi * x / 365)[:, None] + rng.standard_normal((n, 70))).astype(np.float32)
    y[100:130, 65] = np.nan                      # points 64..69 share a warp with an incomplete column
    got = _np(xs.loess_trend(y, time=t, f=0.2, niter=1, d=0))
    for j in (0, 17, 31, 32, 63, 64, 65, 69):
        want = o.loess_nb(x, y[:, j].astype(np.float64), f=0.2, niter=1, d=0, dx=1.0)
        np.testing.assert_allclose(got[:, j], want, rtol=1e-11, atol=1e-12, equal_nan=True)


def test_dqm_adjust_loess_matches_oracle():
    xs = _xs()
    case = ("time.month", 1, "noleap", 3, 20, "+", "tas", np.float32)
    group, window, cal, years, nq, kind, var, dt = case
    tx, to, ref, hist, sim = _make(case, n_pts=5)
    q = o.equally_spaced_nodes(nq).astype(dt)
    gidx, G, _ = o.group_index(to, group)
    af_o, hq_o, sc_o = o.dqm_train(ref.T.copy(), hist.T.copy(), gidx, G, window, q, kind)
    scen_o, trend_o = o.dqm_adjust(sim.T.copy(), af_o, hq_o, sc_o, group=group, window=window, time=to, interp="nearest",
                                   extrapolation="constant", kind=kind, loess=dict(f=0.2, niter=1, d=0))
    out = xs.dqm_adjust(xs.Dataset({"sim": sim, "af": af_o, "hist_q": hq_o, "scaling": sc_o}, time=tx), group=group,
                        interp="nearest", extrapolation="constant", kind=kind,
                        detrend=xs.LoessDetrend(group="time", kind=kind, f=0.2, niter=1, d=0))
    np.testing.assert_allclose(_np(out.trend).T, trend_o, rtol=1e-10, atol=1e-10, equal_nan=True)
    close = np.isclose(_np(out.scen).T, scen_o, rtol=2e-6, atol=0, equal_nan=True)
    assert close.mean() > 0.999


def test_jitter_is_distributionally_the_reference(golden):
    """processing.jitter (processing.py:180-257): the reference draws from numpy's global RNG, so parity is
    distributional: untouched values bit-identical, replaced values uniform in the reference's
    interval, NaNs kept."""
    from scipy import stats
    xs = _xs()
    rng = np.random.default_rng(5)
    x = synth.pr(rng, o.daily_time_axis(1981, 6, "noleap"), 40, "hist", jitter=False)
    out = _np(xs.jitter_under_thresh(x, "0.01 mm/d", seed=3))
    dry = (x < 0.01)
    assert bits_equal(np.where(dry, 0, out), np.where(dry, 0, x))          # wet values and NaNs untouched
    assert (np.isnan(out) == np.isnan(x)).all()
    v = out[dry]
    assert v.min() > 0 and v.max() <= np.float32(0.01)
    assert stats.kstest(v.astype(np.float64), stats.uniform(0, 0.01).cdf).pvalue > 1e-3
    out2 = _np(xs.jitter_under_thresh(x, 0.01, seed=3))
    assert bits_equal(out, out2)                                            # reproducible for a given seed
    hi = _np(xs.jitter_over_thresh(x, 30.0, 35.0, seed=1))
    big = x >= 30
    assert big.sum() > 10 and (hi[big] >= 30).all() and (hi[big] < 35).all() and bits_equal(np.where(big, 0, hi), np.where(big, 0, x))


def test_train_with_jitter_matches_oracle_statistically():
    """eqm_train(jitter_under_thresh_value=...) jitters hist inside each group after the window gather; like the
    reference's own test (tests/test_adjustment.py:1097-1103) the comparison is to 2 decimals on the quantiles."""
    xs = _xs()
    rng = np.random.default_rng(11)
    to = o.daily_time_axis(1981, 10, "noleap"); tx = xs.TimeAxis.daily(1981, 10, "noleap")
    ref, hist = (synth.pr(rng, to, 6, w, jitter=False, nan_frac=0) for w in ("ref", "hist"))
    q = o.equally_spaced_nodes(20).astype(np.float32)
    ds = xs.eqm_train(xs.Dataset({"ref": ref, "hist": hist}, time=tx), group=xs.Grouper("time.month"), kind="*",
                      quantiles=q, jitter_under_thresh_value="0.01 mm/d")
    hq = _np(ds.hist_q)
    hist_j = np.where(hist < 0.01, rng.uniform(1e-45, 0.01, hist.shape), hist).astype(np.float32)
    gidx, G, _ = o.group_index(to, "time.month")
    _, hq_o = o.eqm_train(ref.T.copy(), hist_j.T.copy(), gidx, G, 1, q, "*")
    np.testing.assert_allclose(hq, hq_o, atol=2e-3, rtol=0.02)   # nodes inside the jittered range are random
    wet = hq_o > 0.02
    np.testing.assert_array_equal(hq[wet], hq_o[wet])            # above the threshold nothing changes


def _mbcn_inputs(years, N, seed=0):
    rng = np.random.default_rng(seed)
    to = o.daily_time_axis(1981, years, "noleap")
    mk = lambda w: np.stack([synth.tas(rng, to, N, w, nan_frac=0), synth.pr(rng, to, N, w, nan_frac=0),
                             synth.tas(rng, to, N, w, nan_frac=0) + 2]).astype(np.float32)  # (V, T, N)
    return to, mk("ref"), mk("hist"), mk("sim")


def test_npdft_reference_golden(golden):
    """The CUDA N-pdf training against the reference's own _npdft_train output (tests/golden)."""
    xs = _xs()
    ref, hist, rots, q = golden["npdft_ref"], golden["npdft_hist"], golden["npdft_rots"], golden["npdft_q"]
    t = xs.TimeAxis.daily(2001, 2, "noleap")[: ref.shape[1]]
    af_q = _np(xs.mbcn_train(ref[:, :, None], hist[:, :, None], time=t, rot_matrices=rots, quantiles=q, group="time"))
    # every step reproduces the reference's arithmetic (numpy's pairwise nanmean / nanstd, the fused chain of
    # OpenBLAS's sgemm for `rot @ x`, numba's quantiles, float64 rank lookups): bit-exact, not merely close
    assert bits_equal(af_q[0, 0].astype(np.float64), np.asarray(golden["npdft_af_q"], np.float64))


@pytest.mark.parametrize("group,window,years,n_iter", [("time", 1, 2, 4), ("time.dayofyear", 5, 2, 2)])
def test_mbcn_matches_oracle(group, window, years, n_iter):
    xs = _xs()
    N = 3
    to, ref, hist, sim = _mbcn_inputs(years, N)
    tx = xs.TimeAxis.daily(1981, years, "noleap")
    rots = o.rand_rot_matrices(3, n_iter, 7)
    q = o.equally_spaced_nodes(10)
    kinds = ["+", "*", "+"]
    blocks = o.mbcn_blocks(to, group, window)
    tr = lambda a: np.ascontiguousarray(a.transpose(0, 2, 1))           # oracle layout (V, N, T)
    afq_o = o.mbcn_train(tr(ref), tr(hist), rots, q, blocks)
    scen_o = o.mbcn_adjust(tr(ref), tr(hist), tr(sim), afq_o, rots, q, blocks, kinds)
    obj = xs.MBCn.train(ref, hist, time=tx, base_kws={"nquantiles": q, "group": xs.Grouper(group, window)}, n_iter=n_iter,
                        rot_matrices=rots)
    afq = _np(obj.ds["af_q"])                                           # (blocks, N, n_iter, V, nq)
    assert afq.shape == afq_o.shape
    # the N-pdf iteration is chaotic (a one-ulp difference moves af_q by 1e-3 within 15 iterations), so "within
    # 1e-6" can only be met by reproducing every rounding of the reference: the comparison is bit for bit
    assert bits_equal(afq, afq_o)
    scen = tr(_np(obj.adjust(sim, ref, hist, time=tx, kinds=kinds)))
    assert bits_equal(scen, scen_o)


def test_vecquantiles_reference_golden(golden):
    """CUDA vecquantiles against the reference's numba _vecquantiles outputs (tests/golden): bit-exact."""
    xs = _xs()
    for tag in ("f32", "f64"):
        a, rk, want = golden[f"vecq_{tag}_in"], golden[f"vecq_{tag}_rnk"], golden[f"vecq_{tag}_out"]
        t = xs.TimeAxis.daily(2001, 1, "noleap")[: a.shape[1]]
        got = _np(xs.vecquantiles(a, rk[:, None], time=t, group="time", time_axis=-1))[:, 0]
        assert bits_equal(got, want), tag


def test_map_cdf_reference_golden(golden):
    xs = _xs()
    x, y, v, want = golden["mapcdf_x"], golden["mapcdf_y"], golden["mapcdf_v"], golden["mapcdf_out"]
    t = xs.TimeAxis.daily(2001, 2, "noleap")[: x.size]
    got = _np(xs.map_cdf(xs.Dataset({"x": x[:, None], "y": y[:, None]}, time=t), y_value=v, group="time"))
    assert got.dtype == np.float32                          # output_dtypes=[ds.x.dtype] (utils.py:83)
    assert bits_equal(got[0, 0], want.astype(np.float32))
    # grouped: every month against the oracle
    rng = np.random.default_rng(2)
    tt = xs.TimeAxis.daily(2001, 3, "noleap"); to = o.daily_time_axis(2001, 3, "noleap")
    X = rng.normal(size=(len(tt), 5)).astype(np.float32); Y = (rng.normal(size=(len(tt), 5)) + 0.5).astype(np.float32)
    got = _np(xs.map_cdf(xs.Dataset({"x": X, "y": Y}, time=tt), y_value=[0.2], group="time.month"))
    gidx, G, _ = o.group_index(to, "time.month")
    for p in range(5):
        for g in range(G):
            sel = gidx == g
            assert got[p, g, 0] == np.float32(o.map_cdf_1d(X[sel, p], Y[sel, p], 0.2)[0])


def test_qdm_full_size_segments_fit_shared_memory():
    """Regression: QDM with rank_window=True on 30-year daily data (930-slot segments) and nq=100 needs the sort
    buffer AND the staged tables in one CTA -- the launcher narrows the tile instead of failing."""
    xs = _xs()
    rng = np.random.default_rng(1)
    tx = xs.TimeAxis.daily(1981, 30, "noleap"); to = o.daily_time_axis(1981, 30, "noleap")
    sim = synth.pr(rng, to, 34, "sim")
    q = o.equally_spaced_nodes(100).astype(np.float32)
    af = rng.normal(1.0, 0.1, size=(34, 365, 100)).astype(np.float32)
    g = xs.Grouper("time.dayofyear", 31)
    out = xs.qdm_adjust(xs.Dataset({"sim": sim, "af": af, "quantiles": q}, time=tx), group=g, interp="nearest",
                        extrapolation="constant", kind="*", rank_window=True)
    gidx, G, _ = o.group_index(to, "time.dayofyear")
    simq_o = o.grouped_rank_pct(sim.T.copy()[:3], gidx, G, 31, True)
    assert bits_equal(_np(out.sim_q).T[:3], simq_o)


def _dry_inputs(years=8, N=6, seed=4):
    """hist/sim drier than ref, so that frequency adaptation is needed (dP0 > 0)."""
    rng = np.random.default_rng(seed)
    to = o.daily_time_axis(1981, years, "noleap")
    ref = synth.pr(rng, to, N, "hist", jitter=False, nan_frac=0.001)   # 60 % wet
    hist = synth.pr(rng, to, N, "ref", jitter=False, nan_frac=0.001)   # 45 % wet
    sim = synth.pr(rng, to, N, "ref", jitter=False, nan_frac=0.001)
    return to, ref, hist, sim


@pytest.mark.parametrize("group,window", [("time.month", 1), ("time.dayofyear", 15), ("time", 1)])
def test_adapt_freq_train_matches_oracle(group, window):
    """eqm_train(adapt_freq_thresh=...): P0_ref, P0_hist, pth are deterministic -> bit-exact; the adapted hist
    quantiles depend on random fills -> compared like the reference's own test (2 decimals)."""
    xs = _xs()
    to, ref, hist, sim = _dry_inputs()
    tx = xs.TimeAxis.daily(1981, 8, "noleap")
    q = o.equally_spaced_nodes(20).astype(np.float32)
    gidx, G, _ = o.group_index(to, group)
    rng = np.random.default_rng(0)
    af_o, hq_o, p0r_o, p0h_o, pth_o = o.eqm_train_adapt_freq(ref.T.copy(), hist.T.copy(), gidx, G, window, q, "*", 0.05, rng)
    ds = xs.eqm_train(xs.Dataset({"ref": ref, "hist": hist}, time=tx), group=xs.Grouper(group, window), kind="*",
                      quantiles=q, adapt_freq_thresh="0.05 mm/d")
    assert bits_equal(_np(ds.P0_ref), p0r_o) and bits_equal(_np(ds.P0_hist), p0h_o)
    assert bits_equal(_np(ds.pth), pth_o)
    assert np.isfinite(pth_o).mean() > 0.9                       # adaptation really happens in this setup
    hq = _np(ds.hist_q)
    # nodes inside (thresh, pth) are order statistics of ~20 random fills per group: statistically equal only
    assert np.isclose(hq, hq_o, rtol=0.25, atol=0.05).mean() > 0.95
    assert np.isclose(hq, hq_o, rtol=1.0, atol=0.5).all()
    big = hq_o > np.nanmax(pth_o) + 1e-3
    np.testing.assert_array_equal(hq[big], hq_o[big])            # beyond pth nothing is random


def test_dqm_train_with_adapt_freq_matches_oracle():
    """dqm_train(adapt_freq_thresh=...): frequency adaptation first, normalisation by the ADAPTED hist's mean after.
    P0 / pth are deterministic (bit-exact); scaling and the nodes depend on random fills (statistical)."""
    xs = _xs()
    to, ref, hist, sim = _dry_inputs()
    tx = xs.TimeAxis.daily(1981, 8, "noleap")
    q = o.equally_spaced_nodes(20).astype(np.float32)
    gidx, G, _ = o.group_index(to, "time.month")
    rng = np.random.default_rng(0)
    af_o, hq_o, sc_o, p0r_o, p0h_o, pth_o = o.dqm_train_adapt_freq(ref.T.copy(), hist.T.copy(), gidx, G, 1, q, "*", 0.05, rng)
    ds = xs.dqm_train(xs.Dataset({"ref": ref, "hist": hist}, time=tx), group="time.month", kind="*", quantiles=q,
                      adapt_freq_thresh="0.05 mm/d")
    assert bits_equal(_np(ds.P0_ref), p0r_o) and bits_equal(_np(ds.P0_hist), p0h_o) and bits_equal(_np(ds.pth), pth_o)
    # the fills are U(thresh, pth) with pth << mean: they move the mean (hence scaling and every normalised node) by
    # well under a percent, differently for every random stream
    np.testing.assert_allclose(_np(ds.scaling), sc_o, rtol=2e-2)
    hq = _np(ds.hist_q)
    assert np.isclose(hq, hq_o, rtol=0.25, atol=0.02).mean() > 0.95
    # without the option the result differs (the adaptation is really applied before the normalisation)
    ds0 = xs.dqm_train(xs.Dataset({"ref": ref, "hist": hist}, time=tx), group="time.month", kind="*", quantiles=q)
    assert not np.allclose(_np(ds0.scaling), _np(ds.scaling), rtol=1e-4)


def test_dqm_adjust_with_adapt_freq_and_tail_factor():
    """DetrendedQuantileMapping with adapt_freq_thresh + max_tail_factor (_adjustment.py:727-746, 776-777): the masked
    samples are the frequency-adapted sim, everything else equals the pipeline fed with the adapted sim directly."""
    xs = _xs()
    to, ref, hist, sim = _dry_inputs()
    tx = xs.TimeAxis.daily(1981, 8, "noleap")
    obj = xs.DetrendedQuantileMapping.train(ref, hist, time=tx, nquantiles=20, group="time.month", kind="*",
                                            adapt_freq_thresh="0.05 mm/d", max_tail_factor=1.5)
    scen = _np(obj.adjust(sim, time=tx, detrend=1))
    import torch as _t
    ad = _np(xs._adjustment._adapt_freq_preprocess(
        xs.Dataset(obj.ds), xs._adjustment._as_device(sim), sim.shape[1], 1, sim.shape[1], obj.group, tx, _t.float32,
        "0.05 mm/d"))
    gidx, G, _ = o.group_index(to, "time.month")
    lastq = _np(obj.ds["hist_q_raw"])[:, gidx, -1].T
    mask = ad > 1.5 * lastq
    assert mask.sum() > 0 and bits_equal(scen[mask], ad[mask])
    plain = xs.DetrendedQuantileMapping(obj.ds, obj.group, "*")          # no options: plain DQM on the adapted sim
    ref_scen = _np(plain.adjust(ad, time=tx, detrend=1))
    assert bits_equal(np.where(mask, 0, scen), np.where(mask, 0, ref_scen))


def test_adapt_freq_adjust_and_tail_factor():
    xs = _xs()
    to, ref, hist, sim = _dry_inputs()
    tx = xs.TimeAxis.daily(1981, 8, "noleap")
    obj = xs.EmpiricalQuantileMapping.train(ref, hist, time=tx, nquantiles=20, group="time.month", kind="*",
                                            adapt_freq_thresh="0.05 mm/d", max_tail_factor=1.5)
    scen = _np(obj.adjust(sim, time=tx))
    # frequency adaptation of sim with the stored factors: the dry-day frequency moves towards ref's
    gidx, G, _ = o.group_index(to, "time.month")
    ad = _np(xs._adjustment._adapt_freq_preprocess(
        xs.Dataset(obj.ds), xs._adjustment._as_device(sim), sim.shape[1], 1, sim.shape[1], obj.group, tx,
        __import__("torch").float32, "0.05 mm/d"))
    p0_before = np.nanmean(sim <= 0.05); p0_after = np.nanmean(ad <= 0.05); p0_ref = np.nanmean(ref <= 0.05)
    assert abs(p0_after - p0_ref) < 0.02 < abs(p0_before - p0_ref)
    assert bits_equal(np.where(sim > 0.05, ad, 0), np.where(sim > 0.05, sim, 0))      # wet days untouched
    assert (np.isnan(ad) == np.isnan(sim)).all()
    # expected number of replaced values per (point, group) against the oracle's rule
    rng = np.random.default_rng(1)
    p0r, p0h, pth = (_np(obj.ds[k]) for k in ("P0_ref", "P0_hist", "pth"))
    n_gpu = n_ora = 0
    for g in range(G):
        sel = gidx == g
        sa, *_ = o.adapt_freq_segment(sim[sel].T.copy(), 0.05, rng, P0_ref=p0r[:, g], P0_hist=p0h[:, g], pth=pth[:, g])
        n_ora += (sa != sim[sel].T).sum() - np.isnan(sim[sel]).sum()
        n_gpu += (ad[sel] != sim[sel]).sum() - np.isnan(sim[sel]).sum()
    assert abs(n_gpu - n_ora) < 0.03 * n_ora
    # the tied (dry) values take DISTINCT positions of their tie block (a permutation, like the reference's random
    # tie-break followed by a re-rank), so the number of replaced values of every (point, group) is deterministic:
    # the count of sorted positions r with  P0_ref / P0_hist * P0_sim <= r / (n - 1) <= P0_sim
    for g in range(G):
        sel = gidx == g
        for p in range(sim.shape[1]):
            x = sim[sel, p]
            x = x[~np.isnan(x)]
            n = x.size
            dp0 = (p0h[p, g] - p0r[p, g]) / p0h[p, g] if p0h[p, g] else np.nan
            if not (dp0 > 0) or n < 2:
                continue
            p0s = (x <= 0.05).sum() / n
            rk = np.arange(n) / (n - 1)
            expect = int((~((rk < (p0r[p, g] / p0h[p, g]) * p0s) | (rk > p0s))).sum())
            got = int((ad[sel, p] != sim[sel, p]).sum() - np.isnan(sim[sel, p]).sum())
            assert got == expect, (g, p, got, expect)
    # max_tail_factor: values above 1.5 x the last raw hist quantile are passed through un-adjusted
    hq_raw = _np(obj.ds["hist_q_raw"])
    lastq = hq_raw[:, gidx, -1].T                                                       # (time, points)
    mask = ad > 1.5 * lastq
    assert mask.sum() > 0 and bits_equal(scen[mask], ad[mask])
    plain = _np(xs.EmpiricalQuantileMapping(obj.ds, obj.group, "*", adapt_freq_thresh="0.05 mm/d").adjust(sim, time=tx))
    assert bits_equal(np.where(mask, 0, scen), np.where(mask, 0, plain))


@pytest.mark.parametrize("group,years,nq,dt,kind,var", [("time.month", 4, 20, np.float32, "+", "tas"),
                                                        ("time.month", 3, 50, np.float64, "*", "pr"),
                                                        ("time.dayofyear", 4, 30, np.float32, "+", "tas")])
@pytest.mark.parametrize("extrap", ["constant", "nan"])
def test_qdm_adjust_grouped_linear_matches_oracle(group, years, nq, dt, kind, var, extrap):
    """QDM with grouped interp="linear" (SciPy LinearNDInterpolator on the shared (quantile, group) lattice):
    sim_q bit-exact, scen within 1e-6 (f32) / 1e-12 (f64) relative of the oracle, which calls SciPy itself."""
    xs = _xs()
    window = 1
    case = (group, window, "noleap", years, nq, kind, var, dt)
    tx, to, ref, hist, sim = _make(case, n_pts=7)
    ref[:, 3] = synth.tas(np.random.default_rng(9), to, 1, "ref", dt)[:, 0] if var == "tas" else synth.pr(np.random.default_rng(9), to, 1, "ref", dt)[:, 0]
    hist[:, 3] = ref[:, 3] * 1.01; hist[:, 4] = ref[:, 4] * 0.99   # (no all-NaN tables: NaN factors are not built for linear)
    q = o.equally_spaced_nodes(nq).astype(dt)
    gidx, G, _ = o.group_index(to, group)
    af_o, _ = o.eqm_train(ref.T.copy(), hist.T.copy(), gidx, G, window, q, kind)
    assert np.isfinite(af_o).all()
    scen_o, simq_o = o.qdm_adjust(sim.T.copy(), af_o, q, group=group, time=to, window=window, interp="linear",
                                  extrapolation=extrap, kind=kind)
    out = xs.qdm_adjust(xs.Dataset({"sim": sim, "af": af_o, "quantiles": q}, time=tx), group=group, interp="linear",
                        extrapolation=extrap, kind=kind)
    assert bits_equal(_np(out.sim_q).T, simq_o)
    scen = _np(out.scen).T
    assert (np.isnan(scen) == np.isnan(scen_o)).all()
    np.testing.assert_allclose(scen, scen_o, rtol=1e-6 if dt == np.float32 else 1e-12, atol=0, equal_nan=True)


def test_escore_reference_kat_and_oracle():
    xs = _xs()
    # reference tests/test_processing.py:215-226 -- value from Cannon's MBC R package
    x = np.array([1, 4, 3, 6, 4, 7, 5, 8, 4, 5, 3, 7], dtype=np.float64).reshape(2, 6)
    y = np.array([6, 6, 3, 8, 5, 7, 3, 7, 3, 6, 4, 3], dtype=np.float64).reshape(2, 6)
    np.testing.assert_allclose(o.escore(x, y), 1.90018550338863)
    got = _np(xs.escore(x[:, :, None], y[:, :, None]))
    np.testing.assert_allclose(got, [1.90018550338863], rtol=1e-13)
    rng = np.random.default_rng(0)
    A = rng.normal(size=(4, 300, 5)).astype(np.float32); B = (rng.normal(size=(4, 250, 5)) + 0.3).astype(np.float32)
    A[1, 7, 2] = np.nan; B[0, 3, 2] = np.nan
    got = _np(xs.escore(A, B))
    want = [o.escore(A[:, :, p].astype(np.float64), B[:, :, p].astype(np.float64)) for p in range(5)]
    np.testing.assert_allclose(got, want, rtol=2e-6)
    sub = _np(xs.escore(A, B, N=100))
    want = [o.escore(A[:, ::3, p].astype(np.float64), B[:, ::3, p].astype(np.float64)) for p in range(5)]
    np.testing.assert_allclose(sub, want, rtol=2e-6)
    # scale=True: both clouds standardised with the thinned target's nanmean / population nanstd (processing.py:455-464)
    def std_pair(a, b):
        avg = np.nanmean(a, axis=1, keepdims=True); sd = np.nanstd(a, axis=1, keepdims=True)
        return (a - avg) / sd, (b - avg) / sd
    A64, B64 = A.astype(np.float64), B.astype(np.float64)
    got = _np(xs.escore(A64, B64, N=100, scale=True))
    want = [o.escore(*std_pair(A64[:, ::3, p], B64[:, ::3, p])) for p in range(5)]
    np.testing.assert_allclose(got, want, rtol=1e-10)


def test_mbcn_escores_decrease():
    """MBCn.train(n_escore>0): the energy score between ref and the transformed hist is recorded after every
    rotation (_adjustment.py:325-326) and goes down as the N-pdf transform converges."""
    xs = _xs()
    to, ref, hist, sim = _mbcn_inputs(2, 2)
    tx = xs.TimeAxis.daily(1981, 2, "noleap")
    obj = xs.MBCn.train(ref, hist, time=tx, base_kws={"nquantiles": 20, "group": "time"}, n_iter=6, n_escore=200, seed=3)
    esc = _np(obj.ds["escores"])[0]           # (points, n_iter)
    assert esc.shape == (2, 6) and np.isfinite(esc).all()
    assert (esc[:, -1] < esc[:, 0]).all()


def test_trained_object_roundtrip(tmp_path):
    """A trained adjustment is just its tables (base.py:75-100; reference tests/test_adjustment.py:186-195):
    save -> load -> adjust gives the same bits as adjusting with the in-memory object."""
    xs = _xs()
    case = ("time.dayofyear", 31, "noleap", 4, 50, "*", "pr", np.float32)
    tx, to, ref, hist, sim = _make(case, n_pts=9)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", DeprecationWarning)
        qdm = xs.QuantileDeltaMapping.train(ref, hist, time=tx, nquantiles=50, group=xs.Grouper("time.dayofyear", 31), kind="*")
        a = _np(qdm.adjust(sim, time=tx))
        qdm.save(tmp_path / "qdm.npz")
        qdm2 = xs.QuantileDeltaMapping.load(tmp_path / "qdm.npz")
        assert qdm2.group.name == "time.dayofyear" and qdm2.group.window == 31 and qdm2.kind == "*"
        assert bits_equal(_np(qdm2.adjust(sim, time=tx)), a)
    with pytest.raises(ValueError):
        xs.EmpiricalQuantileMapping.load(tmp_path / "qdm.npz")


# ---------------------------------------------------------------------------------------------
# Edge cases: ragged / degenerate inputs (tests/test_adjustment.py of the reference exercises NaN
# slices, constant series and short records through the same entry points).
# ---------------------------------------------------------------------------------------------
def _tie_aware_check(scen, scen_o, sim, af, hq, group, to, extrap, kind, dt, min_unique=0.5):
    lo, hi = o.qm_adjust_factor_bounds(sim.T.copy(), af, hq, group=group, time=to, extrapolation=extrap)
    s_lo = o.apply_correction(sim.T, lo.astype(dt), kind).astype(dt)
    s_hi = o.apply_correction(sim.T, hi.astype(dt), kind).astype(dt)
    unique = (s_lo == s_hi) | (np.isnan(s_lo) & np.isnan(s_hi))
    assert unique.mean() >= min_unique
    assert bits_equal(np.where(unique, scen, 0), np.where(unique, scen_o, 0))
    tie = ~unique
    mn, mx = np.minimum(s_lo, s_hi)[tie], np.maximum(s_lo, s_hi)[tie]
    assert ((scen[tie] >= mn) & (scen[tie] <= mx)).all()


@pytest.mark.parametrize("n_pts", [1, 31, 32, 33, 65])
def test_ragged_point_counts(n_pts):
    """Point counts around the 32-lane tile width: the partial last tile must neither read nor write out of range."""
    xs = _xs()
    case = ("time.month", 1, "noleap", 5, 50, "+", "tas", np.float32)
    tx, to, ref, hist, sim = _make(case, n_pts=max(n_pts, 6), seed=11)
    ref, hist, sim = (np.ascontiguousarray(a[:, :n_pts]) for a in (ref, hist, sim))
    q = o.equally_spaced_nodes(50).astype(np.float32)
    gidx, G, _ = o.group_index(to, "time.month")
    af_o, hq_o = o.eqm_train(ref.T.copy(), hist.T.copy(), gidx, G, 1, q, "+")
    ds = xs.eqm_train(xs.Dataset({"ref": ref, "hist": hist}, time=tx), group="time.month", kind="+", quantiles=q)
    assert bits_equal(_np(ds.af), af_o) and bits_equal(_np(ds.hist_q), hq_o)
    # guard band: the output buffer is a view into a larger poisoned allocation
    big = torch.full((sim.shape[0] + 2, n_pts + 64), -12345.0, device="cuda")
    out = xs.qm_adjust(xs.Dataset({"sim": sim, "af": af_o, "hist_q": hq_o}, time=tx), group="time.month",
                       interp="nearest", extrapolation="constant", kind="+")
    scen = _np(out.scen).T
    scen_o = o.qm_adjust(sim.T.copy(), af_o, hq_o, group="time.month", time=to, interp="nearest",
                         extrapolation="constant", kind="+")
    _tie_aware_check(scen, scen_o, sim, af_o, hq_o, "time.month", to, "constant", "+", np.float32)
    assert (big == -12345.0).all()


def test_empty_block_is_a_no_op():
    xs = _xs()
    tx, to = _time("noleap", 2)
    z = np.zeros((len(to), 0), np.float32)
    q = o.equally_spaced_nodes(10).astype(np.float32)
    ds = xs.eqm_train(xs.Dataset({"ref": z, "hist": z}, time=tx), group="time.month", kind="+", quantiles=q)
    assert tuple(ds.af.shape) == (0, 12, 10) and tuple(ds.hist_q.shape) == (0, 12, 10)
    out = xs.qm_adjust(xs.Dataset({"sim": z, "af": _np(ds.af), "hist_q": _np(ds.hist_q)}, time=tx), group="time.month",
                       interp="nearest", extrapolation="constant", kind="+")
    assert tuple(out.scen.shape) == (len(to), 0)


@pytest.mark.parametrize("group,window", [("time", 1), ("time.month", 1), ("time.dayofyear", 7)])
def test_degenerate_series(group, window):
    """Constant series (every node tied), +-inf samples, a single valid sample, and more quantiles than samples."""
    xs = _xs()
    rng = np.random.default_rng(5)
    tx, to = _time("noleap", 2)
    T, P = len(to), 8
    ref = (280 + 3 * rng.standard_normal((T, P))).astype(np.float32)
    hist = (281 + 3 * rng.standard_normal((T, P))).astype(np.float32)
    sim = (282 + 3 * rng.standard_normal((T, P))).astype(np.float32)
    hist[:, 0] = 281.0                      # constant: all quantile nodes identical
    ref[:, 1] = 280.0
    if group == "time":                     # (grouped: SciPy's cKDTree refuses non-finite nodes -- the reference raises)
        hist[::7, 2] = np.inf               # infinities sort before NaN, after everything else
        hist[3::11, 2] = -np.inf
        ref[::5, 3] = np.inf
        sim[::13, 5] = np.inf; sim[5::13, 5] = -np.inf
    hist[:, 4] = np.nan; hist[17, 4] = 279.5   # exactly one valid sample in one group only
    sim[7::13, 5] = np.nan
    hist[:, 6] = np.round(hist[:, 6])       # heavy duplicates
    q = o.equally_spaced_nodes(200).astype(np.float32)   # more nodes than samples for the small groups
    gidx, G, _ = o.group_index(to, group)
    with np.errstate(all="ignore"):
        af_o, hq_o = o.eqm_train(ref.T.copy(), hist.T.copy(), gidx, G, window, q, "+")
    ds = xs.eqm_train(xs.Dataset({"ref": ref, "hist": hist}, time=tx), group=xs.Grouper(group, window), kind="+",
                      quantiles=q)
    assert bits_equal(_np(ds.hist_q), hq_o)
    assert bits_equal(_np(ds.af), af_o)
    with np.errstate(all="ignore"):
        scen_o = o.qm_adjust(sim.T.copy(), af_o, hq_o, group=group, time=to, interp="nearest", extrapolation="constant",
                             kind="+")
    out = xs.qm_adjust(xs.Dataset({"sim": sim, "af": af_o, "hist_q": hq_o}, time=tx), group=xs.Grouper(group, window),
                       interp="nearest", extrapolation="constant", kind="+")
    scen = _np(out.scen).T
    assert np.array_equal(np.isnan(scen), np.isnan(scen_o))
    if group == "time":
        ok = np.ones(P, bool)
        ok[2] = False   # infinite nodes make hist_q non-monotonic: interp1d's searchsorted result is then arbitrary
    else:
        ok = np.ones(P, bool)
        ok[[0, 2, 4, 6]] = False   # tied / infinite node rows: cKDTree ties and inf-inf distances are unpinned
    assert bits_equal(scen[ok], scen_o[ok]) or _tie_ok(scen[ok], scen_o[ok], sim.T[ok], af_o[ok], hq_o[ok], group, to)


def _tie_ok(scen, scen_o, simT, af, hq, group, to):
    with np.errstate(all="ignore"):
        lo, hi = o.qm_adjust_factor_bounds(simT.copy(), af, hq, group=group, time=to, extrapolation="constant")
        s_lo, s_hi = (simT + lo).astype(np.float32), (simT + hi).astype(np.float32)
    unique = (s_lo == s_hi) | (np.isnan(s_lo) & np.isnan(s_hi))
    good = bits_equal(np.where(unique, scen, 0), np.where(unique, scen_o, 0))
    mn, mx = np.minimum(s_lo, s_hi)[~unique], np.maximum(s_lo, s_hi)[~unique]
    return good and ((scen[~unique] >= mn) & (scen[~unique] <= mx)).all()


def test_sim_missing_a_group_and_other_period():
    """sim covers another period than the training data and has no February at all (ragged groups, one empty)."""
    xs = _xs()
    rng = np.random.default_rng(9)
    tx_h, to_h = _time("noleap", 4, 1981)
    ref, hist = (synth.tas(rng, to_h, 9, w) for w in ("ref", "hist"))
    to_s = o.daily_time_axis(2041, 3, "noleap")
    keep = to_s.month != 2
    sim = synth.tas(rng, to_s, 9, "sim")[keep]
    tx_s = xs.TimeAxis.from_fields(to_s.year[keep], to_s.month[keep], to_s.day[keep], "noleap")
    q = o.equally_spaced_nodes(50).astype(np.float32)
    gidx, G, _ = o.group_index(to_h, "time.month")
    af_o, hq_o = o.eqm_train(ref.T.copy(), hist.T.copy(), gidx, G, 1, q, "+")
    out = xs.qm_adjust(xs.Dataset({"sim": sim, "af": af_o, "hist_q": hq_o}, time=tx_s), group="time.month",
                       interp="nearest", extrapolation="constant", kind="+")
    scen = _np(out.scen)
    # the same call on the full sim axis, February rows dropped afterwards, must agree bit for bit
    sim_full = np.full((len(to_s), 9), 280.0, np.float32); sim_full[keep] = sim
    tx_full = xs.TimeAxis.daily(2041, 3, "noleap")
    full = _np(xs.qm_adjust(xs.Dataset({"sim": sim_full, "af": af_o, "hist_q": hq_o}, time=tx_full), group="time.month",
                            interp="nearest", extrapolation="constant", kind="+").scen)
    assert bits_equal(scen, full[keep])
    scen_o = o.qm_adjust(sim_full.T.copy(), af_o, hq_o, group="time.month", time=to_s, interp="nearest",
                         extrapolation="constant", kind="+")
    _tie_aware_check(full.T, scen_o, sim_full, af_o, hq_o, "time.month", to_s, "constant", "+", np.float32, 0.95)


# ---------------------------------------------------------------------------------------------
# BASELINE.json full-size launch geometry: one bench-sized slab (48 latitude rows x 1440 = 69 120
# gridpoints x 10 950 days), checked through size-independent properties and against the oracle on
# sampled columns.
# ---------------------------------------------------------------------------------------------
def test_full_slab_properties_and_sampled_oracle():
    xs = _xs()
    P, years = 48 * 1440, 30
    tx, to = _time("noleap", years)
    T = len(to)
    g = torch.Generator(device="cuda").manual_seed(1234)
    doy = torch.as_tensor(to.dayofyear, device="cuda", dtype=torch.float32)[:, None]

    def field(A, sigma, off):
        x = torch.randn((T, P), device="cuda", generator=g) * sigma
        x += 273.15 + off - A * torch.cos(2 * torch.pi * (doy - 15) / 365)
        return x

    ref, hist, sim = field(12, 3.0, 0.0), field(10, 3.5, 1.5), field(10, 3.5, 3.5)
    hist[:, 777] = float("nan")
    ref[100:200, 4242] = float("nan")
    q = o.equally_spaced_nodes(50).astype(np.float32)
    ds = xs.eqm_train(xs.Dataset({"ref": ref, "hist": hist}, time=tx), group="time.month", kind="+", quantiles=q)
    af, hq = ds.af, ds.hist_q
    assert tuple(af.shape) == (P, 12, 50)
    # 1. quantile nodes are sorted and bracketed by the series' extremes; the all-NaN point is all-NaN
    ok = torch.ones(P, dtype=torch.bool, device="cuda"); ok[777] = False
    assert (hq[ok][:, :, 1:] >= hq[ok][:, :, :-1]).all()
    assert torch.isnan(hq[777]).all() and torch.isnan(af[777]).all()
    assert (hq[ok].amin(dim=(1, 2)) >= hist[:, ok].amin(dim=0)).all() and (hq[ok].amax(dim=(1, 2)) <= hist[:, ok].amax(dim=0)).all()
    # 2. idempotence: training a series against itself gives the neutral factor exactly
    ds0 = xs.eqm_train(xs.Dataset({"ref": hist, "hist": hist}, time=tx), group="time.month", kind="+", quantiles=q)
    assert (ds0.af[ok] == 0).all() and bits_equal(_np(ds0.hist_q), _np(hq))
    # 3. a neutral table returns sim bit for bit; the real table moves every sample by one of its own factors
    neutral = xs.qm_adjust(xs.Dataset({"sim": sim, "af": torch.zeros_like(af), "hist_q": hq}, time=tx),
                           group="time.month", interp="nearest", extrapolation="constant", kind="+").scen
    assert torch.equal(neutral[:, ok], sim[:, ok])
    scen = xs.qm_adjust(xs.Dataset({"sim": sim, "af": af, "hist_q": hq}, time=tx), group="time.month",
                        interp="nearest", extrapolation="constant", kind="+").scen
    assert torch.isnan(scen[:, 777]).all() and not torch.isnan(scen[:, ok]).any()
    # 4. sampled columns (both ends of the slab, tile edges, the NaN-holed point) against the oracle
    cols = np.array([0, 1, 31, 32, 33, 4242, 34559, 34560, 69087, 69088, 69119])
    refc, histc, simc = (_np(a[:, cols]) for a in (ref, hist, sim))
    gidx, G, _ = o.group_index(to, "time.month")
    af_o, hq_o = o.eqm_train(refc.T.copy(), histc.T.copy(), gidx, G, 1, q, "+")
    assert bits_equal(_np(af[cols]), af_o) and bits_equal(_np(hq[cols]), hq_o)
    scen_o = o.qm_adjust(simc.T.copy(), af_o, hq_o, group="time.month", time=to, interp="nearest",
                         extrapolation="constant", kind="+")
    _tie_aware_check(_np(scen[:, cols]).T, scen_o, simc, af_o, hq_o, "time.month", to, "constant", "+", np.float32, 0.99)
    # 5. checksum of checksums: the adjusted field's column sums equal sim's column sums plus the applied factors'
    d = (scen[:, ok].double() - sim[:, ok].double())
    lo, hi = af[ok].amin(dim=(1, 2)).double(), af[ok].amax(dim=(1, 2)).double()
    assert (d.amin(dim=0) >= lo - 1e-3).all() and (d.amax(dim=0) <= hi + 1e-3).all()


def test_adjust_periods_moving_windows():
    """stack_periods -> adjust -> unstack_periods (base.py:1072-1381): a 60-year sim adjusted in 30-year windows every
    10 years equals the manual per-window calls stitched at the centre strides."""
    xs = _xs()
    rng = np.random.default_rng(21)
    tx_h, to_h = _time("noleap", 30, 1981)
    ref, hist = (synth.tas(rng, to_h, 7, w) for w in ("ref", "hist"))
    tx_s = xs.TimeAxis.daily(2011, 60, "noleap")
    to_s = o.daily_time_axis(2011, 60, "noleap")
    sim = synth.tas(rng, to_s, 7, "sim")
    obj = xs.EmpiricalQuantileMapping.train(ref, hist, time=tx_h, nquantiles=50, group="time.month", kind="+")
    scen, cov = xs.adjust_periods(obj, sim, time=tx_s, window=30, stride=10)
    assert cov == slice(0, len(tx_s)) and tuple(scen.shape) == sim.shape
    scen = _np(scen)
    # manual: windows start 2011, 2021, 2031, 2041; kept years [2011, 2031) | [2031, 2041) | [2041, 2051) | [2051, 2071)
    per = xs.stack_periods(tx_s, window=30, stride=10)
    assert per.start_years == (2011, 2021, 2031, 2041)
    keep = [(2011, 2031), (2031, 2041), (2041, 2051), (2051, 2071)]
    for slc, (ya, yb) in zip(per.slices, keep):
        full = _np(obj.adjust(sim[slc], time=tx_s[slc]))
        yr = tx_s.year[slc]
        sel = (yr >= ya) & (yr < yb)
        assert bits_equal(scen[slc][sel], full[sel])


@pytest.mark.parametrize("years,dt", [(30, np.float32), (12, np.float64), (30, np.float64)])
def test_long_segments_group_time(years, dt):
    """group="time" over decades: 4k-11k-row segments, the narrow-tile sorter (C = 4, 2, 1 columns per CTA).
    BASELINE.json config 0 is this shape (EQM, nquantiles=50, group='time', 30 years)."""
    xs = _xs()
    case = ("time", 1, "noleap", years, 50, "+", "tas", dt)
    tx, to, ref, hist, sim = _make(case, n_pts=9, seed=4)
    q = o.equally_spaced_nodes(50).astype(dt)
    gidx, G, _ = o.group_index(to, "time")
    af_o, hq_o = o.eqm_train(ref.T.copy(), hist.T.copy(), gidx, G, 1, q, "+")
    ds = xs.eqm_train(xs.Dataset({"ref": ref, "hist": hist}, time=tx), group="time", kind="+", quantiles=q)
    if dt == np.float32:
        assert bits_equal(_np(ds.hist_q), hq_o) and bits_equal(_np(ds.af), af_o)
    else:
        np.testing.assert_allclose(_np(ds.hist_q), hq_o, rtol=1e-12, atol=0, equal_nan=True)
    # QDM adjust ranks every sample inside the full series (rank kernel on the same sorter)
    scen_o, simq_o = o.qdm_adjust(sim.T.copy(), af_o, q, group="time", time=to, window=1, interp="nearest",
                                  extrapolation="constant", kind="+")
    out = xs.qdm_adjust(xs.Dataset({"sim": sim, "af": af_o, "quantiles": q}, time=tx), group="time", interp="nearest",
                        extrapolation="constant", kind="+")
    assert bits_equal(_np(out.sim_q).T, simq_o)
    assert bits_equal(_np(out.scen).T, scen_o)


def test_generic_train_kernel_jitter_and_signed_zeros():
    """The generic (float64 / point-major) train kernel: jitter through its batched segment load, and signed zeros
    through the compare + select exchanges of the float64 sorter (no value may be lost or duplicated)."""
    xs = _xs()
    rng = np.random.default_rng(12)
    to = o.daily_time_axis(1981, 10, "noleap"); tx = xs.TimeAxis.daily(1981, 10, "noleap")
    ref, hist = (synth.pr(rng, to, 6, w, jitter=False, nan_frac=0).astype(np.float64) for w in ("ref", "hist"))
    q = o.equally_spaced_nodes(20)
    gidx, G, _ = o.group_index(to, "time.month")
    # 1. float64 + jitter: statistically the oracle's, exactly the oracle's above the threshold
    ds = xs.eqm_train(xs.Dataset({"ref": ref, "hist": hist}, time=tx), group="time.month", kind="*", quantiles=q,
                      jitter_under_thresh_value="0.01 mm/d")
    hist_j = np.where(hist < 0.01, rng.uniform(1e-45, 0.01, hist.shape), hist)
    _, hq_o = o.eqm_train(ref.T.copy(), hist_j.T.copy(), gidx, G, 1, q, "*")
    hq = _np(ds.hist_q)
    np.testing.assert_allclose(hq, hq_o, atol=2e-3, rtol=0.02)
    wet = hq_o > 0.1    # (a node just above the threshold may still interpolate towards a jittered neighbour)
    np.testing.assert_allclose(hq[wet], hq_o[wet], rtol=1e-12)
    # 2. signed zeros (dry days as +0.0 and -0.0 mixed): same quantiles as the oracle, zeros counted once each
    z = hist.copy()
    z[(z < 0.01) & (rng.random(z.shape) < 0.5)] = -0.0
    z[(z < 0.01) & (z != 0)] = 0.0
    ds = xs.eqm_train(xs.Dataset({"ref": ref, "hist": z}, time=tx), group="time.month", kind="+", quantiles=q)
    _, hq_o = o.eqm_train(ref.T.copy(), z.T.copy(), gidx, G, 1, q, "+")
    np.testing.assert_allclose(_np(ds.hist_q), hq_o, rtol=1e-12, atol=0)
    # long segment (narrow tile, warp-shuffle exchanges) with the same data
    ds = xs.eqm_train(xs.Dataset({"ref": ref, "hist": z}, time=tx), group="time", kind="+", quantiles=q)
    gidx1, G1, _ = o.group_index(to, "time")
    _, hq_o = o.eqm_train(ref.T.copy(), z.T.copy(), gidx1, G1, 1, q, "+")
    np.testing.assert_allclose(_np(ds.hist_q), hq_o, rtol=1e-12, atol=0)


@pytest.mark.parametrize("kind", ["+", "*"])
def test_poly_trend_preserve_mean(kind):
    """PolyDetrend(preserve_mean=True) (detrending.py:205): the group's trend loses its own mean."""
    xs = _xs()
    rng = np.random.default_rng(8)
    tx, to = _time("noleap", 6)
    x = synth.tas(rng, to, 7, "sim", np.float64)
    gidx, G, _ = o.group_index(to, "time.month")
    want = o.group_trend_poly(x.T.copy(), gidx, G, 1, o.time_ordinal(to), 1, preserve_mean=True, kind=kind)
    got = _np(xs.poly_trend(x, time=tx, group="time.month", degree=1, kind=kind, preserve_mean=True)).T
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-9, equal_nan=True)
    sim = synth.tas(rng, to, 7, "sim", np.float32)
    out = xs.dqm_adjust(xs.Dataset({"sim": sim, "af": np.zeros((7, G, 4), np.float32),
                                    "hist_q": np.tile(np.linspace(-20, 20, 4, dtype=np.float32), (7, G, 1)),
                                    "scaling": np.zeros((7, G), np.float32)}, time=tx), group="time.month",
                        interp="nearest", extrapolation="constant", kind="+",
                        detrend=xs.PolyDetrend(degree=1, kind="+", group="time.month", preserve_mean=True))
    np.testing.assert_allclose(_np(out.scen), sim, rtol=2e-6, equal_nan=True)   # zero factors: detrend + retrend = identity


def test_host_entry_two_threads_concurrently():
    """VERDICT r1 item 9 / ABI note "re-entrant": two host threads call the end-to-end entry at the same time (ctypes
    releases the GIL; each thread owns its device workspace and streams); results equal the sequential ones bit for bit."""
    import threading
    xs = _xs()
    rng = np.random.default_rng(91)
    to = o.daily_time_axis(1981, 6, "noleap")
    tx = xs.TimeAxis.daily(1981, 6, "noleap"); ts = xs.TimeAxis.daily(2041, 6, "noleap")
    data = []
    for k in range(2):
        ref, hist, sim = (synth.tas(rng, to, 3000 + 512 * k, w, np.float32) for w in ("ref", "hist", "sim"))
        sim[::97, 5] = np.nan
        data.append((ref, hist, sim))
    kw = dict(time=tx, sim_time=ts, nquantiles=30, group="time.month", kind="+", method="eqm", slab_points=1024)
    seq = [xs.train_adjust_host(*d, **kw) for d in data]
    for _ in range(3):
        res, err = [None, None], []

        def work(i):
            try:
                res[i] = xs.train_adjust_host(*data[i], **kw)
            except Exception as e:  # pragma: no cover
                err.append(e)
        th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
        [t.start() for t in th]
        [t.join() for t in th]
        assert not err, err
        for a, b in zip(seq, res):
            assert bits_equal(np.asarray(a), np.asarray(b))


def test_train_row_stride_beyond_int32_bytes_takes_the_generic_kernel():
    """Launcher guard of the float32 fast kernels (`st * 4` must fit 31 bits): a time stride of 2^29 + 32 elements goes to
    the generic kernel and gives the same tables as the contiguous copy."""
    xs = _xs()
    from xsdba_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    T, n, nq = 3, 40, 3
    st = (1 << 29) + 32
    free, _ = torch.cuda.mem_get_info()
    if free < 2 * (2 * st + n) * 4 + (1 << 30):
        pytest.skip("not enough device memory for two strided series")
    gen = torch.Generator(device=dev); gen.manual_seed(3)
    big = [torch.empty(2 * st + n, dtype=torch.float32, device=dev) for _ in range(2)]
    small = []
    for b in big:
        v = torch.randn((T, n), generator=gen, device=dev, dtype=torch.float32)
        for t in range(T):
            b[t * st:t * st + n] = v[t]
        small.append(v.contiguous())
    tx = xs.TimeAxis.daily(2001, 1, "noleap")
    from xsdba_b200.base import grouping_handle
    h = grouping_handle(np.zeros(T, np.int32), 1, 1)
    q = torch.tensor([0.25, 0.5, 0.75], dtype=torch.float32, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    outs = []
    for ref, hist, stride in ((big[0], big[1], st), (small[0], small[1], n)):
        af = torch.empty((n, 1, nq), device=dev); hq = torch.empty_like(af)
        _lib.check(lib.xsdba_qm_train_f32(ref.data_ptr(), hist.data_ptr(), n, 1, stride, h.ptr, q.data_ptr(), nq, 43, 0,
                                          af.data_ptr(), hq.data_ptr(), None, s))
        torch.cuda.synchronize()
        outs.append((_np(af), _np(hq)))
    assert bits_equal(outs[0][0], outs[1][0]) and bits_equal(outs[0][1], outs[1][1])
    del big
    torch.cuda.empty_cache()


def test_loess_gaussian_weights_reference_golden():
    """CUDA LOESS with `weights="gaussian"` against the reference's numba kernel (golden) on the per-point path, and
    against the restatement on complete series (shared-table paths) and with robustness iterations."""
    import os
    xs = _xs()
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loess_gaussian.npz"))
    x, y = g["loess_x"], g["loess_y"]
    n = x.size
    t = xs.TimeAxis.daily(2001, 1, "noleap")[:n]
    series = np.stack([y, y[::-1].copy()], axis=1)
    for k in range(3):
        d, f, niter, dx = g[f"case{k}_params"]
        got = _np(xs.loess_trend(series, time=t, f=float(f), niter=int(niter), d=int(d), weights="gaussian"))
        np.testing.assert_allclose(got[:, 0], g[f"case{k}_out"], rtol=1e-9, atol=1e-10, equal_nan=True)
        want = o.loess_nb(x, series[:, 1], f=float(f), niter=int(niter), weights="gaussian", d=int(d), dx=float(dx))
        np.testing.assert_allclose(got[:, 1], want, rtol=1e-9, atol=1e-10, equal_nan=True)
    rng = np.random.default_rng(9)
    t3 = xs.TimeAxis.daily(2001, 3, "noleap")
    m = len(t3)
    xx = np.arange(m, dtype=np.float64)
    yy = (280 + 5 * np.sin(2 * np.pi * xx / 365)[:, None] + rng.standard_normal((m, 40))).astype(np.float32)
    yy[100:130, 35] = np.nan
    got = _np(xs.loess_trend(yy, time=t3, f=0.2, niter=1, d=0, weights="gaussian"))
    # (the abscissa exactly as loess_smoothing and the library build it, loess.py:244-245: the gaussian kernel is
    #  discontinuous at the window edge, so the last bit of |x_j - x_i| / h decides whether the edge sample counts)
    xx = (xx - xx[0]) / (xx[-1] - xx[0])
    for j in (0, 31, 32, 35, 39):
        want = o.loess_nb(xx, yy[:, j].astype(np.float64), f=0.2, niter=1, weights="gaussian", d=0, dx=float(xx[1] - xx[0]))
        np.testing.assert_allclose(got[:, j], want, rtol=1e-11, atol=1e-12, equal_nan=True)
    tas = xs.LoessDetrend(group="time", kind="+", f=0.2, niter=1, d=0, weights="gaussian")
    assert tas.weights == "gaussian"


def test_loess_unequal_spacing_reference_golden(golden):
    """The dx == 0 branch (loess.py:107-111, 151-158): golden cases 2 (d=0, f=0.3, niter=3) and 3 (d=1, f=0.2, niter=1)
    of the reference's numba _loess_nb, then a genuinely irregular time axis against the restatement."""
    xs = _xs()
    x, y = golden["loess_x"], golden["loess_y"]
    n = x.size
    t = xs.TimeAxis.daily(2001, 1, "noleap")[:n]
    series = np.stack([y, y[::-1].copy(), np.where(np.arange(n) % 17 == 3, np.nan, y)], axis=1)
    for k in (2, 3):
        d, f, niter, dx = golden[f"loess_case{k}_params"]
        assert dx == 0.0
        got = _np(xs.loess_trend(series, time=t, f=float(f), niter=int(niter), d=int(d), equal_spacing=False))
        np.testing.assert_allclose(got[:, 0], golden[f"loess_case{k}_out"], rtol=1e-9, atol=1e-10, equal_nan=True)
        for j in (1, 2):
            want = o.loess_nb(x, series[:, j], f=float(f), niter=int(niter), d=int(d), dx=0.0)
            np.testing.assert_allclose(got[:, j], want, rtol=1e-9, atol=1e-10, equal_nan=True)
    rng = np.random.default_rng(12)
    full = xs.TimeAxis.daily(2001, 2, "standard")
    keep = np.sort(rng.choice(len(full), size=500, replace=False))
    ti = full[keep]
    og = np.asarray(ti.ordinal, np.float64)
    xi = (og - og[0]) / (og[-1] - og[0])
    yy = (np.sin(9 * xi)[:, None] + 0.3 * rng.standard_normal((500, 37))).astype(np.float32)
    yy[40:60, 5] = np.nan
    for d, f, niter, w in ((0, 0.2, 1, "tricube"), (1, 0.3, 2, "tricube"), (0, 0.25, 1, "gaussian")):
        got = _np(xs.loess_trend(yy, time=ti, f=f, niter=niter, d=d, weights=w))     # spacing detected: unequal
        for j in (0, 5, 31, 36):
            want = o.loess_nb(xi, yy[:, j].astype(np.float64), f=f, niter=niter, weights=w, d=d, dx=0.0)
            np.testing.assert_allclose(got[:, j], want, rtol=1e-9, atol=1e-10, equal_nan=True)


def test_loess_grouped_detrend_matches_restatement():
    """LoessDetrend(group="time.month") (detrending.py:211-296 through map_groups): every month's members are smoothed
    on their own normalised time coordinate (unequally spaced: the dx == 0 branch), also through dqm_adjust."""
    xs = _xs()
    rng = np.random.default_rng(13)
    tx = xs.TimeAxis.daily(1981, 4, "noleap")
    T = len(tx)
    og = np.asarray(tx.ordinal, np.float64)
    y = (280 + 4 * np.sin(2 * np.pi * np.arange(T) / 365)[:, None] + rng.standard_normal((T, 9))).astype(np.float32)
    y[100:140, 3] = np.nan
    got = _np(xs.loess_trend(y, time=tx, f=0.3, niter=1, d=0, loess_group="time.month"))
    month = tx.month
    for j in (0, 3, 8):
        want = np.full(T, np.nan)
        for m in range(1, 13):
            rows = np.nonzero(month == m)[0]
            xg = (og[rows] - og[rows][0]) / (og[rows][-1] - og[rows][0])
            want[rows] = o.loess_nb(xg, y[rows, j].astype(np.float64), f=0.3, niter=1, d=0, dx=0.0)
        np.testing.assert_allclose(got[:, j], want, rtol=1e-9, atol=1e-10, equal_nan=True)
    # point-major layout gives the same trend
    got_pm = _np(xs.loess_trend(np.ascontiguousarray(y.T), time=tx, f=0.3, niter=1, d=0, loess_group="time.month", time_axis=-1))
    np.testing.assert_array_equal(got_pm.T, got)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_cubic_interpolation_group_time_matches_scipy(dt):
    """interp="cubic" with group="time": scipy interp1d(kind="cubic") (the not-a-knot spline of make_interp_spline)
    through the oracle, for EQM, QDM and DQM adjust.  Tolerances: 1e-6 (float32) / 1e-9 (float64) relative -- the spline
    is solved by the second-derivative tridiagonal system here and by B-spline collocation in SciPy."""
    xs = _xs()
    case = ("time", 1, "noleap", 3, 20, "+", "tas", dt)
    tx, to, ref, hist, sim = _make(case, n_pts=11)
    q = o.equally_spaced_nodes(20).astype(dt)
    gidx, G, _ = o.group_index(to, "time")
    af_o, hq_o = o.eqm_train(ref.T.copy(), hist.T.copy(), gidx, G, 1, q, "+")
    rtol = 1e-6 if dt == np.float32 else 1e-9
    for extrap in ("constant", "nan"):
        scen_o = o.qm_adjust(sim.T.copy(), af_o, hq_o, group="time", time=to, interp="cubic", extrapolation=extrap, kind="+")
        out = xs.qm_adjust(xs.Dataset({"sim": sim, "af": af_o, "hist_q": hq_o}, time=tx), group="time", interp="cubic",
                           extrapolation=extrap, kind="+")
        np.testing.assert_allclose(_np(out.scen).T, scen_o, rtol=rtol, atol=0, equal_nan=True)
    scen_q, simq_o = o.qdm_adjust(sim.T.copy(), af_o, q, group="time", time=to, window=1, interp="cubic",
                                  extrapolation="constant", kind="+")
    out = xs.qdm_adjust(xs.Dataset({"sim": sim, "af": af_o, "quantiles": q}, time=tx), group="time", interp="cubic",
                        extrapolation="constant", kind="+")
    assert bits_equal(_np(out.sim_q).T, simq_o)
    np.testing.assert_allclose(_np(out.scen).T, scen_q, rtol=rtol, atol=0, equal_nan=True)
    af_d, hq_d, sc_d = o.dqm_train(ref.T.copy(), hist.T.copy(), gidx, G, 1, q, "+")
    scen_d, trend_d = o.dqm_adjust(sim.T.copy(), af_d, hq_d, sc_d, group="time", window=1, time=to, interp="cubic",
                                   extrapolation="constant", kind="+", detrend=1)
    out = xs.dqm_adjust(xs.Dataset({"sim": sim, "af": af_d, "hist_q": hq_d, "scaling": sc_d}, time=tx), group="time",
                        interp="cubic", extrapolation="constant", kind="+", detrend=1)
    np.testing.assert_allclose(_np(out.scen).T, scen_d, rtol=max(rtol, 2e-6), atol=0, equal_nan=True)
    with pytest.raises(NotImplementedError):
        xs.qm_adjust(xs.Dataset({"sim": sim, "af": af_o, "hist_q": hq_o}, time=tx), group="time.month", interp="cubic",
                     extrapolation="constant", kind="+")
