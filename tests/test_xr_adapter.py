"""The drop-in seam (xsdba_b200/xr_adapter.py): marshalling between the reference's Dataset contract and the array-
level CUDA mirror, `patch()` / `unpatch()`, `.func`.  CPU part: a stub backend records what the adapter hands down
and returns recognisable tables, so dimension order, dtype widening, coordinates and the unit-string thresholds are
checked without a GPU.  GPU part: the real kernels through the adapter against the oracle."""
import types

import numpy as np
import pytest

import qm_oracle as o
import synth
import xr_shim as xr
from conftest import bits_equal

torch = pytest.importorskip("torch")


def _inputs(dt=np.float32, years=3, nlat=3, nlon=4, seed=0):
    import xsdba_b200 as xs
    rng = np.random.default_rng(seed)
    tx = xs.TimeAxis.daily(1981, years, "noleap")
    to = o.daily_time_axis(1981, years, "noleap")
    ref, hist, sim = (synth.tas(rng, to, nlat * nlon, w, dt) for w in ("ref", "hist", "sim"))
    tidx = xr.time_index(tx)
    coords = {"time": tidx, "lat": np.linspace(40, 42, nlat), "lon": np.linspace(-75, -72, nlon)}

    def da(a, dims):   # stored (lat, time, lon): the adapter must not depend on the dimension order
        x = a.reshape(len(tx), nlat, nlon)
        x = np.transpose(x, [("time", "lat", "lon").index(d) for d in dims])
        return xr.DataArray(x, dims, {d: coords[d] for d in dims}, attrs={"units": "K"})
    return tx, to, ref, hist, sim, da


def test_marshalling_with_stub_backend(monkeypatch):
    import xsdba_b200 as xs
    from xsdba_b200 import xr_adapter as A
    tx, to, ref, hist, sim, da = _inputs(np.float32)
    seen = {}

    def fake_train(ds, *, group, kind, quantiles, **kw):
        seen.update(ref=np.asarray(ds["ref"]), hist=np.asarray(ds["hist"]), group=group, kw=kw, time=ds.time)
        n, G, nq = ds["ref"].shape[1], 12, len(quantiles)
        af = torch.arange(n * G * nq, dtype=torch.float64).reshape(n, G, nq)
        return xs.Dataset({"af": af, "hist_q": af + 0.5, "hist_q_raw": None, "quantiles": quantiles})
    monkeypatch.setattr(A.L4, "eqm_train", fake_train)
    ds = xr.Dataset({"ref": da(ref, ("lat", "time", "lon")), "hist": da(hist.astype(np.float64), ("lon", "lat", "time"))})
    q = xs.equally_spaced_nodes(5)
    out = A.eqm_train(ds, group=types.SimpleNamespace(name="time.month", window=1, add_dims=[]), kind="+", quantiles=q,
                      jitter_under_thresh_value="0.01 K")
    # time-major, points flattened in the order of ref's non-time dims (lat, lon); widest dtype wins
    assert seen["ref"].dtype == np.float64 and seen["ref"].shape == (len(tx), 12)
    np.testing.assert_array_equal(seen["ref"], ref.astype(np.float64))
    np.testing.assert_array_equal(seen["hist"], hist.astype(np.float64))
    assert seen["group"].name == "time.month" and seen["kw"]["jitter_under_thresh_value"] == 0.01
    assert (seen["time"].month == tx.month).all() and seen["time"].calendar == "noleap"
    assert out["af"].dims == ("lat", "lon", "month", "quantiles") and out["af"].shape == (3, 4, 12, 5)
    np.testing.assert_array_equal(out["af"].values.ravel(), np.arange(12 * 12 * 5))
    np.testing.assert_array_equal(out["af"].coords["month"], np.arange(1, 13))
    np.testing.assert_array_equal(out["af"].coords["quantiles"], q)
    np.testing.assert_array_equal(out["af"].coords["lat"], ds["lat"].values)
    for dummy, dims in (("hist_q_raw", ("lat", "lon", "month", "quantiles")), ("P0_ref", ("lat", "lon", "month")),
                        ("P0_hist", ("lat", "lon", "month")), ("pth", ("lat", "lon", "month"))):
        assert out[dummy].dims == dims and np.isnan(out[dummy].values).all()   # _adjustment.py:277-286
    assert isinstance(out, xr.Dataset) and isinstance(out["af"], xr.DataArray)
    with pytest.raises(ValueError):
        A.eqm_train(ds, group="time.month", kind="+", quantiles=q, jitter_under_thresh_value="0.01 degC")


def test_add_dims_pooling_layout_with_stub_backend(monkeypatch):
    """Grouper(add_dims=[...]) (base.py:410-415) on CPU: what the adapter hands to the array-level train -- the pooled
    dimension laid end to end along time with window/2 NaN spacer steps that belong to no group -- and the tables it
    repeats over that dimension for adjust."""
    import xsdba_b200 as xs
    from xsdba_b200 import xr_adapter as A
    tx = xs.TimeAxis.daily(1981, 2, "noleap")
    T, R, P = len(tx), 3, 2
    rng = np.random.default_rng(4)
    ref = rng.standard_normal((T, R, P)).astype(np.float32)
    coords = {"time": xr.time_index(tx)}
    da = lambda a, dims: xr.DataArray(a, dims, coords)  # noqa: E731
    seen = {}

    def fake_train(ds, *, group, kind, quantiles, **kw):
        seen.update(ref=np.asarray(ds["ref"]), group=group, time=ds.time)
        n, G, nq = ds["ref"].shape[1], group.n_groups(ds.time), len(quantiles)
        af = torch.zeros((n, G, nq), dtype=torch.float32)
        return xs.Dataset({"af": af, "hist_q": af, "hist_q_raw": None, "quantiles": quantiles})
    monkeypatch.setattr(A.L4, "eqm_train", fake_train)
    ds = xr.Dataset({"ref": da(ref, ("time", "realization", "pt")), "hist": da(np.transpose(ref, (1, 2, 0)), ("realization", "pt", "time"))})
    grp = types.SimpleNamespace(name="time.dayofyear", window=5, add_dims=["realization"], prop="dayofyear")
    out = A.eqm_train(ds, group=grp, kind="+", quantiles=xs.equally_spaced_nodes(4))
    gap = 2
    pooled = seen["ref"]
    assert pooled.shape == (R * (T + gap), P) and len(seen["time"]) == R * (T + gap)
    gidx = seen["group"].zero_based_index(seen["time"]).reshape(R, T + gap)
    for r in range(R):
        np.testing.assert_array_equal(pooled.reshape(R, T + gap, P)[r, :T], ref[:, r, :])
        assert np.isnan(pooled.reshape(R, T + gap, P)[r, T:]).all()
        np.testing.assert_array_equal(gidx[r, :T], tx.dayofyear - 1)
        assert (gidx[r, T:] == -1).all()
    assert seen["group"].n_groups(seen["time"]) == 365 and seen["group"].window == 5 and not seen["group"].add_dims
    assert out["af"].dims == ("pt", "dayofyear", "quantiles")        # the pooled dimension is gone from the tables
    with pytest.raises(NotImplementedError):
        A.eqm_train(ds, group=grp, kind="+", quantiles=xs.equally_spaced_nodes(4), adapt_freq_thresh="1 mm/d")


def test_adjust_marshalling_with_stub_backend(monkeypatch):
    import xsdba_b200 as xs
    from xsdba_b200 import xr_adapter as A
    tx, to, ref, hist, sim, da = _inputs(np.float32)
    seen = {}

    def fake_adjust(ds, *, group, interp, extrapolation, kind, **kw):
        seen.update({k: np.asarray(v) for k, v in ds.items()}, kw=kw, interp=interp)
        return xs.Dataset({"scen": torch.from_numpy(np.asarray(ds["sim"]) + 1)})
    monkeypatch.setattr(A.L4, "qm_adjust", fake_adjust)
    af = np.random.default_rng(1).normal(size=(12, 7, 3, 4)).astype(np.float32)       # (month, quantiles, lat, lon)
    tab = lambda a: xr.DataArray(a, ("month", "quantiles", "lat", "lon"))
    nan_g = xr.DataArray(np.full((3, 4, 12), np.nan, np.float32), ("lat", "lon", "month"))
    ds = xr.Dataset({"af": tab(af), "hist_q": tab(af * 2), "hist_q_raw": tab(af * np.nan), "P0_ref": nan_g, "P0_hist": nan_g,
                     "pth": nan_g, "sim": da(sim, ("lon", "time", "lat"))})
    out = A.qm_adjust(ds, group="time.month", interp="nearest", extrapolation="constant", kind="+")
    np.testing.assert_array_equal(seen["sim"], sim.reshape(len(tx), 3, 4).transpose(0, 2, 1).reshape(len(tx), 12))
    np.testing.assert_array_equal(seen["af"], af.transpose(3, 2, 0, 1).reshape(12, 12, 7))   # points in (lon, lat) order
    assert "hist_q_raw" not in seen and "P0_ref" not in seen     # NaN dummies are not handed down
    assert out["scen"].dims == ("lon", "time", "lat")
    np.testing.assert_array_equal(out["scen"].values, ds["sim"].values + 1)


def test_patch_rebinds_and_restores():
    from xsdba_b200 import xr_adapter as A
    mod_a = types.ModuleType("fake_xsdba_adjustment")
    mod_b = types.ModuleType("fake_xsdba__adjustment")
    for m in (mod_a, mod_b):
        for n in A.NAMES:
            setattr(m, n, f"orig_{n}")
    A.patch([mod_a, mod_b])
    try:
        for n in A.NAMES:
            f = getattr(mod_a, n)
            assert f is getattr(A, n) and getattr(mod_b, n) is f and f.func is f     # .func: base.py:723, 775
    finally:
        A.unpatch()
    assert all(getattr(mod_a, n) == f"orig_{n}" and getattr(mod_b, n) == f"orig_{n}" for n in A.NAMES)


def test_patch_without_xsdba_fails_loudly():
    from xsdba_b200 import xr_adapter as A
    with pytest.raises(ImportError):
        A.patch()


@pytest.mark.gpu
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_adapter_end_to_end_matches_oracle(dt):
    """EQM / QDM / DQM through the adapter, called the way the reference's Adjustment classes call the L4 functions
    (adjustment.py:485-528, 615-668, 727-742), against the oracle."""
    import xsdba_b200 as xs
    from xsdba_b200 import xr_adapter as A
    tx, to, ref, hist, sim, da = _inputs(dt, years=4)
    q = o.equally_spaced_nodes(20).astype(dt)
    grp = types.SimpleNamespace(name="time.month", window=1, add_dims=[], prop="month")
    gidx, G, _ = o.group_index(to, "time.month")
    ds = xr.Dataset({"ref": da(ref, ("lat", "time", "lon")), "hist": da(hist, ("time", "lon", "lat"))})
    tr = A.eqm_train(ds, group=grp, kind="+", quantiles=q)
    af_o, hq_o = o.eqm_train(ref.T.copy(), hist.T.copy(), gidx, G, 1, q, "+")
    assert tr["af"].dtype == dt
    assert bits_equal(tr["af"].values.reshape(12, G, 20), af_o) and bits_equal(tr["hist_q"].values.reshape(12, G, 20), hq_o)
    trained = tr.drop_vars(["P0_ref", "P0_hist", "pth"])                                  # as EQM._train does
    out = A.qm_adjust(trained.assign(sim=da(sim, ("lon", "lat", "time"))), group=grp, interp="nearest",
                      extrapolation="constant", kind="+", adapt_freq_thresh=None, max_tail_factor=None)
    scen_o = o.qm_adjust(sim.T.copy(), af_o, hq_o, group="time.month", time=to, interp="nearest", extrapolation="constant",
                         kind="+")
    scen = out["scen"].transpose("lat", "lon", "time").values.reshape(12, -1)
    lo, hi = o.qm_adjust_factor_bounds(sim.T.copy(), af_o, hq_o, group="time.month", time=to, extrapolation="constant")
    unique = (lo == hi) | (np.isnan(lo) & np.isnan(hi))
    assert unique.mean() > 0.99 and bits_equal(np.where(unique, scen, 0), np.where(unique, scen_o, 0))
    # QDM: the class passes rank_window=None and reads .scen / .sim_q
    out = A.qdm_adjust(trained.assign(sim=da(sim, ("time", "lat", "lon"))), group=grp, interp="nearest",
                       extrapolation="constant", kind="+", adapt_freq_thresh=None, rank_window=None, max_tail_factor=None)
    scen_q, simq_o = o.qdm_adjust(sim.T.copy(), af_o, q, group="time.month", time=to, window=1, interp="nearest",
                                  extrapolation="constant", kind="+")
    assert bits_equal(out["sim_q"].values.reshape(len(tx), 12).T, simq_o)
    # DQM with the default detrend=1 and with a reference-style detrend object
    trd = A.dqm_train(ds, group=grp, kind="+", quantiles=q)
    af_d, hq_d, sc_d = o.dqm_train(ref.T.copy(), hist.T.copy(), gidx, G, 1, q, "+")
    np.testing.assert_allclose(trd["scaling"].values.reshape(12, G), sc_d, rtol=2e-6)
    out = A.dqm_adjust(trd.drop_vars(["P0_ref", "P0_hist", "pth"]).assign(sim=da(sim, ("time", "lat", "lon"))), group=grp,
                       interp="nearest", extrapolation="constant", kind="+", detrend=1)
    assert out["scen"].dims == ("time", "lat", "lon") and out["trend"].shape == out["scen"].shape
    class PolyDetrend:   # what a reference detrend object looks like from outside: class name + parameters dict
        parameters = {"degree": 1, "kind": "+", "group": grp}
    poly = PolyDetrend()
    out2 = A.dqm_adjust(trd.drop_vars(["P0_ref", "P0_hist", "pth"]).assign(sim=da(sim, ("time", "lat", "lon"))), group=grp,
                        interp="nearest", extrapolation="constant", kind="+", detrend=poly)
    assert bits_equal(out2["scen"].values, out["scen"].values)


@pytest.mark.gpu
@pytest.mark.parametrize("group,window", [("time.month", 1), ("time.dayofyear", 31)])
def test_adapter_add_dims_pools_realizations(group, window):
    """Grouper(..., add_dims=["realization"]) (base.py:410-415): train pools the extra dimension into every group's
    sample -- checked against nan_quantile on the concatenated group segments of all realizations -- and adjust applies
    the pooled tables to every realization."""
    from xsdba_b200 import xr_adapter as A
    dt = np.float32
    rng = np.random.default_rng(17)
    years, R, P = 4, 3, 5
    import xsdba_b200 as xs
    tx = xs.TimeAxis.daily(1981, years, "noleap")
    to = o.daily_time_axis(1981, years, "noleap")
    T = len(tx)
    ref = (280 + 3 * rng.standard_normal((T, R, P))).astype(dt)
    hist = (282 + 4 * rng.standard_normal((T, R, P))).astype(dt)
    hist[5:40, 1, 2] = np.nan
    sim = (283 + 4 * rng.standard_normal((T, R, P))).astype(dt)
    q = o.equally_spaced_nodes(15).astype(dt)
    coords = {"time": xr.time_index(tx)}
    da = lambda a, dims: xr.DataArray(a, dims, coords)  # noqa: E731
    grp = types.SimpleNamespace(name=group, window=window, add_dims=["realization"], prop=group.split(".")[1])
    ds = xr.Dataset({"ref": da(ref, ("time", "realization", "pt")), "hist": da(np.transpose(hist, (1, 0, 2)), ("realization", "time", "pt"))})
    tr = A.eqm_train(ds, group=grp, kind="+", quantiles=q)
    assert tr["af"].dims == ("pt", grp.prop, "quantiles")
    gidx, G, _ = o.group_index(to, group)
    hq = tr["hist_q"].values
    af = tr["af"].values
    for g in (0, 1, G // 2, G - 1):
        seg_h = np.concatenate([o.group_segment(hist[:, r, :].T.copy(), gidx, g, window) for r in range(R)], axis=1)
        seg_r = np.concatenate([o.group_segment(ref[:, r, :].T.copy(), gidx, g, window) for r in range(R)], axis=1)
        hq_o = o.nan_quantile(seg_h, q)
        rq_o = o.nan_quantile(seg_r, q)
        assert bits_equal(hq[:, g], hq_o), g
        assert bits_equal(af[:, g], (rq_o - hq_o).astype(dt)), g
    trained = tr.drop_vars(["P0_ref", "P0_hist", "pth", "hist_q_raw"])
    out = A.qm_adjust(trained.assign(sim=da(sim, ("time", "realization", "pt"))), group=grp, interp="nearest",
                      extrapolation="constant", kind="+", adapt_freq_thresh=None, max_tail_factor=None)
    assert out["scen"].dims == ("time", "realization", "pt")
    grp0 = types.SimpleNamespace(name=group, window=window, add_dims=[], prop=grp.prop)
    for r in range(R):   # every realization separately with the same (pooled) tables
        one = A.qm_adjust(trained.assign(sim=da(sim[:, r, :], ("time", "pt"))), group=grp0, interp="nearest",
                          extrapolation="constant", kind="+", adapt_freq_thresh=None, max_tail_factor=None)
        assert bits_equal(out["scen"].values[:, r, :], one["scen"].values)
