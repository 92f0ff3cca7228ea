"""GPU parity of the float32 time-major train kernels: K1b (bucket select, the default) against K1f (sorting
network, XSDBA_B200_TRAIN_ALGO=sort) and against the oracle, on the inputs that stress the bucket map -- exact
zeros, jittered multi-scale precipitation, +-inf, -0.0, constant and heavily tied columns, tiny valid counts, a
group without members -- at the full 30-year segment sizes.  Everything is bit-exact (NaN == NaN)."""
import os

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


def _np(t):
    return t.detach().cpu().numpy()


def _train(xs, ref, hist, tx, group, q, kind, algo, **kw):
    old = os.environ.pop("XSDBA_B200_TRAIN_ALGO", None)
    if algo == "sort":
        os.environ["XSDBA_B200_TRAIN_ALGO"] = "sort"
    try:
        fn = kw.pop("fn", xs.eqm_train)
        ds = fn(xs.Dataset({"ref": ref, "hist": hist}, time=tx), group=group, kind=kind, quantiles=q, **kw)
        torch.cuda.synchronize()
        return ds
    finally:
        os.environ.pop("XSDBA_B200_TRAIN_ALGO", None)
        if old is not None:
            os.environ["XSDBA_B200_TRAIN_ALGO"] = old


def _stress_inputs(rng, to, P, var):
    gen = getattr(synth, var)
    kw = {} if var == "tas" else {"jitter": False}
    ref = gen(rng, to, P, "ref", np.float32, **kw)
    hist = gen(rng, to, P, "hist", np.float32, **kw)
    T = len(to)
    hist[:, 1] = 281.0                                  # constant column: every sample in bucket 0
    ref[:, 2] = np.round(ref[:, 2])                     # heavy duplicates
    hist[::7, 3] = np.inf                               # +-inf samples: infinite range, scale 0
    hist[3::11, 3] = -np.inf
    ref[::5, 4] = np.inf
    ref[2::9, 5] = -np.inf
    hist[:, 6] = np.nan; hist[17, 6] = 279.5            # one valid sample in one group
    hist[:, 7] = np.nan                                 # all-NaN point
    ref[:, 8] = np.where(rng.random(T) < 0.5, 0.0, ref[:, 8])   # exact ties at the minimum (zeros) ...
    ref[:, 8] = np.abs(ref[:, 8])
    z = rng.random(T) < 0.3
    hist[:, 9] = np.where(z, -0.0, np.abs(hist[:, 9]))  # ... and signed zeros
    hist[::2, 9] = np.where(hist[::2, 9] == 0, 0.0, hist[::2, 9])
    ref[:, 10] = ref[:, 10] * 1e-30                     # tiny magnitudes (products near the subnormal range)
    hist[:, 11] = hist[:, 11] * 1e30                    # huge magnitudes
    hist[5, 12] = 1e30                                  # one outlier: everything else lands in bucket 0 / 1
    ref[:, 13] = np.where(rng.random(T) < 0.9, np.nan, ref[:, 13])   # mostly NaN
    dry = rng.random(T) < 0.6                           # multi-scale: jittered dry days under a gamma tail
    hist[:, 14] = np.where(dry, rng.uniform(1e-6, 0.01, T), rng.gamma(0.9, 5.0, T)).astype(np.float32)
    ref[:, 14] = np.where(rng.random(T) < 0.55, 0.0, rng.gamma(0.8, 7.5, T)).astype(np.float32)
    return ref, hist


@pytest.mark.parametrize("group,window,years,nq,var,kind", [
    ("time.month", 1, 30, 50, "tas", "+"),
    ("time.month", 1, 30, 100, "pr", "*"),
    ("time.dayofyear", 31, 30, 100, "pr", "*"),
    ("time.dayofyear", 31, 30, 50, "tas", "+"),
    ("time.season", 1, 10, 128, "tas", "+"),
    ("time.month", 1, 2, 50, "tas", "+"),
])
def test_bucket_kernel_equals_sort_kernel(group, window, years, nq, var, kind):
    xs = _xs()
    rng = np.random.default_rng(21)
    to = o.daily_time_axis(1981, years, "noleap"); tx = xs.TimeAxis.daily(1981, years, "noleap")
    P = 45  # a full tile + a ragged one
    ref, hist = _stress_inputs(rng, to, P, var)
    q = o.equally_spaced_nodes(nq).astype(np.float32)
    grp = xs.Grouper(group, window)
    a = _train(xs, ref, hist, tx, grp, q, kind, "bucket")
    b = _train(xs, ref, hist, tx, grp, q, kind, "sort")
    hq_a, hq_b, af_a, af_b = _np(a.hist_q), _np(b.hist_q), _np(a.af), _np(b.af)
    # -0.0 and +0.0 compare equal in every sorter, so which zero comes out of a tie is unpinned (numba's sort
    # included): compare values where both are zero, bits elsewhere
    zero_node = (hq_a == 0) & (hq_b == 0)
    for x, y in ((hq_a, hq_b), (af_a, af_b)):
        both_zero = ((x == 0) & (y == 0)) | zero_node
        assert bits_equal(np.where(both_zero, 0, x), np.where(both_zero, 0, y))
    assert np.array_equal(np.abs(af_a[zero_node]), np.abs(af_b[zero_node]), equal_nan=True)


@pytest.mark.parametrize("group,window,nq,var,kind", [("time.month", 1, 50, "tas", "+"), ("time.month", 1, 50, "pr", "*"),
                                                      ("time.dayofyear", 31, 100, "pr", "*")])
def test_bucket_kernel_matches_oracle_full_segments(group, window, nq, var, kind):
    xs = _xs()
    rng = np.random.default_rng(22)
    to = o.daily_time_axis(1981, 30, "noleap"); tx = xs.TimeAxis.daily(1981, 30, "noleap")
    P = 33
    ref, hist = _stress_inputs(rng, to, P, var)
    q = o.equally_spaced_nodes(nq).astype(np.float32)
    gidx, G, _ = o.group_index(to, group)
    ds = _train(xs, ref, hist, tx, xs.Grouper(group, window), q, kind, "bucket")
    af, hq = _np(ds.af), _np(ds.hist_q)
    sel = np.arange(G) if G <= 12 else np.array([0, 15, 16, 180, 349, 364])
    with np.errstate(all="ignore"):
        for g in sel:
            ref_q = o.nan_quantile(o.group_segment(ref.T.copy(), gidx, g, window), q)
            hist_q = o.nan_quantile(o.group_segment(hist.T.copy(), gidx, g, window), q)
            af_o = o.get_correction(hist_q, ref_q, kind).astype(np.float32)
            z = (hq[:, g] == 0) & (hist_q == 0)
            assert bits_equal(np.where(z, 0, hq[:, g]), np.where(z, 0, hist_q)), g
            # a zero hist_q node makes af = ref_q / (+-0): its sign follows the unpinned sign of the zero
            z = ((af[:, g] == 0) & (af_o == 0)) | (hist_q == 0)
            assert bits_equal(np.where(z, 0, af[:, g]), np.where(z, 0, af_o)), g
            assert np.array_equal(np.abs(af[:, g][hist_q == 0]), np.abs(af_o[hist_q == 0]), equal_nan=True), g


def test_bucket_kernel_dqm_and_jitter_variants():
    """NORM (dqm_train) and JITTER instantiations: bucket and sort kernels agree bit for bit (same hash draws)."""
    xs = _xs()
    rng = np.random.default_rng(23)
    to = o.daily_time_axis(1981, 30, "noleap"); tx = xs.TimeAxis.daily(1981, 30, "noleap")
    ref, hist = (synth.pr(rng, to, 40, w, jitter=False) for w in ("ref", "hist"))
    q = o.equally_spaced_nodes(50).astype(np.float32)
    for fn, kw in ((xs.dqm_train, {}), (xs.eqm_train, {"jitter_under_thresh_value": "0.01 mm/d"}),
                   (xs.dqm_train, {"jitter_under_thresh_value": "0.01 mm/d"})):
        for grp in (xs.Grouper("time.month"), xs.Grouper("time.dayofyear", 31)):
            a = _train(xs, ref, hist, tx, grp, q, "*", "bucket", fn=fn, **kw)
            b = _train(xs, ref, hist, tx, grp, q, "*", "sort", fn=fn, **kw)
            for k in ("af", "hist_q") + (("scaling",) if fn is xs.dqm_train else ()):
                assert bits_equal(_np(a[k]), _np(b[k])), (fn.__name__, kw, grp, k)


def test_group_without_members_and_short_groups():
    """A month that never occurs (S == 0 branch) and a month with a handful of days, through the fast kernels."""
    xs = _xs()
    rng = np.random.default_rng(24)
    full = xs.TimeAxis.daily(1981, 3, "noleap")
    keep = (full.month != 2) & ~((full.month == 3) & (full.day > 2))   # no February, two days of March per year
    tx = full[np.nonzero(keep)[0]]
    to_full = o.daily_time_axis(1981, 3, "noleap")
    ref = (280 + 3 * rng.standard_normal((len(tx), 40))).astype(np.float32)
    hist = (281 + 3 * rng.standard_normal((len(tx), 40))).astype(np.float32)
    q = o.equally_spaced_nodes(50).astype(np.float32)
    for algo in ("bucket", "sort"):
        ds = _train(xs, ref, hist, tx, xs.Grouper("time.month"), q, "+", algo)
        hq, af = _np(ds.hist_q), _np(ds.af)
        assert np.isnan(hq[:, 1]).all() and np.isnan(af[:, 1]).all()
        gidx = (tx.month - 1).astype(np.int32)
        for g in (0, 2, 11):
            seg_h = hist[gidx == g].T.copy(); seg_r = ref[gidx == g].T.copy()
            hq_o = o.nan_quantile(seg_h, q); rq_o = o.nan_quantile(seg_r, q)
            assert bits_equal(hq[:, g], hq_o), (algo, g)
            assert bits_equal(af[:, g], (rq_o - hq_o).astype(np.float32)), (algo, g)
    del to_full


def _train_window(xs, ref, hist, tx, group, q, kind, window_kernel, mode="train"):
    old = os.environ.pop("XSDBA_B200_NO_WINDOW_KERNEL", None)
    if not window_kernel:
        os.environ["XSDBA_B200_NO_WINDOW_KERNEL"] = "1"
    try:
        if mode == "quantile":
            out = xs.group_quantile(hist, time=tx, group=group, quantiles=q)
            torch.cuda.synchronize()
            return out
        ds = xs.eqm_train(xs.Dataset({"ref": ref, "hist": hist}, time=tx), group=group, kind=kind, quantiles=q)
        torch.cuda.synchronize()
        return ds
    finally:
        os.environ.pop("XSDBA_B200_NO_WINDOW_KERNEL", None)
        if old is not None:
            os.environ["XSDBA_B200_NO_WINDOW_KERNEL"] = old


@pytest.mark.parametrize("cal,years,window,nq,var,kind", [
    ("noleap", 30, 31, 100, "pr", "*"),
    ("noleap", 30, 31, 50, "tas", "+"),
    ("standard", 12, 31, 50, "tas", "+"),      # day 366: a sparse group, windows shifted after Feb 29
    ("360_day", 5, 7, 20, "pr", "*"),
    ("noleap", 3, 5, 200, "tas", "+"),          # more nodes than samples in a window
    ("noleap", 30, 15, 30, "pr", "*"),
])
def test_window_kernel_equals_per_group_kernels(cal, years, window, nq, var, kind):
    """K1w (one ordering per chunk of day-of-year groups) against the per-group kernels: every group, bit for bit."""
    xs = _xs()
    rng = np.random.default_rng(77)
    to = o.daily_time_axis(1981, years, cal); tx = xs.TimeAxis.daily(1981, years, cal)
    P = 21  # two full tiles of 8 and a ragged one
    ref, hist = _stress_inputs(rng, to, P, var)
    q = o.equally_spaced_nodes(nq).astype(np.float32)
    grp = xs.Grouper("time.dayofyear", window)
    a = _train_window(xs, ref, hist, tx, grp, q, kind, True)
    b = _train_window(xs, ref, hist, tx, grp, q, kind, False)
    hq_a, hq_b, af_a, af_b = _np(a.hist_q), _np(b.hist_q), _np(a.af), _np(b.af)
    zero_node = (hq_a == 0) & (hq_b == 0)
    for x, y in ((hq_a, hq_b), (af_a, af_b)):
        both_zero = ((x == 0) & (y == 0)) | zero_node
        assert bits_equal(np.where(both_zero, 0, x), np.where(both_zero, 0, y))
    assert np.array_equal(np.abs(af_a[zero_node]), np.abs(af_b[zero_node]), equal_nan=True)
    qa = _np(_train_window(xs, ref, hist, tx, grp, q, kind, True, mode="quantile"))
    z = (qa == 0) & (hq_b == 0)
    assert bits_equal(np.where(z, 0, qa), np.where(z, 0, hq_b))


def test_window_kernel_matches_oracle_sampled_groups():
    xs = _xs()
    rng = np.random.default_rng(78)
    to = o.daily_time_axis(1981, 30, "noleap"); tx = xs.TimeAxis.daily(1981, 30, "noleap")
    ref, hist = (synth.pr(rng, to, 19, w) for w in ("ref", "hist"))
    hist[:400, 3] = np.nan
    q = o.equally_spaced_nodes(100).astype(np.float32)
    gidx, G, _ = o.group_index(to, "time.dayofyear")
    ds = _train_window(xs, ref, hist, tx, xs.Grouper("time.dayofyear", 31), q, "*", True)
    af, hq = _np(ds.af), _np(ds.hist_q)
    with np.errstate(all="ignore"):
        for g in (0, 1, 14, 15, 16, 37, 38, 39, 180, 349, 350, 363, 364):   # series ends, chunk seams, the middle
            ref_q = o.nan_quantile(o.group_segment(ref.T.copy(), gidx, g, 31), q)
            hist_q = o.nan_quantile(o.group_segment(hist.T.copy(), gidx, g, 31), q)
            assert bits_equal(hq[:, g], hist_q), g
            assert bits_equal(af[:, g], o.get_correction(hist_q, ref_q, "*").astype(np.float32)), g


def _qdm_adjust_window(xs, sim, af, q, tx, group, kind, window_kernel):
    old = os.environ.pop("XSDBA_B200_NO_WINDOW_KERNEL", None)
    if not window_kernel:
        os.environ["XSDBA_B200_NO_WINDOW_KERNEL"] = "1"
    try:
        out = xs.qdm_adjust(xs.Dataset({"sim": sim, "af": af, "quantiles": q}, time=tx), group=group, interp="nearest",
                            extrapolation="constant", kind=kind, rank_window=True)
        torch.cuda.synchronize()
        return _np(out.sim_q), _np(out.scen)
    finally:
        os.environ.pop("XSDBA_B200_NO_WINDOW_KERNEL", None)
        if old is not None:
            os.environ["XSDBA_B200_NO_WINDOW_KERNEL"] = old


@pytest.mark.parametrize("cal,years,window,nq,var,kind", [
    ("noleap", 30, 31, 100, "pr", "*"),
    ("noleap", 30, 31, 50, "tas", "+"),
    ("standard", 12, 31, 50, "tas", "+"),      # day 366: a sparse group, windows shifted after Feb 29
    ("360_day", 5, 7, 20, "pr", "*"),
    ("noleap", 3, 5, 200, "tas", "+"),          # more nodes than the kernel's quantile axis holds: per-group fallback
])
def test_rank_window_kernel_equals_per_group_kernel(cal, years, window, nq, var, kind):
    """K3w (rank_window=True from one ordering per chunk of day-of-year groups) against K3 (one sort per group): ranks and
    adjusted values of every time step, bit for bit -- ties, signed zeros, +-inf, NaN samples, NaN factors at the ends
    and in the middle of a factor row, an all-NaN factor row."""
    xs = _xs()
    rng = np.random.default_rng(79)
    to = o.daily_time_axis(1981, years, cal); tx = xs.TimeAxis.daily(1981, years, cal)
    P = 21
    ref, hist = _stress_inputs(rng, to, P, var)
    sim = hist.copy()
    sim[:, 0] = ref[:, 0]
    q = o.equally_spaced_nodes(nq).astype(np.float32)
    grp = xs.Grouper("time.dayofyear", window)
    af = _np(xs.eqm_train(xs.Dataset({"ref": ref, "hist": hist}, time=tx), group=grp, kind=kind, quantiles=q).af).copy()
    af[3, :, 5:9] = np.nan            # NaN factors inside the rows of one point ...
    af[4, :, :3] = np.nan; af[4, :, -2:] = np.nan   # ... at both ends of another ...
    af[5, 10:20, :] = np.nan          # ... and whole rows (the neighbouring groups answer)
    sq_a, sc_a = _qdm_adjust_window(xs, sim, af, q, tx, grp, kind, True)
    sq_b, sc_b = _qdm_adjust_window(xs, sim, af, q, tx, grp, kind, False)
    assert bits_equal(sq_a, sq_b)
    assert bits_equal(sc_a, sc_b)
