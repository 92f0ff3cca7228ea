"""Pin the CPU oracle against (1) outputs of the reference's own kernels (tests/golden, produced by
oracle/gen_golden.py) and (2) the known-answer tests of the reference's test-suite."""
import numpy as np
import pytest

import qm_oracle as o
from conftest import bits_equal


@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("nq", [50, 100])
def test_quantile_bit_exact_vs_reference(golden, tag, nq):
    a, q, ref = (golden[f"quant_{tag}_{nq}_{k}"] for k in ("in", "q", "out"))
    got = o.nan_quantile(a, q)
    if tag == "f32":
        assert bits_equal(got, ref)
    else:  # python fma emulation is faithful, not exact: 1e-12 (north_star f64 tolerance)
        np.testing.assert_allclose(got, ref, rtol=1e-12, atol=0, equal_nan=True)


def test_quantile_long_rows(golden):
    assert bits_equal(o.nan_quantile(golden["quant_long_in"], golden["quant_long_q"]), golden["quant_long_out"])


def test_quantile_edge_cases():
    # reference tests/test_nbutils.py:23-34: one valid value -> that value; all NaN -> NaN
    q = np.linspace(0, 1, 11)
    a = np.full((1, 100), np.nan)
    a[0, 4] = 1.0
    assert (o.nan_quantile(a, q) == 1.0).all()
    assert np.isnan(o.nan_quantile(np.full((1, 100), np.nan), q)).all()


def test_equally_spaced_nodes(golden):
    # reference tests/test_utils.py:58-65
    x = o.equally_spaced_nodes(5, eps=1e-4)
    assert len(x) == 7
    d = np.diff(x)
    np.testing.assert_almost_equal(d[0], d[1] / 2, 3)
    np.testing.assert_almost_equal(o.equally_spaced_nodes(1)[0], 0.5)
    assert (x == golden["nodes_5_eps"]).all()
    assert (o.equally_spaced_nodes(50) == golden["nodes_50"]).all()


@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("method", ["nearest", "linear"])
@pytest.mark.parametrize("extrap", ["constant", "nan"])
def test_interp1d_vs_reference(golden, tag, method, extrap):
    newx = golden[f"i1_{tag}_newx"]
    for sfx in ("", "_n"):
        got = o.interp_on_quantiles_1d(newx, golden[f"i1_{tag}_oldx{sfx}"], golden[f"i1_{tag}_oldy{sfx}"], method, extrap)
        assert bits_equal(got, golden[f"i1_{tag}_{method}_{extrap}{sfx}"])


@pytest.mark.parametrize("method,expi", [("nearest", 2.9), ("linear", 2.95)])
@pytest.mark.parametrize("extrap,expe", [("constant", 4.4), ("nan", np.nan)])
def test_interp1d_reference_kat(golden, method, expi, extrap, expe):
    # reference tests/test_utils.py:68-113
    out = o.interp_on_quantiles_1d(golden["i1_kat_newx"], golden["i1_kat_oldx"], golden["i1_kat_oldy"], method, extrap)
    assert bits_equal(out, golden[f"i1_kat_{method}_{extrap}"])
    if np.isnan(expe):
        assert np.isnan(out[0])
    else:
        assert out[0] == expe
    np.testing.assert_allclose(out[25], expi)
    assert np.isnan(out[-1])


@pytest.mark.parametrize("tag", ["m32", "m64", "wide32", "d32", "qdm"])
@pytest.mark.parametrize("extrap", ["constant", "nan"])
def test_interp2d_vs_reference(golden, tag, extrap):
    got = o.interp_on_quantiles_2d(golden[f"i2_{tag}_newx"], golden[f"i2_{tag}_newg"], golden[f"i2_{tag}_oldx"],
                                   golden[f"i2_{tag}_oldy"], golden[f"i2_{tag}_oldg"], "nearest", extrap)
    assert bits_equal(got, golden[f"i2_{tag}_nearest_{extrap}"])


def test_rank_bn_vs_reference(golden):
    got = o.rank_bn(golden["rankbn_in"])
    assert bits_equal(got, golden["rankbn_out"])


def test_rank_reference_kat():
    # reference tests/test_utils.py:197-205 : ranks == argsort().argsort() + 1 for tie-free data
    arr = np.random.default_rng(0).random((10, 50))
    np.testing.assert_array_equal(o.nanrankdata(arr), arr.argsort().argsort() + 1)
    # tests/test_utils.py:208-217: average ties
    r = o.nanrankdata(np.array([[1, 26, 2, 4.0, 6, 2, 2]]))[0]
    assert sorted(r) == [1.0, 3.0, 3.0, 3.0, 5.0, 6.0, 7.0]


def test_rank_pct_single_sample_group_is_nan():
    # reference tests/test_adjustment.py:877-882 relies on 0/0 = NaN for one-sample groups
    assert np.isnan(o.rank_pct(np.array([[3.0]]))).all()


@pytest.mark.parametrize("k", [0, 1, 2, 3])
def test_loess_vs_reference(golden, k):
    d, f, niter, dx = golden[f"loess_case{k}_params"]
    got = o.loess_nb(golden["loess_x"], golden["loess_y"], f=f, niter=int(niter), d=int(d), dx=dx)
    np.testing.assert_allclose(got, golden[f"loess_case{k}_out"], rtol=1e-11, atol=1e-13, equal_nan=True)


def test_group_index_reference_kat():
    # reference tests/test_base.py:46-65: 31 March -> month 3 ; interp index 3.5 ; doy
    t = o.daily_time_axis(2000, 2, "noleap")
    i = np.nonzero((t.year == 2001) & (t.month == 3) & (t.day == 31))[0][0]
    gi, G, coord = o.group_index(t, "time.month")
    assert gi[i] + 1 == 3 and G == 12 and (coord == np.arange(1, 13)).all()
    assert o.group_index_interp(t, "time.month")[i] == 3.5
    gi, G, _ = o.group_index(t, "time.dayofyear")
    assert gi[i] + 1 == 90 and G == 365
    t360 = o.daily_time_axis(2000, 1, "360_day")
    assert o.group_index(t360, "time.dayofyear")[1] == 360
    tstd = o.daily_time_axis(2000, 1, "standard")
    assert len(tstd) == 366 and o.group_index(tstd, "time.dayofyear")[1] == 366


def test_window_gather_reference_kat():
    # SURVEY.md A.3, hand-checked against reference tests/test_processing.py:259-281:
    # two 4-day "years", Grouper("time.dayofyear", window=3): doy-1 segment of y=8..1
    y = np.arange(8, 0, -1).astype(float)
    gidx = np.array([0, 1, 2, 3, 0, 1, 2, 3])
    seg = o.group_segment(y[None, :], gidx, 0, 3)[0]
    np.testing.assert_array_equal(seg, [np.nan, 8, 7, 5, 4, 3])


def test_npdft_vs_reference(golden):
    """oracle npdft_train / npdft_adjust against the reference's own _npdft_train / _npdft_adjust
    (exec'ed from its source by oracle/gen_golden.py): bit-exact."""
    af = o.npdft_train(golden["npdft_ref"].copy(), golden["npdft_hist"].copy(), golden["npdft_rots"], golden["npdft_q"])
    assert np.array_equal(af, golden["npdft_af_q"], equal_nan=True)
    adj = o.npdft_adjust(golden["npdft_sim_std"].copy(), golden["npdft_af_q"], golden["npdft_rots"], golden["npdft_q"])
    assert bits_equal(adj, golden["npdft_adjusted"])


def test_reordering_reference_kat():
    # reference tests/test_processing.py:248-257: x = 1..8 reordered by the ranks of y = 8..1 -> reversed
    x = np.arange(1, 9).astype(float)
    y = np.arange(8, 0, -1).astype(float)
    np.testing.assert_array_equal(o.reordering_1d(x, y), x[::-1])


def test_vecquantiles_and_map_cdf_vs_reference(golden):
    for tag in ("f32", "f64"):
        got = o.vecquantiles_numba(golden[f"vecq_{tag}_in"], golden[f"vecq_{tag}_rnk"])
        assert bits_equal(got, golden[f"vecq_{tag}_out"])
    got = o.map_cdf_1d(golden["mapcdf_x"], golden["mapcdf_y"], golden["mapcdf_v"])
    assert bits_equal(got, golden["mapcdf_out"])


def test_loess_gaussian_weights_reference_golden():
    """`weights="gaussian"` (loess.py:16-26): the restatement against the reference's numba _loess_nb with
    _gaussian_weighting (tests/golden/loess_gaussian.npz, oracle/gen_golden_loess_gaussian.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loess_gaussian.npz"))
    x, y = g["loess_x"], g["loess_y"]
    for k in range(3):
        d, f, niter, dx = g[f"case{k}_params"]
        got = o.loess_nb(x, y, f=float(f), niter=int(niter), weights="gaussian", d=int(d), dx=float(dx))
        np.testing.assert_allclose(got, g[f"case{k}_out"], rtol=1e-12, atol=1e-13, equal_nan=True)
