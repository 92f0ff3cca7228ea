"""CPU-side checks (no GPU compute): the C ABI library builds for sm_100a, loads, and exports every
symbol include/xsdba_b200.h declares; host logic (Grouper / TimeAxis / argument validation) mirrors
the reference; the product never routes through the oracle."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from xsdba_b200 import build, _lib
    build.build()
    return _lib.load()


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "xsdba_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(xsdba_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 18
    raw = ctypes.CDLL(os.path.join(ROOT, "xsdba_b200", "libxsdba_b200.so"))
    for n in sorted(names):
        assert hasattr(raw, n), f"{n} declared in include/xsdba_b200.h but not exported"
    from xsdba_b200 import _lib
    assert names == set(_lib.SIGNATURES), names ^ set(_lib.SIGNATURES)


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", os.path.join(ROOT, "xsdba_b200", "libxsdba_b200.so")],
                         capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_headline_train_kernel_is_not_spilling():
    """train_bucket_kernel<false,false> sits at the 64-register cap of a 1024-thread CTA: a code-generation accident
    (nvcc --split-compile, an edit that lengthens a live range) costs 20 %.  The good build keeps
    its stack frame at 80 bytes (the out-of-line sorter's call frame); spilling builds showed 144-330."""
    import shutil
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-res-usage", os.path.join(ROOT, "xsdba_b200", "libxsdba_b200.so")],
                         capture_output=True, text=True).stdout
    m = re.search(r"train_bucket_kernelILb0ELb0[^\n]*\n\s*REG:(\d+) STACK:(\d+)", out)
    assert m, "train_bucket_kernel<false,false> not found in the library"
    assert int(m.group(1)) <= 64 and int(m.group(2)) <= 96, m.group(0)


def test_version_and_status_strings(lib):
    assert lib.xsdba_version() >= 100
    assert b"invalid" in lib.xsdba_status_string(-1)
    assert b"segment" in lib.xsdba_status_string(-3)


def test_no_device_is_a_loud_error_not_a_fallback(lib):
    """Without a GPU the grouping handle cannot be created: the product raises, it never computes on CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import xsdba_b200 as xs
    t = xs.TimeAxis.daily(2000, 2, "noleap")
    x = np.zeros((len(t), 4), np.float32)
    with pytest.raises((ValueError, RuntimeError)):
        xs.eqm_train(xs.Dataset({"ref": x, "hist": x}, time=t), group="time.month", kind="+",
                     quantiles=xs.equally_spaced_nodes(10))


def test_product_does_not_import_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "xsdba_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inc")):
                src = open(os.path.join(dirpath, f)).read()
                assert "qm_oracle" not in src and "oracle/" not in src and "import oracle" not in src, f


def test_grouper_mirrors_reference_semantics():
    import qm_oracle as o
    import xsdba_b200 as xs
    for cal in ("noleap", "standard", "360_day", "all_leap"):
        tx = xs.TimeAxis.daily(1999, 3, cal)
        to = o.daily_time_axis(1999, 3, cal)
        assert len(tx) == len(to)
        for group in ("time", "time.month", "time.dayofyear", "time.season"):
            g = xs.Grouper(group)
            gi, G, coord = o.group_index(to, group)
            np.testing.assert_array_equal(g.zero_based_index(tx), gi)
            assert g.n_groups(tx) == G
        np.testing.assert_array_equal(xs.Grouper("time.month").get_index(tx, interp=True), o.group_index_interp(to, "time.month"))
    # reference tests/test_base.py:46-65
    t = xs.TimeAxis.daily(2000, 2, "noleap")
    i = np.nonzero((t.year == 2001) & (t.month == 3) & (t.day == 31))[0][0]
    assert xs.Grouper("time.month").get_index(t)[i] == 3
    assert xs.Grouper("time.month").get_index(t, interp=True)[i] == 3.5
    assert xs.Grouper("time.dayofyear").get_index(t)[i] == 90
    with pytest.raises(ValueError):  # base.py:151-156
        xs.Grouper("time", window=5)
    g = xs.Grouper("time.month", add_dims=["realization"])   # pooled by xr_adapter; the array-level API has no dim names
    assert g.add_dims == ["realization"]
    with pytest.raises(NotImplementedError):
        g.handle(t)


def test_time_axis_constructors_agree():
    import xsdba_b200 as xs
    a = xs.TimeAxis.daily(1999, 3, "standard")
    d = np.arange("1999-01-01", "2002-01-01", dtype="datetime64[D]")
    b = xs.TimeAxis.from_datetime64(d)
    for f in ("year", "month", "day", "dayofyear", "days_in_month"):
        np.testing.assert_array_equal(getattr(a, f), getattr(b, f))
    c = xs.TimeAxis.from_fields(a.year, a.month, a.day, "standard")
    np.testing.assert_array_equal(c.dayofyear, a.dayofyear)


def test_equally_spaced_nodes_matches_reference(golden):
    import xsdba_b200 as xs
    assert (xs.equally_spaced_nodes(50) == golden["nodes_50"]).all()
    assert (xs.equally_spaced_nodes(5, eps=1e-4) == golden["nodes_5_eps"]).all()


# ---------------------------------------------------------------------------------------------
# stack_periods / unstack_periods index logic (base.py:1072-1381, freq="YS") -- host only
# ---------------------------------------------------------------------------------------------
def test_stack_periods_non_overlapping_and_overlapping():
    from xsdba_b200.calendar import TimeAxis
    from xsdba_b200.periods import stack_periods, unstack_periods
    t = TimeAxis.daily(1950, 150, "noleap")
    p = stack_periods(t, window=30)
    assert len(p) == 5 and p.start_years == (1950, 1980, 2010, 2040, 2070)
    assert all(n == 30 * 365 for n in p.lengths) and p.slices[-1].stop == len(t)
    x = np.arange(len(t), dtype=np.float64)[:, None] * np.ones((1, 3))
    y, cov = unstack_periods([x[s] for s in p.slices], p, t)
    assert cov == slice(0, len(t)) and np.array_equal(y, x)
    p = stack_periods(t, window=30, stride=10)
    assert p.start_years == tuple(range(1950, 2071, 10))          # the last complete window starts in 2070
    y, cov = unstack_periods([x[s] for s in p.slices], p, t)
    assert cov == slice(0, len(t)) and np.array_equal(y, x)        # every day exactly once, in order
    # time-last layout
    y2, _ = unstack_periods([x.T[:, s] for s in p.slices], p, t, time_axis=-1)
    assert np.array_equal(y2, x.T)


def test_stack_periods_docstring_table_and_edges():
    """The example of unstack_periods' docstring (base.py:1293-1307): stride = window / 5, min_length = 4 strides,
    7 strides of data -> 4 periods, the last one shorter; kept strides 0-2 | 3 | 4 | 5-6."""
    from xsdba_b200.calendar import TimeAxis
    from xsdba_b200.periods import stack_periods, unstack_periods
    t = TimeAxis.daily(2000, 7, "noleap")
    p = stack_periods(t, window=5, stride=1, min_length=4)
    assert p.start_years == (2000, 2001, 2002, 2003)
    assert p.lengths == (5 * 365, 5 * 365, 5 * 365, 4 * 365)
    marks = [np.full((n, 1), i) for i, n in enumerate(p.lengths)]
    y, cov = unstack_periods(marks, p, t)
    assert cov == slice(0, 7 * 365)
    assert np.array_equal(y[:, 0], np.repeat([0, 0, 0, 1, 2, 3, 3], 365))
    # incomplete last window is dropped by default; a series ending before 31 December does not close its year
    assert len(stack_periods(t, window=5, stride=1)) == 3
    assert len(stack_periods(t[:-1], window=7)) == 0 and len(stack_periods(t, window=7)) == 1
    # first window must start in January (base.py:1200-1208)
    assert stack_periods(t[40:], window=3, stride=1).start_years[0] == 2001
    with pytest.raises(ValueError):
        stack_periods(TimeAxis.daily(2000, 7, "standard"), window=5)
    with pytest.raises(ValueError):
        stack_periods(t, window=2, stride=3)
    with pytest.raises(NotImplementedError):
        pp = stack_periods(t, window=4, stride=2)
        unstack_periods([np.zeros((n, 1)) for n in pp.lengths], pp, t)
