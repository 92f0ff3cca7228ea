"""Load the reference's own kernel files (nbutils.py, utils.py, loess.py) under stub modules.

TEST INFRASTRUCTURE ONLY.  This module is used by ``oracle/gen_golden.py`` *in the build
container* (where ``/root/reference`` exists) to produce the committed fixtures under
``tests/golden/``.  Nothing in the product, the GPU tests, ``smoke()`` or ``bench.py``
imports it, and it does nothing useful on the GPU box (no ``/root/reference`` there).

The reference cannot be imported as a package here (xarray, dask, bottleneck, cftime, pint,
jsonpickle, boltons are absent), but its numerical kernels only need
numpy / numba / scipy, which are present.  We therefore register minimal stand-ins for the
missing imports and ``exec`` the reference's files from where they lie.  The arithmetic that
runs is the reference's own (numba-compiled ``_nan_quantile_1d`` etc.), nothing is copied.
"""
from __future__ import annotations

import functools
import importlib.util
import os
import sys
import types

REF_SRC = "/root/reference/src/xsdba"


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def available() -> bool:
    return os.path.isdir(REF_SRC)


@functools.lru_cache(maxsize=1)
def load():
    """Return (nbutils, utils, loess) modules of the reference, numba-compiled on first use."""
    if not available():
        raise RuntimeError("reference sources not present (expected only in the build container)")
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/xsdba_b200_numba_cache")
    os.makedirs(os.environ["NUMBA_CACHE_DIR"], exist_ok=True)

    import numpy as np
    from scipy.stats import rankdata

    class _DataArray:  # never instantiated by the kernels we call
        pass

    class _Dataset:
        pass

    def _apply_ufunc(*a, **k):
        raise NotImplementedError("xarray.apply_ufunc stub: call the raw kernel instead")

    def _nanrankdata(arr, axis=None):
        # bottleneck.nanrankdata stand-in: average ties, NaN -> NaN, ALWAYS float64 (bottleneck's
        # documented return dtype; scipy keeps float32 for float32 input with NaNs)
        return np.asarray(rankdata(arr, method="average", axis=axis, nan_policy="omit"), dtype=np.float64)

    xr = _mod("xarray", DataArray=_DataArray, Dataset=_Dataset, apply_ufunc=_apply_ufunc)
    core = _mod("xarray.core")
    cu = _mod("xarray.core.utils", get_temp_dimname=lambda dims, name: name)
    xr.core = core
    core.utils = cu
    _mod("bottleneck", nanrankdata=_nanrankdata)
    bolt = _mod("boltons")
    bolt.funcutils = _mod("boltons.funcutils", wraps=functools.wraps)
    dask = _mod("dask")
    dask.array = _mod("dask.array", Array=type("Array", (), {}))

    pkg = _mod("xsdba")
    pkg.__path__ = [REF_SRC]

    class _Grouper:
        pass

    _mod(
        "xsdba.base",
        Grouper=_Grouper,
        _interpolate_doy_calendar=lambda *a, **k: None,
        ensure_chunk_size=lambda da, **k: da,
        parse_group=lambda f: f,
    )

    out = []
    for name in ("nbutils", "utils", "loess"):
        full = f"xsdba.{name}"
        spec = importlib.util.spec_from_file_location(full, os.path.join(REF_SRC, f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[full] = mod
        spec.loader.exec_module(mod)
        setattr(pkg, name, mod)
        out.append(mod)
    return tuple(out)
