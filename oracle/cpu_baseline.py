"""CPU baseline / reference arm for bench.py -- TEST INFRASTRUCTURE, NOT PRODUCT.

Times the oracle (``qm_oracle``: numpy sort-based quantiles per group + SciPy griddata / interp1d
factor lookup per gridpoint, i.e. the reference's own algorithm and third-party arithmetic) on a
bounded sample of the bench workload, on the host cores.  ``kind`` is "port": /root/reference cannot
be imported as a package in this image (xarray / dask absent) and does not exist on the GPU box.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np

import qm_oracle as o


def _synth(seed, time_axis, n_pts, which):
    rng = np.random.default_rng(seed)
    A, sigma, k, off = {"ref": (12, 3.0, 1.0, 0.0), "hist": (10, 3.5, 1.0, 1.5), "sim": (10, 3.5, 1.1, 3.5)}[which]
    doy = time_axis.dayofyear[None, :]
    yr = (time_axis.year - time_axis.year[0])[None, :]
    lat = np.linspace(-1.2, 1.2, n_pts)[:, None]
    x = 273.15 + 15 * np.cos(lat) - A * np.cos(2 * np.pi * (doy - 15) / 365) + 0.03 * yr * k + off \
        + sigma * rng.standard_normal((n_pts, len(time_axis)))
    return x.astype(np.float32)  # point-major [n_pts, T]


def _work(args):
    seed, n_pts, n_years, nq, group = args
    t_tr = o.daily_time_axis(1981, n_years, "noleap")
    t_sim = o.daily_time_axis(2041, n_years, "noleap")
    ref = _synth(seed, t_tr, n_pts, "ref")
    hist = _synth(seed + 1, t_tr, n_pts, "hist")
    sim = _synth(seed + 2, t_sim, n_pts, "sim")
    q = o.equally_spaced_nodes(nq).astype(np.float32)
    gidx, G, _ = o.group_index(t_tr, group)
    t0 = time.perf_counter()
    af, hq = o.eqm_train(ref, hist, gidx, G, 1, q, "+")
    scen = o.qm_adjust(sim, af, hq, group=group, time=t_sim, interp="nearest", extrapolation="constant", kind="+")
    dt = time.perf_counter() - t0
    return dt, float(np.nansum(scen))


def run(n_pts_total: int, cores: int | None = None, n_years: int = 30, nq: int = 50, group: str = "time.month",
        seed: int = 1234):
    """EQM train+adjust of ``n_pts_total`` gridpoints x ``n_years`` daily years on ``cores`` processes.
    Returns dict(value = gridpoint-days/s (wall), seconds, cores, sample)."""
    cores = cores or os.cpu_count() or 1
    cores = max(1, min(cores, n_pts_total))
    per = [n_pts_total // cores + (1 if i < n_pts_total % cores else 0) for i in range(cores)]
    jobs = [(seed + 10 * i, n, n_years, nq, group) for i, n in enumerate(per) if n > 0]
    T = len(o.daily_time_axis(2041, n_years, "noleap"))
    if cores == 1:
        _work((0, 1, 1, nq, group))  # warm-up: imports SciPy outside the timed region
        t0 = time.perf_counter()
        res = [_work(jobs[0])]
        wall = time.perf_counter() - t0
    else:
        ctx = mp.get_context("fork")
        with ctx.Pool(cores) as pool:
            pool.map(_work, [(0, 1, 1, nq, group)] * cores)  # warm the workers (imports, scipy)
            t0 = time.perf_counter()
            res = pool.map(_work, jobs)
            wall = time.perf_counter() - t0
    return {
        "value": n_pts_total * T / wall,
        "unit": "gridpoint*days/s",
        "seconds": wall,
        "cores": cores,
        "kind": "port",
        "sample": f"{n_pts_total} gridpoints x {T} days, EQM nq={nq} group={group} f32 (numpy/SciPy oracle, "
                  f"synthetic data generated outside the timed region)",
    }


def _synth_pr(seed, time_axis, n_pts, which):
    rng = np.random.default_rng(seed)
    p_wet, shape, scale = {"ref": (0.45, 0.8, 7.5), "hist": (0.60, 0.9, 5.0), "sim": (0.60, 0.9, 5.5)}[which]
    T = len(time_axis)
    x = np.where(rng.random((n_pts, T)) < p_wet, rng.gamma(shape, scale, size=(n_pts, T)), 0.0)
    x = np.where(x < 0.01, rng.uniform(1e-5, 0.01, size=x.shape), x)   # jittered once, before timing
    return x.astype(np.float32)


def _work_cfg3(args):
    seed, n_pts, n_years, nq = args
    t_tr = o.daily_time_axis(1981, n_years, "noleap")
    t_sim = o.daily_time_axis(2041, n_years, "noleap")
    ref, hist, sim = (_synth_pr(seed + i, t, n_pts, w) for i, (t, w) in enumerate(((t_tr, "ref"), (t_tr, "hist"), (t_sim, "sim"))))
    q = o.equally_spaced_nodes(nq).astype(np.float32)
    gidx, G, _ = o.group_index(t_tr, "time.dayofyear")
    t0 = time.perf_counter()
    af, _hq = o.eqm_train(ref, hist, gidx, G, 31, q, "*")
    scen, _ = o.qdm_adjust(sim, af, q, group="time.dayofyear", time=t_sim, window=31, interp="nearest",
                           extrapolation="constant", kind="*", rank_window=False)
    return time.perf_counter() - t0, float(np.nansum(scen))


def run_cfg3(n_pts_total: int, cores: int | None = None, n_years: int = 30, nq: int = 100, seed: int = 4321):
    """BASELINE config 3 (QDM kind='*', Grouper('time.dayofyear', 31), nq=100) on ``cores`` processes."""
    cores = max(1, min(cores or os.cpu_count() or 1, n_pts_total))
    per = [n_pts_total // cores + (1 if i < n_pts_total % cores else 0) for i in range(cores)]
    jobs = [(seed + 10 * i, n, n_years, nq) for i, n in enumerate(per) if n > 0]
    T = len(o.daily_time_axis(2041, n_years, "noleap"))
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_work_cfg3, [(0, 1, 1, nq)] * cores)
        t0 = time.perf_counter()
        pool.map(_work_cfg3, jobs)
        wall = time.perf_counter() - t0
    return {"value": n_pts_total * T / wall, "unit": "gridpoint*days/s", "seconds": wall, "cores": cores, "kind": "port",
            "sample": f"{n_pts_total} gridpoints x {T} days, QDM nq={nq} Grouper(time.dayofyear, 31) kind=* f32 "
                      "(numpy/SciPy oracle)"}


if __name__ == "__main__":
    import sys
    print(run(int(sys.argv[1]) if len(sys.argv) > 1 else 64, int(sys.argv[2]) if len(sys.argv) > 2 else None))
