"""Synthetic inputs for the parity tests (numpy, host): the generators of SURVEY.md section 8d at
small sizes.  Series are returned time-major: (time, points)."""
import numpy as np


def tas(rng, time, n_pts, which, dtype=np.float32, nan_frac=0.001):
    A, sigma, k, off = {"ref": (12, 3.0, 1.0, 0.0), "hist": (10, 3.5, 1.0, 1.5), "sim": (10, 3.5, 1.1, 3.5)}[which]
    doy = time.dayofyear[:, None]
    yr = (time.year - time.year[0])[:, None]
    lat = np.linspace(-1.2, 1.2, n_pts)[None, :]
    x = 273.15 + 15 * np.cos(lat) - A * np.cos(2 * np.pi * (doy - 15) / 365) + 0.03 * yr * k + off \
        + sigma * rng.standard_normal((len(time), n_pts))
    x = x.astype(dtype)
    if nan_frac:
        x[rng.random(x.shape) < nan_frac] = np.nan
    return x


def pr(rng, time, n_pts, which, dtype=np.float32, nan_frac=0.001, jitter=True):
    p_wet, shape, scale = {"ref": (0.45, 0.8, 7.5), "hist": (0.60, 0.9, 5.0), "sim": (0.60, 0.9, 5.5)}[which]
    wet = rng.random((len(time), n_pts)) < p_wet
    x = np.where(wet, rng.gamma(shape, scale, size=wet.shape), 0.0)
    if jitter:  # deterministic stand-in for jitter_under_thresh("0.01 mm/d") (processing.py:124-148)
        dry = x < 0.01
        x = np.where(dry, rng.uniform(1e-5, 0.01, size=x.shape), x)
    x = x.astype(dtype)
    if nan_frac:
        x[rng.random(x.shape) < nan_frac] = np.nan
    return x
