"""Golden vectors for LOESS with gaussian weights from the REFERENCE's own numba kernel (loess._loess_nb with
loess._gaussian_weighting, loess.py:16-26, 49-179), same series as reference_kernels.npz.
TEST INFRASTRUCTURE ONLY; runs in the build container only (needs /root/reference, read-only).
Usage:  python oracle/gen_golden_loess_gaussian.py  ->  tests/golden/loess_gaussian.npz"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402


def main():
    _, _, lo = ref_loader.load()
    base = np.load(os.path.join(os.path.dirname(HERE), "tests", "golden", "reference_kernels.npz"))
    y = base["loess_y"]
    # the abscissa as loess_smoothing builds it from a daily time coordinate (loess.py:244-245): (t - t0) / (t_last - t0),
    # a quotient per sample.  (np.linspace of reference_kernels.npz differs from it in the last bit, which the gaussian
    # kernel -- 0.146 just inside the window edge, 0 on it -- turns into 1e-3 differences.)
    t = np.arange(y.size, dtype=np.float64)
    x = (t - t[0]) / (t[-1] - t[0])
    g = {"loess_x": x, "loess_y": y}
    dx = float(x[1] - x[0])
    for k, (d, f, niter) in enumerate(((0, 0.2, 1), (1, 0.3, 1), (0, 0.5, 2))):
        rf = {0: lo._constant_regression, 1: lo._linear_regression}[d]
        g[f"case{k}_params"] = np.array([d, f, niter, dx])
        g[f"case{k}_out"] = lo._loess_nb(x, y.copy(), f=f, niter=niter, weight_func=lo._gaussian_weighting, reg_func=rf,
                                         dx=dx, skipna=True)
    out = os.path.join(os.path.dirname(HERE), "tests", "golden", "loess_gaussian.npz")
    np.savez_compressed(out, **g)
    print(out, {k: v.shape for k, v in g.items()})


if __name__ == "__main__":
    main()
