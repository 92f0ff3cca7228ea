"""Generate the committed golden fixtures under tests/golden/ from the REFERENCE's own kernels.

TEST INFRASTRUCTURE ONLY; runs in the build container only (needs /root/reference, read-only).
Usage:  python oracle/gen_golden.py

What runs is the reference's code: numba-compiled ``nbutils._quantile`` / ``_extrapolate_on_quantiles``
/ ``loess._loess_nb`` and the SciPy-backed ``utils._interp_on_quantiles_1D`` /
``utils._interp_on_quantiles_2d`` / ``utils.equally_spaced_nodes`` / ``utils._rank_bn``, loaded by
``oracle/ref_loader.py``.  Inputs are seeded; every array needed to replay a case is stored with its
output so the tests never need the reference again.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    nbu, u, lo = ref_loader.load()
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20260117)
    g = {}

    # ---- quantiles (nbutils.py:108-148, 198-221) ------------------------------------------------
    for dt, tag in ((np.float32, "f32"), (np.float64, "f64")):
        for nq in (50, 100):
            rows, S = 24, 930
            a = (rng.gamma(0.8, 5.0, size=(rows, S)) + rng.normal(size=(rows, S)) * 0.01).astype(dt)
            a[rng.random(a.shape) < 0.03] = np.nan
            a[0, :] = np.nan                 # all-NaN row            (tests/test_nbutils.py:29-34)
            a[1, 1:] = np.nan                # one valid value        (tests/test_nbutils.py:23-27)
            a[2, 2:] = np.nan                # two valid values
            a[3, :] = np.round(a[3, :])      # many ties
            a[4, 400:] = np.nan              # half NaN
            a[5, :] = -a[5, :]               # negative values
            q = u.equally_spaced_nodes(nq).astype(dt)
            g[f"quant_{tag}_{nq}_in"] = a
            g[f"quant_{tag}_{nq}_q"] = q
            g[f"quant_{tag}_{nq}_out"] = nbu._quantile(a.copy(), q, 1)
    # group="time"-sized rows
    a = (rng.normal(size=(3, 10950)) * 3.5 + 280).astype(np.float32)
    a[1, rng.random(10950) < 0.001] = np.nan
    q = u.equally_spaced_nodes(50).astype(np.float32)
    g["quant_long_in"], g["quant_long_q"], g["quant_long_out"] = a, q, nbu._quantile(a.copy(), q, 1)

    g["nodes_5_eps"] = u.equally_spaced_nodes(5, eps=1e-4)
    g["nodes_50"] = u.equally_spaced_nodes(50)

    # ---- 1-D factor lookup (utils.py:350-377) ----------------------------------------------------
    for dt, tag in ((np.float32, "f32"), (np.float64, "f64")):
        nq = 50
        oldx = np.sort(rng.normal(size=nq) * 3 + 280).astype(dt)
        oldy = (rng.normal(size=nq)).astype(dt)
        oldx_n, oldy_n = oldx.copy(), oldy.copy()
        oldx_n[[0, 17]] = np.nan
        oldy_n[[-1, 30]] = np.nan
        newx = (rng.normal(size=400) * 4 + 280).astype(dt)
        newx[[3, 50]] = np.nan
        newx[10] = oldx[5]                       # exactly on a node
        newx[11] = oldx[0]; newx[12] = oldx[-1]  # exactly on the bounds
        g[f"i1_{tag}_oldx"], g[f"i1_{tag}_oldy"], g[f"i1_{tag}_newx"] = oldx, oldy, newx
        g[f"i1_{tag}_oldx_n"], g[f"i1_{tag}_oldy_n"] = oldx_n, oldy_n
        for method in ("nearest", "linear"):
            for extrap in ("constant", "nan"):
                g[f"i1_{tag}_{method}_{extrap}"] = u._interp_on_quantiles_1D(newx, oldx, oldy, method, extrap)
                g[f"i1_{tag}_{method}_{extrap}_n"] = u._interp_on_quantiles_1D(newx, oldx_n, oldy_n, method, extrap)
    # the reference's own KAT (tests/test_utils.py:68-113): nearest 2.9 / linear 2.95 / constant 4.4
    oldx = np.linspace(205, 229, num=25); oldy = np.linspace(2, 4.4, num=25)
    newx = np.linspace(240, 200, num=41) - 0.5
    newx = np.where(newx > 201, newx, np.nan)
    g["i1_kat_newx"], g["i1_kat_oldx"], g["i1_kat_oldy"] = newx, oldx, oldy
    for method in ("nearest", "linear"):
        for extrap in ("constant", "nan"):
            g[f"i1_kat_{method}_{extrap}"] = u._interp_on_quantiles_1D(newx, oldx, oldy, method, extrap)

    # ---- 2-D (grouped) factor lookup (utils.py:380-400; nbutils.py:375-416) ----------------------
    def table(G, nq, dt, spread):
        base = np.sort(rng.normal(size=(G, nq)) * spread, axis=1) + 10 * np.cos(np.arange(G) / G * 2 * np.pi)[:, None]
        af = rng.normal(size=(G, nq))
        return base.astype(dt), af.astype(dt)

    for tag, G, nq, dt, spread in (("m32", 12, 50, np.float32, 3.0), ("m64", 12, 50, np.float64, 3.0),
                                   ("wide32", 12, 20, np.float32, 25.0), ("d32", 365, 100, np.float32, 3.0)):
        xq, yq = table(G, nq, dt, spread)
        if tag == "m32":
            xq[3, :4] = np.nan; yq[3, :4] = np.nan     # NaN nodes at the start of one row
            yq[7, -2:] = np.nan                         # NaN factors at the end of another
        oldx = np.concatenate([xq[-1:], xq, xq[:1]])   # add_cyclic_bounds (utils.py:305-313)
        oldy = np.concatenate([yq[-1:], yq, yq[:1]])
        oldg = np.broadcast_to(np.arange(0, G + 2, dtype=np.float64)[:, None], oldx.shape).copy()
        T = 1500
        newg = rng.integers(1, G + 1, size=T)
        newx = (rng.normal(size=T) * spread * 1.3 + 10 * np.cos((newg - 1) / G * 2 * np.pi)).astype(dt)
        newx[[5, 99]] = np.nan
        g[f"i2_{tag}_oldx"], g[f"i2_{tag}_oldy"], g[f"i2_{tag}_oldg"] = oldx, oldy, oldg
        g[f"i2_{tag}_newx"], g[f"i2_{tag}_newg"] = newx, newg
        for extrap in ("constant", "nan"):
            g[f"i2_{tag}_nearest_{extrap}"] = u._interp_on_quantiles_2d(newx, newg, oldx, oldy, oldg, "nearest", extrap)
    # QDM flavour: x-axis is the quantile coordinate, identical rows (_adjustment.py:873-880)
    G, nq = 12, 50
    qv = u.equally_spaced_nodes(nq).astype(np.float32)
    oldx = np.broadcast_to(qv, (G + 2, nq)).copy()
    _, yq = table(G, nq, np.float32, 1.0)
    oldy = np.concatenate([yq[-1:], yq, yq[:1]])
    oldg = np.broadcast_to(np.arange(0, G + 2, dtype=np.float64)[:, None], oldx.shape).copy()
    newg = rng.integers(1, G + 1, size=800)
    newx = rng.random(800); newx[:3] = [0.0, 1.0, np.nan]
    g["i2_qdm_oldx"], g["i2_qdm_oldy"], g["i2_qdm_oldg"], g["i2_qdm_newx"], g["i2_qdm_newg"] = oldx, oldy, oldg, newx, newg
    for extrap in ("constant", "nan"):
        g[f"i2_qdm_nearest_{extrap}"] = u._interp_on_quantiles_2d(newx, newg, oldx, oldy, oldg, "nearest", extrap)

    # ---- ranks (utils.py:641-646; bottleneck stand-in = scipy rankdata) --------------------------
    a = rng.normal(size=(6, 40)); a[1, 3:7] = np.nan; a[2] = np.round(a[2]); a[3, 1:] = np.nan
    g["rankbn_in"], g["rankbn_out"] = a, u._rank_bn(a, axis=-1)

    # ---- LOESS (loess.py:49-179) -----------------------------------------------------------------
    n = 240
    x = np.linspace(0, 1, n)
    y = np.sin(5 * x) + rng.normal(size=n) * 0.3
    y[[7, 120, 121]] = np.nan
    g["loess_x"], g["loess_y"] = x, y
    k = 0
    for d, f, niter, eq in ((0, 0.2, 1, True), (1, 0.5, 2, True), (0, 0.3, 3, False), (1, 0.2, 1, False)):
        rf = {0: lo._constant_regression, 1: lo._linear_regression}[d]
        dx = float(x[1] - x[0]) if eq else 0.0
        g[f"loess_case{k}_params"] = np.array([d, f, niter, dx])
        g[f"loess_case{k}_out"] = lo._loess_nb(x, y.copy(), f=f, niter=niter, weight_func=lo._tricube_weighting,
                                               reg_func=rf, dx=dx, skipna=True)
        k += 1

    # ---- N-pdf transform (_adjustment.py:289-328, 426-464): the two functions are pure numpy + nbutils +
    # utils, so they are exec'ed from the reference's source text with the stub-loaded modules ------------
    import ast
    src = open(os.path.join(ref_loader.REF_SRC, "_adjustment.py")).read()
    ns = {"np": np, "nbu": nbu, "u": u}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name in ("_npdft_train", "_npdft_adjust"):
            exec(compile(ast.Module([node], []), "_adjustment.py", "exec"), ns)
    V, Tn, n_iter = 3, 500, 4
    rots = np.stack([np.linalg.qr(rng.standard_normal((V, V)))[0] for _ in range(n_iter)]).astype(np.float32)
    mk = lambda off: (rng.standard_normal((V, Tn)) * np.array([[3.0], [1.0], [7.0]]) + off).astype(np.float32)
    ref, hist, sim = mk(0.0), mk(1.0), mk(1.5)
    hist[1, 7] = np.nan
    qn = u.equally_spaced_nodes(20)
    af_q, _ = ns["_npdft_train"](ref.copy(), hist.copy(), rots, qn, "nearest", "constant", -1, True)
    sim_std = ((sim - sim.mean(-1, keepdims=True)) / sim.std(-1, keepdims=True)).astype(np.float32)
    adj = ns["_npdft_adjust"](sim_std.copy(), af_q, rots, qn, "nearest", "constant")
    g["npdft_ref"], g["npdft_hist"], g["npdft_sim_std"], g["npdft_rots"], g["npdft_q"] = ref, hist, sim_std, rots, qn
    g["npdft_af_q"], g["npdft_adjusted"] = af_q, adj

    # ---- vecquantiles (numba guvectorize, nbutils.py:151-161) and map_cdf_1d (utils.py:35-44) ----------------
    for dt, tag in ((np.float32, "f32"), (np.float64, "f64")):
        a = rng.gamma(2.0, 3.0, size=(40, 300)).astype(dt)
        a[rng.random(a.shape) < 0.05] = np.nan
        a[0, 1:] = np.nan
        rk = rng.random(40).astype(dt); rk[3] = np.nan; rk[4] = 0.0; rk[5] = 1.0
        g[f"vecq_{tag}_in"], g[f"vecq_{tag}_rnk"] = a, rk
        g[f"vecq_{tag}_out"] = nbu._vecquantiles(a, rk)
    xx = rng.normal(size=400).astype(np.float32); yy = (rng.normal(size=400) * 2 + 1).astype(np.float32)
    xx[7] = np.nan; yy[11] = np.nan
    yv = np.array([-1.0, 0.3, 2.5])
    g["mapcdf_x"], g["mapcdf_y"], g["mapcdf_v"], g["mapcdf_out"] = xx, yy, yv, u.map_cdf_1d(xx, yy, yv)

    np.savez_compressed(os.path.join(OUT, "reference_kernels.npz"), **g)
    sz = os.path.getsize(os.path.join(OUT, "reference_kernels.npz"))
    print(f"wrote {len(g)} arrays, {sz/1024:.0f} KiB")


if __name__ == "__main__":
    main()
