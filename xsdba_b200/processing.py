"""Host mirror of the pre-processing helpers the QM classes use (xsdba.processing)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def _quantity(v):
    """Threshold kwargs reach the reference as unit strings (e.g. "0.01 mm/d") and are converted to the
    data's units with pint.  Units handling is out of scope here: a number, or the number
    in front of a unit string that is assumed to be the data's unit already."""
    if v is None:
        return None
    if isinstance(v, str):
        return float(v.split()[0])
    return float(v)


def jitter_params(dtype, lower=None, upper=None, minimum=None, maximum=None) -> np.ndarray:
    """{lower, minimum, upper, maximum} as the C ABI wants them, with the reference's next-after rules
    (processing.py:222-224, 243-247)."""
    npdt = np.float32 if dtype in (torch.float32, np.float32) else np.float64
    lower, upper, minimum, maximum = map(_quantity, (lower, upper, minimum, maximum))
    out = np.array([np.nan, 0.0, np.nan, 0.0], np.float64)
    if lower is not None:
        out[0] = lower
        out[1] = float(np.nextafter(npdt(minimum if minimum is not None else 0), npdt(np.inf), dtype=npdt))
    if upper is not None:
        if maximum is None:
            raise ValueError("If 'upper' is given, so must 'maximum'.")  # processing.py:238-239
        out[2] = upper
        out[3] = float(np.nextafter(npdt(maximum), npdt(-np.inf), dtype=npdt)) if npdt == np.float32 else maximum
    return out


def jitter(x, lower=None, upper=None, minimum=None, maximum=None, seed: int = 0):
    """``xsdba.processing.jitter`` (processing.py:180-257) on a CUDA tensor / numpy array -> CUDA tensor."""
    from ._adjustment import _as_device, _stream
    lib = _lib.load()
    xt = _as_device(x).contiguous()
    if xt.dtype not in (torch.float32, torch.float64):
        xt = xt.to(torch.float32)
    j4 = jitter_params(xt.dtype, lower, upper, minimum, maximum)
    out = torch.empty_like(xt)
    fn = getattr(lib, "xsdba_jitter_f32" if xt.dtype == torch.float32 else "xsdba_jitter_f64")
    _lib.check(fn(xt.data_ptr(), xt.numel(), j4.ctypes.data_as(_lib.c_f64p), C.c_uint64(seed), out.data_ptr(), _stream()),
               "jitter")
    return out


def jitter_under_thresh(x, thresh, seed: int = 0):
    """processing.py:124-148."""
    return jitter(x, lower=thresh, seed=seed)


def jitter_over_thresh(x, thresh, upper_bnd, seed: int = 0):
    """processing.py:151-177."""
    return jitter(x, upper=thresh, maximum=upper_bnd, seed=seed)


def escore(tgt, sim, N: int = 0, scale: bool = False):
    """``xsdba.processing.escore`` (processing.py:393-489): energy score between the multivariate clouds ``tgt`` and
    ``sim``, arrays (variable, time, *points) -> (*points,).  ``scale=True`` standardises both clouds with the
    (sub-sampled) target's NaN-skipping mean and population standard deviation first (processing.py:462-464)."""
    from ._adjustment import _as_device, _stream
    lib = _lib.load()
    t = _as_device(tgt)
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float32)
    s_ = _as_device(sim, t.dtype)
    if scale:
        if N > 0:  # the statistics are those of the thinned series (processing.py:455-460 come first)
            t = t[:, :: -(-t.shape[1] // N)]
            s_ = s_[:, :: -(-s_.shape[1] // N)]
            N = 0
        avg = torch.nanmean(t, dim=1, keepdim=True)
        std = torch.sqrt(torch.nanmean((t - avg) ** 2, dim=1, keepdim=True))
        t = (t - avg) / std
        s_ = (s_ - avg) / std
    pshape = tuple(t.shape[2:])
    t = t.reshape(t.shape[0], t.shape[1], -1).contiguous()
    s_ = s_.reshape(s_.shape[0], s_.shape[1], -1).contiguous()
    V, Tt, Np = t.shape
    Ts = s_.shape[1]
    out = torch.empty((Np,), dtype=t.dtype, device=t.device)
    fn = getattr(lib, "xsdba_escore_f32" if t.dtype == torch.float32 else "xsdba_escore_f64")
    _lib.check(fn(t.data_ptr(), s_.data_ptr(), Np, 1, Np, Tt, Ts, V, Tt * Np, Ts * Np, int(N), out.data_ptr(), _stream()),
               "escore")
    return out.reshape(pshape)
