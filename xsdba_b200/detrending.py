"""Host mirror of ``xsdba.detrending`` for the DQM path (detrending.py:165-296): parameter objects with
the reference's names; the fits run in the CUDA library."""
from __future__ import annotations

from .base import parse_group


class BaseDetrend:
    def __init__(self, *, group="time", kind="+", **kw):
        self.group = parse_group(group)
        self.kind = kind
        self.params = kw


class PolyDetrend(BaseDetrend):
    """``PolyDetrend(group, kind, degree, preserve_mean)`` (detrending.py:165-208)."""

    def __init__(self, group="time", kind="+", degree=4, preserve_mean=False, mult_skip_zeros=False):
        if mult_skip_zeros:
            raise NotImplementedError("mult_skip_zeros is not built in xsdba_b200 yet")
        if not 0 <= int(degree) <= 4:
            raise NotImplementedError("xsdba_b200 fits polynomial trends of degree 0..4")
        super().__init__(group=group, kind=kind, degree=int(degree), preserve_mean=bool(preserve_mean))
        self.degree = int(degree)
        self.preserve_mean = bool(preserve_mean)


class LoessDetrend(BaseDetrend):
    """``LoessDetrend(group, kind, f, niter, d, weights, ...)`` (detrending.py:211-272)."""

    def __init__(self, group="time", kind="+", f=0.2, niter=1, d=0, weights="tricube", equal_spacing=None,
                 skipna=True, mult_skip_zeros=False):
        if mult_skip_zeros or not skipna or weights not in ("tricube", "gaussian"):
            raise NotImplementedError("only skipna LOESS with tricube / gaussian weights is built in xsdba_b200")
        super().__init__(group=group, kind=kind, f=f, niter=niter, d=d, weights=weights)
        self.f, self.niter, self.d, self.weights, self.equal_spacing = float(f), int(niter), int(d), weights, equal_spacing
