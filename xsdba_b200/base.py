"""Host mirror of ``xsdba.base.Grouper`` for the hot path (base.py:118-501).

Same constructor, names and meaning as the reference's ``Grouper``; group membership is computed on
the host from a :class:`~xsdba_b200.calendar.TimeAxis` and handed to the CUDA library as a small
index table (``xsdba_grouping_create``).
"""
from __future__ import annotations

import ctypes as C
import weakref
from collections import OrderedDict

import numpy as np

from . import _lib
from .calendar import MAX_DOY, TimeAxis, is_leap_year


class GroupingHandle:
    """Owns one ``xsdba_grouping_t`` (device-side membership / window tables)."""

    def __init__(self, gidx: np.ndarray, n_groups: int, window: int):
        lib = _lib.load()
        gidx = np.ascontiguousarray(gidx, dtype=np.int32)
        ptr = C.c_void_p()
        st = lib.xsdba_grouping_create(C.byref(ptr), gidx.ctypes.data_as(_lib.c_i32p), gidx.size, n_groups, window)
        _lib.check(st, "xsdba_grouping_create")
        self.ptr = ptr
        self.n_groups = n_groups
        self.window = window
        self.n_time = int(gidx.size)
        self.max_segment = int(lib.xsdba_grouping_max_segment(ptr))
        self._finalizer = weakref.finalize(self, lib.xsdba_grouping_destroy, ptr)


_HANDLES: "OrderedDict[tuple, GroupingHandle]" = OrderedDict()
_HANDLES_LOCK = __import__("threading").Lock()   # dask may call blocks from several threads (base.py:715)


def grouping_handle(gidx: np.ndarray, n_groups: int, window: int) -> GroupingHandle:
    import torch
    dev = torch.cuda.current_device() if torch.cuda.is_available() else -1   # the tables live on one device
    key = (gidx.astype(np.int32).tobytes(), int(n_groups), int(window), dev)
    with _HANDLES_LOCK:
        h = _HANDLES.get(key)
        if h is None:
            h = GroupingHandle(gidx, n_groups, window)
            _HANDLES[key] = h
            while len(_HANDLES) > 16:
                _HANDLES.popitem(last=False)
        else:
            _HANDLES.move_to_end(key)
        return h


class Grouper:
    """``Grouper(group, window=1, add_dims=None)`` -- see base.py:128-175.

    ``group`` is ``"time"`` or ``"time.<prop>"`` with prop in month / dayofyear / season.
    """

    def __init__(self, group: str, window: int = 1, add_dims=None):
        if "." in group:
            dim, prop = group.split(".")
        else:
            dim, prop = group, "group"
        if dim != "time":
            raise NotImplementedError("xsdba_b200 groups along 'time' only")
        if not isinstance(window, (int, np.integer)) or window < 1:
            raise ValueError("window must be a positive integer")
        if window > 1 and "." not in group:
            # base.py:151-156
            raise ValueError("Grouping windows are meant for grouping along a time accessor (e.g. 'time.dayofyear')")
        if prop not in ("group", "month", "dayofyear", "season"):
            raise NotImplementedError(f"grouping on time.{prop} is not supported")
        # add_dims (base.py:410-415): dimension NAMES pooled with time in the group-wise reductions.  Only the
        # Dataset-level seam (xr_adapter) knows dimension names; it pools them into the time axis before it calls the
        # array-level functions, which refuse a Grouper that still carries add_dims (see handle()).
        self.dim, self.prop, self.name, self.window = dim, prop, group, int(window)
        self.add_dims = list(add_dims or [])

    def __repr__(self):
        return f"Grouper(name='{self.name}', window={self.window})"

    # base.py:207-230
    def get_coordinate(self, time: TimeAxis | None = None) -> np.ndarray:
        if self.prop == "month":
            return np.arange(1, 13)
        if self.prop == "season":
            return np.array(["DJF", "MAM", "JJA", "SON"])
        if self.prop == "dayofyear":
            mdoy = MAX_DOY[time.calendar] if time is not None else 365
            return np.arange(1, mdoy + 1)
        return np.array([1])

    def n_groups(self, time: TimeAxis | None = None) -> int:
        return int(self.get_coordinate(time).shape[0])

    # base.py:274-345
    def get_index(self, time: TimeAxis, interp: bool | None = None) -> np.ndarray:
        if self.prop == "group":
            return np.ones(len(time), dtype=int)
        if interp:
            if self.prop == "month":
                return time.month - 0.5 + time.day / time.days_in_month
            if self.prop == "season":
                cal = time.calendar
                length_year = 360 if cal == "360_day" else 365 + (0 if cal == "noleap" else is_leap_year(time.year, cal))
                return time.dayofyear / length_year * 4 - 1 / 6
            return time.dayofyear
        if self.prop == "season":
            return time.month % 12 // 3
        return getattr(time, self.prop)

    def zero_based_index(self, time: TimeAxis) -> np.ndarray:
        """0-based row of each time step in the trained tables."""
        if self.prop == "group":
            return np.zeros(len(time), np.int32)
        idx = self.get_index(time)
        return (idx if self.prop == "season" else idx - 1).astype(np.int32)

    def handle(self, time: TimeAxis, with_window: bool = True) -> GroupingHandle:
        if self.add_dims:
            raise NotImplementedError("add_dims pooling needs named dimensions: go through xsdba_b200.xr_adapter "
                                      "(base.py:410-415)")
        return grouping_handle(self.zero_based_index(time), self.n_groups(time), self.window if with_window else 1)


def parse_group(group, window: int = 1) -> Grouper:
    """base.py:504-538: accept a str (+ window kwarg) or a Grouper."""
    return group if isinstance(group, Grouper) else Grouper(group, window=window)
