"""Moving multi-year windows over a daily series: ``xsdba.base.stack_periods`` / ``unstack_periods``
(base.py:1072-1381) for ``freq="YS"``, the production pattern of adjusting a 150-year simulation in 30-year windows
(SURVEY.md section 8f rank 4).

Without xarray a "stacked" array is a list of time slices of the original series plus the parameters needed to put
the pieces back; :func:`adjust_periods` runs a trained adjustment on every window and reassembles the result.
Only index logic lives here -- the adjustment itself goes through the CUDA path.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .calendar import TimeAxis, month_lengths

#: calendars whose years all have the same length (xsdba.base.uniform_calendars)
UNIFORM_CALENDARS = ("noleap", "365_day", "all_leap", "366_day", "360_day")


@dataclass(frozen=True)
class Periods:
    """What ``stack_periods`` records on the ``period`` coordinate (start of each window, window, stride) plus the
    slices of the source time axis (base.py:1213-1232)."""
    slices: tuple
    window: int
    stride: int
    start_years: tuple = field(default=())

    def __len__(self) -> int:
        return len(self.slices)

    @property
    def lengths(self):
        return tuple(s.stop - s.start for s in self.slices)


def _ends_on_year_end(time: TimeAxis) -> bool:
    y, m, d = int(time.year[-1]), int(time.month[-1]), int(time.day[-1])
    return m == 12 and d == int(month_lengths(y, time.calendar)[11])


def stack_periods(time: TimeAxis, window: int = 30, stride: int | None = None, min_length: int | None = None,
                  align_days: bool = True) -> Periods:
    """Windows of ``window`` years taken every ``stride`` years (``freq="YS"``), as slices of ``time``.

    Follows base.py:1150-1211: strides are year bins starting at the first time step's year; a window must be
    complete (``min_length`` years, default ``window``) -- the iteration stops at the first window that runs past the
    end of the series, where "past the end" is judged with one extra time step appended (base.py:1178-1188), so a
    series ending on 31 December closes its last window; a first window that does not start in January is skipped
    when ``min_length == window`` (base.py:1200-1208).
    """
    stride = window if stride is None else stride
    min_length = window if min_length is None else min_length
    if stride > window:
        raise ValueError(f"Stride must be less than or equal to window. Got {stride} > {window}.")
    if align_days and time.calendar not in UNIFORM_CALENDARS:
        raise ValueError(
            f"Stacking {window}YS periods will result in unaligned day-of-year. Consider converting the calendar of "
            "your data to one with uniform year lengths, or pass `align_days=False` to disable this check.")
    year = np.asarray(time.year)
    n = len(time)
    if n == 0:
        return Periods((), window, stride)
    y0, y_last = int(year[0]), int(year[-1])
    # with the extra step of `time2` the series is known up to (excluding) this year boundary
    closed_until = y_last + 1 if _ends_on_year_end(time) else y_last

    def first_index_of_year(y):
        return int(np.searchsorted(year, y, side="left"))

    slices, starts = [], []
    k = 0
    while True:
        ys = y0 + k * stride
        s = first_index_of_year(ys)
        if s >= n:
            break
        if ys + min_length > closed_until:  # the (minimum-length) window is open-ended: stop
            break
        k += 1
        if s == 0 and min_length == window and int(time.month[0]) != 1:
            continue  # fractional first period
        e = first_index_of_year(ys + window) if ys + window <= y_last else n
        slices.append(slice(s, e))
        starts.append(ys)
    return Periods(tuple(slices), window, stride, tuple(starts))


def unstack_periods(pieces, periods: Periods, time: TimeAxis, time_axis: int = 0):
    """Inverse of :func:`stack_periods` for per-window arrays ``pieces[i]`` (base.py:1274-1381): with
    ``stride == window`` the windows are concatenated; otherwise ``window / stride`` must be odd and only the
    centre-most stride of each window is kept, except the beginning of the first and the end of the last window.
    Returns the reassembled array and the slice of ``time`` it covers."""
    if len(pieces) != len(periods):
        raise ValueError("one array per period is needed")
    if len(periods) == 0:
        raise ValueError("no period to unstack")
    window, stride = periods.window, periods.stride
    cat = _concat
    if window == stride:
        return cat(list(pieces), time_axis), slice(periods.slices[0].start, periods.slices[-1].stop)
    if (window / stride) % 2 != 1:
        raise NotImplementedError(
            "`unstack_periods` can't work with strides that do not divide the window into an odd number of parts."
            f"Got {window} / {stride} which is not an odd integer.")
    n_win = window // stride
    mid = (n_win - 1) // 2
    year = np.asarray(time.year)
    out = []
    first = last = None
    for i, (slc, ys) in enumerate(zip(periods.slices, periods.start_years)):
        yrs = year[slc]
        length = slc.stop - slc.start
        sec_start = lambda j: int(np.searchsorted(yrs, ys + j * stride, side="left"))  # noqa: E731
        if i == 0:
            a, b = 0, min(sec_start(mid + 1), length)
        elif i == len(periods) - 1:
            a, b = sec_start(mid), length
        else:
            a, b = sec_start(mid), min(sec_start(mid + 1), length)
        out.append(_take(pieces[i], a, b, time_axis))
        if first is None:
            first = slc.start + a
        last = slc.start + b
    return cat(out, time_axis), slice(first, last)


def _take(a, lo, hi, axis):
    idx = [slice(None)] * a.ndim
    idx[axis] = slice(lo, hi)
    return a[tuple(idx)]


def _concat(parts, axis):
    if hasattr(parts[0], "detach"):
        import torch
        return torch.cat(parts, dim=axis)
    return np.concatenate(parts, axis=axis)


def adjust_periods(obj, sim, *, time: TimeAxis, window: int = 30, stride: int | None = None, time_axis: int = 0,
                   **adjust_kw):
    """``unstack_periods(obj.adjust(stack_periods(sim)))``: adjust ``sim`` window by window with a trained
    adjustment object and stitch the centre strides back together.  Returns (scen, covered time slice)."""
    periods = stack_periods(time, window=window, stride=stride)
    if len(periods) == 0:
        raise ValueError("the series is shorter than one window")
    pieces = []
    for slc in periods.slices:
        pieces.append(obj.adjust(_take(sim, slc.start, slc.stop, time_axis), time=time[slc], time_axis=time_axis,
                                 **adjust_kw))
    return unstack_periods(pieces, periods, time, time_axis=time_axis)
