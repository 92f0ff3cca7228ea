"""Daily time coordinates without xarray/cftime: the fields ``ds.time.dt`` gives the reference.

The reference derives group membership from ``ds.indexes['time']`` (``.month``, ``.dayofyear``,
``.day``, ``.days_in_month``; base.py:302-329) and the number of day-of-year groups from the calendar
(``max_doy``, base.py:105-115).  ``TimeAxis`` carries exactly those fields.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

#: number of day-of-year groups per calendar (base.py:105-115)
MAX_DOY = {"standard": 366, "gregorian": 366, "proleptic_gregorian": 366, "julian": 366, "noleap": 365,
           "365_day": 365, "all_leap": 366, "366_day": 366, "360_day": 360}

_MONTH_LEN = np.array([31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31], dtype=np.int64)


def is_leap_year(year, calendar: str):
    year = np.asarray(year)
    if calendar in ("noleap", "365_day", "360_day"):
        return np.zeros(year.shape, bool)
    if calendar in ("all_leap", "366_day"):
        return np.ones(year.shape, bool)
    if calendar == "julian":
        return year % 4 == 0
    return (year % 4 == 0) & ((year % 100 != 0) | (year % 400 == 0))


def month_lengths(year: int, calendar: str) -> np.ndarray:
    if calendar == "360_day":
        return np.full(12, 30, np.int64)
    ml = _MONTH_LEN.copy()
    if bool(is_leap_year(year, calendar)):
        ml[1] = 29
    return ml


@dataclass(frozen=True)
class TimeAxis:
    year: np.ndarray
    month: np.ndarray
    day: np.ndarray
    dayofyear: np.ndarray
    days_in_month: np.ndarray
    calendar: str = "standard"

    def __len__(self) -> int:
        return int(self.year.shape[0])

    def __getitem__(self, sl) -> "TimeAxis":
        return TimeAxis(self.year[sl], self.month[sl], self.day[sl], self.dayofyear[sl], self.days_in_month[sl],
                        self.calendar)

    @property
    def ordinal(self) -> np.ndarray:
        """Days since the first element (float64) -- the x-axis of the trend fits."""
        if self.calendar == "360_day":
            per_year = np.full(self.year.shape, 360)
        else:
            per_year = None
        y0 = int(self.year.min())
        years = np.arange(y0, int(self.year.max()) + 1)
        if per_year is None:
            ylen = np.array([month_lengths(int(y), self.calendar).sum() for y in years])
        else:
            ylen = np.full(years.shape, 360)
        start = np.concatenate([[0], np.cumsum(ylen)[:-1]])
        o = start[self.year - y0] + self.dayofyear - 1
        return (o - o[0]).astype(np.float64)

    @staticmethod
    def daily(start_year: int, n_years: int, calendar: str = "noleap") -> "TimeAxis":
        ys, ms, ds, doys, dims = [], [], [], [], []
        for y in range(start_year, start_year + n_years):
            ml = month_lengths(y, calendar)
            doy = 1
            for m in range(12):
                n = int(ml[m])
                ys.append(np.full(n, y)); ms.append(np.full(n, m + 1)); ds.append(np.arange(1, n + 1))
                doys.append(np.arange(doy, doy + n)); dims.append(np.full(n, n))
                doy += n
        cat = lambda parts: np.concatenate(parts).astype(np.int64)  # noqa: E731
        return TimeAxis(cat(ys), cat(ms), cat(ds), cat(doys), cat(dims), calendar)

    @staticmethod
    def from_fields(year, month, day, calendar: str = "standard") -> "TimeAxis":
        year = np.asarray(year, np.int64); month = np.asarray(month, np.int64); day = np.asarray(day, np.int64)
        dim = np.empty_like(year); doy = np.empty_like(year)
        for y in np.unique(year):
            ml = month_lengths(int(y), calendar)
            before = np.concatenate([[0], np.cumsum(ml)[:-1]])
            sel = year == y
            dim[sel] = ml[month[sel] - 1]
            doy[sel] = before[month[sel] - 1] + day[sel]
        return TimeAxis(year, month, day, doy, dim, calendar)

    @staticmethod
    def from_datetime64(t) -> "TimeAxis":
        t = np.asarray(t).astype("datetime64[D]")
        y = t.astype("datetime64[Y]").astype(np.int64) + 1970
        m = t.astype("datetime64[M]").astype(np.int64) % 12 + 1
        d = (t - t.astype("datetime64[M]")).astype(np.int64) + 1
        return TimeAxis.from_fields(y, m, d, "standard")
