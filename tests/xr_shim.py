"""A minimal duck type of xarray's Dataset / DataArray (xarray is absent from this image): just what the L4
functions of the reference touch -- ``ds[name]``, ``ds.data_vars``, ``ds.indexes["time"]``, ``ds.assign``,
``da.dims / .values / .transpose / .attrs``, attribute access, keyword constructors.  Test infrastructure only."""
import numpy as np


class DataArray:
    def __init__(self, data, dims=None, coords=None, name=None, attrs=None):
        self.values = np.asarray(data)
        self.dims = tuple(dims) if dims is not None else tuple(f"dim_{i}" for i in range(self.values.ndim))
        assert len(self.dims) == self.values.ndim, (self.dims, self.values.shape)
        self.coords = dict(coords or {})
        self.name = name
        self.attrs = dict(attrs or {})

    @property
    def dtype(self):
        return self.values.dtype

    @property
    def shape(self):
        return self.values.shape

    @property
    def units(self):
        return self.attrs["units"]

    def transpose(self, *dims):
        if Ellipsis in dims:
            i = dims.index(Ellipsis)
            rest = [d for d in self.dims if d not in dims]
            dims = tuple(dims[:i]) + tuple(rest) + tuple(dims[i + 1:])
        assert sorted(dims) == sorted(self.dims), (dims, self.dims)
        return DataArray(np.transpose(self.values, [self.dims.index(d) for d in dims]), dims, self.coords, self.name, self.attrs)

    def astype(self, dt):
        return DataArray(self.values.astype(dt), self.dims, self.coords, self.name, self.attrs)


class Dataset:
    def __init__(self, data_vars=None, coords=None, attrs=None):
        self.data_vars = {}
        self.coords = dict(coords or {})
        self.attrs = dict(attrs or {})
        for k, v in (data_vars or {}).items():
            self[k] = v

    def __setitem__(self, k, v):
        if not isinstance(v, DataArray):
            v = DataArray(v[1], v[0])
        for d, c in v.coords.items():
            self.coords.setdefault(d, DataArray(np.asarray(c), (d,)))
        self.data_vars[k] = v

    def __getitem__(self, k):
        if k in self.data_vars:
            return self.data_vars[k]
        c = self.coords[k]
        return c if isinstance(c, DataArray) else DataArray(np.asarray(c), (k,))

    def __contains__(self, k):
        return k in self.data_vars

    def __iter__(self):
        return iter(self.data_vars)

    def __getattr__(self, k):
        try:
            return self.__dict__["data_vars"][k]
        except KeyError as e:
            raise AttributeError(k) from e

    @property
    def indexes(self):
        return {k: (v.values if isinstance(v, DataArray) else v) for k, v in self.coords.items()}

    def assign(self, **kw):
        out = Dataset(dict(self.data_vars), dict(self.coords), dict(self.attrs))
        for k, v in kw.items():
            out[k] = v
        return out

    def drop_vars(self, names):
        return Dataset({k: v for k, v in self.data_vars.items() if k not in names}, dict(self.coords), dict(self.attrs))


class CFTime:
    """cftime-like element: year / month / day / calendar."""

    def __init__(self, y, m, d, calendar):
        self.year, self.month, self.day, self.calendar = y, m, d, calendar


def time_index(time_axis):
    """TimeAxis -> array of cftime-like objects (what ds.indexes['time'] yields for a CFTimeIndex)."""
    return np.array([CFTime(int(y), int(m), int(d), time_axis.calendar)
                     for y, m, d in zip(time_axis.year, time_axis.month, time_axis.day)], dtype=object)
