"""Device-resident columns that survive across calls and DataFrame.sort_values (item 3).

The reference re-unboxes pandas objects on every jitted call (`sdc/hiframes/boxing.py:89-140, 232`; strings are
copied element by element under the GIL, `sdc/str_arr_ext.py:1197-1238`), so every operator pays the host -> device
copy again.  `DeviceFrame` / `DeviceSeries` keep the columns in HBM -- numeric columns as 1-D CUDA tensors, string
columns as device `StringArray`s in the reference's str_arr layout -- and chain the same kernels:

    dev = DeviceFrame.from_pandas(df)                       # one H2D
    out = dev.merge(other, on="k").groupby("g").sum()       # join -> gather -> aggregate, all in HBM
    out.to_pandas()                                         # one D2H

Every method is the device form of the host mirror of the same name (`groupby.py`, `join.py`, `rolling.py`,
`sort.py`, `str_arr.py`): same argument meaning, same results; nothing here computes on the host.
  * groupby(by).<agg>() ........ sdc/datatypes/hpat_pandas_groupby_functions.py:361-480 (one key column, `:3066-3067`)
  * merge(how=inner|left|right|outer)  pandas.merge (the reference has no overload, sdc/hiframes/join.py:34-43)
  * rolling(w[, minp]).<op>() ... sdc/datatypes/hpat_pandas_dataframe_rolling_functions.py:167-189 (per column)
  * sort_values(by, ascending, na_position) ... the DataFrame form the reference's tests describe
    (sdc/tests/test_dataframe.py:908-1000); Series form sdc/datatypes/hpat_pandas_series_functions.py:3853-3959
"""
import numpy as np

from . import native
from .rolling import check_rolling_args


def _torch():
    import torch
    return torch


def _is_str(col):
    from .str_arr import StringArray
    return isinstance(col, StringArray)


def _to_device(values):
    """NumPy / pandas column -> device column (numeric tensor or StringArray)."""
    torch = _torch()
    from .str_arr import StringArray
    a = np.asarray(values)
    if a.dtype.kind in "OUS":
        return StringArray.from_pylist(list(a)).cuda()
    if a.dtype.kind == "b":
        a = a.astype(np.int64)
    if a.dtype.kind not in "iuf":
        raise TypeError("DeviceFrame: unsupported column dtype %s" % a.dtype)
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _to_host(col):
    if _is_str(col):
        return np.asarray(col.to_pylist(), dtype=object)
    return col.cpu().numpy()


def _take(col, idx, nullable=False):
    """col[idx] on the device; nullable: idx < 0 gives NaN / NA (numeric columns become float64)."""
    from .join import take_nullable_device
    from .sort import gather_device
    torch = _torch()
    if _is_str(col):
        return col.take(idx)
    if nullable:
        c = col if col.dtype in (torch.int64, torch.float64) else col.to(torch.float64 if col.dtype.is_floating_point else torch.int64)
        return take_nullable_device(c.contiguous(), idx)
    return gather_device(col.contiguous(), idx)


def _length(col):
    return col.n if _is_str(col) else col.shape[0]


class DeviceSeries:
    def __init__(self, data, name=None):
        self.data, self.name = data, name

    def __len__(self):
        return _length(self.data)

    @classmethod
    def from_pandas(cls, s):
        return cls(_to_device(s.to_numpy()), s.name)

    def to_pandas(self):
        import pandas as pd
        return pd.Series(_to_host(self.data), name=self.name)

    def rolling(self, window, min_periods=None, center=False, win_type=None, on=None, axis=0, closed=None):
        minp = check_rolling_args(window, min_periods, center, win_type, on, axis, closed)
        return _DeviceRolling({self.name: self.data}, int(window), minp, series=True)

    def sort_values(self, ascending=True, na_position="last"):
        """-> (sorted DeviceSeries, int64 device tensor of the original positions)."""
        idx = _sort_indexer(self.data, ascending, na_position)
        return DeviceSeries(_take(self.data, idx), self.name), idx

    def groupby(self, by, sort=True):
        keys = by.data if isinstance(by, DeviceSeries) else by
        if _length(keys) != len(self):
            raise ValueError("Method Series.groupby(). The object by\n expected: array of the same length")
        return _DeviceGroupBy({self.name: self.data}, keys, getattr(by, "name", None), sort)


def _sort_indexer(col, ascending, na_position):
    from .sort import sort_values_indexer_device
    if na_position not in ("last", "first"):
        raise ValueError("Method sort_values(). Unsupported parameter. Given na_position != 'last', 'first'")
    if _is_str(col):
        torch = _torch()
        idx = col.argsort(bool(ascending))
        # NA strings are stored empty: they sort first ascending / last descending; move the block where asked
        n_na = col.n - int(np.count_nonzero(col.valid_mask())) if col.n else 0
        if n_na:
            valid = torch.from_numpy(col.valid_mask()).to(idx.device)
            isna = ~valid[idx]
            idx = torch.cat([idx[isna], idx[~isna]]) if na_position == "first" else torch.cat([idx[~isna], idx[isna]])
        return idx
    return sort_values_indexer_device(col.contiguous(), bool(ascending), na_position)


class _DeviceRolling:
    def __init__(self, cols, window, minp, series=False):
        self._cols, self._w, self._minp, self._series = cols, window, minp, series

    def _apply(self, op, ddof=1):
        from .rolling import rolling_device
        torch = _torch()
        out = {}
        for name, col in self._cols.items():
            if _is_str(col):
                continue
            c = col if col.dtype in (torch.int64, torch.float64) else col.to(torch.float64 if col.dtype.is_floating_point else torch.int64)
            out[name] = rolling_device(c.contiguous(), op, self._w, self._minp, ddof)
        if self._series:
            (name, data), = out.items()
            return DeviceSeries(data, name)
        return DeviceFrame(out)

    def sum(self): return self._apply("sum")
    def mean(self): return self._apply("mean")
    def min(self): return self._apply("min")
    def max(self): return self._apply("max")
    def count(self): return self._apply("count")
    def var(self, ddof=1): return self._apply("var", int(ddof))
    def std(self, ddof=1): return self._apply("std", int(ddof))


class _DeviceGroupBy:
    def __init__(self, cols, keys, key_name, sort):
        self._cols, self._keys, self._key_name, self._sort = cols, keys, key_name, bool(sort)

    def agg(self, funcs, ddof=1):
        """-> DeviceFrame: the key column first, then one column per (value column, aggregate) named `col` for a
        single aggregate and `col_agg` for several."""
        from .groupby import groupby_device
        from .str_arr import groupby_str_device
        torch = _torch()
        funcs = (funcs,) if isinstance(funcs, str) else tuple(funcs)
        names = [n for n, c in self._cols.items() if not _is_str(c)]
        cols = []
        for n in names:
            c = self._cols[n]
            cols.append((c if c.dtype in (torch.int64, torch.float64)
                         else c.to(torch.float64 if c.dtype.is_floating_point else torch.int64)).contiguous())
        if _is_str(self._keys):
            rk, res = groupby_str_device(self._keys, cols, funcs, self._sort, ddof)
        else:
            k = self._keys
            k = (k if k.dtype in (torch.int64, torch.float64)
                 else k.to(torch.float64 if k.dtype.is_floating_point else torch.int64)).contiguous()
            rk, res = groupby_device(k, cols, funcs, self._sort, ddof)
        out = {self._key_name if self._key_name is not None else "key": rk}
        for a in funcs:
            for n, t in zip(names, res[a]):
                out[n if len(funcs) == 1 else "%s_%s" % (n, a)] = t
        return DeviceFrame(out, index=self._key_name if self._key_name is not None else "key")

    def sum(self): return self.agg("sum")
    def mean(self): return self.agg("mean")
    def min(self): return self.agg("min")
    def max(self): return self.agg("max")
    def count(self): return self.agg("count")
    def prod(self): return self.agg("prod")
    def median(self): return self.agg("median")
    def var(self, ddof=1): return self.agg("var", int(ddof))
    def std(self, ddof=1): return self.agg("std", int(ddof))


class DeviceFrame:
    """Ordered mapping name -> device column.  `index`: the column to_pandas() turns into the index (groupby results)."""

    def __init__(self, cols, index=None):
        self._cols = dict(cols)
        self._index = index
        lens = {_length(c) for c in self._cols.values()}
        if len(lens) > 1:
            raise ValueError("DeviceFrame: columns of different lengths %s" % sorted(lens))

    @classmethod
    def from_pandas(cls, df):
        return cls({c: _to_device(df[c].to_numpy()) for c in df.columns})

    def to_pandas(self):
        import pandas as pd
        df = pd.DataFrame({n: _to_host(c) for n, c in self._cols.items()})
        return df.set_index(self._index) if self._index is not None else df

    @property
    def columns(self):
        return list(self._cols)

    def __len__(self):
        return _length(next(iter(self._cols.values()))) if self._cols else 0

    def __getitem__(self, item):
        if isinstance(item, str):
            return DeviceSeries(self._cols[item], item)
        return DeviceFrame({c: self._cols[c] for c in item})

    def take(self, idx, nullable=False):
        return DeviceFrame({n: _take(c, idx, nullable) for n, c in self._cols.items()})

    # ---- operators ----
    def groupby(self, by, sort=True):
        if not isinstance(by, str) or by not in self._cols:
            raise TypeError("Method groupby(). The object by\n expected: one column name")
        return _DeviceGroupBy({n: c for n, c in self._cols.items() if n != by}, self._cols[by], by, sort)

    def rolling(self, window, min_periods=None, center=False, win_type=None, on=None, axis=0, closed=None):
        minp = check_rolling_args(window, min_periods, center, win_type, on, axis, closed)
        return _DeviceRolling(self._cols, int(window), minp)

    def sort_values(self, by, ascending=True, na_position="last", kind="quicksort"):
        """Rows ordered by ONE column (stable; NaN / NA block last or first), every column gathered on the device."""
        if not isinstance(by, str) or by not in self._cols:
            raise KeyError(by)
        if kind not in ("quicksort", "mergesort"):
            raise ValueError("Method sort_values(). Unsupported parameter. Given kind != 'quicksort', 'mergesort'")
        return self.take(_sort_indexer(self._cols[by], ascending, na_position))

    def merge(self, right, how="inner", on=None, left_on=None, right_on=None):
        """pandas.merge on one numeric key column; payload columns gathered on the device, `_x` / `_y` suffixes for
        overlapping names, the key column coalesced for rows without a left partner."""
        from .join import join_device
        torch = _torch()
        if how not in native.JOIN_HOW:
            raise ValueError("merge: how=%r is not supported (inner / left / right / outer)" % (how,))
        if on is not None:
            if left_on is not None or right_on is not None:
                raise ValueError('Can only pass argument "on" OR "left_on" and "right_on", not a combination of both.')
            left_on = right_on = on
        if left_on is None or right_on is None:
            common = [c for c in self._cols if c in right._cols]
            if len(common) != 1:
                raise ValueError("merge: exactly one key column is supported")
            left_on = right_on = common[0]
        lk, rk = self._cols[left_on], right._cols[right_on]
        if _is_str(lk) or _is_str(rk):
            raise TypeError("merge: key dtype str is not supported (int64 / float64)")
        if lk.dtype != rk.dtype or lk.dtype not in (torch.int64, torch.float64):
            both_int = not lk.dtype.is_floating_point and not rk.dtype.is_floating_point
            lk, rk = lk.to(torch.int64 if both_int else torch.float64), rk.to(torch.int64 if both_int else torch.float64)
        li, ri = join_device(lk.contiguous(), rk.contiguous(), how)
        left_null, right_null = how in ("right", "outer"), how in ("left", "outer")
        shared = left_on == right_on
        out = {}
        for n, c in self._cols.items():
            name = "%s_x" % n if (n in right._cols and not (shared and n == left_on)) else n
            if shared and n == left_on and left_null:
                # the key of a row without a left partner comes from the right side
                lv = _take(c.to(rk.dtype) if c.dtype != rk.dtype else c, li.clamp(min=0))
                rv = _take(right._cols[right_on].to(rk.dtype) if right._cols[right_on].dtype != rk.dtype else right._cols[right_on],
                           ri.clamp(min=0))
                out[name] = torch.where(li < 0, rv, lv)          # plumbing: picks between two gathered columns
                continue
            out[name] = _take(c, li, left_null)
        for n, c in right._cols.items():
            if shared and n == right_on:
                continue
            out["%s_y" % n if n in self._cols else n] = _take(c, ri, right_null)
        return DeviceFrame(out)
