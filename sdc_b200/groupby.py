"""DataFrame.groupby(by).<agg>() / Series.groupby(arr).<agg>() on the B200 backend.

Host-side mirror of the reference's groupby interface:
  * DataFrame.groupby(by) -- `by` must be ONE column name ... sdc/datatypes/hpat_pandas_dataframe_functions.py:2993-3106
  * Series.groupby(by)    -- `by` a 1-D array of equal length . sdc/datatypes/hpat_pandas_series_functions.py:4720-4811
  * gb['B'] -> SeriesGroupBy, gb['B','C'] -> column subset, a second [] raises IndexError
                                                               sdc/datatypes/hpat_pandas_groupby_functions.py:144-191
  * count / max / mean / median / min / prod / std(ddof) / sum / var(ddof) ... groupby_functions.py:361-480, 483-600
The aggregation runs in libsdc_b200.so (csrc/groupby.cu: one hash-aggregate pass); nothing here
touches a row.  torch only holds device buffers.
"""
import contextlib
import ctypes as ct

import numpy as np

from . import native

AGGS = ("sum", "mean", "min", "max", "count", "var", "std", "prod", "median")


def _torch():
    import torch
    return torch


def _tname(t):
    return str(t.dtype).replace("torch.", "")


def groupby_device(keys, cols, aggs, sort=True, ddof=1, expected_groups=0, gid_range=0):
    """Hash groupby-aggregate of device columns.

    keys : 1-D contiguous CUDA tensor, int64 or float64 (NaN keys are skipped, -0.0 == +0.0);
           gid_range > 0: int64 group ids in [0, gid_range), negative ids skip the row (SDCB200_GID)
    cols : list of 1-D contiguous CUDA tensors (int64 / float64) of the same length
    aggs : iterable of names from AGGS -- every aggregate is produced for every column
    -> (group_keys tensor[ng], {agg: [tensor[ng] per column]})
    Result dtypes: float64, except count -> int64 and sum/min/max of an int64 column -> int64.
    """
    torch = _torch()
    lib = native.load()
    aggs = tuple(aggs)
    assert keys.is_cuda and keys.dim() == 1 and keys.is_contiguous()
    kname = _tname(keys)
    if kname not in ("int64", "float64"):
        raise TypeError("groupby: key dtype %s is not supported (int64 / float64 / str_arr)" % kname)
    n = keys.shape[0]
    mask = 0
    for a in aggs:
        if a not in native.AGG_BITS:
            raise ValueError("groupby: unknown aggregate %r" % (a,))
        mask |= native.AGG_BITS[a]
    ptrs = (ct.c_void_p * max(1, len(cols)))()
    dts = (ct.c_int32 * max(1, len(cols)))()
    for i, c in enumerate(cols):
        assert c.is_cuda and c.dim() == 1 and c.is_contiguous() and c.shape[0] == n
        cn = _tname(c)
        if cn not in ("int64", "float64"):
            raise TypeError("groupby: value column dtype %s is not supported" % cn)
        ptrs[i] = c.data_ptr()
        dts[i] = native.DTYPE_CODES[cn]
    handle, ng = ct.c_int64(0), ct.c_int64(0)
    # switching the current device costs ~10 us in torch: only when the keys live on another one
    with (torch.cuda.device(keys.device) if torch.cuda.current_device() != keys.device.index else contextlib.nullcontext()):
        stream = native.current_stream_ptr()
        if gid_range > 0:
            if kname != "int64":
                raise TypeError("groupby: group ids must be int64")
            kcode, hint = native.DTYPE_GID, int(gid_range)
        else:
            kcode, hint = native.DTYPE_CODES[kname], int(expected_groups)
        native.check(lib.sdcb200_groupby(kcode, keys.data_ptr(), n, len(cols), ptrs, dts, mask,
                                         1 if sort else 0, int(ddof), hint, stream,
                                         ct.byref(handle), ct.byref(ng)))
        g = ng.value
        owner = _ResultHandle(handle.value)
        if g == 0:
            out_keys = torch.empty(0, dtype=keys.dtype, device=keys.device)
            return out_keys, {a: [torch.empty(0, dtype=_out_dtype(torch, a, c), device=keys.device) for c in cols]
                              for a in aggs}
        # the result stays where the library wrote it: tensors view the handle's buffers and keep it alive
        kp, vp, stride = ct.c_void_p(), ct.c_void_p(), ct.c_int64(0)
        native.check(lib.sdcb200_groupby_result(handle.value, ct.byref(kp), ct.byref(vp), ct.byref(stride)))
        # one zero-copy view per output row (a view over the interface costs ~2 us, slicing a matrix view and
        # reinterpreting its dtype three times that); every view keeps the handle alive
        dev = keys.device
        out_keys = torch.as_tensor(_DeviceView(kp.value, (g,), "<f8" if kname == "float64" else "<i8", owner), device=dev)
        ranks = sorted(set(aggs), key=lambda a: native.AGG_BITS[a])
        row_bytes = 8 * stride.value
        res = {}
        for a in aggs:
            base = vp.value + ranks.index(a) * len(cols) * row_bytes if cols else 0
            res[a] = [torch.as_tensor(_DeviceView(base + i * row_bytes, (g,),
                                                  "<i8" if _out_dtype(torch, a, c) == torch.int64 else "<f8", owner), device=dev)
                      for i, c in enumerate(cols)]
    return out_keys, res


class PendingGroupby:
    """A groupby whose kernel is in flight (`sdcb200_groupby_begin`); result() waits and returns what groupby_device returns.
    The key and value columns are kept alive until then (the library reruns an overflowed optimistic attempt from them)."""

    def __init__(self, handle, keys, cols, aggs):
        self._handle, self._keys, self._cols, self._aggs = handle, keys, cols, aggs
        self._done = None

    def result(self):
        if self._done is None:
            torch = _torch()
            lib = native.load()
            ng = ct.c_int64(0)
            keys, cols, aggs = self._keys, self._cols, self._aggs
            with (torch.cuda.device(keys.device) if torch.cuda.current_device() != keys.device.index else contextlib.nullcontext()):
                native.check(lib.sdcb200_groupby_end(self._handle, ct.byref(ng)))
                self._done = _wrap_result(lib, torch, self._handle, ng.value, keys, cols, aggs, _tname(keys))
            self._keys = self._cols = None
        return self._done


def groupby_device_async(keys, cols, aggs, sort=True, ddof=1, expected_groups=0):
    """groupby_device without the wait: enqueues the aggregate on torch's current stream and returns a PendingGroupby.
    Back-to-back small aggregates (C1: 10 M rows) then pipeline -- the host prepares the next call while the kernel of
    the previous one runs -- instead of paying a stream synchronisation each."""
    torch = _torch()
    lib = native.load()
    aggs = tuple(aggs)
    mask, ptrs, dts, kname = _marshal(keys, cols, aggs)
    handle = ct.c_int64(0)
    with (torch.cuda.device(keys.device) if torch.cuda.current_device() != keys.device.index else contextlib.nullcontext()):
        native.check(lib.sdcb200_groupby_begin(native.DTYPE_CODES[kname], keys.data_ptr(), keys.shape[0], len(cols), ptrs, dts, mask,
                                               1 if sort else 0, int(ddof), int(expected_groups), native.current_stream_ptr(),
                                               ct.byref(handle)))
    return PendingGroupby(handle.value, keys, list(cols), aggs)


def _marshal(keys, cols, aggs):
    assert keys.is_cuda and keys.dim() == 1 and keys.is_contiguous()
    kname = _tname(keys)
    if kname not in ("int64", "float64"):
        raise TypeError("groupby: key dtype %s is not supported (int64 / float64 / str_arr)" % kname)
    n = keys.shape[0]
    mask = 0
    for a in aggs:
        if a not in native.AGG_BITS:
            raise ValueError("groupby: unknown aggregate %r" % (a,))
        mask |= native.AGG_BITS[a]
    ptrs = (ct.c_void_p * max(1, len(cols)))()
    dts = (ct.c_int32 * max(1, len(cols)))()
    for i, c in enumerate(cols):
        assert c.is_cuda and c.dim() == 1 and c.is_contiguous() and c.shape[0] == n
        cn = _tname(c)
        if cn not in ("int64", "float64"):
            raise TypeError("groupby: value column dtype %s is not supported" % cn)
        ptrs[i] = c.data_ptr()
        dts[i] = native.DTYPE_CODES[cn]
    return mask, ptrs, dts, kname


def _wrap_result(lib, torch, handle, g, keys, cols, aggs, kname):
    owner = _ResultHandle(handle)
    if g == 0:
        out_keys = torch.empty(0, dtype=keys.dtype, device=keys.device)
        return out_keys, {a: [torch.empty(0, dtype=_out_dtype(torch, a, c), device=keys.device) for c in cols] for a in aggs}
    kp, vp, stride = ct.c_void_p(), ct.c_void_p(), ct.c_int64(0)
    native.check(lib.sdcb200_groupby_result(handle, ct.byref(kp), ct.byref(vp), ct.byref(stride)))
    dev = keys.device
    out_keys = torch.as_tensor(_DeviceView(kp.value, (g,), "<f8" if kname == "float64" else "<i8", owner), device=dev)
    ranks = sorted(set(aggs), key=lambda a: native.AGG_BITS[a])
    row_bytes = 8 * stride.value
    res = {}
    for a in aggs:
        base = vp.value + ranks.index(a) * len(cols) * row_bytes if cols else 0
        res[a] = [torch.as_tensor(_DeviceView(base + i * row_bytes, (g,),
                                              "<i8" if _out_dtype(torch, a, c) == torch.int64 else "<f8", owner), device=dev)
                  for i, c in enumerate(cols)]
    return out_keys, res


def _out_dtype(torch, agg, col):
    is_int = agg == "count" or (col.dtype == torch.int64 and agg in ("sum", "min", "max", "prod"))
    return torch.int64 if is_int else torch.float64


class _ResultHandle:
    """Owns a groupby result handle; frees it when the last tensor viewing it is gone."""

    def __init__(self, handle):
        self.handle = handle

    def __del__(self):
        try:
            native.load().sdcb200_groupby_free(self.handle)
        except Exception:
            pass


class _DeviceView:
    """Device buffer exposed through __cuda_array_interface__ (zero-copy into torch)."""

    def __init__(self, ptr, shape, typestr, owner):
        self.owner = owner
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2,
                                         "strides": None}


def _value_column(values):
    """NumPy column -> contiguous int64 / float64 as the kernels read it."""
    a = np.ascontiguousarray(values)
    if a.dtype.kind in "iub":
        return a.astype(np.int64, copy=False)
    if a.dtype.kind == "f":
        return a.astype(np.float64, copy=False)
    raise TypeError("groupby: unsupported value column dtype %s" % a.dtype)


def groupby_host(keys, columns, aggs, sort=True, ddof=1, expected_groups=0):
    """NumPy in / NumPy out: (result_keys, {agg: {name: array}}).  String keys go through str_arr."""
    torch = _torch()
    names = list(columns)
    cols = [torch.from_numpy(_value_column(columns[c])).cuda() for c in names]
    k = np.asarray(keys)
    if k.dtype.kind in "OUS":
        from . import str_arr
        rkeys, res = str_arr.groupby_str_device(str_arr.StringArray.from_pylist(list(keys)).cuda(), cols, aggs, sort, ddof)
        rk = rkeys.to_pylist()
    else:
        if k.dtype.kind in "iub":
            kk = np.ascontiguousarray(k).astype(np.int64, copy=False)
        elif k.dtype.kind == "f":
            kk = np.ascontiguousarray(k).astype(np.float64, copy=False)
        else:
            raise TypeError("groupby: unsupported key dtype %s" % k.dtype)
        dk, res = groupby_device(torch.from_numpy(kk).cuda(), cols, aggs, sort, ddof, expected_groups)
        rk = dk.cpu().numpy().astype(k.dtype if k.dtype.kind != "b" else np.int64, copy=False)
    out = {a: {names[i]: t.cpu().numpy() for i, t in enumerate(ts)} for a, ts in res.items()}
    return rk, out


class _GroupByBase:
    def __init__(self, sort):
        self._sort = bool(sort)

    def _check_ddof(self, name, ddof):
        if not isinstance(ddof, (int, np.integer)) or isinstance(ddof, bool):
            raise TypeError("Method GroupBy.{}(). The object ddof\n given: {}\n expected: int".format(
                name, type(ddof).__name__))

    def count(self):
        return self._apply("count")

    def max(self):
        return self._apply("max")

    def mean(self, *args):
        return self._apply("mean")

    def median(self):
        return self._apply("median")

    def min(self):
        return self._apply("min")

    def prod(self):
        return self._apply("prod")

    def sum(self):
        return self._apply("sum")

    def std(self, ddof=1, *args):
        self._check_ddof("std", ddof)
        return self._apply("std", int(ddof))

    def var(self, ddof=1, *args):
        self._check_ddof("var", ddof)
        return self._apply("var", int(ddof))

    def agg(self, funcs):
        """Several aggregates in ONE pass over the rows (C4: sum/mean/min/max/count)."""
        return self._apply_multi(tuple(funcs))


class DataFrameGroupBy(_GroupByBase):
    """`df.groupby('A')` -- lazy, like the reference's DataFrameGroupByType (groupby_types.py:34-115)."""

    def __init__(self, df, by, sort=True, columns=None, selected=False):
        super().__init__(sort)
        if not isinstance(by, str) or by not in df.columns:
            raise TypeError("Method groupby(). The object by\n expected: one column name")   # :3066-3067
        self._df, self._by = df, by
        self._columns = [c for c in df.columns if c != by] if columns is None else list(columns)
        self._selected = selected

    def __getitem__(self, item):
        if self._selected:
            raise IndexError("DataFrame.GroupBy.getitem(): Columns already selected")      # :169-170
        if isinstance(item, str):
            if item not in self._df.columns:
                raise KeyError(item)
            return SeriesGroupBy(self._df[item], self._df[self._by].to_numpy(), self._sort, key_name=self._by)
        cols = list(item)
        for c in cols:
            if c not in self._df.columns:
                raise KeyError(c)
        return DataFrameGroupBy(self._df, self._by, self._sort, columns=cols, selected=True)

    def _run(self, aggs, ddof=1):
        import pandas as pd
        keys = self._df[self._by]
        kv = keys.to_numpy()
        if kv.dtype.kind == "O":
            kv = list(kv)
        cols = {c: self._df[c].to_numpy() for c in self._columns}
        rk, res = groupby_host(kv, cols, aggs, self._sort, ddof)
        return pd.Index(rk, name=self._by), res

    def _apply(self, agg, ddof=1):
        import pandas as pd
        index, res = self._run((agg,), ddof)
        return pd.DataFrame({c: res[agg][c] for c in self._columns}, index=index, columns=self._columns)

    def _apply_multi(self, aggs):
        import pandas as pd
        index, res = self._run(aggs)
        data = {(c, a): res[a][c] for c in self._columns for a in aggs}
        return pd.DataFrame(data, index=index, columns=pd.MultiIndex.from_tuples(list(data)))


class SeriesGroupBy(_GroupByBase):
    """`S.groupby(arr)` / `df.groupby('A')['B']` (series_functions.py:4720-4811)."""

    def __init__(self, series, by, sort=True, key_name=None):
        super().__init__(sort)
        by_arr = by if isinstance(by, list) else np.asarray(by)
        if len(by_arr) != len(series):
            raise ValueError("Method Series.groupby(). The object by\n expected: array of the same length")  # :4797-4798
        self._s, self._byv, self._key_name = series, by_arr, key_name

    def _apply(self, agg, ddof=1):
        import pandas as pd
        kv = self._byv
        if not isinstance(kv, list) and kv.dtype.kind == "O":
            kv = list(kv)
        rk, res = groupby_host(kv, {"v": self._s.to_numpy()}, (agg,), self._sort, ddof)
        return pd.Series(res[agg]["v"], index=pd.Index(rk, name=self._key_name), name=self._s.name)

    def _apply_multi(self, aggs):
        import pandas as pd
        kv = self._byv
        if not isinstance(kv, list) and kv.dtype.kind == "O":
            kv = list(kv)
        rk, res = groupby_host(kv, {"v": self._s.to_numpy()}, aggs, self._sort)
        return pd.DataFrame({a: res[a]["v"] for a in aggs}, index=pd.Index(rk, name=self._key_name))


def groupby(obj, by, sort=True):
    """`obj.groupby(by, sort=sort)` for a pandas DataFrame (by = column name) or Series (by = array)."""
    import pandas as pd
    if isinstance(obj, pd.DataFrame):
        return DataFrameGroupBy(obj, by, sort)
    if isinstance(obj, pd.Series):
        return SeriesGroupBy(obj, by, sort)
    raise TypeError("groupby: expected a pandas DataFrame or Series")
