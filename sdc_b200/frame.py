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


def _as_key(col):
    """numeric key column as the kernels read it (int64 / float64, contiguous)"""
    torch = _torch()
    if _is_str(col):
        raise TypeError("multi-key groupby / merge: key dtype str is not supported (int64 / float64)")
    if col.dtype not in (torch.int64, torch.float64):
        col = col.to(torch.float64 if col.dtype.is_floating_point else torch.int64)
    return col.contiguous()


def _dense_ids(col, domain=None):
    """Order-preserving dense ids of a key column, from the library's own kernels: the sorted distinct keys of
    `domain` (default: the column) come from a value-less groupby, a row's id is its key's position among them --
    the right indexer of a left join against that unique list (-1: NaN key, or a key `domain` does not hold).
    -> (distinct keys, int64 ids)"""
    from .groupby import groupby_device
    from .join import join_device
    uniq, _ = groupby_device(col if domain is None else domain, [], ("count",), True)
    _, ids = join_device(col, uniq, "left")
    return uniq, ids


def _combine_ids(a, b, nb, invalid=-1):
    """a * nb + b per row, `invalid` where either id is negative (csrc/groupby.cu: sdcb200_combine_ids)"""
    torch = _torch()
    lib = native.load()
    out = torch.empty_like(a)
    with torch.cuda.device(a.device):
        native.check(lib.sdcb200_combine_ids(a.data_ptr(), b.data_ptr(), int(nb), int(invalid), out.data_ptr(), a.shape[0],
                                             native.current_stream_ptr()))
    return out


def _composite_ids(cols, domains=None, invalid=-1):
    """Several key columns -> one int64 id column whose order is the lexicographic order of the key tuples.
    -> (ids, [distinct keys per column])"""
    uniqs, gid, total = [], None, 1
    for j, c in enumerate(cols):
        c = _as_key(c)
        d = None if domains is None else _as_key(domains[j])
        if d is not None and d.dtype != c.dtype:
            torch = _torch()
            c, d = c.to(torch.float64), d.to(torch.float64)
        u, ids = _dense_ids(c, d)
        nu = max(int(u.shape[0]), 1)
        total *= nu
        if total >= 1 << 62:
            raise OverflowError("multi-key: the product of the distinct key counts does not fit an int64 id")
        gid = _combine_ids(ids, ids, 0, invalid) if gid is None else _combine_ids(gid, ids, nu, invalid)
        uniqs.append(u)
    return gid, uniqs


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

    def _apply_moment(self, op, other=None, ddof=1):
        """skew / kurt / cov / corr (csrc/rolling_moments.cu); `other`: a DeviceSeries (every column against it), a
        DeviceFrame (column by column, missing columns skipped) or None (each column against itself)."""
        from .rolling import rolling_moments_device
        torch = _torch()
        out = {}
        for name, col in self._cols.items():
            if _is_str(col):
                continue
            oth = None
            if isinstance(other, DeviceFrame):
                if name not in other._cols:
                    continue
                oth = other._cols[name]
            elif isinstance(other, DeviceSeries):
                oth = other.data
            elif other is not None:
                oth = other
            x = col.to(torch.float64).contiguous()
            y = None if oth is None else oth.to(torch.float64).contiguous()
            out[name] = rolling_moments_device(x, op, self._w, self._minp, y, ddof)
        if self._series:
            (name, data), = out.items()
            return DeviceSeries(data, name)
        return DeviceFrame(out)

    def skew(self): return self._apply_moment("skew")
    def kurt(self): return self._apply_moment("kurt")
    def cov(self, other=None, pairwise=None, ddof=1): return self._apply_moment("cov", other, int(ddof))
    def corr(self, other=None, pairwise=None): return self._apply_moment("corr", other)


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


class _DeviceGroupByMulti(_DeviceGroupBy):
    """groupby([k1, k2, ...]) (the reference accepts ONE key column,
    hpat_pandas_dataframe_functions.py:3066-3067): the key tuples become one dense id column (`_composite_ids`),
    the aggregate runs on the ids (SDCB200_GID while the id space fits a direct table, plain int64 keys beyond),
    and the group ids are decoded back into one key column each.  sort=True orders the groups lexicographically,
    rows with a NaN in any key are dropped (pandas dropna=True)."""
    _GID_LIMIT = 1 << 24

    def __init__(self, cols, key_cols, key_names, sort):
        self._cols, self._key_cols, self._key_names, self._sort = cols, key_cols, list(key_names), bool(sort)

    def agg(self, funcs, ddof=1):
        from .groupby import groupby_device
        torch = _torch()
        funcs = (funcs,) if isinstance(funcs, str) else tuple(funcs)
        names = [n for n, c in self._cols.items() if not _is_str(c)]
        cols = [_as_key(self._cols[n]) for n in names]
        gid, uniqs = _composite_ids(self._key_cols)
        sizes = [max(int(u.shape[0]), 1) for u in uniqs]
        total = 1
        for x in sizes:
            total *= x
        if total <= self._GID_LIMIT:
            rk, res = groupby_device(gid, cols, funcs, self._sort, ddof, gid_range=total)
        else:
            rk, res = groupby_device(gid, cols, funcs, self._sort, ddof)
            keep = rk >= 0                                   # the rows with a NaN key collapsed into group -1
            if not bool(keep.all()):
                rk = rk[keep]
                res = {a: [t[keep] for t in ts] for a, ts in res.items()}
        out, rem = {}, rk
        for name, u, size in reversed(list(zip(self._key_names, uniqs, sizes))):
            out[name] = u[rem % size] if u.shape[0] else u
            rem = rem // size
        out = {n: out[n] for n in self._key_names}
        for a in funcs:
            for n, t in zip(names, res[a]):
                out[n if len(funcs) == 1 else "%s_%s" % (n, a)] = t
        return DeviceFrame(out, index=list(self._key_names))


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

    # ---- columnar I/O: Arrow tables become device columns without a pandas detour ----
    @classmethod
    def from_arrow(cls, table):
        """pyarrow.Table -> device columns.  Numeric columns are copied from their Arrow buffers (nulls become NaN,
        which makes an integer column float64 -- pandas' rule); string columns keep Arrow's offsets / data / validity
        buffers, which ARE the str_arr layout (`StringArray.from_arrow`); bool columns become int64.  The reference
        reads CSV through Arrow as well (sdc/native/arrow_reader.cpp, sdc/io/csv_ext.py) but boxes every column into
        pandas / NumPy first."""
        import pyarrow as pa
        from .str_arr import StringArray
        torch = _torch()
        cols = {}
        for name, col in zip(table.column_names, table.columns):
            t = col.type
            if pa.types.is_dictionary(t):
                col = col.cast(t.value_type)
                t = col.type
            if pa.types.is_string(t) or pa.types.is_large_string(t):
                cols[name] = StringArray.from_arrow(col).cuda()
                continue
            if pa.types.is_boolean(t):
                col, t = col.cast(pa.int64()), pa.int64()
            if not (pa.types.is_integer(t) or pa.types.is_floating(t)):
                raise TypeError("DeviceFrame.from_arrow: unsupported column type %s (%s)" % (t, name))
            if col.null_count:
                col = col.cast(pa.float64())
                a = col.to_numpy(zero_copy_only=False) if not isinstance(col, pa.ChunkedArray) else col.combine_chunks().to_numpy(zero_copy_only=False)
            else:
                a = (col.combine_chunks() if isinstance(col, pa.ChunkedArray) else col).to_numpy(zero_copy_only=False)
            a = np.ascontiguousarray(a)
            if a.dtype.kind in "iu" and a.dtype != np.int64:
                a = a.astype(np.int64)
            elif a.dtype.kind == "f" and a.dtype != np.float64:
                a = a.astype(np.float64)
            cols[name] = torch.from_numpy(a.copy() if not a.flags.writeable else a).cuda()
        return cls(cols)

    def to_arrow(self):
        import pyarrow as pa
        arrays, names = [], []
        for n, c in self._cols.items():
            arrays.append(c.to_arrow() if _is_str(c) else pa.array(c.cpu().numpy()))
            names.append(n)
        return pa.table(arrays, names=names)

    @classmethod
    def read_csv(cls, path, read_options=None, parse_options=None, convert_options=None, engine="auto", sep=",",
                 delimiter=None, names=None, usecols=None, dtype=None, skiprows=None):
        """CSV file -> device columns (what `pd.read_csv` is lowered to in the reference: sdc/io/csv_ext.py +
        native/arrow_reader.cpp; census_benchmark.py:84-85 starts this way).
        engine="device": the file's bytes are copied to the GPU and decoded there (sdc_b200/csv.py, csrc/csv.cu;
        numeric columns, bit-identical to Arrow's conversion); engine="arrow": Arrow's multithreaded host reader, then
        one copy per column -- string columns keep Arrow's buffers, which are the str_arr layout; engine="auto": the
        device reader unless Arrow options are passed or the file turns out to hold a non-numeric column.  Like pandas,
        an empty field of a string column is a missing value unless `convert_options` says otherwise."""
        import pyarrow.csv as pacsv
        arrow_opts = read_options is not None or parse_options is not None or convert_options is not None
        if engine not in ("auto", "device", "arrow"):
            raise ValueError("read_csv: engine must be 'auto', 'device' or 'arrow'")
        if engine == "device" or (engine == "auto" and not arrow_opts):
            from .csv import NotNumericCSV, read_csv_device
            try:
                cols, _ = read_csv_device(path, sep=sep, delimiter=delimiter, names=names, usecols=usecols, dtype=dtype,
                                          skiprows=skiprows)
                return cls(cols)
            except NotNumericCSV:
                if engine == "device":
                    raise
        if not arrow_opts:
            delim = delimiter if delimiter is not None else sep
            read_options = pacsv.ReadOptions(skip_rows=int(skiprows or 0), column_names=list(names) if names is not None else None)
            parse_options = pacsv.ParseOptions(delimiter=delim)
            import pyarrow as pa
            ctypes_ = None
            if isinstance(dtype, dict):
                ctypes_ = {k: pa.from_numpy_dtype(np.dtype(v)) for k, v in dtype.items()}
            convert_options = pacsv.ConvertOptions(strings_can_be_null=True, column_types=ctypes_,
                                                   include_columns=list(usecols) if usecols is not None else None)
        if convert_options is None:
            convert_options = pacsv.ConvertOptions(strings_can_be_null=True)
        return cls.from_arrow(pacsv.read_csv(path, read_options=read_options, parse_options=parse_options,
                                             convert_options=convert_options))

    @classmethod
    def read_parquet(cls, path, columns=None):
        import pyarrow.parquet as pq
        return cls.from_arrow(pq.read_table(path, columns=columns))

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
        if isinstance(by, (list, tuple)) and len(by) == 1:
            by = by[0]
        if isinstance(by, (list, tuple)):
            for b in by:
                if not isinstance(b, str) or b not in self._cols:
                    raise KeyError(b)
            return _DeviceGroupByMulti({n: c for n, c in self._cols.items() if n not in by}, [self._cols[b] for b in by], by, sort)
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
        """pandas.merge on numeric key columns (one, or a list through composite ids); payload columns gathered on the
        device, `_x` / `_y` suffixes for overlapping names, key columns coalesced for rows without a left partner."""
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
        lkeys = [left_on] if isinstance(left_on, str) else list(left_on)
        rkeys = [right_on] if isinstance(right_on, str) else list(right_on)
        if len(lkeys) != len(rkeys) or not lkeys:
            raise ValueError("len(right_on) must equal len(left_on)")
        if len(lkeys) == 1:
            lk, rk = self._cols[lkeys[0]], right._cols[rkeys[0]]
            if _is_str(lk) or _is_str(rk):
                raise TypeError("merge: key dtype str is not supported (int64 / float64)")
            if lk.dtype != rk.dtype or lk.dtype not in (torch.int64, torch.float64):
                both_int = not lk.dtype.is_floating_point and not rk.dtype.is_floating_point
                lk, rk = lk.to(torch.int64 if both_int else torch.float64), rk.to(torch.int64 if both_int else torch.float64)
            li, ri = join_device(lk.contiguous(), rk.contiguous(), how, ordered=True)      # pandas keeps left-row order
        else:
            # several key columns: both sides get ids over the RIGHT side's distinct keys, column
            # by column; a left key tuple the right side does not hold (or a NaN on either side) gets an id that
            # matches nothing (-1 left, -2 right), then the single-key hash join runs on the composite ids
            rcols = [right._cols[k] for k in rkeys]
            rgid, _ = _composite_ids(rcols, invalid=-2)
            lgid, _ = _composite_ids([self._cols[k] for k in lkeys], domains=rcols, invalid=-1)
            li, ri = join_device(lgid, rgid, how, ordered=True)
        left_null, right_null = how in ("right", "outer"), how in ("left", "outer")
        shared = {l for l, r in zip(lkeys, rkeys) if l == r}           # key columns that appear once in the result
        out = {}
        for n, c in self._cols.items():
            name = "%s_x" % n if (n in right._cols and n not in shared) else n
            if n in shared and left_null:
                # the key of a row without a left partner comes from the right side
                rc = right._cols[n]
                kd = torch.int64 if (c.dtype == torch.int64 and rc.dtype == torch.int64) else torch.float64
                lv = _take(c.to(kd) if c.dtype != kd else c, li.clamp(min=0))
                rv = _take(rc.to(kd) if rc.dtype != kd else rc, ri.clamp(min=0))
                out[name] = torch.where(li < 0, rv, lv)          # plumbing: picks between two gathered columns
                continue
            out[name] = _take(c, li, left_null)
        for n, c in right._cols.items():
            if n in shared:
                continue
            out["%s_y" % n if n in self._cols else n] = _take(c, ri, right_null)
        return DeviceFrame(out)

    def merge_asof(self, right, on):
        """pd.merge_asof(left, right, on=key) with the defaults (direction='backward', exact matches allowed): every
        left row keeps its place and gets the last right row whose key is <= its own (NaN payload when there is none).
        Both frames must be sorted by the key, as pandas requires."""
        from .join import asof_backward_device
        torch = _torch()
        lk, rk = _as_key(self._cols[on]), _as_key(right._cols[on])
        if lk.dtype != rk.dtype:
            lk, rk = lk.to(torch.float64), rk.to(torch.float64)
        if lk.shape[0] > 1 and not bool((lk[1:] >= lk[:-1]).all()):
            raise ValueError("left keys must be sorted")
        if rk.shape[0] > 1 and not bool((rk[1:] >= rk[:-1]).all()):
            raise ValueError("right keys must be sorted")
        ri = asof_backward_device(lk, rk)
        out = {}
        for n, c in self._cols.items():
            out["%s_x" % n if (n in right._cols and n != on) else n] = c
        for n, c in right._cols.items():
            if n != on:
                out["%s_y" % n if n in self._cols else n] = _take(c, ri, True)
        return DeviceFrame(out)
