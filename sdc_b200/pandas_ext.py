"""The Pandas surface of the hot path INSIDE nopython code: `s.rolling(w).mean()`, `s.groupby(arr).sum()`,
`df.groupby('A').sum()`, `df.rolling(w).mean()`, `s.sort_values()`, `pd.merge(l, r, on='k', how='inner')` typed for
Numba 0.65 and lowered to the C-ABI host entry points of libsdc_b200.so (numba_ext.py).

What this replaces in the reference (the typing / boxing layer and the overload registrations of the path -- a minimal
version of them: numeric float64 / int64 columns, default or numeric index):
  * Series / DataFrame types + (un)boxing ..... sdc/hiframes/pd_series_type.py, sdc/hiframes/pd_dataframe_type.py:39-54,
                                                sdc/hiframes/boxing.py:89-140 (unbox_dataframe), :232 (box)
  * RollingType + `.rolling(...)` checks ...... sdc/datatypes/hpat_pandas_rolling_types.py:34-176
  * Series.rolling methods .................... sdc/datatypes/hpat_pandas_series_rolling_functions.py:886-1144
  * DataFrame.rolling = per-column Series ..... sdc/datatypes/hpat_pandas_dataframe_rolling_functions.py:167-189
  * GroupBy types + init ...................... sdc/datatypes/hpat_pandas_groupby_types.py:34-115,
                                                sdc/datatypes/hpat_pandas_groupby_functions.py:76-141
  * groupby methods ........................... sdc/datatypes/hpat_pandas_groupby_functions.py:361-600
  * Series.sort_values ........................ sdc/datatypes/hpat_pandas_series_functions.py:3853-3959
  * pd.merge .................................. absent in the reference (sdc/hiframes/join.py:34-43 stub); pandas semantics
                                                for inner / left on one numeric key column
Importing this module registers the types: afterwards `numba.njit` functions accept and return pandas Series /
DataFrames of float64 / int64 columns.  Every method body is one call into the library (no window, hash table or
sort is computed here); a non-zero status raises -- there is no CPU fallback.
"""
import operator

import numpy as np
import pandas as pd
from numba import types
from numba.core import cgutils
from numba.core.errors import TypingError
from numba.extending import (NativeValue, box, intrinsic, make_attribute_wrapper, models, overload, overload_method,
                             register_model, typeof_impl, unbox)

from . import numba_ext as nx

_NUMERIC = (types.float64, types.int64)


def _np_dtype_to_numba(dt):
    if dt == np.float64:
        return types.float64
    if dt == np.int64:
        return types.int64
    return None


# ================================================================== Series
class SeriesType(types.Type):
    """pandas.Series of float64 / int64 values; index: none (default RangeIndex) or a float64 / int64 array."""

    def __init__(self, dtype, index=types.none):
        self.dtype = dtype
        self.index = index
        self.data = types.Array(dtype, 1, "C")
        super().__init__(name="sdcb200.Series(%s, index=%s)" % (dtype, index))


@register_model(SeriesType)
class SeriesModel(models.StructModel):
    def __init__(self, dmm, fe_type):
        models.StructModel.__init__(self, dmm, fe_type, [("data", fe_type.data), ("index", fe_type.index)])


make_attribute_wrapper(SeriesType, "data", "_data")
make_attribute_wrapper(SeriesType, "index", "_index")


def _index_type_of(idx):
    if isinstance(idx, pd.RangeIndex) and idx.start == 0 and idx.step == 1:
        return types.none
    nt = _np_dtype_to_numba(idx.dtype)
    if nt is None:
        raise TypingError("sdc_b200: index dtype %s is not supported (default, int64 or float64 index)" % idx.dtype)
    return types.Array(nt, 1, "C")


@typeof_impl.register(pd.Series)
def _typeof_series(val, c):
    nt = _np_dtype_to_numba(val.dtype)
    if nt is None:
        raise TypingError("sdc_b200: Series dtype %s is not supported (float64 / int64)" % val.dtype)
    return SeriesType(nt, _index_type_of(val.index))


@intrinsic
def init_series(typingctx, data, index):
    """Series from a values array and an index (None or an array) -- init_series, sdc/hiframes/api.py."""
    if not isinstance(data, types.Array):
        return None
    ret = SeriesType(data.dtype, index if isinstance(index, types.Array) else types.none)

    def codegen(context, builder, sig, args):
        s = cgutils.create_struct_proxy(sig.return_type)(context, builder)
        s.data = args[0]
        s.index = args[1] if isinstance(sig.args[1], types.Array) else context.get_dummy_value()
        context.nrt.incref(builder, sig.args[0], args[0])
        if isinstance(sig.args[1], types.Array):
            context.nrt.incref(builder, sig.args[1], args[1])
        return s._getvalue()
    return ret(data, index), codegen


def _contig(a):
    return np.ascontiguousarray(a)


def _unbox_array(c, obj, attr, arr_type):
    """obj.<attr> -> contiguous ndarray -> native array."""
    vals = c.pyapi.object_getattr_string(obj, attr)
    fn = c.pyapi.unserialize(c.pyapi.serialize_object(_contig))
    arr = c.pyapi.call_function_objargs(fn, (vals,))
    native = c.unbox(arr_type, arr)
    c.pyapi.decref(vals)
    c.pyapi.decref(fn)
    c.pyapi.decref(arr)
    return native.value


@unbox(SeriesType)
def _unbox_series(typ, obj, c):
    s = cgutils.create_struct_proxy(typ)(c.context, c.builder)
    s.data = _unbox_array(c, obj, "values", typ.data)
    if isinstance(typ.index, types.Array):
        idx = c.pyapi.object_getattr_string(obj, "index")
        s.index = _unbox_array(c, idx, "values", typ.index)
        c.pyapi.decref(idx)
    else:
        s.index = c.context.get_dummy_value()
    return NativeValue(s._getvalue(), is_error=cgutils.is_not_null(c.builder, c.pyapi.err_occurred()))


@box(SeriesType)
def _box_series(typ, val, c):
    s = cgutils.create_struct_proxy(typ)(c.context, c.builder, value=val)
    data = c.box(typ.data, s.data)
    cls = c.pyapi.unserialize(c.pyapi.serialize_object(pd.Series))
    if isinstance(typ.index, types.Array):
        index = c.box(typ.index, s.index)
        res = c.pyapi.call_function_objargs(cls, (data, index))
        c.pyapi.decref(index)
    else:
        res = c.pyapi.call_function_objargs(cls, (data,))
    c.pyapi.decref(data)
    c.pyapi.decref(cls)
    return res


@overload_method(SeriesType, "__len__")
def _series_len(self):
    return lambda self: len(self._data)


@overload(len)
def _series_len2(s):
    if isinstance(s, SeriesType):
        return lambda s: len(s._data)


# ================================================================== rolling
class RollingType(types.Type):
    """`obj.rolling(window, min_periods)` of a Series or a DataFrame (RollingType, hpat_pandas_rolling_types.py:34-63)."""

    def __init__(self, obj):
        self.obj = obj
        super().__init__(name="sdcb200.Rolling(%s)" % obj)


@register_model(RollingType)
class RollingModel(models.StructModel):
    def __init__(self, dmm, fe_type):
        models.StructModel.__init__(self, dmm, fe_type, [("obj", fe_type.obj), ("window", types.int64), ("minp", types.int64)])


make_attribute_wrapper(RollingType, "obj", "_obj")
make_attribute_wrapper(RollingType, "window", "_window")
make_attribute_wrapper(RollingType, "minp", "_minp")


@intrinsic
def _init_rolling(typingctx, obj, window, minp):
    ret = RollingType(obj)

    def codegen(context, builder, sig, args):
        r = cgutils.create_struct_proxy(sig.return_type)(context, builder)
        r.obj, r.window, r.minp = args
        context.nrt.incref(builder, sig.args[0], args[0])
        return r._getvalue()
    return ret(obj, types.int64, types.int64), codegen


def _rolling_ctor_overload(self, window, min_periods=None, center=False, win_type=None, on=None, axis=0, closed=None):
    # typing errors and value errors with the reference's messages (hpat_pandas_rolling_types.py:108-170)
    if not isinstance(window, types.Integer):
        raise TypingError("Method rolling(). The object window\n given: %s\n expected: int" % window)
    if not isinstance(min_periods, (types.Integer, types.NoneType, types.Omitted)) and min_periods is not None:
        raise TypingError("Method rolling(). The object min_periods\n given: %s\n expected: None, int" % min_periods)
    minp_none = min_periods is None or isinstance(min_periods, (types.NoneType, types.Omitted))

    def impl(self, window, min_periods=None, center=False, win_type=None, on=None, axis=0, closed=None):
        if window < 0:
            raise ValueError('window must be non-negative')
        if minp_none:
            minp = window
        else:
            minp = min_periods
        if minp < 0:
            raise ValueError('min_periods must be >= 0')
        if minp > window:
            raise ValueError('min_periods must be <= window')
        if center != False:  # noqa: E712  (the reference's own comparison)
            raise ValueError('Method rolling(). The object center\n expected: False')
        if win_type is not None:
            raise ValueError('Method rolling(). The object win_type\n expected: None')
        if on is not None:
            raise ValueError('Method rolling(). The object on\n expected: None')
        if axis != 0:
            raise ValueError('Method rolling(). The object axis\n expected: 0')
        if closed is not None:
            raise ValueError('Method rolling(). The object closed\n expected: None')
        return _init_rolling(self, window, minp)
    return impl


overload_method(SeriesType, "rolling")(_rolling_ctor_overload)

_ROLL_FNS = {"sum": nx.rolling_sum, "mean": nx.rolling_mean, "count": nx.rolling_count, "min": nx.rolling_min,
             "max": nx.rolling_max, "var": nx.rolling_var, "std": nx.rolling_std}


def _register_rolling_method(name, with_ddof):
    fn = _ROLL_FNS[name]

    def series_impl_factory():
        if with_ddof:
            def impl(self, ddof=1):
                out = fn(self._obj._data, self._window, self._minp, ddof)
                return init_series(out, self._obj._index)
        else:
            def impl(self):
                out = fn(self._obj._data, self._window, self._minp)
                return init_series(out, self._obj._index)
        return impl

    if with_ddof:
        def ovld(self, ddof=1):
            if isinstance(self.obj, SeriesType):
                if not isinstance(ddof, (types.Integer, types.Omitted, int)):
                    raise TypingError("Method rolling.%s(). The object ddof\n given: %s\n expected: int" % (name, ddof))
                return series_impl_factory()
            if isinstance(self.obj, DataFrameType):
                return _df_rolling_impl(self.obj, name, True)
    else:
        def ovld(self):
            if isinstance(self.obj, SeriesType):
                return series_impl_factory()
            if isinstance(self.obj, DataFrameType):
                return _df_rolling_impl(self.obj, name, False)
    overload_method(RollingType, name)(ovld)


# ================================================================== DataFrame
class DataFrameType(types.Type):
    """pandas.DataFrame of float64 / int64 columns (DataFrameType, sdc/hiframes/pd_dataframe_type.py:39-54): the column
    names are part of the type, the columns a tuple of arrays."""

    def __init__(self, columns, dtypes, index=types.none):
        self.columns = tuple(columns)
        self.dtypes = tuple(dtypes)
        self.index = index
        self.data = types.Tuple([types.Array(d, 1, "C") for d in self.dtypes])
        super().__init__(name="sdcb200.DataFrame(%s, %s, index=%s)" % (self.columns, self.dtypes, index))


@register_model(DataFrameType)
class DataFrameModel(models.StructModel):
    def __init__(self, dmm, fe_type):
        models.StructModel.__init__(self, dmm, fe_type, [("data", fe_type.data), ("index", fe_type.index)])


make_attribute_wrapper(DataFrameType, "data", "_data")
make_attribute_wrapper(DataFrameType, "index", "_index")


@typeof_impl.register(pd.DataFrame)
def _typeof_df(val, c):
    cols, dts = [], []
    for name in val.columns:
        if not isinstance(name, str):
            raise TypingError("sdc_b200: DataFrame column names must be strings")
        nt = _np_dtype_to_numba(val[name].dtype)
        if nt is None:
            raise TypingError("sdc_b200: column %r has dtype %s (float64 / int64 are supported)" % (name, val[name].dtype))
        cols.append(name)
        dts.append(nt)
    return DataFrameType(cols, dts, _index_type_of(val.index))


@unbox(DataFrameType)
def _unbox_df(typ, obj, c):
    df = cgutils.create_struct_proxy(typ)(c.context, c.builder)
    arrs = []
    for name, dt in zip(typ.columns, typ.dtypes):
        key = c.pyapi.unserialize(c.pyapi.serialize_object(name))
        col = c.pyapi.object_getitem(obj, key)
        arrs.append(_unbox_array(c, col, "values", types.Array(dt, 1, "C")))
        c.pyapi.decref(key)
        c.pyapi.decref(col)
    df.data = c.context.make_tuple(c.builder, typ.data, arrs)
    if isinstance(typ.index, types.Array):
        idx = c.pyapi.object_getattr_string(obj, "index")
        df.index = _unbox_array(c, idx, "values", typ.index)
        c.pyapi.decref(idx)
    else:
        df.index = c.context.get_dummy_value()
    return NativeValue(df._getvalue(), is_error=cgutils.is_not_null(c.builder, c.pyapi.err_occurred()))


def _make_frame(columns, arrays, index):
    d = dict(zip(columns, arrays))
    return pd.DataFrame(d, index=index, columns=list(columns))


@box(DataFrameType)
def _box_df(typ, val, c):
    df = cgutils.create_struct_proxy(typ)(c.context, c.builder, value=val)
    arrs = []
    for i, dt in enumerate(typ.dtypes):
        a = c.builder.extract_value(df.data, i)
        c.context.nrt.incref(c.builder, types.Array(dt, 1, "C"), a)
        arrs.append(c.box(types.Array(dt, 1, "C"), a))
    tup = c.pyapi.tuple_pack(arrs)
    cols = c.pyapi.unserialize(c.pyapi.serialize_object(typ.columns))
    if isinstance(typ.index, types.Array):
        c.context.nrt.incref(c.builder, typ.index, df.index)
        index = c.box(typ.index, df.index)
    else:
        index = c.pyapi.make_none()
    fn = c.pyapi.unserialize(c.pyapi.serialize_object(_make_frame))
    res = c.pyapi.call_function_objargs(fn, (cols, tup, index))
    for a in arrs:
        c.pyapi.decref(a)
    for o in (tup, cols, index, fn):
        c.pyapi.decref(o)
    c.context.nrt.decref(c.builder, typ, val)
    return res


def _make_init_dataframe(columns, dtypes, index_is_array):
    """An intrinsic building DataFrameType(columns, dtypes) from that many arrays (+ an index array)."""
    n = len(columns)

    @intrinsic
    def init_df(typingctx, data, index):
        idx_t = index if isinstance(index, types.Array) else types.none
        ret = DataFrameType(columns, dtypes, idx_t)

        def codegen(context, builder, sig, args):
            df = cgutils.create_struct_proxy(sig.return_type)(context, builder)
            df.data = args[0]
            df.index = args[1] if isinstance(sig.args[1], types.Array) else context.get_dummy_value()
            context.nrt.incref(builder, sig.args[0], args[0])
            if isinstance(sig.args[1], types.Array):
                context.nrt.incref(builder, sig.args[1], args[1])
            return df._getvalue()
        return ret(data, index), codegen
    assert n >= 0
    return init_df


def _exec_impl(text, glbs, name="impl"):
    loc = {}
    exec(text, glbs, loc)
    return loc[name]


overload_method(DataFrameType, "rolling")(_rolling_ctor_overload)


def _df_rolling_impl(df_t, name, with_ddof):
    """DataFrame.rolling(...).<op>(): the Series method per column, default index kept
    (df_rolling_method_main_codegen, hpat_pandas_dataframe_rolling_functions.py:167-189)."""
    n = len(df_t.columns)
    init_df = _make_init_dataframe(df_t.columns, (types.float64,) * n, isinstance(df_t.index, types.Array))
    args = "self, ddof=1" if with_ddof else "self"
    extra = ", ddof" if with_ddof else ""
    lines = ["def impl(%s):" % args]
    for i in range(n):
        lines.append("    c%d = fn(self._obj._data[%d], self._window, self._minp%s)" % (i, i, extra))
    lines.append("    return init_df((%s), self._obj._index)" % "".join("c%d, " % i for i in range(n)))
    return _exec_impl("\n".join(lines), {"fn": _ROLL_FNS[name], "init_df": init_df})


for _name in ("sum", "mean", "count", "min", "max"):
    _register_rolling_method(_name, False)
for _name in ("var", "std"):
    _register_rolling_method(_name, True)


# ================================================================== groupby
class SeriesGroupByType(types.Type):
    """Series.groupby(by): (values, by, sort) -- SeriesGroupByType, hpat_pandas_groupby_types.py:75-115."""

    def __init__(self, series, by):
        self.series = series
        self.by = by
        super().__init__(name="sdcb200.SeriesGroupBy(%s, by=%s)" % (series, by))


@register_model(SeriesGroupByType)
class SeriesGroupByModel(models.StructModel):
    def __init__(self, dmm, fe_type):
        models.StructModel.__init__(self, dmm, fe_type, [("series", fe_type.series), ("by", fe_type.by), ("sort", types.boolean)])


make_attribute_wrapper(SeriesGroupByType, "series", "_series")
make_attribute_wrapper(SeriesGroupByType, "by", "_by")
make_attribute_wrapper(SeriesGroupByType, "sort", "_sort")


class DataFrameGroupByType(types.Type):
    """df.groupby('A'): (frame, key column position, sort) -- DataFrameGroupByType, hpat_pandas_groupby_types.py:34-72.
    `selected`: positions of the value columns (all but the key unless `gb['B']` / `gb['B', 'C']` picked some)."""

    def __init__(self, frame, key, selected=None, as_series=False):
        self.frame = frame
        self.key_col = key
        self.selected = tuple(selected) if selected is not None else tuple(i for i in range(len(frame.columns)) if i != key)
        self.as_series = as_series
        self.explicit = selected is not None
        super().__init__(name="sdcb200.DataFrameGroupBy(%s, key=%d, cols=%s, series=%s)" % (frame, key, self.selected, as_series))


@register_model(DataFrameGroupByType)
class DataFrameGroupByModel(models.StructModel):
    def __init__(self, dmm, fe_type):
        models.StructModel.__init__(self, dmm, fe_type, [("frame", fe_type.frame), ("sort", types.boolean)])


make_attribute_wrapper(DataFrameGroupByType, "frame", "_frame")
make_attribute_wrapper(DataFrameGroupByType, "sort", "_sort")


@intrinsic
def _init_series_groupby(typingctx, series, by, sort):
    ret = SeriesGroupByType(series, by)

    def codegen(context, builder, sig, args):
        g = cgutils.create_struct_proxy(sig.return_type)(context, builder)
        g.series, g.by, g.sort = args
        context.nrt.incref(builder, sig.args[0], args[0])
        context.nrt.incref(builder, sig.args[1], args[1])
        return g._getvalue()
    return ret(series, by, types.boolean), codegen


def _make_init_df_groupby(key, selected=None, as_series=False):
    @intrinsic
    def init_gb(typingctx, frame, sort):
        ret = DataFrameGroupByType(frame, key, selected, as_series)

        def codegen(context, builder, sig, args):
            g = cgutils.create_struct_proxy(sig.return_type)(context, builder)
            g.frame, g.sort = args
            context.nrt.incref(builder, sig.args[0], args[0])
            return g._getvalue()
        return ret(frame, types.boolean), codegen
    return init_gb


@overload_method(SeriesType, "groupby")
def _series_groupby(self, by=None, axis=0, level=None, as_index=True, sort=True, group_keys=True, squeeze=False, observed=False):
    # sdc_pandas_series_groupby, hpat_pandas_series_functions.py:4720-4811: `by` is an array of the same length
    if isinstance(by, SeriesType):
        def impl_s(self, by=None, axis=0, level=None, as_index=True, sort=True, group_keys=True, squeeze=False, observed=False):
            if len(by._data) != len(self._data):
                raise ValueError('Method Series.groupby(). The object by\n expected: array of the same length')
            return _init_series_groupby(self, by._data, sort)
        return impl_s
    if not (isinstance(by, types.Array) and by.ndim == 1 and by.dtype in _NUMERIC):
        raise TypingError("Method Series.groupby(). The object by\n given: %s\n expected: 1-D int64 / float64 array" % (by,))

    def impl(self, by=None, axis=0, level=None, as_index=True, sort=True, group_keys=True, squeeze=False, observed=False):
        if len(by) != len(self._data):
            raise ValueError('Method Series.groupby(). The object by\n expected: array of the same length')
        return _init_series_groupby(self, np.ascontiguousarray(by), sort)
    return impl


@overload_method(DataFrameType, "groupby", prefer_literal=True)
def _df_groupby(self, by=None, axis=0, level=None, as_index=True, sort=True, group_keys=True, squeeze=False, observed=False):
    # sdc_pandas_dataframe_groupby, hpat_pandas_dataframe_functions.py:2993-3106: `by` is ONE literal column name
    if not isinstance(by, types.StringLiteral):
        raise TypingError("Method DataFrame.groupby(). The object by\n given: %s\n expected: a literal column name" % (by,))
    name = by.literal_value
    if name not in self.columns:
        raise TypingError("Method DataFrame.groupby(). The object by\n column %r is not in the frame" % name)
    init_gb = _make_init_df_groupby(self.columns.index(name))

    def impl(self, by=None, axis=0, level=None, as_index=True, sort=True, group_keys=True, squeeze=False, observed=False):
        return init_gb(self, sort)
    return impl


@overload(operator.getitem, prefer_literal=True)
def _df_groupby_getitem(gb, item):
    """gb['B'] -> one column (a Series result), gb['B', 'C'] -> a column subset; a second [] raises IndexError
    (hpat_pandas_groupby_functions.py:144-191)."""
    if not isinstance(gb, DataFrameGroupByType):
        return None
    if gb.explicit:
        def impl_err(gb, item):
            raise IndexError("DataFrame.GroupBy.getitem: Columns already selected")
        return impl_err
    if isinstance(item, types.StringLiteral):
        names, as_series = (item.literal_value,), True
    elif isinstance(item, types.BaseTuple) and all(isinstance(t, types.StringLiteral) for t in item.types):
        names, as_series = tuple(t.literal_value for t in item.types), False
    else:
        raise TypingError("DataFrame.GroupBy.getitem: literal column name(s) expected, given %s" % (item,))
    for nme in names:
        if nme not in gb.frame.columns:
            raise TypingError("DataFrame.GroupBy.getitem: column %r is not in the frame" % nme)
    init_gb = _make_init_df_groupby(gb.key_col, tuple(gb.frame.columns.index(nme) for nme in names), as_series)

    def impl(gb, item):
        return init_gb(gb._frame, gb._sort)
    return impl


_GB_FNS = {"sum": nx.groupby_sum, "mean": nx.groupby_mean, "min": nx.groupby_min, "max": nx.groupby_max,
           "count": nx.groupby_count, "var": nx.groupby_var, "std": nx.groupby_std}


def _gb_out_dtype(agg, col_dtype):
    if agg == "count":
        return types.int64
    if col_dtype == types.int64 and agg in ("sum", "min", "max"):
        return types.int64
    return types.float64


def _register_groupby_method(agg, with_ddof):
    fn = _GB_FNS[agg]

    def series_ovld_impl():
        if with_ddof:
            def impl(self, ddof=1):
                k, v = fn(self._by, self._series._data, self._sort, ddof)
                return init_series(v, k)
        else:
            def impl(self):
                k, v = fn(self._by, self._series._data, self._sort)
                return init_series(v, k)
        return impl

    def frame_impl(gb_t):
        df_t = gb_t.frame
        cols = [df_t.columns[i] for i in gb_t.selected]
        out_dts = tuple(_gb_out_dtype(agg, df_t.dtypes[i]) for i in gb_t.selected)
        args = "self, ddof=1" if with_ddof else "self"
        extra = ", ddof" if with_ddof else ""
        lines = ["def impl(%s):" % args, "    key = self._frame._data[%d]" % gb_t.key_col]
        for j, i in enumerate(gb_t.selected):
            lines.append("    k%d, v%d = fn(key, self._frame._data[%d], self._sort%s)" % (j, j, i, extra))
        if not gb_t.selected:
            raise TypingError("DataFrame.groupby: no value columns")
        if gb_t.as_series:
            lines.append("    return init_series(v0, k0)")
            return _exec_impl("\n".join(lines), {"fn": fn, "init_series": init_series})
        init_df = _make_init_dataframe(tuple(cols), out_dts, True)
        lines.append("    return init_df((%s), k0)" % "".join("v%d, " % j for j in range(len(cols))))
        return _exec_impl("\n".join(lines), {"fn": fn, "init_df": init_df})

    if with_ddof:
        def s_ovld(self, ddof=1):
            return series_ovld_impl()

        def d_ovld(self, ddof=1):
            return frame_impl(self)
    else:
        def s_ovld(self):
            return series_ovld_impl()

        def d_ovld(self):
            return frame_impl(self)
    overload_method(SeriesGroupByType, agg)(s_ovld)
    overload_method(DataFrameGroupByType, agg)(d_ovld)


for _agg in ("sum", "mean", "min", "max", "count"):
    _register_groupby_method(_agg, False)
for _agg in ("var", "std"):
    _register_groupby_method(_agg, True)


# ================================================================== sort_values
@overload_method(SeriesType, "sort_values")
def _series_sort_values(self, axis=0, ascending=True, inplace=False, kind="quicksort", na_position="last"):
    """hpat_pandas_series_sort_values, hpat_pandas_series_functions.py:3853-3959: NaNs split off, stable argsort of the
    rest through the library, NaN block placed per na_position, values and index gathered."""
    has_index = isinstance(self.index, types.Array)
    is_float = self.dtype == types.float64

    def impl(self, axis=0, ascending=True, inplace=False, kind="quicksort", na_position="last"):
        if inplace:
            raise ValueError("Method sort_values(). The object inplace\n expected: False")
        if kind != "quicksort" and kind != "mergesort":
            raise ValueError("Method sort_values(). Unsupported parameter. Given kind != 'quicksort' or 'mergesort'")
        if na_position != "last" and na_position != "first":
            raise ValueError("Method sort_values(). Unsupported parameter. Given na_position != 'last' or 'first'")
        data = self._data
        n = len(data)
        order = nx.parallel_stable_argsort(data, ascending)          # NaN last in both directions (native/utils.cpp:160-182)
        if is_float and na_position == "first":
            nn = 0
            for i in range(n):
                if data[i] != data[i]:
                    nn += 1
            if nn > 0:
                rot = np.empty(n, np.int64)
                rot[:nn] = order[n - nn:]
                rot[nn:] = order[:n - nn]
                order = rot
        vals = data[order]
        if has_index:
            return init_series(vals, self._index[order])
        return init_series(vals, order)
    return impl


# ================================================================== pd.merge
@overload(pd.merge, prefer_literal=True)
def _pd_merge(left, right, how="inner", on=None):
    """pd.merge(left, right, on='k', how='inner' | 'left') for frames of numeric columns: the joined rows' indexers
    come from the library (sdcb200_host_join); unmatched right rows of a left join become NaN (setitem_arr_nan,
    sdc/hiframes/join.py:34-43).  The key column comes first, then left's and right's other columns -- pandas' order."""
    if not (isinstance(left, DataFrameType) and isinstance(right, DataFrameType)):
        return None
    if not isinstance(on, types.StringLiteral):
        raise TypingError("pd.merge: `on` must be a literal column name")
    if isinstance(how, types.StringLiteral):
        how_v = how.literal_value
    elif isinstance(how, (str, types.Omitted)):
        how_v = how if isinstance(how, str) else how.value
    else:
        raise TypingError("pd.merge: `how` must be a literal string")
    if how_v not in ("inner", "left"):
        raise TypingError("pd.merge: how=%r is not supported in nopython code (inner / left)" % (how_v,))
    key = on.literal_value
    if key not in left.columns or key not in right.columns:
        raise TypingError("pd.merge: column %r must be in both frames" % key)
    lk, rk = left.columns.index(key), right.columns.index(key)
    if left.dtypes[lk] != right.dtypes[rk]:
        raise TypingError("pd.merge: key dtypes differ")
    lcols = [i for i in range(len(left.columns)) if i != lk]
    rcols = [i for i in range(len(right.columns)) if i != rk]
    names = [left.columns[i] for i in lcols] + [right.columns[i] for i in rcols]
    if len(set(names)) != len(names) or key in names:
        raise TypingError("pd.merge: overlapping column names are not supported in nopython code")
    left_join = how_v == "left"
    out_cols = (key,) + tuple(names)
    out_dts = (left.dtypes[lk],) + tuple(left.dtypes[i] for i in lcols) + tuple(types.float64 if left_join else right.dtypes[i] for i in rcols)
    init_df = _make_init_dataframe(out_cols, out_dts, False)
    lines = ["def impl(left, right, how='inner', on=None):",
             "    li, ri = merge_indexers(left._data[%d], right._data[%d], %r)" % (lk, rk, how_v),
             "    c0 = left._data[%d][li]" % lk]
    outs = ["c0"]
    for j, i in enumerate(lcols):
        lines.append("    l%d = left._data[%d][li]" % (j, i))
        outs.append("l%d" % j)
    for j, i in enumerate(rcols):
        if left_join:
            lines.append("    r%d = take_nan(right._data[%d], ri)" % (j, i))
        else:
            lines.append("    r%d = right._data[%d][ri]" % (j, i))
        outs.append("r%d" % j)
    lines.append("    return init_df((%s), None)" % "".join(o + ", " for o in outs))
    return _exec_impl("\n".join(lines), {"merge_indexers": nx.merge_indexers, "init_df": init_df, "take_nan": _take_nan})


def _take_nan(col, idx):
    raise NotImplementedError("call from nopython code")


@overload(_take_nan)
def _take_nan_ovld(col, idx):
    def impl(col, idx):
        out = np.empty(len(idx), np.float64)
        for i in range(len(idx)):
            j = idx[i]
            out[i] = np.nan if j < 0 else col[j]
        return out
    return impl
