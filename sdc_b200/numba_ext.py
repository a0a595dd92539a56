"""Numba glue: the C-ABI library called from nopython code, the way SDC binds its natives.

FFI idiom A of the reference (sdc/functions/sort.py:39-72): a raw function address is cast to a ctypes
CFUNCTYPE and the resulting object is called inside `@njit` code with `arr.ctypes` / `len(arr)`.  The
reference takes the addresses from its CPython extension module `sdc.concurrent_sort`
(sdc/native/module.cpp:118-176); here they come from `libsdc_b200.so` through `native.fn_address`.
The functions below have the reference's Python-level names and argument meaning:

    parallel_sort(arr) / parallel_stable_sort(arr)                sdc/functions/sort.py:243-256
    parallel_argsort(arr, ascending) / parallel_stable_argsort    sdc/functions/sort.py:289-310
    rolling_<op>(arr, window, min_periods[, ddof])                gen_sdc_pandas_series_rolling_impl instantiations,
                                                                  sdc/datatypes/hpat_pandas_series_rolling_functions.py:736-754

They run on HOST arrays (what a jitted SDC function holds) through the `sdcb200_host_*` / reference-named
entry points, so the H2D / D2H copies are inside the call.  A non-zero status raises in the jitted code:
there is no CPU fallback.
"""
import ctypes as ct

import numba
import numpy as np
from numba import types
from numba.extending import overload

from . import native

_SORT_T = ct.CFUNCTYPE(None, ct.c_void_p, ct.c_uint64)                                   # parallel_sort_t_sig
_ARGSORT_T = ct.CFUNCTYPE(None, ct.c_void_p, ct.c_void_p, ct.c_uint64, ct.c_uint8)        # parallel_argsort_t_sig
_ROLLING_T = ct.CFUNCTYPE(ct.c_int32, ct.c_int32, ct.c_int32, ct.c_void_p, ct.c_void_p, ct.c_int64, ct.c_int64,
                          ct.c_int64, ct.c_int64)
_ERR_T = ct.CFUNCTYPE(ct.c_int64)

_NAMES = {types.int8: "i8", types.uint8: "u8", types.int16: "i16", types.uint16: "u16", types.int32: "i32",
          types.uint32: "u32", types.int64: "i64", types.uint64: "u64", types.float32: "f32", types.float64: "f64"}


def bind(sym, sig):
    """ctypes binding to an exported symbol (the reference's `bind`, sdc/functions/sort.py:39-42)."""
    return ct.cast(native.fn_address(sym), sig)


_host_rolling = bind("sdcb200_host_rolling", _ROLLING_T)
_launch_count = bind("sdcb200_launch_count", _ERR_T)
_sort_fns = {t: bind("parallel_sort_" + s, _SORT_T) for t, s in _NAMES.items()}
_stable_sort_fns = {t: bind("parallel_stable_sort_" + s, _SORT_T) for t, s in _NAMES.items()}
_argsort_fns = {t: bind("parallel_argsort_u64" + s, _ARGSORT_T) for t, s in _NAMES.items()}
_stable_argsort_fns = {t: bind("parallel_stable_argsort_u64" + s, _ARGSORT_T) for t, s in _NAMES.items()}


# ------------------------------------------------------------------ sort / argsort
def parallel_sort(arr):
    raise NotImplementedError("call from nopython code")


def parallel_stable_sort(arr):
    raise NotImplementedError("call from nopython code")


def parallel_argsort(arr, ascending=True):
    raise NotImplementedError("call from nopython code")


def parallel_stable_argsort(arr, ascending=True):
    raise NotImplementedError("call from nopython code")


def _sort_overload(table):
    def ovld(arr):
        if not (isinstance(arr, types.Array) and arr.dtype in table):
            return None
        fn = table[arr.dtype]

        def impl(arr):
            before = _launch_count()
            fn(arr.ctypes, len(arr))
            if len(arr) > 1 and _launch_count() == before:        # the natives return void (module.cpp:31-92)
                raise RuntimeError("libsdc_b200: sort did not run on the GPU (see stderr)")
            return arr
        return impl
    return ovld


def _argsort_overload(table):
    def ovld(arr, ascending=True):
        if not (isinstance(arr, types.Array) and arr.dtype in table):
            return None
        fn = table[arr.dtype]

        def impl(arr, ascending=True):
            index = np.empty(len(arr), np.int64)                   # sdc/functions/sort.py:294
            before = _launch_count()
            fn(index.ctypes, arr.ctypes, len(arr), types.uint8(ascending))
            if len(arr) > 0 and _launch_count() == before:
                raise RuntimeError("libsdc_b200: argsort did not run on the GPU (see stderr)")
            return index
        return impl
    return ovld


overload(parallel_sort)(_sort_overload(_sort_fns))
overload(parallel_stable_sort)(_sort_overload(_stable_sort_fns))
overload(parallel_argsort)(_argsort_overload(_argsort_fns))
overload(parallel_stable_argsort)(_argsort_overload(_stable_argsort_fns))


# ------------------------------------------------------------------ rolling
def _make_rolling(op_code, with_ddof):
    def stub(arr, window, min_periods, ddof=1):
        raise NotImplementedError("call from nopython code")

    def ovld(arr, window, min_periods, ddof=1):
        if not (isinstance(arr, types.Array) and arr.dtype in (types.float64, types.int64)):
            return None
        dcode = 9 if arr.dtype == types.float64 else 6

        def impl(arr, window, min_periods, ddof=1):
            # argument errors are raised before the call, like gen_sdc_pandas_rolling_overload_body
            # (sdc/datatypes/hpat_pandas_rolling_types.py:144-170)
            if window < 0:
                raise ValueError('window must be non-negative')
            if min_periods < 0:
                raise ValueError('min_periods must be >= 0')
            if min_periods > window:
                raise ValueError('min_periods must be <= window')
            out = np.empty(len(arr), np.float64)
            status = _host_rolling(op_code, dcode, arr.ctypes, out.ctypes, len(arr), window, min_periods, ddof)
            if status != 0:
                raise RuntimeError("libsdc_b200: rolling failed (sdcb200_last_error has the message)")
            return out
        return impl

    overload(stub)(ovld)
    return stub


rolling_sum = _make_rolling(native.ROLL_OPS["sum"], False)
rolling_mean = _make_rolling(native.ROLL_OPS["mean"], False)
rolling_var = _make_rolling(native.ROLL_OPS["var"], True)
rolling_std = _make_rolling(native.ROLL_OPS["std"], True)
rolling_count = _make_rolling(native.ROLL_OPS["count"], False)
rolling_min = _make_rolling(native.ROLL_OPS["min"], False)
rolling_max = _make_rolling(native.ROLL_OPS["max"], False)


# ------------------------------------------------------------------ groupby / merge on host arrays
# groupby_<agg>(keys, values, sort=True) -> (group keys, aggregate)   -- df.groupby(k)[v].<agg>() / Series.groupby(keys).<agg>()
#                                           (sdc/datatypes/hpat_pandas_groupby_functions.py:483-600, skipna rules of
#                                           sdc/functions/numpy_like.py:504-532, 636-676, 775-792, 850-872)
# merge_indexers(left_keys, right_keys, how) -> (left rows, right rows, -1 = no partner)
#                                           (the indexer convention of _sdc_internal_join, sdc/datatypes/common_functions.py:225-336)
# The stream arguments of the C ABI are declared as integers here so that nopython code can pass 0 (the NULL stream).
_HOST_GB_T = ct.CFUNCTYPE(ct.c_int32, ct.c_int32, ct.c_void_p, ct.c_int64, ct.c_int32, ct.c_void_p, ct.c_void_p, ct.c_uint32,
                          ct.c_int32, ct.c_int64, ct.c_int64, ct.c_void_p, ct.c_void_p)
_GB_KEYS_T = ct.CFUNCTYPE(ct.c_int32, ct.c_int64, ct.c_void_p, ct.c_int32, ct.c_size_t)
_GB_VALS_T = ct.CFUNCTYPE(ct.c_int32, ct.c_int64, ct.c_int32, ct.c_uint32, ct.c_void_p, ct.c_int32, ct.c_size_t)
_GB_FREE_T = ct.CFUNCTYPE(ct.c_int32, ct.c_int64)
_HOST_JOIN_T = ct.CFUNCTYPE(ct.c_int32, ct.c_int32, ct.c_int32, ct.c_void_p, ct.c_int64, ct.c_void_p, ct.c_int64, ct.c_void_p,
                            ct.c_void_p)
_JOIN_IDX_T = ct.CFUNCTYPE(ct.c_int32, ct.c_int64, ct.c_void_p, ct.c_void_p, ct.c_int32, ct.c_size_t)
_JOIN_FREE_T = ct.CFUNCTYPE(ct.c_int32, ct.c_int64, ct.c_size_t)

_host_groupby = bind("sdcb200_host_groupby", _HOST_GB_T)
_groupby_keys = bind("sdcb200_groupby_keys", _GB_KEYS_T)
_groupby_values = bind("sdcb200_groupby_values", _GB_VALS_T)
_groupby_free = bind("sdcb200_groupby_free", _GB_FREE_T)
_host_join = bind("sdcb200_host_join", _HOST_JOIN_T)
_join_indexers = bind("sdcb200_join_indexers", _JOIN_IDX_T)
_join_free = bind("sdcb200_join_free", _JOIN_FREE_T)


def _make_groupby(agg):
    bit = native.AGG_BITS[agg]

    def stub(keys, values, sort=True, ddof=1):
        raise NotImplementedError("call from nopython code")

    def ovld(keys, values, sort=True, ddof=1):
        if not (isinstance(keys, types.Array) and keys.dtype in (types.int64, types.float64)):
            return None
        if not (isinstance(values, types.Array) and values.dtype in (types.int64, types.float64)):
            return None
        kcode = 9 if keys.dtype == types.float64 else 6
        vcode = 9 if values.dtype == types.float64 else 6
        kdt = np.float64 if keys.dtype == types.float64 else np.int64
        int_out = agg == "count" or (values.dtype == types.int64 and agg in ("sum", "min", "max", "prod"))
        odt = np.int64 if int_out else np.float64

        def impl(keys, values, sort=True, ddof=1):
            if len(values) != len(keys):
                raise ValueError('Method Series.groupby(). The object by\n expected: array of the same length')
            k = np.ascontiguousarray(keys)
            v = np.ascontiguousarray(values)
            ptrs = np.empty(1, np.uintp)
            ptrs[0] = v.ctypes.data
            dts = np.full(1, vcode, np.int32)
            out = np.zeros(2, np.int64)                      # handle, group count
            status = _host_groupby(kcode, k.ctypes, len(k), 1, ptrs.ctypes, dts.ctypes, bit, 1 if sort else 0, ddof, 0,
                                   out[0:1].ctypes, out[1:2].ctypes)
            if status != 0:
                raise RuntimeError("libsdc_b200: groupby failed (sdcb200_last_error has the message)")
            rk = np.empty(out[1], kdt)
            rv = np.empty(out[1], odt)
            s1 = _groupby_keys(out[0], rk.ctypes, 1, 0)
            s2 = _groupby_values(out[0], 0, bit, rv.ctypes, 1, 0)
            _groupby_free(out[0])
            if s1 != 0 or s2 != 0:
                raise RuntimeError("libsdc_b200: groupby result copy failed")
            return rk, rv
        return impl

    overload(stub)(ovld)
    return stub


groupby_sum = _make_groupby("sum")
groupby_mean = _make_groupby("mean")
groupby_min = _make_groupby("min")
groupby_max = _make_groupby("max")
groupby_count = _make_groupby("count")
groupby_var = _make_groupby("var")
groupby_std = _make_groupby("std")


def merge_indexers(left_keys, right_keys, how="inner"):
    raise NotImplementedError("call from nopython code")


@overload(merge_indexers)
def _merge_indexers_ovld(left_keys, right_keys, how="inner"):
    if not (isinstance(left_keys, types.Array) and isinstance(right_keys, types.Array)):
        return None
    if left_keys.dtype != right_keys.dtype or left_keys.dtype not in (types.int64, types.float64):
        return None
    kcode = 9 if left_keys.dtype == types.float64 else 6

    def impl(left_keys, right_keys, how="inner"):
        if how == "inner":
            hcode = 0
        elif how == "left":
            hcode = 1
        elif how == "right":
            hcode = 2
        elif how == "outer":
            hcode = 3
        else:
            raise ValueError("merge: how is not supported (inner / left / right / outer)")
        lk = np.ascontiguousarray(left_keys)
        rk = np.ascontiguousarray(right_keys)
        out = np.zeros(2, np.int64)                          # handle, pair count
        status = _host_join(hcode, kcode, lk.ctypes, len(lk), rk.ctypes, len(rk), out[0:1].ctypes, out[1:2].ctypes)
        if status != 0:
            raise RuntimeError("libsdc_b200: join failed (sdcb200_last_error has the message)")
        li = np.empty(out[1], np.int64)
        ri = np.empty(out[1], np.int64)
        s1 = 0
        if out[1] > 0:
            s1 = _join_indexers(out[0], li.ctypes, ri.ctypes, 1, 0)
        _join_free(out[0], 0)
        if s1 != 0:
            raise RuntimeError("libsdc_b200: join result copy failed")
        return li, ri
    return impl
