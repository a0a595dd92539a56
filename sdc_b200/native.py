"""ctypes binding of libsdc_b200.so (the C ABI declared in include/sdc_b200.h).

This is the analogue of the reference's FFI idiom A (sdc/functions/sort.py:39-72:
`ct.cast(address, CFUNCTYPE)` over symbols published by a native module) -- here the
symbols come straight from a plain shared library.  There is no CPU fallback: if the
library is missing or a CUDA call fails, an exception is raised.
"""
import ctypes as ct
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SDCB200_LIB_PATH") or os.path.join(_HERE, "lib", "libsdc_b200.so")   # override: A/B builds (tools/)

OK = 0
JOIN_HOW = {"inner": 0, "left": 1, "right": 2, "outer": 3}
JOIN_ANY_ORDER = 0x100          # SDCB200_JOIN_ANY_ORDER: how="inner" returns the pairs as a row set, any order
AGG_BITS = {"sum": 1, "mean": 2, "min": 4, "max": 8, "count": 16, "var": 32, "std": 64, "prod": 128, "median": 256}
ROLL_OPS = {"sum": 0, "mean": 1, "var": 2, "std": 3, "count": 4, "min": 5, "max": 6}
ROLL_MOMENT_OPS = {"skew": 7, "kurt": 8, "cov": 9, "corr": 10}
DTYPE_CODES = {"int8": 0, "uint8": 1, "int16": 2, "uint16": 3, "int32": 4, "uint32": 5,
               "int64": 6, "uint64": 7, "float32": 8, "float64": 9}
DTYPE_GID = 100     # SDCB200_GID: dense int64 group ids as a groupby key column
ROUTE_HASH, ROUTE_MOD, ROUTE_RANGE = 0, 1, 2

_i32, _i64, _vp, _u8 = ct.c_int32, ct.c_int64, ct.c_void_p, ct.c_uint8

# name -> (restype, argtypes); every symbol include/sdc_b200.h declares
SIGNATURES = {
    "sdcb200_last_error": (ct.c_char_p, []),
    "sdcb200_version": (ct.c_char_p, []),
    "sdcb200_launch_count": (_i64, []),
    "sdcb200_device_info": (_i32, [ct.POINTER(_i32), ct.POINTER(_i64), ct.POINTER(_i32), ct.POINTER(_i32)]),
    "sdcb200_host_alloc": (_i32, [ct.POINTER(_vp), _i64]),
    "sdcb200_host_free": (_i32, [_vp]),
    "sdcb200_dev_alloc": (_i32, [ct.POINTER(_vp), _i64, _vp]),
    "sdcb200_dev_free": (_i32, [_vp, _vp]),
    "sdcb200_memcpy_h2d": (_i32, [_vp, _vp, _i64, _vp]),
    "sdcb200_memcpy_d2h": (_i32, [_vp, _vp, _i64, _vp]),
    "sdcb200_stream_sync": (_i32, [_vp]),
    "sdcb200_rolling": (_i32, [_i32, _i32, _vp, _vp, _i64, _i64, _i64, _i64, _vp, _i64, _i64, _vp]),
    "sdcb200_host_rolling": (_i32, [_i32, _i32, _vp, _vp, _i64, _i64, _i64, _i64]),
    "sdcb200_halo_mailbox_bytes": (_i64, []),
    "sdcb200_rolling_linked": (_i32, [_i32, _i32, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _vp, ct.c_uint32, _vp]),
    "sdcb200_rolling_moments": (_i32, [_i32, _vp, _i64, _vp, _i64, _vp, _i64, _i64, _i64, _vp]),
    "sdcb200_sort": (_i32, [_i32, _vp, _i64, _vp]),
    "sdcb200_argsort": (_i32, [_i32, _vp, _i64, _i32, _vp, _vp]),
    "sdcb200_gather": (_i32, [_i32, _vp, _vp, _i64, _vp, _vp]),
    "sdcb200_host_sort": (_i32, [_i32, _vp, _i64]),
    "sdcb200_host_argsort": (_i32, [_i32, _vp, _i64, _i32, _vp]),
    "set_number_of_threads": (None, [ct.c_uint64]),
    "sdcb200_groupby": (_i32, [_i32, _vp, _i64, _i32, ct.POINTER(_vp), ct.POINTER(_i32), ct.c_uint32, _i32, _i64,
                               _i64, _vp, ct.POINTER(_i64), ct.POINTER(_i64)]),
    "sdcb200_groupby_begin": (_i32, [_i32, _vp, _i64, _i32, ct.POINTER(_vp), ct.POINTER(_i32), ct.c_uint32, _i32, _i64,
                                     _i64, _vp, ct.POINTER(_i64)]),
    "sdcb200_groupby_end": (_i32, [_i64, ct.POINTER(_i64)]),
    "sdcb200_groupby_keys": (_i32, [_i64, _vp, _i32, _vp]),
    "sdcb200_groupby_values": (_i32, [_i64, _i32, ct.c_uint32, _vp, _i32, _vp]),
    "sdcb200_groupby_result": (_i32, [_i64, ct.POINTER(_vp), ct.POINTER(_vp), ct.POINTER(_i64)]),
    "sdcb200_groupby_free": (_i32, [_i64]),
    "sdcb200_host_groupby": (_i32, [_i32, _vp, _i64, _i32, ct.POINTER(_vp), ct.POINTER(_i32), ct.c_uint32, _i32, _i64,
                                    _i64, ct.POINTER(_i64), ct.POINTER(_i64)]),
    "sdcb200_host_join": (_i32, [_i32, _i32, _vp, _i64, _vp, _i64, ct.POINTER(_i64), ct.POINTER(_i64)]),
    "sdcb200_combine_ids": (_i32, [_vp, _vp, _i64, _i64, _vp, _i64, _vp]),
    "sdcb200_join": (_i32, [_i32, _i32, _vp, _i64, _vp, _i64, _vp, ct.POINTER(_i64), ct.POINTER(_i64)]),
    "sdcb200_join_ids": (_i32, [_i32, _i32, _vp, _i64, _vp, _i64, _vp, _vp, _vp, ct.POINTER(_i64), ct.POINTER(_i64)]),
    "sdcb200_join_indexers": (_i32, [_i64, _vp, _vp, _i32, _vp]),
    "sdcb200_join_result": (_i32, [_i64, ct.POINTER(_vp), ct.POINTER(_vp)]),
    "sdcb200_join_free": (_i32, [_i64, _vp]),
    "sdcb200_take_nullable_f64": (_i32, [_i32, _vp, _vp, _i64, _vp, _vp]),
    "sdcb200_asof_backward": (_i32, [_i32, _vp, _i64, _vp, _i64, _vp, _vp]),
    "sdcb200_str_argsort": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp]),
    "sdcb200_str_factorize": (_i32, [_vp, _vp, _vp, _i64, _i64, _vp, ct.POINTER(_i64), ct.POINTER(_i64),
                                     ct.POINTER(_i64), _vp]),
    "sdcb200_str_group_rows": (_i32, [_i64, _i64, _vp, _i64, _vp, _vp]),
    "sdcb200_str_table_free": (_i32, [_i64, _vp]),
    "sdcb200_str_gather_sizes": (_i32, [_vp, _vp, _i64, _vp, ct.POINTER(_i64), _vp]),
    "sdcb200_str_gather_chars": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "sdcb200_str_profile": (_i32, [_vp, _vp, _i64, ct.POINTER(_i64), _vp]),
    "sdcb200_str_pack_words": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "sdcb200_str_unpack_words": (_i32, [_vp, _i64, _i32, _vp, _vp, ct.POINTER(_i64), _vp]),
    "sdcb200_str_words_order": (_i32, [_vp, _i64, _i32, _vp, _vp]),
    "sdcb200_csv_open": (_i32, [_vp, _i64, ct.POINTER(_i64), ct.POINTER(_i64), _vp]),
    "sdcb200_csv_classify": (_i32, [_i64, _i32, _i32, _vp]),
    "sdcb200_csv_parse": (_i32, [_i64, _i32, _i32, _vp, _vp]),
    "sdcb200_csv_close": (_i32, [_i64]),
    "sdcb200_str_hash": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "sdcb200_str_pack_lens": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "sdcb200_str_unpack_lens": (_i32, [_vp, _i64, _vp, _vp, ct.POINTER(_i64), _vp]),
    "stable_argsort": (None, [_vp, _vp, ct.c_uint64, ct.c_int8, _vp]),
    "sdcb200_hash_partition": (_i32, [_i32, _i32, _vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    "sdcb200_combine_mean": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "sdcb200_pack_partials": (_i32, [_vp, _i32, _i64, _i64, _vp, _vp]),
    "sdcb200_combine_partials": (_i32, [_vp, _i32, _i32, _vp, _i64, _vp, _vp, _vp]),
    "sdcb200_ipc_alloc": (_i32, [ct.POINTER(_vp), _i64]),
    "sdcb200_ipc_free": (_i32, [_vp]),
    "sdcb200_ipc_get_handle": (_i32, [_vp, _vp]),
    "sdcb200_ipc_open": (_i32, [_vp, ct.POINTER(_vp)]),
    "sdcb200_ipc_close": (_i32, [_vp]),
    "sdcb200_shuffle_plan": (_i32, [_i32, _i32, _vp, _i64, _i32, _vp, _vp, _vp]),
    "sdcb200_shuffle_send": (_i32, [_i32, _i32, _i32, ct.POINTER(_vp), _i32, _i64, _i64, _i32, _vp, ct.POINTER(_i64),
                                    ct.POINTER(_vp), _i64, _vp]),
    "sdcb200_shuffle_plan_range": (_i32, [_i32, ct.POINTER(ct.c_uint64), _i32, _i32, _vp, _i64, _i32, _vp, _vp, _vp]),
    "sdcb200_shuffle_send_range": (_i32, [_i32, ct.POINTER(ct.c_uint64), _i32, _i32, _i32, ct.POINTER(_vp), _i32, _i64, _i64, _i32,
                                          _vp, ct.POINTER(_i64), ct.POINTER(_vp), _i64, _vp]),
    "sdcb200_affine_i64": (_i32, [_vp, _i64, _i64, _i64, _i32, _vp]),
    "sdcb200_gather_ids": (_i32, [_vp, _vp, _i64, _vp, _vp]),
}

# the reference's sdc.concurrent_sort entry points (sdc/native/module.cpp:31-92)
REF_SORT_TYPES = ("i8", "u8", "i16", "u16", "i32", "u32", "i64", "u64", "f32", "f64")
for _t in REF_SORT_TYPES:
    SIGNATURES["parallel_sort_" + _t] = (None, [_vp, ct.c_uint64])
    SIGNATURES["parallel_stable_sort_" + _t] = (None, [_vp, ct.c_uint64])
    SIGNATURES["parallel_argsort_u64" + _t] = (None, [_vp, _vp, ct.c_uint64, _u8])
    SIGNATURES["parallel_stable_argsort_u64" + _t] = (None, [_vp, _vp, ct.c_uint64, _u8])


class NativeError(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library once; raise if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                "libsdc_b200.so not found at %s -- build it with "
                "`python sdc_b200/csrc/build.py` (there is no CPU fallback)" % LIB_PATH)
        lib = ct.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)        # AttributeError here == ABI drift
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status):
    if status != OK:
        msg = load().sdcb200_last_error().decode("utf-8", "replace")
        raise NativeError("libsdc_b200 status %d: %s" % (status, msg))


def launch_count():
    return int(load().sdcb200_launch_count())


def fn_address(name):
    """Raw address of an exported function (for Numba: ctypes CFUNCTYPE / llvmlite.add_symbol)."""
    return ct.cast(getattr(load(), name), ct.c_void_p).value


def current_stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
