"""oracle/_ref loader (TEST INFRASTRUCTURE): the reference's OWN compiled code, not a restatement.

`oracle/ref_recipe/Makefile` compiles, with g++ and from the sources where they lie under /root/reference,
  * sdc/native/str_ext/sdc_str_arr.hpp -- stable_argsort (:565-593, the string sort behind Series.sort_values and
    the string-key groupby's ordering), convert_len_arr_to_offset (:242-252, the shuffle's lengths -> offsets step),
    is_na (:310-314, the validity bitmap convention), set_string_array_range (:213-240);
  * sdc/_distributed.h:77-102 -- the block partition of a row range over ranks
into oracle/_ref/libsdc_ref_str.so (git-ignored, travels to the GPU box with the snapshot).  The numeric sort natives
(sdc/native/sort.cpp, stable_sort.cpp) include TBB unconditionally and TBB is absent, and the Numba-compiled kernels need
numba 0.54: those stay restatements.

available() is False when the library was not built (a checkout without /root/reference and without the prebuilt file);
tests that use it skip then."""
import ctypes as ct
import os

import numpy as np

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libsdc_ref_str.so")
_lib = None


def available():
    return os.path.exists(_PATH)


def _load():
    global _lib
    if _lib is None:
        lib = ct.CDLL(_PATH)
        lib.stable_argsort.argtypes = [ct.c_void_p, ct.c_void_p, ct.c_int64, ct.c_int8, ct.c_void_p]
        lib.stable_argsort.restype = None
        lib.convert_len_arr_to_offset.argtypes = [ct.c_void_p, ct.c_int64]
        lib.convert_len_arr_to_offset.restype = None
        lib.is_na.argtypes = [ct.c_void_p, ct.c_int64]
        lib.is_na.restype = ct.c_bool
        for name in ("ref_dist_get_start", "ref_dist_get_end", "ref_dist_get_node_portion"):
            getattr(lib, name).argtypes = [ct.c_int64, ct.c_int, ct.c_int]
            getattr(lib, name).restype = ct.c_int64
        lib.ref_index_rank.argtypes = [ct.c_int64, ct.c_int, ct.c_int]
        lib.ref_index_rank.restype = ct.c_int64
        _lib = lib
    return _lib


def stable_argsort(offsets, chars, ascending=True):
    """The reference's stable_argsort(char* data, uint32* offsets, int64 len, int8 ascending, uint64* result)."""
    offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
    chars = np.ascontiguousarray(chars, dtype=np.uint8)
    n = len(offsets) - 1
    out = np.empty(n, dtype=np.uint64)
    buf = chars if len(chars) else np.zeros(1, dtype=np.uint8)
    _load().stable_argsort(buf.ctypes.data, offsets.ctypes.data, n, 1 if ascending else 0, out.ctypes.data)
    return out


def len_arr_to_offsets(lens):
    """The reference's convert_len_arr_to_offset over a copy of `lens` (uint32[n] -> uint32[n + 1])."""
    n = len(lens)
    buf = np.zeros(n + 1, dtype=np.uint32)
    buf[:n] = lens
    _load().convert_len_arr_to_offset(buf.ctypes.data, n)
    return buf


def is_na(bitmap, i):
    bitmap = np.ascontiguousarray(bitmap, dtype=np.uint8)
    return bool(_load().is_na(bitmap.ctypes.data, i))


def block_range(total, num_pes, node_id):
    """(start, end) of rank node_id's block: hpat_dist_get_start / _get_end."""
    lib = _load()
    return int(lib.ref_dist_get_start(total, num_pes, node_id)), int(lib.ref_dist_get_end(total, num_pes, node_id))


def index_rank(total, num_pes, index):
    return int(_load().ref_index_rank(total, num_pes, index))
