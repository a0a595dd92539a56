"""Oracle (test infrastructure): str_arr layout and the reference's string ordering on the CPU.

  * layout: uint32 offsets[n+1], uint8 chars, validity bitmap LSB-first with 1 = valid, NA stored empty
        sdc/str_arr_type.py:33-101, sdc/native/str_ext/sdc_str_arr.hpp:160-176, 310-314, 389-393
  * stable_argsort: every string is copied to a std::string and (string, idx) pairs go through
    std::stable_sort with `<` (ascending) or `>` (descending): unsigned byte-lexicographic order on the
    UTF-8 bytes, a proper prefix first, ties in input order both ways ... sdc_str_arr.hpp:565-593
  * convert_len_arr_to_offset: exclusive prefix sum of lengths .......... sdc_str_arr.hpp:242-252
Pinned in tests/test_oracle_strings.py against pandas on the string cases of sdc/tests/test_series.py:3856-4072.
"""
import functools

import numpy as np


def to_str_arr(values):
    """Python list (None = NA) -> (offsets uint32[n+1], chars uint8[total], bitmap uint8[ceil(n/8)])."""
    enc = [b"" if v is None else v.encode("utf-8") for v in values]
    lens = np.asarray([len(e) for e in enc], dtype=np.uint32)
    offsets = len_arr_to_offsets(lens)
    chars = np.frombuffer(b"".join(enc), dtype=np.uint8).copy()
    bitmap = np.packbits(np.asarray([v is not None for v in values], dtype=bool), bitorder="little")
    return offsets, chars, bitmap


def len_arr_to_offsets(lens):
    offsets = np.zeros(len(lens) + 1, dtype=np.uint32)
    acc = 0
    for i, l in enumerate(lens):       # the reference's serial loop
        offsets[i] = acc
        acc += int(l)
    offsets[len(lens)] = acc
    return offsets


def is_na(bitmap, i):
    return ((bitmap[i >> 3] >> (i & 7)) & 1) == 0


def stable_argsort(offsets, chars, ascending=True):
    """uint64 positions, exactly the reference's std::stable_sort over (std::string, idx)."""
    buf = chars.tobytes()
    strs = [buf[offsets[i]:offsets[i + 1]] for i in range(len(offsets) - 1)]     # bytes compare unsigned, prefix first
    if ascending:
        order = sorted(range(len(strs)), key=lambda i: strs[i])
    else:
        cmp = lambda a, b: (strs[a] < strs[b]) - (strs[a] > strs[b])               # `greater`: ties stay in input order
        order = sorted(range(len(strs)), key=functools.cmp_to_key(cmp))
    return np.asarray(order, dtype=np.uint64)
