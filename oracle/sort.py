"""Oracle (test infrastructure): sort / argsort / Series.sort_values on the CPU.

Restates the ORDER the reference's natives produce, on top of NumPy's stable sort
(the reference's TBB sort is an absent third-party dependency; its results are fully
determined by the comparator for the stable variants and up to tie order otherwise):
  * comparators: NaN sorts last for `less` AND for `greater`; -0.0 == +0.0
        sdc/native/utils.cpp:160-182
  * parallel_stable_argsort_u64<t>(index, data, len, ascending): sort the index array with
    IndexCompare{data, less | greater}; stable => ties keep their original order in both
    directions ........ sdc/native/stable_sort.cpp:283-293, sdc/native/utils.hpp:85-100
  * string stable_argsort: std::stable_sort of (std::string, idx): unsigned-byte lexicographic,
    shorter prefix first, descending = greater with ties in original order
        sdc/native/str_ext/sdc_str_arr.hpp:565-593
  * Series.sort_values: NaNs split off, argsort of the rest, NaN block first / last
        sdc/datatypes/hpat_pandas_series_functions.py:3924-3957
Pinned by tests/test_oracle_sort.py against pandas / NumPy mergesort on the reference's own
inputs (sdc/tests/test_sdc_numpy.py:311-425, sdc/tests/test_series.py:3856-4072).
"""
import numpy as np


def _descending_key(a):
    """A key whose ascending stable order is `greater` with NaN last and ties in input order."""
    if a.dtype.kind == "f":
        return -a                      # NaN stays NaN (sorted last by NumPy), -0.0/+0.0 stay tied
    if a.dtype.kind in "iu":
        return ~a                      # monotone decreasing, no overflow at the type minimum
    if a.dtype.kind == "b":
        return ~a
    raise TypeError(a.dtype)


def stable_argsort(a, ascending=True):
    """Index array the reference's parallel_stable_argsort_u64<t> writes (as int64)."""
    a = np.asarray(a)
    key = a if ascending else _descending_key(a)
    return np.argsort(key, kind="stable").astype(np.int64)


def sort(a):
    """parallel_sort_<t> / parallel_stable_sort_<t>: ascending, NaN last (values only)."""
    return np.sort(np.asarray(a), kind="stable")


def str_stable_argsort(strings, ascending=True):
    """`stable_argsort` over a str_arr: byte-wise (UTF-8) order; `strings` is a list of str."""
    enc = [s.encode("utf-8") for s in strings]
    idx = list(range(len(enc)))
    if ascending:
        idx.sort(key=lambda i: enc[i])                 # list.sort is stable
    else:
        # stable sort with `greater`: descending, equal strings keep input order
        idx.sort(key=lambda i: enc[i], reverse=True)   # reverse=True preserves stability in CPython
    return np.asarray(idx, dtype=np.int64)


def sort_values_indexer(data, ascending=True, na_position="last", is_na=None):
    """Positions `sorted_index` of Series.sort_values(kind='mergesort') (reference :3935-3952).

    `data` numeric ndarray, or list of str/None (then NA = None).  Returns int64 positions.
    """
    if na_position not in ("last", "first"):
        raise ValueError("Method sort_values(). Unsupported parameter. Given na_position != 'last', 'first'")
    if isinstance(data, np.ndarray) and data.dtype.kind in "iufb":
        mask = np.isnan(data) if data.dtype.kind == "f" else np.zeros(len(data), dtype=bool)
        good = np.flatnonzero(~mask)
        order = stable_argsort(data[good], ascending)
    else:
        mask = np.asarray([v is None or (isinstance(v, float) and np.isnan(v)) for v in data], dtype=bool)
        good = np.flatnonzero(~mask)
        order = str_stable_argsort([data[i] for i in good], ascending)
    nans = np.flatnonzero(mask)
    if na_position == "last":
        return np.concatenate([good[order], nans]).astype(np.int64)
    return np.concatenate([nans, good[order]]).astype(np.int64)
