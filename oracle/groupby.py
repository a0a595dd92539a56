"""Oracle (test infrastructure): DataFrame.groupby(key).<agg>() / Series.groupby(arr).<agg>() on the CPU.

Restates the reference's index-list groupby:
  * per-chunk Dict[key -> List[int64 row]] built over `parallel_chunks`, NaN / None keys skipped
        sdc/datatypes/hpat_pandas_dataframe_functions.py:3076-3097  (Series: hpat_pandas_series_functions.py:4720-4811)
  * serial merge of the chunk dicts in chunk order (rows stay ascending inside a key; key order =
    global first-appearance order) ........ sdc/datatypes/hpat_pandas_groupby_functions.py:58-73
  * result: keys from dict iteration, optional stable argsort (sort=True), then per column and per
    group `take` + Series.<agg>() ......... sdc/datatypes/hpat_pandas_groupby_functions.py:194-243
  * reducers (skipna): nansum functions/numpy_like.py:504-532, nanmean :775-792, nanmin/nanmax
    :636-676, nanvar :850-872, Series.var/std hpat_pandas_series_functions.py:1605-1620, :1322-1324,
    Series.count :3635-3658 (integer columns: len), nanprod numpy_like.py:683-705 (all-NaN -> 1, integer
    columns accumulate in intp), Series.median -> numpy.nanmedian hpat_pandas_series_functions.py:3721-3730
Pinned by tests/test_oracle_groupby.py against the literal outputs of
examples/dataframe/groupby/*.py, examples/series/series_groupby.py and pandas on the inputs of
sdc/tests/test_groupby.py.
"""
import numba
import numpy as np
from numba import types
from numba.typed import Dict, List

from .prange_utils import chunk_bounds

AGGS = ("sum", "mean", "min", "max", "count", "var", "std", "prod", "median")


_ROWS_T = types.ListType(types.int64)


@numba.njit(cache=True)
def _build_chunk_dicts_i64(keys, bounds):
    parts = [Dict.empty(types.int64, _ROWS_T) for _ in range(len(bounds) - 1)]
    for c in range(len(bounds) - 1):       # the reference runs this loop under prange
        d = parts[c]
        for j in range(bounds[c], bounds[c + 1]):
            k = keys[j]
            if k in d:
                d[k].append(j)
            else:
                lst = List.empty_list(types.int64)
                lst.append(j)
                d[k] = lst
    return parts


@numba.njit(cache=True)
def _merge_and_flatten_i64(parts, n):
    """merge_groupby_dicts_inplace in chunk order, then CSR: (keys, offsets, rows)."""
    res = parts[0]
    for c in range(1, len(parts)):
        for k, lst in parts[c].items():
            if k in res:
                res[k].extend(lst)
            else:
                res[k] = lst
    ng = len(res)
    gkeys = np.empty(ng, dtype=np.int64)
    offs = np.zeros(ng + 1, dtype=np.int64)
    rows = np.empty(n, dtype=np.int64)
    g = 0
    pos = 0
    for k, lst in res.items():
        gkeys[g] = k
        for r in lst:
            rows[pos] = r
            pos += 1
        g += 1
        offs[g] = pos
    return gkeys, offs, rows[:pos]


def group_rows(keys, nchunks=None):
    """-> (group_keys in first-appearance order, offsets[ng+1], rows) for int64 / float64 / str keys."""
    if nchunks is None:
        nchunks = numba.config.NUMBA_NUM_THREADS
    if isinstance(keys, np.ndarray) and keys.dtype.kind in "iu":
        k64 = np.ascontiguousarray(keys).astype(np.int64)
        b = chunk_bounds(len(k64), nchunks)
        if len(k64) == 0:
            return k64[:0], np.zeros(1, dtype=np.int64), np.zeros(0, dtype=np.int64)
        gk, offs, rows = _merge_and_flatten_i64(_build_chunk_dicts_i64(k64, b), len(k64))
        return gk.astype(keys.dtype), offs, rows
    # float keys (NaN skipped, -0.0 == 0.0 hash to one key) and string keys (None skipped): a
    # Python dict has the same first-appearance / equality behaviour as Numba's typed Dict
    groups = {}
    for j, k in enumerate(keys.tolist() if isinstance(keys, np.ndarray) else keys):
        if k is None or (isinstance(k, float) and k != k):
            continue
        groups.setdefault(k, []).append(j)
    gk = list(groups.keys())
    offs = np.zeros(len(gk) + 1, dtype=np.int64)
    rows = np.empty(sum(len(v) for v in groups.values()), dtype=np.int64)
    for g, k in enumerate(gk):
        v = groups[k]
        rows[offs[g]:offs[g] + len(v)] = v
        offs[g + 1] = offs[g] + len(v)
    if isinstance(keys, np.ndarray):
        gk = np.asarray(gk, dtype=keys.dtype)
    return gk, offs, rows


@numba.njit(cache=True)
def _reduce_f64(vals, offs, rows, order, op, ddof):
    """op: 0 sum 1 mean 2 min 3 max 4 count 5 var 6 std 7 prod 8 median -- per group in `order`, NaN skipped."""
    ng = len(order)
    out = np.empty(ng, dtype=np.float64)
    for jj in range(ng):
        g = order[jj]
        lo, hi = offs[g], offs[g + 1]
        if op == 7:
            pr = 1.0
            for p in range(lo, hi):
                v = vals[rows[p]]
                if not np.isnan(v):
                    pr *= v
            out[jj] = pr
            continue
        if op == 8:
            tmp = np.empty(hi - lo, dtype=np.float64)
            m = 0
            for p in range(lo, hi):
                v = vals[rows[p]]
                if not np.isnan(v):
                    tmp[m] = v
                    m += 1
            if m == 0:
                out[jj] = np.nan
            else:
                srt = np.sort(tmp[:m])
                out[jj] = (srt[(m - 1) // 2] + srt[m // 2]) / 2.0
            continue
        s = 0.0
        cnt = 0
        mn = np.inf
        mx = -np.inf
        for p in range(lo, hi):
            v = vals[rows[p]]
            if not np.isnan(v):
                s += v
                cnt += 1
                mn = min(mn, v)
                mx = max(mx, v)
        if op == 0:
            out[jj] = s
        elif op == 1:
            out[jj] = s / cnt if cnt else np.nan
        elif op == 2:
            out[jj] = mn if cnt else np.nan
        elif op == 3:
            out[jj] = mx if cnt else np.nan
        elif op == 4:
            out[jj] = cnt
        else:
            if cnt <= ddof:
                out[jj] = np.nan
            else:
                m = s / cnt
                ssd = 0.0
                for p in range(lo, hi):
                    v = vals[rows[p]]
                    if not np.isnan(v):
                        ssd += (v - m) * (v - m)
                var = (ssd / cnt) * cnt / (cnt - ddof)
                out[jj] = var ** 0.5 if op == 6 else var
    return out


@numba.njit(cache=True)
def _reduce_i64(vals, offs, rows, order, op):
    """integer column: 0 sum 2 min 3 max 4 count 7 prod (exact int64, wrap-around)."""
    ng = len(order)
    out = np.empty(ng, dtype=np.int64)
    for jj in range(ng):
        g = order[jj]
        lo, hi = offs[g], offs[g + 1]
        if op == 7:
            pr = np.int64(1)
            for p in range(lo, hi):
                pr = pr * vals[rows[p]]
            out[jj] = pr
            continue
        s = 0
        mn = vals[rows[lo]]
        mx = mn
        for p in range(lo, hi):
            v = vals[rows[p]]
            s += v
            mn = min(mn, v)
            mx = max(mx, v)
        if op == 0:
            out[jj] = s
        elif op == 2:
            out[jj] = mn
        elif op == 3:
            out[jj] = mx
        else:
            out[jj] = hi - lo
    return out


def key_order(gkeys, sort):
    """sort=True: stable argsort of the group keys (groupby_functions.py:209-210); else dict order."""
    if not sort:
        return np.arange(len(gkeys), dtype=np.int64)
    if isinstance(gkeys, np.ndarray):
        return np.argsort(gkeys, kind="stable").astype(np.int64)
    enc = [k.encode("utf-8") for k in gkeys]
    return np.asarray(sorted(range(len(enc)), key=lambda i: enc[i]), dtype=np.int64)


def groupby_agg(keys, columns, agg, sort=True, ddof=1, nchunks=None):
    """-> (result_keys, {name: result array}).  `columns` maps name -> 1-D numeric ndarray.

    Result dtypes follow pandas: float columns -> float64 (count -> int64);
    integer columns: sum / min / max / count / prod -> int64, mean / var / std / median -> float64.
    """
    assert agg in AGGS
    gkeys, offs, rows = group_rows(keys, nchunks)
    order = key_order(gkeys, sort)
    opc = AGGS.index(agg)
    out = {}
    for name, col in columns.items():
        col = np.ascontiguousarray(col)
        if col.dtype.kind in "iub" and agg in ("sum", "min", "max", "count", "prod"):
            out[name] = _reduce_i64(col.astype(np.int64), offs, rows, order, opc)
        else:
            r = _reduce_f64(col.astype(np.float64), offs, rows, order, opc, ddof)
            out[name] = r.astype(np.int64) if agg == "count" else r
    rkeys = gkeys[order] if isinstance(gkeys, np.ndarray) else [gkeys[i] for i in order]
    return rkeys, out
