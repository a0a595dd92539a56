"""Oracle (test infrastructure): Series.rolling(window, min_periods).<op>() on the CPU.

Restates the reference's chunked add-new / drop-old window accumulator
(reference: sdc/datatypes/hpat_pandas_series_rolling_functions.py):
  * chunk driver -- prelude / interlude / steady state ........ :586-633, :636-683, :686-733
  * sum / mean accumulators (non-finite values are skipped) .... :434-451, :478-484, :539-545
  * var / std: sum and sum of squares, ddof correction ......... :454-475, :559-583
  * count: counts non-NaN, "seen" counter never decreases ...... :237-255
  * min / max: finite only, rescan when the extreme leaves ..... :332-405
Argument checks restate sdc/datatypes/hpat_pandas_rolling_types.py:144-170.
Chunks come from oracle.prange_utils (reference prange_utils.py:48-69).

The driver is written once with an integer op code instead of the reference's
generated closures; the arithmetic and its order inside a chunk are the same, so
for a given chunk count the floating-point results follow the reference's
add-then-subtract rounding.  Pinned by tests/test_oracle_rolling.py against the
literal outputs in examples/series/rolling/*.py and against pandas on the
reference's own test vectors (tests/golden/rolling_*.json).
"""
import numba
import numpy as np

from .prange_utils import chunk_bounds

OP_SUM, OP_MEAN, OP_VAR, OP_STD, OP_COUNT, OP_MIN, OP_MAX = range(7)
OPS = {"sum": OP_SUM, "mean": OP_MEAN, "var": OP_VAR, "std": OP_STD,
       "count": OP_COUNT, "min": OP_MIN, "max": OP_MAX}


def check_rolling_args(window, min_periods=None, center=False, win_type=None,
                       on=None, axis=0, closed=None):
    """Validation + default of min_periods (rolling_types.py:144-170). Returns minp."""
    if window < 0:
        raise ValueError('window must be non-negative')
    minp = window if min_periods is None else min_periods
    if minp < 0:
        raise ValueError('min_periods must be >= 0')
    if minp > window:
        raise ValueError('min_periods must be <= window')
    if center is not False:
        raise ValueError('Method rolling(). The object center\n expected: False')
    if win_type is not None:
        raise ValueError('Method rolling(). The object win_type\n expected: None')
    if on is not None:
        raise ValueError('Method rolling(). The object on\n expected: None')
    if axis != 0:
        raise ValueError('Method rolling(). The object axis\n expected: 0')
    if closed is not None:
        raise ValueError('Method rolling(). The object closed\n expected: None')
    return minp


@numba.njit(cache=True, inline='always')
def _emit(op, n, minp, s1, s2, ddof):
    """Window state -> output value (result_or_nan & friends, :478-484, :539-583)."""
    if op == OP_SUM or op == OP_COUNT or op == OP_MIN or op == OP_MAX:
        if n < minp:
            return np.nan
        return s1
    if op == OP_MEAN:
        if n == 0 or n < minp:
            return np.nan
        return s1 / n
    # var / std
    if n < max(1, minp):
        return np.nan
    v = (s2 - s1 * s1 / n) / n
    if n - ddof < 1:
        v = np.nan
    else:
        v = v * n / (n - ddof)
    if op == OP_STD:
        return v ** 0.5
    return v


@numba.njit(cache=True, inline='always')
def _rescan(op, a, idx, win):
    """Recompute the finite extreme of a[idx-win+1 .. idx] (calc_min/calc_max, :332-341, :370-379)."""
    lo = max(0, idx - win + 1)
    n = 0
    r = np.nan
    for j in range(lo, idx + 1):
        v = a[j]
        if np.isfinite(v):
            n += 1
            if np.isnan(r) or (v > r if op == OP_MAX else v < r):
                r = v
    return n, r


@numba.njit(cache=True, parallel=True)
def _rolling(a, win, minp, ddof, op, nchunks):
    length = len(a)
    out = np.empty(length, dtype=np.float64)
    b = chunk_bounds(length, nchunks)
    for c in numba.prange(len(b) - 1):
        start = b[c]
        stop = b[c + 1]
        n = 0                       # "nfinite" slot (for count: elements seen so far)
        s1 = 0.0
        s2 = 0.0
        if op == OP_MIN or op == OP_MAX:
            s1 = np.nan
        if win == 0:                # :604-607 -- every output is the empty-window result
            for i in range(start, stop):
                out[i] = _emit(op, n, minp, s1, s2, ddof)
            continue
        pre_lo = max(0, start - win + 1)
        inter_hi = min(pre_lo + win, stop)
        for i in range(pre_lo, stop):
            v = a[i]
            # ---- put(a[i]) ----
            if op == OP_COUNT:
                n += 1
                if not np.isnan(v):
                    s1 += 1.0
            elif np.isfinite(v):
                n += 1
                if op == OP_MIN:
                    if np.isnan(s1) or v < s1:
                        s1 = v
                elif op == OP_MAX:
                    if np.isnan(s1) or v > s1:
                        s1 = v
                else:
                    s1 += v
                    if op == OP_VAR or op == OP_STD:
                        s2 += v * v
            if i < start:           # prelude: accumulate only (:609-617)
                continue
            # ---- pop(a[i - win]) in the steady state (:624-629) ----
            if i >= inter_hi:
                w = a[i - win]
                if op == OP_COUNT:
                    if not np.isnan(w):
                        s1 -= 1.0
                elif np.isfinite(w):
                    n -= 1
                    if op == OP_MIN or op == OP_MAX:
                        if n:
                            if w == s1:
                                n, s1 = _rescan(op, a, i, win)
                        else:
                            s1 = np.nan
                    else:
                        s1 -= w
                        if op == OP_VAR or op == OP_STD:
                            s2 -= w * w
            out[i] = _emit(op, n, minp, s1, s2, ddof)
    return out


def rolling(a, op, window, min_periods=None, ddof=1, nchunks=None):
    """`pd.Series(a).rolling(window, min_periods).<op>()` values as float64[n]."""
    minp = check_rolling_args(window, min_periods)
    arr = np.ascontiguousarray(a)
    if arr.dtype != np.float64:
        arr = arr.astype(np.float64)      # reference reads ints as numbers; output is float64 (:596)
    if nchunks is None:
        nchunks = numba.config.NUMBA_NUM_THREADS
    return _rolling(arr, window, minp, ddof, OPS[op], nchunks)
