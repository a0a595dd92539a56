"""Oracle (test infrastructure): Series.rolling(window, min_periods).skew() / .kurt() / .cov(other) / .corr(other).

Restates the reference's chunked add-new / drop-old accumulators over sums of powers and cross products
(reference: sdc/datatypes/hpat_pandas_series_rolling_functions.py):
  * skew: put/pop of (sum, sum2, sum3) .......... :408-431, result :548-556, formula sdc/functions/statistics.py:33-41
  * kurt: put/pop of (sum .. sum4) .............. :303-331, result :520-536
  * cov:  put/pop of (sum_x, sum_y, sum_xy, count), align_finiteness=True (both overloads pass it) :261-300,
          result :509-517, ddof :559-565
  * corr: put/pop of (sum_x, sum_y, sum_xy, sum_xx, sum_yy) :207-236, result :487-506,
          _almost_equal = |x - y| <= eps (sdc/datatypes/common_functions.py:584-600)
  * two-series driver (different lengths: only rows below min(len) enter a window, outputs past
    min(len) + window are NaN) .................. :824-879 (corr), :925-992 (cov)
A row takes part only if it is finite in BOTH series.  Pinned by tests/test_oracle_rolling_moments.py against pandas
on the reference's test vectors (test_rolling.py:1606-1610, test_utils.py:56-62) where the window is not degenerate.
"""
import numba
import numpy as np

from .prange_utils import chunk_bounds
from .rolling import check_rolling_args

OP_SKEW, OP_KURT, OP_COV, OP_CORR = 7, 8, 9, 10
OPS = {"skew": OP_SKEW, "kurt": OP_KURT, "cov": OP_COV, "corr": OP_CORR}
_EPS = np.finfo(np.float64).eps


@numba.njit(cache=True, inline='always')
def _emit(op, n, minp, s, ddof):
    nan = np.nan
    if op == OP_SKEW:
        if n < max(3, minp):
            return nan
        m2 = (s[1] - s[0] * s[0] / n) / n
        m3 = (s[2] - 3. * s[0] * s[1] / n + 2. * s[0] * s[0] * s[0] / n / n) / n
        res = nan if m2 == 0 else m3 / m2 ** 1.5
        if n > 2 and m2 > 0:
            res = np.sqrt((n - 1.) * n) / (n - 2.) * m3 / m2 ** 1.5
        return res
    if op == OP_KURT:
        if n < max(4, minp):
            return nan
        m2 = (s[1] - s[0] * s[0] / n) / n
        m4 = (s[3] - 4 * s[0] * s[2] / n + 6 * s[0] * s[0] * s[1] / n / n - 3 * s[0] * s[0] * s[0] * s[0] / n / n / n) / n
        res = 0. if m2 == 0 else m4 / m2 ** 2.0
        if n > 2 and m2 > 0:
            res = 1.0 / (n - 2) / (n - 3) * ((n ** 2 - 1.0) * m4 / m2 ** 2.0 - 3 * (n - 1) ** 2.0)
        return res
    if op == OP_COV:
        if n < max(1, minp):
            return nan
        res = (s[2] - s[0] * s[1] / n) / n
        if n - ddof < 1:
            return nan
        return res * n / (n - ddof)
    # corr
    if n < max(1, minp):
        return nan
    var_x = s[3] - s[0] * s[0] / n
    if abs(var_x) <= _EPS:
        return nan
    var_y = s[4] - s[1] * s[1] / n
    if abs(var_y) <= _EPS:
        return nan
    cov_xy = s[2] - s[0] * s[1] / n
    return cov_xy / np.sqrt(var_x * var_y)


@numba.njit(cache=True, inline='always')
def _update(op, s, x, y, sign):
    """put (sign = +1) / pop (sign = -1) of one row that is finite in both series."""
    if op == OP_SKEW or op == OP_KURT:
        s[0] += sign * x
        s[1] += sign * (x * x)
        s[2] += sign * (x * x * x)
        if op == OP_KURT:
            s[3] += sign * (x * x * x * x)
    else:
        s[0] += sign * x
        s[1] += sign * y
        s[2] += sign * (x * y)
        if op == OP_CORR:
            s[3] += sign * (x * x)
            s[4] += sign * (y * y)


@numba.njit(cache=True, parallel=True)
def _rolling_moments(a, b, win, minp, ddof, op, nchunks):
    min_length = min(len(a), len(b))
    length = max(len(a), len(b))
    out = np.empty(length, dtype=np.float64)
    bounds = chunk_bounds(length, nchunks)
    for c in numba.prange(len(bounds) - 1):
        start = bounds[c]
        stop = bounds[c + 1]
        n = 0
        s = np.zeros(5)
        if win == 0:
            for i in range(start, stop):
                out[i] = _emit(op, n, minp, s, ddof)
            continue
        pre_lo = max(0, start - win + 1)
        inter_hi = min(pre_lo + win, stop)
        for i in range(pre_lo, stop):
            if i < min_length:                                   # put
                x = a[i]
                y = b[i]
                if np.isfinite(x) and np.isfinite(y):
                    n += 1
                    _update(op, s, x, y, 1.0)
            if i < start:
                continue
            if i >= inter_hi and i - win < min_length:           # pop
                x = a[i - win]
                y = b[i - win]
                if np.isfinite(x) and np.isfinite(y):
                    n -= 1
                    _update(op, s, x, y, -1.0)
            if i >= min_length + win:
                out[i] = np.nan
            else:
                out[i] = _emit(op, n, minp, s, ddof)
    return out


def rolling_moments(a, op, window, min_periods=None, other=None, ddof=1, nchunks=None):
    """`pd.Series(a).rolling(window, min_periods).<op>([other])` values as float64[max(len)]."""
    minp = check_rolling_args(window, min_periods)
    x = np.ascontiguousarray(a, dtype=np.float64)
    y = x if other is None else np.ascontiguousarray(other, dtype=np.float64)
    if nchunks is None:
        nchunks = numba.config.NUMBA_NUM_THREADS
    return _rolling_moments(x, y, window, minp, ddof, OPS[op], nchunks)
