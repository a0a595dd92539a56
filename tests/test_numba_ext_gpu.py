"""The C-ABI library called from nopython code through the reference's FFI idiom A (ctypes CFUNCTYPE
inside @njit, sdc/functions/sort.py:39-72, 243-310) -- parity with the oracle."""
import numba
import numpy as np
import pytest

from oracle import rolling as orolling
from oracle import sort as osort

pytestmark = pytest.mark.gpu


def test_jitted_rolling_and_sort():
    from sdc_b200 import numba_ext as nx

    @numba.njit
    def f(x, w):
        return nx.rolling_mean(x, w, w), nx.rolling_std(x, w, 1, 1), nx.rolling_count(x, w, 0)

    x = np.random.default_rng(0).random(50_000)
    x[::97] = np.nan
    m, s, c = f(x, 100)
    np.testing.assert_allclose(m, orolling.rolling(x, "mean", 100, 100, 1), rtol=1e-9, equal_nan=True)
    np.testing.assert_allclose(s, orolling.rolling(x, "std", 100, 1, 1), rtol=1e-7, atol=1e-12, equal_nan=True)
    np.testing.assert_array_equal(c, orolling.rolling(x, "count", 100, 0, 1))

    @numba.njit
    def g(a):
        return nx.parallel_stable_argsort(a, True), nx.parallel_stable_argsort(a, False), nx.parallel_sort(a.copy())

    for dt in (np.float64, np.int32, np.uint8, np.float32):
        a = (np.random.default_rng(1).standard_normal(20_000) * 50).astype(dt)
        up, down, srt = g(a)
        np.testing.assert_array_equal(up, osort.stable_argsort(a, True))
        np.testing.assert_array_equal(down, osort.stable_argsort(a, False))
        np.testing.assert_array_equal(srt, np.sort(a, kind="stable"))

    @numba.njit
    def bad(x):
        return nx.rolling_sum(x, 3, 5)

    with pytest.raises(ValueError, match="min_periods must be <= window"):
        bad(x)
