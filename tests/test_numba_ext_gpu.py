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


def test_jitted_groupby_and_merge():
    """df.groupby(k)[v].<agg>() and the merge indexers from nopython code on HOST arrays (sdcb200_host_groupby /
    sdcb200_host_join): parity with the groupby / join oracles."""
    from oracle import groupby as ogroupby
    from oracle import join as ojoin
    from sdc_b200 import numba_ext as nx

    @numba.njit
    def f(k, v):
        ks, s = nx.groupby_sum(k, v)
        _, m = nx.groupby_mean(k, v)
        _, c = nx.groupby_count(k, v)
        kf, mx = nx.groupby_max(k, v, False)
        return ks, s, m, c, kf, mx

    rng = np.random.default_rng(5)
    k = rng.integers(-20, 300, 80_000)
    v = rng.standard_normal(80_000)
    v[::53] = np.nan
    ks, s, m, c, kf, mx = f(k, v)
    for agg, got_k, got, sort in (("sum", ks, s, True), ("mean", ks, m, True), ("count", ks, c, True), ("max", kf, mx, False)):
        wk, want = ogroupby.groupby_agg(k, {"v": v}, agg, sort)
        np.testing.assert_array_equal(got_k, wk)
        if agg in ("count", "max"):
            np.testing.assert_array_equal(got, want["v"])
        else:
            np.testing.assert_allclose(got, want["v"], rtol=1e-9, atol=1e-12)
    assert c.dtype == np.int64 and s.dtype == np.float64

    @numba.njit
    def g(l, r):
        return nx.merge_indexers(l, r, "inner"), nx.merge_indexers(l, r, "left")

    l = rng.integers(0, 500, 30_000)
    r = rng.permutation(700)[:400].astype(np.int64)
    (li, ri), (ll, lr) = g(l, r)
    for how, a, b in (("inner", li, ri), ("left", ll, lr)):
        wl, wr = ojoin.join_indexers(l, r, how)
        got = sorted(zip(a.tolist(), b.tolist()))
        assert got == sorted(zip(np.asarray(wl).tolist(), np.asarray(wr).tolist())), how
