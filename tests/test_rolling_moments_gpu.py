"""GPU parity: csrc/rolling_moments.cu (rolling skew / kurt / cov / corr) against oracle/rolling_moments.py.

Tolerance.  Both sides evaluate the reference's formulas on sums of powers; they differ in where the sliding
sums are restarted (every 15 / 31 outputs on the GPU, at chunk starts in the oracle), so the results agree up to
the cancellation of those formulas: rtol 1e-7 on standard-normal data wherever the window's variance exceeds 1e-3
(m3 / m2^1.5 and m4 / m2^2 amplify the rounding of the sums by 1 / m2^1.5 and 1 / m2^2: a 3-row window with variance
1e-7 differs by 5e-6 between the two summation orders), NaN pattern exact on the same windows (the reference's
`m2 == 0` / `_almost_equal(var, 0)` decisions are rounding-dependent by construction)."""
import numpy as np
import pandas as pd
import pytest

from oracle.rolling_moments import rolling_moments as oracle_moments

pytestmark = pytest.mark.gpu

OPS = ("skew", "kurt", "cov", "corr")


def _gpu(x, op, win, minp, other=None, ddof=1):
    from sdc_b200.rolling import rolling_moments_host
    return rolling_moments_host(x, op, win, win if minp is None else minp, other, ddof)


def _cond_mask(x, win, other=None):
    """Outputs whose window variance (over the rows that are finite in BOTH series) is clear of zero."""
    n_out = len(x) if other is None else max(len(x), len(other))
    nmin = len(x) if other is None else min(len(x), len(other))
    m = np.ones(n_out, dtype=bool)
    if nmin:
        a = np.asarray(x[:nmin], dtype=np.float64)
        b = a if other is None else np.asarray(other[:nmin], dtype=np.float64)
        both = np.isfinite(a) & np.isfinite(b)
        for col in (a, b):
            v = pd.Series(np.where(both, col, np.nan)).rolling(win, min_periods=1).var(ddof=0).to_numpy()
            m[:nmin] &= ~(v <= 1e-3)
    return m


def _check(got, want, mask, msg, rtol=1e-7, atol=1e-9):
    assert got.shape == want.shape, msg
    assert np.array_equal(np.isnan(got[mask]), np.isnan(want[mask])), msg + " NaN pattern"
    m = mask & ~np.isnan(want) & ~np.isnan(got)
    if m.any():
        np.testing.assert_allclose(got[m], want[m], rtol=rtol, atol=atol, err_msg=msg)


@pytest.mark.parametrize("op", OPS)
@pytest.mark.parametrize("n", [1, 7, 1000, 100_003])
def test_random_against_oracle(op, n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n)
    y = 0.3 * x + rng.standard_normal(n)
    if n > 100:
        x[rng.integers(0, n, n // 50)] = np.nan
        y[rng.integers(0, n, n // 70)] = np.nan
        x[rng.integers(0, n, 5)] = np.inf
        y[rng.integers(0, n, 5)] = -np.inf
    for win, minp in ((1, 1), (3, 1), (5, None), (10, 4), (47, 47), (48, 20), (100, None), (1000, 500)):
        if win > 4 * n + 8:
            continue
        others = (None, y) if op in ("cov", "corr") else (None,)
        for other in others:
            want = oracle_moments(x, op, win, minp, other)
            got = _gpu(x, op, win, minp, other)
            _check(got, want, _cond_mask(x, win, other), "%s n=%d w=%d mp=%s other=%s" % (op, n, win, minp, other is not None))


@pytest.mark.parametrize("op", ["cov", "corr"])
def test_series_of_different_lengths(op):
    rng = np.random.default_rng(3)
    x = rng.standard_normal(20_000)
    y = rng.standard_normal(12_345)
    for a, b in ((x, y), (y, x)):
        for win in (10, 100):
            want = oracle_moments(a, op, win, None, b)
            got = _gpu(a, op, win, None, b)
            assert got.shape[0] == 20_000
            assert np.isnan(got[12_345 + win:]).all()          # past min(len) + window: NaN (reference :874-877)
            _check(got, want, _cond_mask(a, win, b), "%s w=%d" % (op, win))


def test_cov_ddof_window_zero_and_limits():
    x = np.random.default_rng(9).standard_normal(5000)
    for ddof in (0, 1, 3):
        _check(_gpu(x, "cov", 6, 2, None, ddof), oracle_moments(x, "cov", 6, 2, None, ddof), np.ones(5000, bool), "ddof=%d" % ddof)
    for op in OPS:
        assert np.isnan(_gpu(x, op, 0, 0)).all()
    from sdc_b200.rolling import rolling_moments_host
    with pytest.raises(Exception):
        rolling_moments_host(x, "skew", 5000, 1)               # beyond the tiled limit: an error, never a fallback


def test_mirror_matches_pandas():
    """Rolling.skew / kurt / cov / corr on pandas objects (index kept for one series, default index for two)."""
    from sdc_b200.rolling import rolling
    rng = np.random.default_rng(21)
    s = pd.Series(rng.standard_normal(3000), index=np.arange(3000) * 2, name="s")
    o = pd.Series(rng.standard_normal(3000))
    for op in ("skew", "kurt"):
        got = getattr(rolling(s, 20), op)()
        want = getattr(s.rolling(20), op)()
        assert got.index.equals(want.index) and got.name == "s"
        np.testing.assert_allclose(got.to_numpy(), want.to_numpy(), rtol=1e-7, atol=1e-9, equal_nan=True)
    np.testing.assert_allclose(rolling(s, 20).cov(o).to_numpy(), s.reset_index(drop=True).rolling(20).cov(o).to_numpy(),
                               rtol=1e-7, atol=1e-12, equal_nan=True)
    np.testing.assert_allclose(rolling(s, 20).corr(o).to_numpy(), s.reset_index(drop=True).rolling(20).corr(o).to_numpy(),
                               rtol=1e-7, atol=1e-12, equal_nan=True)
    df = pd.DataFrame({"a": rng.standard_normal(500), "b": rng.standard_normal(500)})
    got = rolling(df, 9, 3).kurt()
    want = df.rolling(9, 3).kurt()
    np.testing.assert_allclose(got.to_numpy(), want.to_numpy(), rtol=1e-7, atol=1e-9, equal_nan=True)
    with pytest.raises(TypeError):
        rolling(s, 5).cov(o, ddof=1.5)
