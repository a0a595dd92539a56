"""Pin the skew / kurt / cov / corr oracle (oracle/rolling_moments.py) against pandas on the reference's test
vectors and on random data, wherever the window is not degenerate (the reference's NaN decisions on a zero
variance -- `m2 == 0`, `_almost_equal(var, 0)` -- depend on the rounding of the sums)."""
import numpy as np
import pandas as pd
import pytest

from helpers import golden_inputs
from oracle.rolling_moments import rolling_moments


def _pandas(x, op, win, minp, other=None, ddof=1):
    r = pd.Series(x).rolling(win, min_periods=minp)
    if op == "cov":
        return r.cov(None if other is None else pd.Series(other), ddof=ddof).to_numpy()
    if op == "corr":
        return r.corr(None if other is None else pd.Series(other)).to_numpy()
    return getattr(r, op)().to_numpy()


def _well_conditioned(x, win, other=None):
    """Windows whose variance is far from zero in both series and that hold no +-inf (pandas keeps inf as a
    value in some of these reductions; the reference skips every non-finite row)."""
    def ok(a):
        a = np.asarray(a, dtype=np.float64)
        fin = np.where(np.isfinite(a), a, np.nan)
        v = pd.Series(fin).rolling(win, min_periods=1).var(ddof=0).to_numpy()
        inf = pd.Series(np.isinf(a).astype(float)).rolling(win, min_periods=1).sum().to_numpy() > 0
        return (v > 1e-6) & ~inf
    m = ok(x)
    if other is not None:
        n = min(len(x), len(other))
        m2 = np.zeros(max(len(x), len(other)), dtype=bool)
        m2[:n] = ok(x[:n])[:n] & ok(other[:n])[:n]
        return m2
    return m


def _check(got, want, mask, msg, rtol):
    mask = mask & ~(np.isnan(got) & np.isnan(want))
    assert np.array_equal(np.isnan(got[mask]), np.isnan(want[mask])), msg + " NaN pattern"
    m = mask & ~np.isnan(want)
    if m.any():
        np.testing.assert_allclose(got[m], want[m], rtol=rtol, atol=1e-9, err_msg=msg)


@pytest.mark.parametrize("op", ["skew", "kurt", "cov", "corr"])
def test_reference_vectors_vs_pandas(op):
    # sdc/tests/test_rolling.py:1606-1610 / test_utils.py:56-62 data, the windows of :820-868
    series = [np.asarray(d, dtype=np.float64) for d in golden_inputs.ROLLING_SUM_DATA + golden_inputs.GLOBAL_FLOAT64]
    n = 0
    for x in series:
        if len(x) < 4 or not np.isfinite(x).any() or np.nanmax(np.abs(np.where(np.isfinite(x), x, 0))) > 1e6:
            continue                   # huge magnitudes: the power sums cancel completely, nothing to pin
        for win in (3, 4, 5, 8):
            for minp in sorted({win, max(1, win // 2)}):
                want = _pandas(x, op, win, minp)
                got = rolling_moments(x, op, win, minp)
                _check(got, want, _well_conditioned(x, win), "%s w=%d mp=%d" % (op, win, minp), 1e-6)
                n += 1
    assert n > 8


@pytest.mark.parametrize("op", ["skew", "kurt", "cov", "corr"])
@pytest.mark.parametrize("nchunks", [1, 7])
def test_random_vs_pandas(op, nchunks):
    rng = np.random.default_rng(5)
    n = 5000
    x = rng.standard_normal(n)
    y = 0.5 * x + rng.standard_normal(n)
    x[rng.integers(0, n, 60)] = np.nan
    y[rng.integers(0, n, 60)] = np.nan
    for win, minp in ((10, None), (10, 4), (100, 50), (7, 7)):
        others = (None, y, y[:3000]) if op in ("cov", "corr") else (None,)
        for other in others:
            want = _pandas(x, op, win, minp, other)
            got = rolling_moments(x, op, win, minp, other, nchunks=nchunks)
            assert got.shape == want.shape
            _check(got, want, _well_conditioned(x, win, other), "%s w=%d mp=%s" % (op, win, minp), 1e-7)


def test_cov_ddof_and_empty_window():
    x = np.arange(20.0) ** 1.5
    for ddof in (0, 1, 2):
        want = _pandas(x, "cov", 5, 5, None, ddof)
        np.testing.assert_allclose(rolling_moments(x, "cov", 5, 5, None, ddof), want, rtol=1e-9, equal_nan=True)
    assert np.isnan(rolling_moments(x, "cov", 0, 0)).all()           # win == 0: nfinite 0 < max(1, minp)
    assert np.isnan(rolling_moments(x, "skew", 2, 1)).all()          # fewer than 3 rows never give a skew
