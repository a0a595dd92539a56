"""Pin the rolling oracle: example-script literals + pandas on the reference's test vectors."""
import numpy as np
import pytest

from helpers import assert_close, golden_inputs, rolling_golden_cases
from oracle import rolling as orolling


def test_examples_literal_outputs():
    # examples/series/rolling/series_rolling_*.py print 6 decimals
    for op, (data, win, want) in golden_inputs.EXAMPLES_ROLLING.items():
        # series_rolling_count.py's comment (1, 2, 3, 2, 2) is pandas<=1.x's count(), which
        # ignored the min_periods=None -> window default; the reference CODE applies it
        # (rolling_types.py:147-150 + result_or_nan) and yields NaN, NaN, 3, 2, 2 like
        # pandas>=2.  The literal is therefore checked with min_periods=0.
        minp = 0 if op == "count" else None
        got = orolling.rolling(np.asarray(data, dtype=np.float64), op, win, minp)
        assert_close(got, want, rtol=0, atol=6e-7, msg=op)


def test_example_dataframe_rolling_mean():
    ex = golden_inputs.EXAMPLE_DF_ROLLING_MEAN
    for col, data in ex["input"].items():
        got = orolling.rolling(np.asarray(data), "mean", ex["window"])
        assert_close(got, ex["expected"][col], rtol=0, atol=6e-7, msg=col)


@pytest.mark.parametrize("nchunks", [1, 3, 8])
def test_pandas_golden(nchunks):
    n = 0
    for op, data, w, mp, ddof, want in rolling_golden_cases():
        got = orolling.rolling(data, op, w, mp, ddof, nchunks=nchunks)
        assert_close(got, want, rtol=1e-9, atol=1e-9, msg="%s w=%d mp=%d ddof=%d" % (op, w, mp, ddof))
        n += 1
    assert n > 1000


def test_argument_errors():
    # messages pinned by sdc/tests/test_rolling.py:426-475
    with pytest.raises(ValueError, match="window must be non-negative"):
        orolling.check_rolling_args(-1)
    with pytest.raises(ValueError, match="min_periods must be >= 0"):
        orolling.check_rolling_args(3, -1)
    with pytest.raises(ValueError, match="min_periods must be <= window"):
        orolling.check_rolling_args(3, 4)
    with pytest.raises(ValueError, match="center"):
        orolling.check_rolling_args(3, 2, center=True)


def test_large_random_vs_pandas():
    import pandas as pd
    x = np.random.default_rng(1).random(200_000)
    x[::1000] = np.nan
    x[5::7777] = np.inf
    s = pd.Series(x)
    for op in ("sum", "mean", "std", "var", "min", "max"):
        want = getattr(s.rolling(100), op)().to_numpy()
        got = orolling.rolling(x, op, 100)
        # pandas treats inf as a value; the reference skips it (numpy.isfinite) -> compare
        # only where the window holds no inf
        has_inf = pd.Series(np.isinf(x).astype(float)).rolling(100, 1).sum().to_numpy() > 0
        assert_close(got[~has_inf], want[~has_inf], rtol=1e-9, atol=1e-9, msg=op)
