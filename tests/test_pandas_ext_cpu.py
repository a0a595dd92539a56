"""The typed Pandas surface for nopython code (sdc_b200/pandas_ext.py) without a GPU: (un)boxing round trips, the
reference's argument errors raised BEFORE any native call (hpat_pandas_rolling_types.py:144-170,
hpat_pandas_series_functions.py:4797-4798), typing errors, and no CPU fallback behind a valid call."""
import numba
import numpy as np
import pandas as pd
import pytest
from numba.core.errors import TypingError

from sdc_b200 import pandas_ext  # noqa: F401  (registers the types)


@numba.njit
def _ident(x):
    return x


def test_series_and_dataframe_round_trip():
    s = pd.Series([1.0, np.nan, 3.5])
    pd.testing.assert_series_equal(_ident(s), s)
    si = pd.Series([3, 1, 2], index=[10, 20, 30])
    pd.testing.assert_series_equal(_ident(si), si)
    df = pd.DataFrame({"A": [1, 2, 1], "B": [0.5, 1.5, 2.5]}, index=[0.5, 1.5, 2.5])
    pd.testing.assert_frame_equal(_ident(df), df)
    # a strided view is made contiguous on the way in
    base = pd.DataFrame(np.arange(12.0).reshape(4, 3), columns=list("xyz"))
    pd.testing.assert_frame_equal(_ident(base), base)
    with pytest.raises(TypingError):
        _ident(pd.Series(["a", "b"]))


def test_rolling_argument_errors_before_any_native_call():
    s = pd.Series(np.arange(10.0))

    @numba.njit
    def f(s, w, mp):
        return s.rolling(w, mp).sum()

    with pytest.raises(ValueError, match="window must be non-negative"):
        f(s, -1, 0)
    with pytest.raises(ValueError, match="min_periods must be >= 0"):
        f(s, 3, -1)
    with pytest.raises(ValueError, match="min_periods must be <= window"):
        f(s, 3, 5)

    @numba.njit
    def g(s):
        return s.rolling(3, center=True).mean()

    with pytest.raises(ValueError, match="The object center"):
        g(s)

    @numba.njit
    def h(s):
        return s.rolling(2.5).mean()

    with pytest.raises(TypingError):
        h(s)


def test_groupby_argument_errors():
    s = pd.Series(np.arange(6.0))

    @numba.njit
    def f(s, k):
        return s.groupby(k).sum()

    with pytest.raises(ValueError, match="array of the same length"):
        f(s, np.array([1, 2, 3]))
    df = pd.DataFrame({"A": [1, 2], "B": [1.0, 2.0]})

    @numba.njit
    def g(df):
        return df.groupby("Z").sum()

    with pytest.raises(TypingError):
        g(df)

    @numba.njit
    def twice(df):
        df.groupby("A")["B"]["B"]          # a second [] raises (hpat_pandas_groupby_functions.py:169-170)
        return 0

    with pytest.raises(IndexError, match="Columns already selected"):
        twice(df)


def test_valid_call_without_a_gpu_raises_instead_of_falling_back():
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a box without a GPU")

    @numba.njit
    def f(s):
        return s.rolling(3).mean()

    with pytest.raises(RuntimeError, match="libsdc_b200"):
        f(pd.Series(np.arange(10.0)))

    @numba.njit
    def g(df):
        return df.groupby("A").sum()

    with pytest.raises(RuntimeError, match="libsdc_b200"):
        g(pd.DataFrame({"A": [1, 2, 1], "B": [1.0, 2.0, 3.0]}))
