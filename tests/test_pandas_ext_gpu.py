"""The Pandas surface of the path inside nopython code (sdc_b200/pandas_ext.py): jitted `s.rolling(100).mean()`,
`df.groupby('A').sum()`, ... against the oracle and pandas -- what north_star calls "the @sdc.jit / Pandas API surface
stays exactly as it is"."""
import numba
import numpy as np
import pandas as pd
import pytest

from oracle import groupby as ogroupby
from oracle import rolling as orolling

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _register():
    from sdc_b200 import pandas_ext  # noqa: F401


def test_jitted_series_rolling_against_oracle():
    @numba.njit
    def f(s):
        return s.rolling(100).mean(), s.rolling(100, 10).std(), s.rolling(7, 1).max(), s.rolling(100).var(0)

    x = np.random.default_rng(0).random(200_000)
    x[::977] = np.nan
    s = pd.Series(x, index=np.arange(5, 5 + len(x)))
    mean, std, mx, var0 = f(s)
    for got, op, w, mp, ddof in ((mean, "mean", 100, 100, 1), (std, "std", 100, 10, 1), (mx, "max", 7, 1, 1), (var0, "var", 100, 100, 0)):
        want = orolling.rolling(x, op, w, mp, ddof)
        assert isinstance(got, pd.Series)
        np.testing.assert_array_equal(got.index.to_numpy(), s.index.to_numpy())      # index preserved (:631-632)
        np.testing.assert_allclose(got.to_numpy(), want, rtol=1e-9, atol=1e-12, equal_nan=True)
    # the reference's example: [4, 3, 5, 2, 6].rolling(3).mean() -> NaN, NaN, 4.0, 3.333333, 4.333333
    ex = numba.njit(lambda s: s.rolling(3).mean())(pd.Series([4.0, 3.0, 5.0, 2.0, 6.0]))
    np.testing.assert_allclose(ex.to_numpy(), [np.nan, np.nan, 4.0, 10 / 3, 13 / 3], rtol=1e-12, equal_nan=True)


def test_jitted_dataframe_groupby_against_oracle_and_pandas():
    rng = np.random.default_rng(1)
    n = 300_000
    df = pd.DataFrame({"A": rng.integers(0, 1000, n), "B": rng.random(n), "C": rng.integers(-50, 50, n)})
    df.loc[::53, "B"] = np.nan

    @numba.njit
    def f(df):
        return df.groupby("A").sum(), df.groupby("A").mean(), df.groupby("A", sort=False).max(), df.groupby("A").count()

    gsum, gmean, gmax, gcnt = f(df)
    for got, agg, sort in ((gsum, "sum", True), (gmean, "mean", True), (gmax, "max", False), (gcnt, "count", True)):
        wk, want = ogroupby.groupby_agg(df["A"].to_numpy(), {"B": df["B"].to_numpy(), "C": df["C"].to_numpy()}, agg, sort)
        assert list(got.columns) == ["B", "C"]
        np.testing.assert_array_equal(got.index.to_numpy(), wk)
        for c in ("B", "C"):
            if agg in ("max", "count") or (c == "C" and agg == "sum"):
                np.testing.assert_array_equal(got[c].to_numpy(), want[c])
            else:
                np.testing.assert_allclose(got[c].to_numpy(), want[c], rtol=1e-9, atol=1e-12)
    pd.testing.assert_frame_equal(gsum, df.groupby("A").sum().rename_axis(None), check_exact=False, rtol=1e-9)
    # the reference's example (examples/dataframe/groupby/dataframe_groupby_sum.py): B [0, 6, 14], C [5, 16, 24], index [1, 2, 3]
    ex = pd.DataFrame({"A": [1, 1, 2, 2, 3, 3, 3], "B": [0, 0, 2, 4, 3, 5, 6], "C": [2, 3, 9, 7, 6, 8, 10]})
    got = numba.njit(lambda d: d.groupby("A").sum())(ex)
    np.testing.assert_array_equal(got.index.to_numpy(), [1, 2, 3])
    np.testing.assert_array_equal(got["B"].to_numpy(), [0, 6, 14])
    np.testing.assert_array_equal(got["C"].to_numpy(), [5, 16, 24])


def test_jitted_column_selection_series_groupby_dataframe_rolling():
    rng = np.random.default_rng(2)
    n = 50_000
    df = pd.DataFrame({"k": rng.integers(0, 40, n), "x": rng.standard_normal(n), "y": rng.standard_normal(n)})

    @numba.njit
    def f(df, s, by):
        return df.groupby("k")["x"].mean(), df.groupby("k")["y", "x"].min(), s.groupby(by).std(), df.rolling(20, 5).mean()

    one, two, sg, roll = f(df, df["x"], df["k"].to_numpy())
    pd.testing.assert_series_equal(one, df.groupby("k")["x"].mean().rename_axis(None).rename(None), check_exact=False, rtol=1e-9)
    assert list(two.columns) == ["y", "x"]
    np.testing.assert_array_equal(two["y"].to_numpy(), df.groupby("k")["y"].min().to_numpy())
    np.testing.assert_allclose(sg.to_numpy(), df["x"].groupby(df["k"].to_numpy()).std().to_numpy(), rtol=1e-9)
    want = df.rolling(20, 5).mean()
    for c in df.columns:
        np.testing.assert_allclose(roll[c].to_numpy(), want[c].to_numpy(), rtol=1e-9, atol=1e-12, equal_nan=True)


def test_jitted_sort_values_and_merge():
    rng = np.random.default_rng(3)
    v = rng.standard_normal(30_000).round(1)
    v[::41] = np.nan
    s = pd.Series(v)

    @numba.njit
    def f(s):
        return s.sort_values(), s.sort_values(ascending=False, kind="mergesort", na_position="first")

    up, down = f(s)
    pd.testing.assert_series_equal(up, s.sort_values(kind="mergesort"))
    pd.testing.assert_series_equal(down, s.sort_values(ascending=False, kind="mergesort", na_position="first"))

    left = pd.DataFrame({"k": rng.integers(0, 500, 20_000), "a": rng.random(20_000)})
    right = pd.DataFrame({"k": rng.permutation(700)[:400].astype(np.int64), "b": rng.random(400), "c": rng.integers(0, 9, 400)})

    @numba.njit
    def g(l, r):
        return pd.merge(l, r, on="k", how="inner"), pd.merge(l, r, on="k", how="left")

    inner, lj = g(left, right)
    for got, how in ((inner, "inner"), (lj, "left")):
        want = pd.merge(left, right, on="k", how=how)
        key = ["k", "a"]
        got_s = got.sort_values(key).reset_index(drop=True)
        want_s = want.sort_values(key).reset_index(drop=True)
        want_s["c"] = want_s["c"].astype(got_s["c"].dtype)
        pd.testing.assert_frame_equal(got_s, want_s, check_exact=False, rtol=1e-12)
