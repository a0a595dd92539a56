"""Pin the groupby oracle: example-script literals + pandas on the reference's test inputs."""
import numpy as np
import pandas as pd
import pytest

from helpers import assert_close, golden_inputs
from oracle import groupby as ogroupby


def _cols(df, key):
    return {c: df[c].to_numpy() for c in df.columns if c != key}


def test_example_literals():
    # examples/dataframe/groupby/dataframe_groupby_{sum,mean,min,max,std,var,count}.py
    for agg, want in golden_inputs.EXAMPLES_GROUPBY.items():
        src = golden_inputs.EXAMPLE_GROUPBY_COUNT_DF if agg == "count" else golden_inputs.EXAMPLE_GROUPBY_DF
        df = pd.DataFrame(src)
        keys, res = ogroupby.groupby_agg(df["A"].to_numpy(), _cols(df, "A"), agg)
        np.testing.assert_array_equal(keys, [1, 2, 3])
        for c in ("B", "C"):
            assert_close(res[c], want[c], rtol=0, atol=6e-7, msg="%s %s" % (agg, c))
    ex = golden_inputs.EXAMPLE_SERIES_GROUPBY            # examples/series/series_groupby.py
    keys, res = ogroupby.groupby_agg(np.asarray(ex["by"]), {"S": np.asarray(ex["S"])}, "mean")
    np.testing.assert_array_equal(keys, ex["index"])
    assert_close(res["S"], ex["mean"], rtol=1e-15, atol=0)


@pytest.mark.parametrize("agg", ogroupby.AGGS)
@pytest.mark.parametrize("nchunks", [1, 3])
def test_default_df_against_pandas(agg, nchunks):
    # sdc/tests/test_groupby.py:58-64, 123-132, 152-212, 244-316
    df = pd.DataFrame(golden_inputs.GROUPBY_DEFAULT_DF)
    keys, res = ogroupby.groupby_agg(df["A"].to_numpy(), _cols(df, "A"), agg, nchunks=nchunks)
    want = getattr(df.groupby("A"), agg)()
    np.testing.assert_array_equal(keys, want.index.to_numpy())
    for c in want.columns:
        w = want[c].to_numpy()
        g = res[c]
        if agg in ("min", "max", "sum", "mean", "var", "std") and c == "E":
            # pandas propagates inf-inf -> NaN the same way; only compare where finite
            pass
        assert_close(np.asarray(g, dtype=np.float64), np.asarray(w, dtype=np.float64), rtol=1e-12, atol=1e-12,
                     msg="%s %s" % (agg, c))
        if agg in ("sum", "min", "max") and c == "B" or agg == "count":
            assert g.dtype == np.int64, (agg, c, g.dtype)


def test_key_dtypes_count():
    # sdc/tests/test_groupby.py:83-100: int / float-with-NaN / string-with-None keys
    for kind, key in golden_inputs.GROUPBY_KEY_DTYPES.items():
        df = pd.DataFrame(golden_inputs.GROUPBY_DEFAULT_DF)
        df["A"] = key
        k = df["A"].to_numpy() if kind != "string" else list(key)
        keys, res = ogroupby.groupby_agg(k, _cols(df, "A"), "count")
        want = df.groupby("A").count()
        assert list(keys) == list(want.index), kind
        for c in want.columns:
            np.testing.assert_array_equal(res[c], want[c].to_numpy(), err_msg="%s %s" % (kind, c))


def test_sort_false_is_first_appearance_order():
    rng = np.random.RandomState(0)
    a = rng.choice(np.arange(20), 1000)
    keys, res = ogroupby.groupby_agg(a, {"B": np.arange(1000)}, "min", sort=False, nchunks=4)
    want = pd.DataFrame({"A": a, "B": np.arange(1000)}).groupby("A", sort=False).min()
    np.testing.assert_array_equal(keys, want.index.to_numpy())
    np.testing.assert_array_equal(res["B"], want["B"].to_numpy())


def test_random_against_pandas():
    rng = np.random.default_rng(0)
    n = 200_000
    a = rng.integers(0, 1000, n)
    b = rng.random(n)
    b[rng.integers(0, n, n // 100)] = np.nan
    df = pd.DataFrame({"A": a, "B": b})
    for agg in ogroupby.AGGS:
        keys, res = ogroupby.groupby_agg(a, {"B": b}, agg)
        want = getattr(df.groupby("A"), agg)()
        np.testing.assert_array_equal(keys, want.index.to_numpy())
        assert_close(res["B"].astype(np.float64), want["B"].to_numpy().astype(np.float64), rtol=1e-9, atol=1e-12, msg=agg)
