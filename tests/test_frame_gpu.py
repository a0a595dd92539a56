"""GPU: device-resident DeviceFrame / DeviceSeries (sdc_b200/frame.py) chain the kernels without leaving HBM and agree
with pandas on the reference's test shapes (groupby: sdc/tests/test_groupby.py:58-64; sort_values:
sdc/tests/test_dataframe.py:908-1000; merge: sdc/tests/test_join.py:67-81, 237-335; rolling: sdc/tests/test_rolling.py)."""
import numpy as np
import pandas as pd
import pytest

from helpers import golden_inputs

pytestmark = pytest.mark.gpu


def _frames():
    from sdc_b200.frame import DeviceFrame
    rng = np.random.default_rng(3)
    n = 50_000
    df = pd.DataFrame({"k": rng.integers(0, 300, n), "g": rng.integers(0, 40, n), "x": rng.standard_normal(n),
                       "i": rng.integers(-50, 50, n)})
    df.loc[rng.integers(0, n, 500), "x"] = np.nan
    return df, DeviceFrame.from_pandas(df)


def test_round_trip_and_getitem():
    from sdc_b200.frame import DeviceFrame
    df = pd.DataFrame({"a": [3, 1, 2], "s": ["x", None, "zz"], "f": [0.5, np.nan, -1.0]})
    dev = DeviceFrame.from_pandas(df)
    back = dev.to_pandas()
    assert back["a"].tolist() == [3, 1, 2]
    assert [None if pd.isna(v) else v for v in back["s"]] == ["x", None, "zz"]      # pandas shows a missing string as NaN
    np.testing.assert_array_equal(back["f"].to_numpy(), df["f"].to_numpy())
    assert dev[["a", "f"]].columns == ["a", "f"] and len(dev["a"]) == 3


@pytest.mark.parametrize("agg", ["sum", "mean", "min", "max", "count", "std", "median", "prod"])
def test_groupby(agg):
    df, dev = _frames()
    got = getattr(dev[["g", "x", "i"]].groupby("g"), agg)().to_pandas()
    want = getattr(df[["g", "x", "i"]].groupby("g"), agg)()
    pd.testing.assert_frame_equal(got, want, check_dtype=False, rtol=1e-9, check_exact=False)
    # string key, first-appearance order
    ds = pd.DataFrame({"s": np.array(["a", "bb", None, "ccc"], dtype=object)[np.arange(len(df)) % 4], "x": df["x"]})
    from sdc_b200.frame import DeviceFrame
    got = DeviceFrame.from_pandas(ds).groupby("s", sort=False).sum().to_pandas()
    want = ds.groupby("s", sort=False).sum()
    pd.testing.assert_frame_equal(got, want, check_dtype=False, rtol=1e-9)


@pytest.mark.parametrize("how", ["inner", "left", "right", "outer"])
def test_merge_then_groupby_stays_on_device(how):
    from sdc_b200.frame import DeviceFrame
    df, dev = _frames()
    rng = np.random.default_rng(8)
    dim = pd.DataFrame({"k": rng.permutation(400)[:250], "w": rng.random(250), "i": rng.integers(0, 9, 250)})
    got = dev.merge(DeviceFrame.from_pandas(dim), how=how, on="k").to_pandas()
    want = df.merge(dim, how=how, on="k")
    cols = list(want.columns)
    assert list(got.columns) == cols
    pd.testing.assert_frame_equal(got.sort_values(cols).reset_index(drop=True), want.sort_values(cols).reset_index(drop=True),
                                  check_dtype=False)
    if how == "inner":
        chained = dev.merge(DeviceFrame.from_pandas(dim), on="k")[["g", "x", "w"]].groupby("g").sum().to_pandas()
        pd.testing.assert_frame_equal(chained, df.merge(dim, on="k")[["g", "x", "w"]].groupby("g").sum(), rtol=1e-9)


@pytest.mark.parametrize("ascending", [True, False])
@pytest.mark.parametrize("na_position", ["last", "first"])
def test_sort_values(ascending, na_position):
    df, dev = _frames()
    got = dev.sort_values("x", ascending=ascending, na_position=na_position).to_pandas()
    want = df.sort_values("x", ascending=ascending, na_position=na_position, kind="mergesort").reset_index(drop=True)
    pd.testing.assert_frame_equal(got, want, check_dtype=False)
    # string column as the sort key
    from sdc_b200.frame import DeviceFrame
    ds = pd.DataFrame({"s": ["b", None, "a", "ab", "", "b", None], "v": np.arange(7)})
    got = DeviceFrame.from_pandas(ds).sort_values("s", ascending=ascending, na_position=na_position).to_pandas()
    want = ds.sort_values("s", ascending=ascending, na_position=na_position, kind="mergesort").reset_index(drop=True)
    norm = lambda col: [None if pd.isna(v) else v for v in col]
    assert norm(got["s"]) == norm(want["s"]) and got["v"].tolist() == want["v"].tolist()
    s_sorted, idx = dev["i"].sort_values(ascending=ascending)
    np.testing.assert_array_equal(s_sorted.to_pandas().to_numpy(), np.sort(df["i"].to_numpy(), kind="stable")[::1 if ascending else -1])


def test_rolling():
    df, dev = _frames()
    got = dev[["x", "i"]].rolling(100).mean().to_pandas()
    want = df[["x", "i"]].rolling(100).mean()
    pd.testing.assert_frame_equal(got, want, rtol=1e-9, atol=1e-9)
    got = dev["x"].rolling(7, min_periods=2).std().to_pandas()
    pd.testing.assert_series_equal(got, df["x"].rolling(7, min_periods=2).std(), rtol=1e-9, atol=1e-9)
    with pytest.raises(ValueError):
        dev.rolling(3, center=True)


@pytest.mark.parametrize("sort", [True, False])
def test_groupby_several_keys(sort, monkeypatch):
    """groupby([k1, k2]) -- the shape of sdc/tests/test_groupby.py:526-558 (skipped there: one `by` column only)."""
    from sdc_b200 import frame
    from sdc_b200.frame import DeviceFrame
    rng = np.random.default_rng(12)
    n = 40_000
    df = pd.DataFrame({"a": rng.integers(-5, 20, n), "b": rng.integers(0, 7, n).astype(np.float64) / 2,
                       "c": rng.integers(100, 103, n), "x": rng.standard_normal(n), "i": rng.integers(-9, 9, n)})
    df.loc[rng.integers(0, n, 300), "b"] = np.nan          # NaN keys drop the row (pandas dropna=True)
    df.loc[rng.integers(0, n, 300), "x"] = np.nan
    dev = DeviceFrame.from_pandas(df)
    for keys in (["a", "b"], ["c", "a", "b"]):
        for agg in ("sum", "mean", "min", "count"):
            got = getattr(dev.groupby(keys, sort=sort), agg)().to_pandas()
            want = getattr(df.groupby(keys, sort=sort), agg)()
            if not sort:
                got, want = got.sort_index(), want.sort_index()
            pd.testing.assert_frame_equal(got, want, check_dtype=False, rtol=1e-9, check_index_type=False)
    if sort:
        # beyond the direct-table limit the composite ids are aggregated as plain int64 keys
        monkeypatch.setattr(frame._DeviceGroupByMulti, "_GID_LIMIT", 16)
        got = dev.groupby(["a", "b"]).sum().to_pandas()
        pd.testing.assert_frame_equal(got, df.groupby(["a", "b"]).sum(), check_dtype=False, rtol=1e-9, check_index_type=False)
        first = dev.groupby(["a", "b"], sort=False).agg(["sum", "max"]).to_pandas()
        assert list(first.columns) == ["c_sum", "x_sum", "i_sum", "c_max", "x_max", "i_max"]


@pytest.mark.parametrize("how", ["inner", "left", "right", "outer"])
def test_merge_on_several_keys(how):
    """merge(on=[k1, k2]) -- the shape of sdc/tests/test_join.py:106-160 (skipped there)."""
    from sdc_b200.frame import DeviceFrame
    rng = np.random.default_rng(31)
    n, m = 30_000, 500
    left = pd.DataFrame({"k1": rng.integers(0, 30, n), "k2": rng.integers(0, 25, n), "x": rng.standard_normal(n)})
    pairs = rng.permutation(40 * 30)[:m]                   # unique (k1, k2) pairs, some absent on the left
    right = pd.DataFrame({"k1": pairs // 30, "k2": pairs % 30, "w": rng.random(m)})
    got = DeviceFrame.from_pandas(left).merge(DeviceFrame.from_pandas(right), how=how, on=["k1", "k2"]).to_pandas()
    want = left.merge(right, how=how, on=["k1", "k2"])
    cols = list(want.columns)
    assert list(got.columns) == cols
    pd.testing.assert_frame_equal(got.sort_values(cols).reset_index(drop=True), want.sort_values(cols).reset_index(drop=True),
                                  check_dtype=False)
    # duplicate pairs on the build side: cross product per pair
    right2 = pd.concat([right, right.assign(w=right["w"] + 1)], ignore_index=True)
    got = DeviceFrame.from_pandas(left).merge(DeviceFrame.from_pandas(right2), how=how, left_on=["k1", "k2"],
                                              right_on=["k1", "k2"]).to_pandas()
    want = left.merge(right2, how=how, on=["k1", "k2"])
    pd.testing.assert_frame_equal(got.sort_values(cols).reset_index(drop=True), want.sort_values(cols).reset_index(drop=True),
                                  check_dtype=False)


def test_rolling_moments_on_device_frames():
    df, dev = _frames()
    got = dev[["x"]].rolling(30, 10).skew().to_pandas()
    pd.testing.assert_frame_equal(got, df[["x"]].rolling(30, 10).skew(), rtol=1e-7, atol=1e-9)
    other = dev["i"]
    got = dev["x"].rolling(50).corr(other).to_pandas()
    want = df["x"].rolling(50).corr(df["i"])
    np.testing.assert_allclose(got.to_numpy(), want.to_numpy(), rtol=1e-7, atol=1e-9, equal_nan=True)


@pytest.mark.parametrize("kind", ["int", "float"])
def test_merge_asof(kind):
    """pd.merge_asof(df1, df2, on='time') -- sdc/tests/test_join.py:237-262 (skipped there)."""
    from sdc_b200.frame import DeviceFrame
    rng = np.random.default_rng(17)
    t1 = np.sort(rng.integers(0, 10_000, 20_000))
    t2 = np.sort(rng.integers(500, 9_000, 3_000))
    if kind == "float":
        t1, t2 = t1 / 4.0, t2 / 4.0
    left = pd.DataFrame({"time": t1, "a": rng.standard_normal(len(t1))})
    right = pd.DataFrame({"time": t2, "b": rng.standard_normal(len(t2)), "a": rng.integers(0, 9, len(t2))})
    got = DeviceFrame.from_pandas(left).merge_asof(DeviceFrame.from_pandas(right), on="time").to_pandas()
    want = pd.merge_asof(left, right, on="time")
    pd.testing.assert_frame_equal(got, want, check_dtype=False)
    with pytest.raises(ValueError):
        DeviceFrame.from_pandas(left.iloc[::-1].reset_index(drop=True)).merge_asof(DeviceFrame.from_pandas(right), on="time")


def test_arrow_csv_parquet_ingest(tmp_path):
    """CSV / Parquet -> Arrow -> device columns, then the kernels, then back."""
    import pyarrow as pa
    import pyarrow.parquet as pq
    from sdc_b200.frame import DeviceFrame
    rng = np.random.default_rng(4)
    n = 5000
    df = pd.DataFrame({"k": rng.integers(0, 50, n), "s": np.array(["aa", "b", "ccc", "dddd"], dtype=object)[rng.integers(0, 4, n)],
                       "x": rng.standard_normal(n), "flag": rng.integers(0, 2, n).astype(bool)})
    df.loc[rng.integers(0, n, 60), "x"] = np.nan
    df.loc[rng.integers(0, n, 40), "s"] = None
    csv, parquet = tmp_path / "t.csv", tmp_path / "t.parquet"
    df.to_csv(csv, index=False)
    pq.write_table(pa.Table.from_pandas(df, preserve_index=False), parquet)
    for dev in (DeviceFrame.read_csv(str(csv)), DeviceFrame.read_parquet(str(parquet))):
        assert dev.columns == ["k", "s", "x", "flag"] and len(dev) == n
        back = dev.to_pandas()
        np.testing.assert_array_equal(back["k"].to_numpy(), df["k"].to_numpy())
        np.testing.assert_allclose(back["x"].to_numpy(), df["x"].to_numpy(), rtol=1e-15, equal_nan=True)
        assert [None if pd.isna(v) else v for v in back["s"]] == [None if pd.isna(v) else v for v in df["s"]]
        got = dev[["s", "x"]].groupby("s").sum().to_pandas()
        pd.testing.assert_frame_equal(got, df[["s", "x"]].groupby("s").sum(), check_dtype=False, rtol=1e-9)
        tab = dev.to_arrow()
        assert tab.column("s").to_pylist() == [None if pd.isna(v) else v for v in df["s"]]
