"""GPU parity: csrc/groupby.cu (hash groupby-aggregate) against the oracle's dict-of-row-lists restatement.

Keys, counts, min/max and integer sums are compared exactly; float sum/mean/var/std to 1e-9 relative
(north_star) -- the reduction order differs (atomics vs the reference's per-group serial loop).
"""
import numpy as np
import pandas as pd
import pytest

from helpers import assert_close, golden_inputs
from oracle import groupby as ogroupby

pytestmark = pytest.mark.gpu

EXACT = ("min", "max", "count")


def _compare(got_keys, got, want_keys, want, agg, msg=""):
    np.testing.assert_array_equal(np.asarray(got_keys), np.asarray(want_keys), err_msg=msg + " keys")
    for c in want:
        g, w = np.asarray(got[c]), np.asarray(want[c])
        assert g.dtype == w.dtype, "%s %s %s dtype %s vs %s" % (msg, agg, c, g.dtype, w.dtype)
        if agg in EXACT or g.dtype.kind in "iu":
            np.testing.assert_array_equal(g, w, err_msg="%s %s %s" % (msg, agg, c))
        else:
            # inf - inf style NaNs appear on both sides the same way (sum of +inf and -inf)
            assert_close(g, w, rtol=1e-9, atol=1e-12, msg="%s %s %s" % (msg, agg, c))


def _host(keys, cols, agg, sort=True, ddof=1, expected_groups=0):
    from sdc_b200.groupby import groupby_host
    rk, res = groupby_host(keys, cols, (agg,), sort, ddof, expected_groups)
    return rk, res[agg]


def test_example_literals():
    # examples/dataframe/groupby/dataframe_groupby_{sum,mean,min,max,std,var,count}.py
    for agg, want in golden_inputs.EXAMPLES_GROUPBY.items():
        src = golden_inputs.EXAMPLE_GROUPBY_COUNT_DF if agg == "count" else golden_inputs.EXAMPLE_GROUPBY_DF
        df = pd.DataFrame(src)
        keys, res = _host(df["A"].to_numpy(), {c: df[c].to_numpy() for c in "BC"}, agg)
        np.testing.assert_array_equal(keys, [1, 2, 3])
        for c in ("B", "C"):
            assert_close(res[c], want[c], rtol=0, atol=6e-7, msg="%s %s" % (agg, c))
    ex = golden_inputs.EXAMPLE_SERIES_GROUPBY            # examples/series/series_groupby.py
    keys, res = _host(np.asarray(ex["by"]), {"S": np.asarray(ex["S"])}, "mean")
    np.testing.assert_array_equal(keys, ex["index"])
    assert_close(res["S"], ex["mean"], rtol=1e-15, atol=0)


@pytest.mark.parametrize("agg", ogroupby.AGGS)
@pytest.mark.parametrize("sort", [True, False])
def test_default_df(agg, sort):
    # sdc/tests/test_groupby.py:58-64, 123-132, 152-212, 244-316: int and float columns with NaN / +-inf
    df = pd.DataFrame(golden_inputs.GROUPBY_DEFAULT_DF)
    cols = {c: df[c].to_numpy() for c in "BCDE"}
    gk, got = _host(df["A"].to_numpy(), cols, agg, sort)
    wk, want = ogroupby.groupby_agg(df["A"].to_numpy(), cols, agg, sort)
    _compare(gk, got, wk, want, agg)


def test_key_dtypes_float_nan_keys_skipped():
    # sdc/tests/test_groupby.py:83-100
    df = pd.DataFrame(golden_inputs.GROUPBY_DEFAULT_DF)
    for kind in ("int", "float"):
        key = np.asarray(golden_inputs.GROUPBY_KEY_DTYPES[kind])
        cols = {c: df[c].to_numpy() for c in "BCDE"}
        for agg in ("count", "sum", "max"):
            gk, got = _host(key, cols, agg)
            wk, want = ogroupby.groupby_agg(key, cols, agg)
            assert gk.dtype == key.dtype
            _compare(gk, got, wk, want, agg, kind)
    # -0.0 and +0.0 are one group; +-inf are ordinary keys
    key = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1.5, -0.0, np.inf])
    v = {"v": np.arange(8.0)}
    gk, got = _host(key, v, "sum")
    wk, want = ogroupby.groupby_agg(key, v, "sum")
    _compare(gk, got, wk, want, "sum", "zeros")
    assert len(gk) == 4


def test_sort_flag_1000_rows_20_groups():
    # sdc/tests/test_groupby.py:102-121 (seed 0); sort=False = first-appearance order (dict order)
    rng = np.random.RandomState(0)
    a = rng.choice(np.arange(20), 1000)
    cols = {"B": np.arange(1000)}
    for sort in (True, False):
        gk, got = _host(a, cols, "min", sort)
        wk, want = ogroupby.groupby_agg(a, cols, "min", sort, nchunks=4)
        _compare(gk, got, wk, want, "min", "sort=%s" % sort)


@pytest.mark.parametrize("n,groups", [(1, 1), (31, 7), (4095, 1000), (4096, 3), (200_000, 1000), (300_000, 5000),
                                      (500_000, 200_000), (250_000, 5), (250_000, 70)])
def test_random_sizes_all_aggs(n, groups):
    rng = np.random.default_rng(n)
    a = rng.integers(-groups // 2, groups - groups // 2, n)
    b = rng.standard_normal(n) * 100
    if n > 100:
        b[rng.integers(0, n, n // 100)] = np.nan
    c = rng.integers(-10**12, 10**12, n)
    cols = {"b": b, "c": c}
    for agg in ogroupby.AGGS:
        for sort in (True, False):
            gk, got = _host(a, cols, agg, sort)
            wk, want = ogroupby.groupby_agg(a, cols, agg, sort)
            _compare(gk, got, wk, want, agg, "n=%d g=%d sort=%s" % (n, groups, sort))


def test_expected_groups_hint_paths_agree():
    """Shared-memory pre-aggregation, direct global table and a too-small hint (regrow) give one answer."""
    rng = np.random.default_rng(7)
    n = 400_000
    a = rng.integers(0, 50_000, n)
    cols = {"b": rng.random(n)}
    wk, want = ogroupby.groupby_agg(a, cols, "sum")
    for hint in (0, 10, 1000, 50_000, 10_000_000):
        gk, got = _host(a, cols, "sum", True, 1, hint)
        _compare(gk, got, wk, want, "sum", "hint=%d" % hint)


@pytest.mark.parametrize("keykind", ["int", "float"])
def test_sort_and_segment_path(keykind, monkeypatch):
    """Very high cardinality goes through radix sort + segmented reduction (csrc/groupby.cu run_groupby_sorted);
    the thresholds are lowered so that the oracle finishes in seconds."""
    monkeypatch.setenv("SDCB200_GB_SORT_ROWS", "1000")
    monkeypatch.setenv("SDCB200_GB_SORT_GROUPS", "1000")
    rng = np.random.default_rng(11)
    n, groups = 300_000, 120_000
    a = rng.integers(-groups // 2, groups - groups // 2, n)
    if keykind == "float":
        a = a.astype(np.float64) * 0.5
        a[rng.integers(0, n, 500)] = np.nan          # NaN keys are skipped
        a[rng.integers(0, n, 50)] = -0.0             # -0.0 == +0.0
        a[rng.integers(0, n, 50)] = np.inf
    b = rng.standard_normal(n) * 100
    b[rng.integers(0, n, n // 100)] = np.nan
    c = rng.integers(-10**12, 10**12, n)
    cols = {"b": b, "c": c}
    for agg in ogroupby.AGGS:
        for sort in (True, False):
            gk, got = _host(a, cols, agg, sort, 1, groups)           # hint above the threshold: sort path at once
            wk, want = ogroupby.groupby_agg(a, cols, agg, sort)
            _compare(gk, got, wk, want, agg, "sorted path %s %s sort=%s" % (keykind, agg, sort))
    # unknown hint: shared table overflows, 4 M-slot table is skipped by the tiny threshold -> sort path as the fallback
    gk, got = _host(a, cols, "sum", True, 1, 10)
    wk, want = ogroupby.groupby_agg(a, cols, "sum", True)
    _compare(gk, got, wk, want, "sum", "sorted path fallback")


def test_empty_and_degenerate():
    from sdc_b200.groupby import groupby_host
    rk, res = groupby_host(np.zeros(0, dtype=np.int64), {"v": np.zeros(0)}, ("sum", "count"))
    assert len(rk) == 0 and len(res["sum"]["v"]) == 0 and res["count"]["v"].dtype == np.int64
    # all keys NaN -> no groups
    rk, res = groupby_host(np.full(100, np.nan), {"v": np.ones(100)}, ("sum",))
    assert len(rk) == 0
    # an all-NaN group: sum 0, mean NaN, min/max NaN, count 0, var NaN
    k = np.array([1, 1, 2, 2])
    v = {"v": np.array([np.nan, np.nan, 1.0, 3.0])}
    for agg in ogroupby.AGGS:
        gk, got = _host(k, v, agg)
        wk, want = ogroupby.groupby_agg(k, v, agg)
        _compare(gk, got, wk, want, agg, "all-nan group")
    # the table's empty-slot sentinel is a legal int64 key
    sentinel = np.array([0x7ff8dead0000beef], dtype=np.uint64).view(np.int64)[0]
    k = np.array([sentinel, 5, sentinel, -1, 5], dtype=np.int64)
    v = {"v": np.array([1.0, 2.0, 3.0, 4.0, 5.0])}
    for sort in (True, False):
        gk, got = _host(k, v, "sum", sort)
        wk, want = ogroupby.groupby_agg(k, v, "sum", sort)
        _compare(gk, got, wk, want, "sum", "sentinel key")
    # no value columns: keys only
    rk, res = groupby_host(np.array([3, 1, 3]), {}, ("count",))
    np.testing.assert_array_equal(rk, [1, 3])


def test_dense_window_outliers_and_extreme_keys():
    """Dense int64 keys take the direct-addressed shared table; keys outside the sampled window (and
    INT64_MIN / INT64_MAX) must still aggregate exactly through the global table."""
    rng = np.random.default_rng(21)
    n = 300_000
    a = rng.integers(-400, 600, n)
    a[rng.integers(0, n, 50)] = rng.integers(10**9, 10**12, 50)
    a[7] = np.iinfo(np.int64).min
    a[8] = np.iinfo(np.int64).max
    a[9] = np.iinfo(np.int64).max
    cols = {"b": rng.standard_normal(n), "c": rng.integers(-5, 5, n)}
    for agg in ("sum", "min", "count", "mean"):
        for sort in (True, False):
            gk, got = _host(a, cols, agg, sort, 1, 1000)
            wk, want = ogroupby.groupby_agg(a, cols, agg, sort)
            _compare(gk, got, wk, want, agg, "dense+outliers sort=%s" % sort)
    # keys clustered at the top of the int64 range (window clamp)
    a2 = np.iinfo(np.int64).max - rng.integers(0, 100, n)
    gk, got = _host(a2, cols, "sum", True, 1, 100)
    wk, want = ogroupby.groupby_agg(a2, cols, "sum", True)
    _compare(gk, got, wk, want, "sum", "top of range")
    # sorted keys: the strided sample still sees the whole range
    a3 = np.sort(rng.integers(0, 2000, n))
    gk, got = _host(a3, cols, "max", True, 1, 2000)
    wk, want = ogroupby.groupby_agg(a3, cols, "max", True)
    _compare(gk, got, wk, want, "max", "sorted keys")


def test_multi_agg_one_pass_and_many_columns():
    from sdc_b200.groupby import groupby_host
    rng = np.random.default_rng(3)
    n = 100_000
    a = rng.integers(0, 300, n)
    cols = {"c%d" % i: rng.standard_normal(n) for i in range(12)}
    aggs = ("sum", "mean", "min", "max", "count")
    rk, res = groupby_host(a, cols, aggs)
    for agg in aggs:
        wk, want = ogroupby.groupby_agg(a, cols, agg)
        _compare(rk, res[agg], wk, want, agg, "multi")


def test_dataframe_and_series_api():
    from sdc_b200.groupby import groupby
    df = pd.DataFrame(golden_inputs.GROUPBY_DEFAULT_DF)
    pd.testing.assert_frame_equal(groupby(df, "A").count(), df.groupby("A").count())
    pd.testing.assert_frame_equal(groupby(df, "A").max(), df.groupby("A").max())
    pd.testing.assert_frame_equal(groupby(df, "A")["C", "D"].min(), df.groupby("A")[["C", "D"]].min())
    pd.testing.assert_series_equal(groupby(df, "A")["D"].mean(), df.groupby("A")["D"].mean())
    pd.testing.assert_frame_equal(groupby(df, "A", sort=False).sum()[["B", "C", "D"]],
                                  df.groupby("A", sort=False).sum()[["B", "C", "D"]])
    pd.testing.assert_frame_equal(groupby(df, "A").std()[["B", "C", "D"]], df.groupby("A").std()[["B", "C", "D"]])
    # sdc/tests/test_groupby.py:244-300, 446-498: median / prod (DataFrame and Series forms)
    pd.testing.assert_frame_equal(groupby(df, "A").median()[["B", "C", "D"]], df.groupby("A").median()[["B", "C", "D"]])
    pd.testing.assert_frame_equal(groupby(df, "A").prod()[["B", "C", "D"]], df.groupby("A").prod()[["B", "C", "D"]])
    pd.testing.assert_series_equal(groupby(df, "A")["B"].median(), df.groupby("A")["B"].median())
    pd.testing.assert_series_equal(groupby(df, "A")["B"].prod(), df.groupby("A")["B"].prod())
    s = pd.Series([390., 350., 30., 20.], name="S")
    pd.testing.assert_series_equal(groupby(s, [0, 1, 0, 1]).mean(), s.groupby(np.array([0, 1, 0, 1])).mean())
    with pytest.raises(IndexError):
        groupby(df, "A")["C", "D"]["C"]
    with pytest.raises(ValueError):
        groupby(s, [0, 1])
    with pytest.raises(TypeError):
        groupby(df, ["A", "B"])


def test_c1_full_size_sum_and_multi():
    """BASELINE config C1: 10 M rows, int64 key with 1 K groups, one float64 value column (1 % NaN variant)."""
    import torch
    from sdc_b200.groupby import groupby_device
    rng = np.random.default_rng(0)
    n = 10_000_000
    a = rng.integers(0, 1000, n, dtype=np.int64)
    b = rng.random(n)
    for variant in ("plain", "nan"):
        if variant == "nan":
            b = b.copy()
            b[rng.integers(0, n, n // 100)] = np.nan
        dk, res = groupby_device(torch.from_numpy(a).cuda(), [torch.from_numpy(b).cuda()],
                                 ("sum", "mean", "min", "max", "count"), True, 1, 1000)
        want = pd.DataFrame({"A": a, "B": b}).groupby("A")["B"].agg(["sum", "mean", "min", "max", "count"])
        np.testing.assert_array_equal(dk.cpu().numpy(), want.index.to_numpy())
        for agg in ("sum", "mean"):
            np.testing.assert_allclose(res[agg][0].cpu().numpy(), want[agg].to_numpy(), rtol=1e-9, atol=0)
        for agg in ("min", "max", "count"):
            np.testing.assert_array_equal(res[agg][0].cpu().numpy(), want[agg].to_numpy())


@pytest.mark.parametrize("keys_kind", ["dense", "skewed", "sparse", "float", "outliers", "wide", "few", "zipf", "hotnan", "nanzero",
                                       "sparse_many", "sparse_nanzero", "float_zipf", "marker_key"])
def test_single_f64_column_specialised_kernels(keys_kind):
    """One float64 column with sum / mean / the five-aggregate set takes gb_smem_kernel's compile-time specialised
    loops (SPEC 1 / 3 / 15): dense keys inside the sampled direct window, a hot key, hashed keys, float keys, keys
    outside the sampled span (rerun with hashing), more groups than shared slots (global-table fallback); a handful
    of groups (the direct window is copied 32 times across the shared table, one copy per lane), Zipf-distributed keys
    (the most frequent sampled keys are aggregated in registers) and a hot key whose values are all NaN (the group
    must still appear: sum 0, count 0, min / max / mean NaN).  Sparse int64 keys and float64 keys take the HASHED lean
    loop (slot from a probe of the CTA's shared key table): more sparse groups than the shared table holds (overflow 3,
    rerun with the generic loop), groups whose values are all NaN / zeros, skewed float keys with -0.0 == +0.0, and a
    key equal to the table's empty marker."""
    rng = np.random.default_rng(23)
    n = 300_000
    if keys_kind == "dense":
        a = rng.integers(0, 1000, n)
    elif keys_kind == "skewed":
        a = np.where(rng.random(n) < 0.8, 7, rng.integers(0, 1000, n))
    elif keys_kind == "sparse":
        a = rng.integers(0, 1000, n) * 1000003 - 5
    elif keys_kind == "float":
        a = rng.integers(0, 700, n).astype(np.float64) / 4
        a[rng.integers(0, n, 50)] = np.nan
    elif keys_kind == "outliers":
        a = rng.integers(100, 600, n)
        a[rng.integers(0, n, 5)] = 1300          # inside the table window, outside the sampled span
        a[rng.integers(0, n, 5)] = 2
    elif keys_kind == "few":
        a = rng.integers(-3, 4, n)
    elif keys_kind == "zipf":
        a = np.minimum(rng.zipf(1.3, n), 1000) - 1
    elif keys_kind == "hotnan":
        a = np.where(rng.random(n) < 0.3, 41, rng.integers(0, 1000, n))
    elif keys_kind == "nanzero":
        a = rng.integers(0, 1000, n)
    elif keys_kind == "sparse_many":
        a = rng.integers(0, 3000, n) * 1000003 - 5
    elif keys_kind == "sparse_nanzero":
        a = rng.integers(0, 900, n) * 7_000_000_019 + 11
    elif keys_kind == "float_zipf":
        a = (np.minimum(rng.zipf(1.3, n), 500) - 1).astype(np.float64) * 0.5 - 3.0
        a[a == 0.0] = np.where(rng.random(int((a == 0.0).sum())) < 0.5, 0.0, -0.0)
        a[rng.integers(0, n, 50)] = np.nan
    elif keys_kind == "marker_key":
        a = rng.integers(0, 500, n) * 1000003
        a[rng.integers(0, n, 40)] = 0x7ff8dead0000beef          # csrc/groupby.cu kEmptyKey
    else:
        a = rng.integers(0, 3000, n)
    b = rng.standard_normal(n) * 50
    b[rng.integers(0, n, n // 50)] = np.nan
    b[rng.integers(0, n, 20)] = np.inf
    if keys_kind == "hotnan":
        b[a == 41] = np.nan
    if keys_kind == "sparse_nanzero":
        u = np.unique(a)
        b[a == u[5]] = np.nan
        b[a == u[6]] = 0.0
        b[a == u[7]] = -0.0
    if keys_kind == "nanzero":        # groups the lean loop's implicit presence (count / sum changed) cannot see
        b[a == 5] = np.nan
        b[a == 6] = 0.0
        b[a == 7] = -0.0
        b[(a == 8) & (np.arange(n) % 2 == 0)] = 1.5
        b[(a == 8) & (np.arange(n) % 2 == 1)] = -1.5
    cols = {"b": b}
    from sdc_b200.groupby import groupby_host
    for aggs in (("sum",), ("mean",), ("sum", "mean", "min", "max", "count"), ("min",)):
        for hint in (1000, 0):
            rk, res = groupby_host(a, cols, aggs, True, 1, hint)
            for agg in aggs:
                wk, want = ogroupby.groupby_agg(a, cols, agg)
                _compare(rk, res[agg], wk, want, agg, "%s %s hint=%d" % (keys_kind, aggs, hint))


@pytest.mark.parametrize("keys_kind", ["uniform", "zipf", "few", "wide", "sparse"])
def test_several_f64_columns_take_the_lean_loop_per_column(keys_kind):
    """>= 4 Mi rows, several float64 columns, low-cardinality int64 keys: run_groupby_int_keys aggregates column by
    column with the one-column lean loop and assembles the result rows; "wide" keys (no direct window) must fall back
    to the generic all-columns kernel.  Against pandas (the oracle's dict-of-lists takes minutes at this size)."""
    import torch
    from sdc_b200.groupby import groupby_device
    rng = np.random.default_rng(77)
    n = 5_000_000
    if keys_kind == "uniform":
        a = rng.integers(-500, 500, n)
    elif keys_kind == "zipf":
        a = np.minimum(rng.zipf(1.3, n), 1000) - 1
    elif keys_kind == "few":
        a = rng.integers(0, 6, n)
    elif keys_kind == "sparse":                           # no direct window: the hashed lean loop, column by column
        a = rng.integers(0, 800, n) * 1000003 - 17
    else:
        a = rng.integers(0, 3000, n) * 1000003
    cols = [rng.standard_normal(n) * 10 for _ in range(3)]
    cols[1][rng.integers(0, n, n // 100)] = np.nan
    cols[2][a == a[0]] = np.nan                           # one group with no valid value in the last column
    df = pd.DataFrame({"a": a, "c0": cols[0], "c1": cols[1], "c2": cols[2]})
    dk = torch.from_numpy(a).cuda()
    dc = [torch.from_numpy(c).cuda() for c in cols]
    for aggs in (("sum",), ("mean",), ("count",), ("sum", "mean", "min", "max", "count")):
        for hint in (1000, 0):
            gk, res = groupby_device(dk, dc, aggs, True, 1, hint)
            want = df.groupby("a")
            np.testing.assert_array_equal(gk.cpu().numpy(), np.sort(df["a"].unique()))
            for agg in aggs:
                w = getattr(want, agg)()
                for i in range(3):
                    g = res[agg][i].cpu().numpy()
                    if agg in EXACT:
                        np.testing.assert_array_equal(g, w["c%d" % i].to_numpy(), err_msg="%s %s c%d" % (keys_kind, agg, i))
                    else:
                        assert_close(g, w["c%d" % i].to_numpy(), rtol=1e-9, atol=1e-12, msg="%s %s c%d hint=%d" % (keys_kind, agg, i, hint))


def test_measurement_switches_give_the_same_result(monkeypatch):
    """The A/B switches of DESIGN 6 (read on every call) select other code paths of the same kernel family -- window
    copies off, hot keys off, control block through a copy-engine read-back, the generic loop for several columns:
    every one of them must give what the default path gives (sparse keys: the hashed lean loop against the generic one)."""
    import torch
    from sdc_b200.groupby import groupby_device
    rng = np.random.default_rng(5)
    n = 4_500_000
    a = torch.from_numpy(np.minimum(rng.zipf(1.4, n), 300) - 1).cuda()
    cols = [torch.from_numpy(rng.standard_normal(n)).cuda() for _ in range(2)]
    cols[0][::77] = float("nan")
    aggs = ("sum", "mean", "min", "max", "count")

    sparse = a * 1_000_003 + 17
    for ncols, keys in ((1, a), (2, a), (1, sparse)):
        def run(ncols, keys=keys):
            k, r = groupby_device(keys, cols[:ncols], aggs, True, 1, 0)
            return k.clone(), {g: [t.clone() for t in r[g]] for g in aggs}
        base_k, base = run(ncols)
        for name, value in (("SDCB200_GB_RF", "1"), ("SDCB200_GB_HOT", "0"), ("SDCB200_GB_MAPPED", "0"),
                            ("SDCB200_GB_SPLIT_COLS", "0"), ("SDCB200_GB_LEAN_HASH", "0")):
            monkeypatch.setenv(name, value)
            k, r = run(ncols)
            monkeypatch.delenv(name)
            assert torch.equal(k, base_k), name
            for g in aggs:
                for x, y in zip(r[g], base[g]):
                    if g in ("sum", "mean"):
                        assert torch.allclose(x, y, rtol=1e-9, atol=1e-12, equal_nan=True), (name, g)
                    elif g == "count":
                        assert torch.equal(x, y), (name, g)
                    else:
                        assert torch.equal(torch.nan_to_num(x, nan=12345.0), torch.nan_to_num(y, nan=12345.0)), (name, g)



def test_dense_high_cardinality_takes_direct_table_and_matches_sort_path(monkeypatch):
    """Tens of millions of expected groups: dense int64 keys (what a ROUTE_MOD shuffle delivers) aggregate in a
    direct-addressed window table (csrc/groupby.cu run_groupby_window, with and without the key-range partition pass in
    front), anything else goes to sort-and-segment -- same result either way.  Thresholds lowered so the test stays small."""
    import torch
    from sdc_b200.groupby import groupby_device
    rng = np.random.default_rng(9)
    n = 600_000
    dense = rng.integers(-40_000, 160_000, n)
    sparse = dense * 1_000_003
    v = rng.standard_normal(n)
    v[::91] = np.nan
    v[dense == 77] = np.nan               # a group without a single value: sum 0, count 0, mean / min / max NaN
    v[dense == 78] = -0.0                 # a group whose sum never leaves -0.0
    v[dense == 79] = 0.0
    dv = torch.from_numpy(v).cuda()
    monkeypatch.setenv("SDCB200_GB_SORT_ROWS", "1000")
    monkeypatch.setenv("SDCB200_GB_SORT_GROUPS", "1000")
    all_aggs = ("sum", "count", "min", "max", "mean")
    modes = (("0", "0"), (str(256 << 20), "0"), (str(256 << 20), "1"))      # (dense threshold, partition pass)
    for keys in (dense, sparse):
        dk = torch.from_numpy(keys).cuda()
        df = pd.DataFrame({"k": keys, "v": v})
        for aggs in (all_aggs, ("sum",), ("count",), ("min", "max"), ("mean",)):
            for sort in (True, False):
                want = df.groupby("k", sort=sort)["v"]
                for dg, part in modes:
                    monkeypatch.setenv("SDCB200_GB_DENSE_GROUPS", dg)   # 0: never try the window table
                    monkeypatch.setenv("SDCB200_GB_WIN_PART", part)
                    gk, res = groupby_device(dk, [dv], aggs, sort, 1, 200_000)
                    msg = "aggs=%s sort=%s dense=%s part=%s" % (aggs, sort, dg, part)
                    np.testing.assert_array_equal(gk.cpu().numpy(), want.sum().index.to_numpy(), err_msg="keys " + msg)
                    for a in aggs:
                        x = res[a][0].cpu().numpy()
                        w = getattr(want, a)().to_numpy()
                        if a in ("sum", "mean"):
                            assert_close(x, w, rtol=1e-9, atol=1e-12, msg=a + " " + msg)
                        else:
                            np.testing.assert_array_equal(x, w, err_msg=a + " " + msg)
                        if a == "sum":
                            assert not np.signbit(x[w == 0]).any(), "sum of nothing / of -0.0 is +0.0 like the reference's nansum: " + msg


def test_async_begin_end_matches_the_blocking_call():
    """sdcb200_groupby_begin / _end: several aggregates in flight at once, every result equal to the blocking call's --
    including a request whose optimistic attempt overflows (more groups than hinted) and is rerun in _end, and requests
    that complete inside _begin (var, many groups)."""
    import torch
    from sdc_b200.groupby import groupby_device, groupby_device_async
    rng = np.random.default_rng(21)
    n = 1_000_000
    cases = []
    for groups, hint, aggs in ((1000, 1000, ("sum",)), (50, 0, ("sum", "mean", "min", "max", "count")), (3000, 100, ("sum", "count")),
                               (200_000, 0, ("sum",)), (500, 500, ("var",))):
        k = torch.from_numpy(rng.integers(-5, groups - 5, n)).cuda()
        v = torch.from_numpy(rng.standard_normal(n)).cuda()
        v[::97] = float("nan")
        cases.append((k, v, aggs, hint))
    kf = torch.from_numpy(rng.integers(0, 40, n).astype(np.float64)).cuda()
    cases.append((kf, cases[0][1], ("sum", "max"), 64))
    pend = [groupby_device_async(k, [v], aggs, True, 1, hint) for k, v, aggs, hint in cases]      # all in flight
    for (k, v, aggs, hint), p in zip(cases, pend):
        gk, res = p.result()
        wk, want = groupby_device(k, [v], aggs, True, 1, hint)
        assert torch.equal(gk, wk)
        for a in aggs:
            x, y = res[a][0], want[a][0]
            if a in ("sum", "mean", "var"):
                assert torch.allclose(x, y, rtol=1e-9, atol=1e-12, equal_nan=True), a
            else:
                assert torch.equal(torch.nan_to_num(x.double(), nan=1e300), torch.nan_to_num(y.double(), nan=1e300)), a
    # a handle dropped without result() releases its table after the kernel
    groupby_device_async(cases[0][0], [cases[0][1]], ("sum",), True, 1, 1000)
    torch.cuda.synchronize()
