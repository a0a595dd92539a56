"""GPU parity: csrc/rolling.cu (through the C ABI) against the oracle and the golden fixtures."""
import numpy as np
import pandas as pd
import pytest

from helpers import assert_close, golden_inputs, rolling_golden_cases
from oracle import rolling as orolling

pytestmark = pytest.mark.gpu

OPS = ("sum", "mean", "var", "std", "count", "min", "max")
EPS = np.finfo(np.float64).eps


def _dev(x):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def check(op, got, want, x, win, msg=""):
    """count/min/max: exact.  sum/mean/var/std: 1e-9 relative (north_star) plus the conditioning
    floor of the formula itself -- both sides evaluate window sums whose rounding error is about
    eps * win * max|x| (sum, mean) or eps * win * max|x|^2 (the s2 - s1^2/n of var; std is
    compared through its square), so results that cancel to ~0 cannot agree to 1e-9 relative."""
    if op in ("count", "min", "max"):
        return assert_close(got, want, rtol=0.0, atol=0.0, msg=msg)
    x = np.asarray(x, dtype=np.float64)
    fin = x[np.isfinite(x)]
    scale = float(np.max(np.abs(fin))) if fin.size else 1.0
    floor = 64 * EPS * max(win, 256)      # GPU prefixes span a 256-row segment; the reference drifts over a whole chunk
    if op in ("sum", "mean"):
        return assert_close(got, want, rtol=1e-9, atol=floor * scale, msg=msg)
    if op == "std":
        # var that cancels to ~0 can round to -tiny on one side (sqrt -> NaN) and +tiny on the other:
        # a NaN facing a value whose square is below the floor counts as that ~0
        got, want = np.square(np.asarray(got, dtype=np.float64)), np.square(np.asarray(want, dtype=np.float64))
        tiny = floor * scale * scale
        g_only = np.isnan(got) & ~np.isnan(want) & (np.nan_to_num(want, nan=np.inf) <= tiny)
        w_only = np.isnan(want) & ~np.isnan(got) & (np.nan_to_num(got, nan=np.inf) <= tiny)
        got = np.where(g_only, want, got)
        want = np.where(w_only, got, want)
    assert_close(got, want, rtol=2e-9, atol=floor * scale * scale, msg=msg)


def test_golden_through_host_abi():
    from sdc_b200.rolling import rolling_host
    n = 0
    for op, data, w, mp, ddof, want in rolling_golden_cases():
        got = rolling_host(data, op, w, mp, ddof)
        check(op, got, want, data, w, msg="%s w=%d mp=%d ddof=%d %s" % (op, w, mp, ddof, data))
        n += 1
    assert n > 1000


def test_examples_literals():
    from sdc_b200.rolling import rolling
    for op, (data, win, want) in golden_inputs.EXAMPLES_ROLLING.items():
        minp = 0 if op == "count" else None       # see tests/test_oracle_rolling.py
        got = getattr(rolling(pd.Series(data), win, minp), op)()
        assert_close(got.to_numpy(), want, rtol=0, atol=6e-7, msg=op)
    ex = golden_inputs.EXAMPLE_DF_ROLLING_MEAN
    got = rolling(pd.DataFrame(ex["input"]), ex["window"]).mean()
    for col in ex["input"]:
        assert_close(got[col].to_numpy(), ex["expected"][col], rtol=0, atol=6e-7, msg=col)


def _mixed(n, seed):
    rng = np.random.default_rng(seed)
    x = rng.random(n) * 200 - 100
    if n > 10:
        x[rng.integers(0, n, max(1, n // 50))] = np.nan
        x[rng.integers(0, n, max(1, n // 200))] = np.inf
        x[rng.integers(0, n, max(1, n // 200))] = -np.inf
        x[rng.integers(0, n, max(1, n // 100))] = 0.0
        x[rng.integers(0, n, max(1, n // 300))] = -0.0
    return x


@pytest.mark.parametrize("n", [1, 2, 3, 63, 64, 65, 1947, 1948, 1949, 2048, 4097, 100003])
def test_device_vs_oracle_sizes(n):
    from sdc_b200.rolling import rolling_device
    x = _mixed(n, n)
    d = _dev(x)
    for win in (1, 2, 3, 5, 31, 32, 33, 63, 64, 65, 100, 127, 128, 129, 191, 192, 193, 194, 200, 257, 1000, 1025, 4096, 4097):
        for mp in sorted({0, 1, win // 2, win}):
            for op in OPS:
                got = rolling_device(d, op, win, mp, 1).cpu().numpy()
                want = orolling.rolling(x, op, win, mp, 1, nchunks=1)
                check(op, got, want, x, win, msg="%s n=%d w=%d mp=%d" % (op, n, win, mp))


def test_window_zero_and_ddof():
    from sdc_b200.rolling import rolling_device
    x = _mixed(1000, 5)
    d = _dev(x)
    for op in OPS:
        got = rolling_device(d, op, 0, 0, 1).cpu().numpy()
        want = orolling.rolling(x, op, 0, 0, 1, nchunks=1)
        assert_close(got, want, rtol=0, atol=0, msg=op)
    for ddof in (0, 1, 2, 5, 100):
        for op in ("var", "std"):
            got = rolling_device(d, op, 10, 3, ddof).cpu().numpy()
            want = orolling.rolling(x, op, 10, 3, ddof, nchunks=1)
            check(op, got, want, x, 10, msg="%s ddof=%d" % (op, ddof))


def test_int64_column():
    from sdc_b200.rolling import rolling_device
    x = np.random.default_rng(3).integers(-1000, 1000, 50_000)
    d = _dev(x)
    for op in OPS:
        got = rolling_device(d, op, 100, 100, 1).cpu().numpy()
        want = orolling.rolling(x, op, 100, 100, 1, nchunks=1)
        check(op, got, want, x, 100, msg=op)


def test_unaligned_pointers():
    import torch
    from sdc_b200.rolling import rolling_device
    x = _mixed(10_001, 9)
    base = _dev(np.concatenate([[0.0], x]))
    col = base[1:]                       # 8-byte aligned only: no TMA, plain loads
    outbuf = torch.empty(10_002, dtype=torch.float64, device="cuda")
    for op in OPS:
        got = rolling_device(col, op, 100, 50, 1, out=outbuf[1:]).cpu().numpy()
        want = orolling.rolling(x, op, 100, 50, 1, nchunks=1)
        check(op, got, want, x, 100, msg=op)


@pytest.mark.parametrize("split", [1, 50, 99, 100, 5000])
def test_row_sharded_halo(split):
    """Second shard + the previous shard's last win-1 rows == the tail of the whole series."""
    from sdc_b200.rolling import rolling_device
    x = _mixed(20_000, 11)
    win = 100
    d = _dev(x)
    for op in OPS:
        for mp in (0, 100):
            whole = rolling_device(d, op, win, mp, 1).cpu().numpy()
            h0 = max(0, split - (win - 1))
            got = rolling_device(d[split:].clone(), op, win, mp, 1, halo=d[h0:split].clone(),
                                 row_offset=split).cpu().numpy()
            check(op, got, whole[split:], x, win, msg="%s split=%d" % (op, split))


@pytest.mark.parametrize("n", [300, 4098, 5000, 20_000, 250_003])
def test_windows_beyond_the_tile_limit(n):
    """win > 4097 (the reference accepts any window up to the series' length): chunk-decomposed windows over a segment
    tree of chunk totals (csrc/rolling.cu rl_*_kernel), every op, windows longer than the series included; the oracle's
    sliding accumulators drift over the whole chunk, so sums are held to the same conditioning floor as elsewhere."""
    from sdc_b200.rolling import rolling_device
    x = _mixed(n, n + 1)
    d = _dev(x)
    for win in (4098, 4099, 6144, 10_000, 65_537, 300_000):
        if win > 4 * n + 5000:
            continue
        for mp in sorted({0, 1, win // 2, win}):
            for op in OPS:
                got = rolling_device(d, op, win, mp, 1).cpu().numpy()
                want = orolling.rolling(x, op, win, mp, 1, nchunks=1)
                check(op, got, want, x, win, msg="%s n=%d w=%d mp=%d" % (op, n, win, mp))


def test_large_window_halo_int64_and_host():
    from sdc_b200.rolling import rolling_device, rolling_host
    x = _mixed(60_000, 21)
    d = _dev(x)
    win = 7000
    for op in OPS:
        whole = rolling_device(d, op, win, 3000, 1).cpu().numpy()
        for split in (1, 2047, 6999, 7000, 30_001):
            h0 = max(0, split - (win - 1))
            got = rolling_device(d[split:].clone(), op, win, 3000, 1, halo=d[h0:split].clone(), row_offset=split).cpu().numpy()
            check(op, got, whole[split:], x, win, msg="%s split=%d" % (op, split))
    xi = np.random.default_rng(3).integers(-1000, 1000, 30_000)
    di = _dev(xi)
    for op in OPS:
        got = rolling_device(di, op, 5000, 100, 1).cpu().numpy()
        check(op, got, orolling.rolling(xi, op, 5000, 100, 1, nchunks=1), xi, 5000, msg="int64 " + op)
    xh = np.random.default_rng(4).random(3_000_000)
    xh[::1009] = np.nan
    for op in ("mean", "max", "count", "std"):
        check(op, rolling_host(xh, op, 50_000, 100, 1), orolling.rolling(xh, op, 50_000, 100, 1), xh, 50_000, msg="host " + op)


def test_host_abi_chunked_large():
    """Host entry point with several pipeline chunks (n > 8 Mi rows)."""
    from sdc_b200.rolling import rolling_host
    n = (1 << 23) * 2 + 12345
    x = np.random.default_rng(1).random(n)
    x[::100003] = np.nan
    for op in ("mean", "max", "count"):
        got = rolling_host(x, op, 100, 100, 1)
        want = orolling.rolling(x, op, 100, 100, 1)
        check(op, got, want, x, 100, msg=op)


def test_c2_full_size_mean_sum_std():
    """BASELINE config C2: 100 M float64, window 100 -- whole-column parity with the oracle."""
    from sdc_b200.rolling import rolling_device
    n = 100_000_000
    x = np.random.default_rng(1).random(n)
    d = _dev(x)
    for op in ("mean", "sum", "std"):
        got = rolling_device(d, op, 100, 100, 1).cpu().numpy()
        want = orolling.rolling(x, op, 100, 100, 1)
        assert np.array_equal(np.isnan(got), np.isnan(want))
        ok = ~np.isnan(want)
        err = np.max(np.abs(got[ok] - want[ok]) / np.maximum(np.abs(want[ok]), 1e-300))
        assert err < 1e-9, (op, err)


def test_std_is_the_square_root_of_var():
    """The std kernel takes a run's square roots together with a branch-free form of sqrt()'s own chain
    (csrc/rolling.cu sqrt_batch).  It must give what sqrt() gives: the var kernel's output through torch.sqrt -- clean
    tiles (the batched roots) and NaN-bearing tiles (sqrt() itself), variances that are subnormal, huge, exactly zero
    and negative from cancellation (-> NaN, as the reference's var ** 0.5)."""
    import torch
    from sdc_b200.rolling import rolling_device
    rng = np.random.default_rng(11)
    n = 3_000_000
    base = rng.standard_normal(n)
    cases = [base * s for s in (1.0, 1e-160, 1e150)]
    cases.append(np.repeat(rng.standard_normal(n // 500), 500))          # constant runs: variance 0 or a rounding residue
    holes = base.copy()
    holes[rng.integers(0, n, n // 1000)] = np.nan
    holes[rng.integers(0, n, 5)] = np.inf
    cases.append(holes)
    for x in cases:
        d = _dev(x)
        for win, minp in ((2, 2), (7, 3), (100, 100)):
            var = rolling_device(d, "var", win, minp, 1)
            std = rolling_device(d, "std", win, minp, 1)
            want = torch.sqrt(var)
            assert torch.equal(torch.isnan(std), torch.isnan(want))
            ok = ~torch.isnan(want)
            ulp = (std[ok].view(torch.int64) - want[ok].view(torch.int64)).abs()
            assert int(ulp.max()) <= 1, (win, int(ulp.max()))


def test_adversarial_reference_perf_data():
    """The reference perf generator's data (choice over test_global_input_data_float64): exact ops."""
    from sdc_b200.rolling import rolling_device
    vals = np.asarray(sum(golden_inputs.GLOBAL_FLOAT64, []), dtype=np.float64)
    x = np.random.default_rng(1).choice(vals, 1_000_000)
    d = _dev(x)
    for op in ("count", "min", "max"):
        got = rolling_device(d, op, 100, 100, 1).cpu().numpy()
        want = orolling.rolling(x, op, 100, 100, 1)
        assert_close(got, want, rtol=0, atol=0, msg=op)
    # sum / mean: rows with fewer than min_periods finite values in the window must be NaN (the
    # non-finite counters).  Elsewhere finfo.max + finfo.max overflows in BOTH implementations (the
    # reference's running sum turns inf and stays inf for the rest of its chunk, :1237, :1297), so
    # the values themselves are not comparable on this data ...
    nfin = orolling.rolling(np.where(np.isfinite(x), 1.0, np.nan), "count", 100, 0, 1)
    for op in ("sum", "mean"):
        got = rolling_device(d, op, 100, 30, 1).cpu().numpy()
        assert np.all(np.isnan(got[nfin < 30])), op
    # ... so the same draw with the two overflowing magnitudes scaled down checks the full pattern
    # and the values
    y = np.where(np.abs(x) > 1e300, np.sign(x) * 1e10, x)
    dy = _dev(y)
    for op in ("sum", "mean", "var", "std"):
        got = rolling_device(dy, op, 100, 30, 1).cpu().numpy()
        want = orolling.rolling(y, op, 100, 30, 1)
        check(op, got, want, y, 100, msg=op)


def test_rolling_class_api_and_errors():
    from sdc_b200.rolling import rolling
    s = pd.Series([4., 3., 5., 2., 6.], index=[4, 3, 2, 1, 0], name="A")
    r = rolling(s, 3).mean()
    pd.testing.assert_series_equal(r, s.rolling(3).mean())
    df = pd.DataFrame({"A": [4, 3, 5, 2, 6], "B": [-4., -3., -5., -2., np.nan]})
    pd.testing.assert_frame_equal(rolling(df, 2, 1).sum(), df.rolling(2, 1).sum())
    with pytest.raises(ValueError, match="window must be non-negative"):
        rolling(s, -1)
    with pytest.raises(ValueError, match="min_periods must be >= 0"):
        rolling(s, 1, -1)
    with pytest.raises(ValueError, match="min_periods must be <= window"):
        rolling(s, 1, 2)
    with pytest.raises(ValueError, match="center"):
        rolling(s, 1, 1, center=True)
    with pytest.raises(TypeError, match="ddof"):
        rolling(s, 3).std(ddof="1")
