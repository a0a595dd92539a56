"""GPU parity: csrc/strings.cu (str_arr argsort / factorize / gather) + string-key groupby -- exact."""
import numpy as np
import pandas as pd
import pytest

from oracle import groupby as ogroupby
from oracle import strings as ostr
from test_oracle_strings import CASES

pytestmark = pytest.mark.gpu


def _rand_strings(rng, n, maxlen, alphabet="abcXYZ019"):
    al = np.array(list(alphabet))
    return ["".join(rng.choice(al, rng.integers(0, maxlen + 1))) for _ in range(n)]


def _check_argsort(vals):
    from sdc_b200.str_arr import StringArray
    offsets, chars, _ = ostr.to_str_arr(vals)
    sa = StringArray.from_pylist(vals)
    np.testing.assert_array_equal(sa.offsets, offsets)
    np.testing.assert_array_equal(sa.chars, chars)
    for asc in (True, False):
        got = sa.argsort(asc).cpu().numpy().astype(np.uint64)
        want = ostr.stable_argsort(offsets, chars, asc)
        np.testing.assert_array_equal(got, want, err_msg="asc=%s n=%d" % (asc, len(vals)))


def test_argsort_reference_cases():
    for vals in CASES:
        _check_argsort([v for v in vals if v is not None])
    _check_argsort(["x"])
    _check_argsort(["", ""])
    _check_argsort(["same"] * 100)
    _check_argsort(["ab", "ab\0", "ab\0\0", "ab", "a\0b", "ab\0c"])


@pytest.mark.parametrize("n,maxlen", [(2, 3), (1000, 5), (4097, 8), (50_000, 11), (20_000, 40)])
def test_argsort_random(n, maxlen):
    _check_argsort(_rand_strings(np.random.default_rng(n), n, maxlen))


def test_reference_symbol_stable_argsort_host_buffers():
    """`stable_argsort(data_ptr, offset_ptr, n, ascending, result_ptr)` exactly as sdc/str_arr_ext.py:94-99 binds it."""
    from sdc_b200 import native
    lib = native.load()
    vals = _rand_strings(np.random.default_rng(9), 5000, 9)
    offsets, chars, _ = ostr.to_str_arr(vals)
    for asc in (1, 0):
        out = np.empty(len(vals), dtype=np.uint64)
        lib.stable_argsort(chars.ctypes.data, offsets.ctypes.data, len(vals), asc, out.ctypes.data)
        np.testing.assert_array_equal(out, ostr.stable_argsort(offsets, chars, bool(asc)))


def test_stable_argsort_against_the_compiled_reference():
    """Same symbol, same buffers, into the reference's OWN stable_argsort (oracle/_ref, compiled from
    sdc/native/str_ext/sdc_str_arr.hpp:565-593) and into this library's: 300 K strings, both directions, bit for bit."""
    from oracle import ref as oref
    if not oref.available():
        pytest.skip("oracle/_ref not built")
    from sdc_b200 import native
    lib = native.load()
    rng = np.random.default_rng(31)
    for vals in (_rand_strings(rng, 300_000, 8), _rand_strings(rng, 100_000, 30, "ab"), _rand_strings(rng, 50_000, 12, "aé日Z0")):
        offsets, chars, _ = ostr.to_str_arr(vals)
        for asc in (1, 0):
            out = np.empty(len(vals), dtype=np.uint64)
            lib.stable_argsort(chars.ctypes.data, offsets.ctypes.data, len(vals), asc, out.ctypes.data)
            np.testing.assert_array_equal(out, oref.stable_argsort(offsets, chars, bool(asc)))


def test_take_and_round_trip():
    import torch
    from sdc_b200.str_arr import StringArray
    vals = ['ac', 'c', None, 'cb', '', 'ddd', None, 'é日本']
    sa = StringArray.from_pylist(vals).cuda()
    assert sa.to_pylist() == vals
    rows = torch.tensor([7, 0, -1, 2, 4, 5, 5], dtype=torch.int64, device="cuda")
    assert sa.take(rows).to_pylist() == ['é日本', 'ac', None, None, '', 'ddd', 'ddd']


def test_series_sort_values_strings():
    # sdc/tests/test_series.py:3856-4072: strings with None and '', both na_position values, both directions
    from sdc_b200.sort import series_sort_values
    s = pd.Series(['ac', 'c', 'cb', 'ca', None, 'da', 'cc', 'ddd', 'd', ''], index=list(range(10, 0, -1)), name="S")
    for asc in (True, False):
        for nap in ("last", "first"):
            got = series_sort_values(s, ascending=asc, kind="mergesort", na_position=nap)
            want = s.sort_values(ascending=asc, kind="mergesort", na_position=nap)
            pd.testing.assert_series_equal(got, want)


def test_groupby_string_keys():
    from sdc_b200.groupby import groupby_host
    from helpers import golden_inputs
    # sdc/tests/test_groupby.py:83-100: string keys with None
    df = pd.DataFrame(golden_inputs.GROUPBY_DEFAULT_DF)
    key = golden_inputs.GROUPBY_KEY_DTYPES["string"]
    cols = {c: df[c].to_numpy() for c in "BCDE"}
    for agg in ("count", "sum", "min", "mean"):
        for sort in (True, False):
            rk, res = groupby_host(list(key), cols, (agg,), sort)
            wk, want = ogroupby.groupby_agg(list(key), cols, agg, sort)
            assert list(rk) == list(wk), (agg, sort, rk, wk)
            for c in want:
                np.testing.assert_allclose(res[agg][c].astype(np.float64), want[c].astype(np.float64), rtol=1e-12,
                                           equal_nan=True, err_msg="%s %s" % (agg, c))
    # perf-test shape (tests_perf/test_perf_df_groupby.py:164-204): 3-char keys, 200 groups
    rng = np.random.default_rng(4)
    pool = _rand_strings(rng, 200, 3, "abcdefghijklmnopqrstuvwxyz")
    keys = [pool[i] for i in rng.integers(0, len(pool), 100_000)]
    v = {"v": rng.standard_normal(100_000), "w": rng.integers(0, 100, 100_000)}
    for sort in (True, False):
        rk, res = groupby_host(keys, v, ("sum", "mean", "min", "max", "count"), sort)
        for agg in ("sum", "mean", "min", "max", "count"):
            wk, want = ogroupby.groupby_agg(keys, v, agg, sort)
            assert list(rk) == list(wk)
            for c in want:
                if agg in ("min", "max", "count") or want[c].dtype.kind == "i":
                    np.testing.assert_array_equal(res[agg][c], want[c])
                else:
                    np.testing.assert_allclose(res[agg][c], want[c], rtol=1e-9)


def test_groupby_string_keys_inline_and_hashed_slots():
    """csrc/strings.cu str_factorize: strings of <= 8 bytes ARE the slot's key word (zero padded, so "ab" and
    "ab\\0" differ only by length), longer ones hash into it and are confirmed byte by byte; the 8-byte string of
    0xFF collides with the empty marker and lives in the spare slot.  Groups must be exact in every case."""
    import torch
    from sdc_b200.str_arr import StringArray, groupby_str_device
    rng = np.random.default_rng(9)
    pool = ["", "a", "ab", "ab\x00", "ab\x00\x00", "abcdefgh", "abcdefg", "abcdefghi", "abcdefghij" * 3,
            "abcdefghij" * 3 + "k", "x" * 8, "x" * 9, "x" * 16, "x" * 17, "\u00e9t\u00e9", "\u00e9t\u00e9s longs et accentu\u00e9s"]
    keys = [pool[i] for i in rng.integers(0, len(pool), 50_000)]
    keys[7] = None
    keys[4242] = None
    v = {"v": rng.standard_normal(50_000)}
    from sdc_b200.groupby import groupby_host
    for sort in (True, False):
        rk, res = groupby_host(keys, v, ("sum", "count"), sort)
        for agg in ("sum", "count"):
            wk, want = ogroupby.groupby_agg(keys, v, agg, sort)
            assert list(rk) == list(wk), (sort, rk, wk)
            if agg == "count":
                np.testing.assert_array_equal(res[agg]["v"], want["v"])
            else:
                np.testing.assert_allclose(res[agg]["v"], want["v"], rtol=1e-9)
    # raw bytes (not UTF-8): 8 x 0xFF is the table's empty marker -> spare slot; 8 x 0xFE is an ordinary key
    codes = np.full((1000, 8), 0xFE, dtype=np.uint8)
    codes[::3] = 0xFF
    codes[1::7, 3] = 0x41
    sa = StringArray.from_fixed_width(codes).cuda()
    val = torch.ones(1000, dtype=torch.float64, device="cuda")
    rk, res = groupby_str_device(sa, [val], ("count",), True, 1, 0)
    as_int = np.ascontiguousarray(codes).view(">u8").reshape(-1)
    uniq, cnt = np.unique(as_int, return_counts=True)
    assert rk.n == len(uniq)
    got = rk.cpu()
    got_int = got.chars.reshape(-1, 8).view(">u8").reshape(-1)
    np.testing.assert_array_equal(got_int, uniq)                     # byte-lexicographic order == big-endian value order
    np.testing.assert_array_equal(res["count"][0].cpu().numpy(), cnt)


@pytest.mark.parametrize("kind", ["upto7", "fixed8", "mixed8", "upto7_na", "fixed8_unaligned_view"])
def test_groupby_short_string_keys_as_words(kind):
    """Strings that all fit one 64-bit word go through the numeric hash groupby as int64 words (str_arr.key_words:
    exactly 8 bytes = a view of the chars, at most 7 bytes = bytes | length << 56 so that "a" and "a\0" stay apart);
    anything else -- an 8-byte string among shorter ones, NA rows -- is factorized.  Every path against the oracle and
    against the factorize path, sort=True and first-appearance order."""
    import torch
    from sdc_b200.str_arr import StringArray, groupby_str_device
    rng = np.random.default_rng(17)
    n = 60_000
    if kind in ("upto7", "upto7_na"):
        pool = ["", "a", "a\x00", "a\x00\x00", "\x00", "\x00\x00a", "ab", "abcdefg", "abcdef\x00", "zzzzzzz", "\u00e9t\u00e9", "Zq9"]
        pool += _rand_strings(rng, 300, 7, "abcXYZ019")
    elif kind == "mixed8":
        pool = ["abcdefgh", "abcdefg", "abcdefg\x00", "a", ""] + _rand_strings(rng, 200, 8, "abcXYZ019")
    else:
        pool = ["abcdefgh", "abcdefg\x00", "\x00" * 8, "\u00e9t\u00e9xy!"]
        pool += ["".join(rng.choice(list("abcXYZ019"), 8)) for _ in range(400)]
        assert all(len(p.encode()) == 8 for p in pool)
    keys = [pool[i] for i in rng.integers(0, len(pool), n)]
    if kind == "upto7_na":
        keys[11] = None
        keys[4000] = None
    v = rng.standard_normal(n)
    v[rng.integers(0, n, n // 30)] = np.nan
    sa = StringArray.from_pylist(keys).cuda()
    if kind == "fixed8_unaligned_view":            # chars at an address that is not 16-byte aligned: no zero-copy view
        buf = torch.empty(sa.chars.shape[0] + 8, dtype=torch.uint8, device="cuda")
        buf[8:] = sa.chars
        sa = StringArray(sa.offsets, buf[8:], sa.bitmap, sa.n)
    kw = sa.key_words()
    assert (kw is not None) == (kind in ("upto7", "fixed8")), kind
    if kw is not None:
        assert kw[1] == (kind == "fixed8")
    val = torch.from_numpy(v).cuda()
    aggs = ("sum", "mean", "min", "max", "count")
    for sort in (True, False):
        wants = {agg: ogroupby.groupby_agg(keys, {"v": v}, agg, sort) for agg in aggs}
        for hint in (0, len(pool)):
            rk, res = groupby_str_device(sa, [val], aggs, sort, 1, hint)
            fk, fres = groupby_str_device(sa, [val], aggs, sort, 1, hint, short_keys=False)
            got_keys = rk.to_pylist()
            assert got_keys == fk.to_pylist(), (kind, sort, hint)
            assert got_keys == list(wants["count"][0]), (kind, sort, hint)
            for agg in aggs:
                want = wants[agg][1]
                g = res[agg][0].cpu().numpy()
                if agg in ("min", "max", "count"):
                    np.testing.assert_array_equal(g, want["v"], err_msg="%s %s" % (kind, agg))
                    np.testing.assert_array_equal(g, fres[agg][0].cpu().numpy())
                else:
                    np.testing.assert_allclose(g, want["v"], rtol=1e-9, atol=1e-12, equal_nan=True, err_msg="%s %s" % (kind, agg))


def test_str_profile_and_word_round_trip():
    """sdcb200_str_profile (shortest / longest / NA count) against numpy, str_pack_words -> str_unpack_words is the identity."""
    from sdc_b200.str_arr import StringArray
    rng = np.random.default_rng(3)
    for n, maxlen, na in ((1, 3, 0), (5, 7, 2), (4099, 7, 0), (100_003, 5, 17), (33, 40, 1)):
        vals = _rand_strings(rng, n, maxlen)
        for i in rng.integers(0, n, na):
            vals[i] = None
        sa = StringArray.from_pylist(vals)
        lens = np.diff(sa.offsets.astype(np.int64))
        assert sa.cuda().profile() == (int(lens.min()), int(lens.max()), sum(x is None for x in vals)), (n, maxlen)
        if maxlen <= 7 and not any(x is None for x in vals):
            d = sa.cuda()
            words, fixed8 = d.key_words()
            assert not fixed8
            back = StringArray.from_key_words(words, False)
            assert back.to_pylist() == vals


def test_groupby_string_keys_skewed_one_float_column():
    """Half of the rows share one string: the dense ids go through the lean loop, whose hot-key search samples the ids
    (group-id keys have no window search of their own) and keeps the hot group's sum / count in registers."""
    from sdc_b200.groupby import groupby_host
    rng = np.random.default_rng(41)
    n = 120_000
    pool = _rand_strings(rng, 60, 5, "abcdefghijklmnopqrstuvwxyz")
    pick = np.where(rng.random(n) < 0.5, 7, rng.integers(0, len(pool), n))
    keys = [pool[i] for i in pick]
    v = rng.standard_normal(n) * 3
    v[rng.integers(0, n, n // 40)] = np.nan
    for aggs in (("sum",), ("mean",), ("sum", "mean", "min", "max", "count")):
        rk, res = groupby_host(keys, {"v": v}, aggs, True)
        for agg in aggs:
            wk, want = ogroupby.groupby_agg(keys, {"v": v}, agg, True)
            assert list(rk) == list(wk)
            if agg in ("min", "max", "count"):
                np.testing.assert_array_equal(res[agg]["v"], want["v"])
            else:
                np.testing.assert_allclose(res[agg]["v"], want["v"], rtol=1e-9)


def test_c4_shape_scaled_properties():
    """BASELINE config C4 (string keys, sort + multi-agg) at 20 M rows: fixed-width 8-char keys built natively;
    checked through invariants -- sorted order is non-decreasing, counts sum to n, sums match a float64 total."""
    import torch
    from sdc_b200.str_arr import StringArray, groupby_str_device
    n, G = 20_000_000, 1000
    rng = np.random.default_rng(4)
    al = np.frombuffer(b"abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789", dtype=np.uint8)
    pool = al[rng.integers(0, len(al), (G, 8))]
    which = rng.integers(0, G, n)
    sa = StringArray.from_fixed_width(pool[which]).cuda()
    val = torch.from_numpy(rng.random(n)).cuda()
    order = sa.argsort(True)
    keys64 = torch.from_numpy(np.ascontiguousarray(pool[which]).view(">u8").astype(np.uint64).reshape(-1).view(np.int64)).cuda()
    sorted_keys = keys64[order]
    # big-endian 8-byte value order == byte-lexicographic order; compare as unsigned via the sign flip
    flipped = sorted_keys ^ torch.tensor(-2**63, dtype=torch.int64, device="cuda")
    assert bool((flipped[1:] >= flipped[:-1]).all())
    # stability: equal keys keep ascending row order
    same = flipped[1:] == flipped[:-1]
    assert bool((order[1:][same] > order[:-1][same]).all())
    rk, res = groupby_str_device(sa, [val], ("sum", "mean", "min", "max", "count"), True, 1, G)
    assert rk.n == len(np.unique(which))
    assert int(res["count"][0].sum().item()) == n
    np.testing.assert_allclose(float(res["sum"][0].sum().item()), float(val.sum().item()), rtol=1e-9)
    got_keys = rk.to_pylist()
    assert got_keys == sorted(got_keys, key=lambda s: s.encode())
    # one group checked exactly against numpy
    g0 = got_keys[0].encode()
    rows = np.flatnonzero((pool[which] == np.frombuffer(g0, dtype=np.uint8)).all(axis=1))
    assert int(res["count"][0][0].item()) == len(rows)
    np.testing.assert_allclose(float(res["min"][0][0].item()), val.cpu().numpy()[rows].min(), rtol=0)


def test_c4_full_size_properties():
    """BASELINE config C4 at its full 200 M rows (8-char string keys, G = 1 K; built on the device like
    tools/bench_ops.py): sortedness + stability of the argsort, and the multi-aggregate checked through totals."""
    import torch
    from sdc_b200.str_arr import StringArray, groupby_str_device
    n, G = 200_000_000, 1000
    rng = np.random.default_rng(4)
    al = np.frombuffer(b"abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789", dtype=np.uint8)
    pool = al[rng.integers(0, len(al), (G, 8))]
    dic = torch.from_numpy(pool).cuda()
    ids = torch.randint(0, G, (n,), device="cuda", generator=torch.Generator(device="cuda").manual_seed(4))
    chars = dic[ids].reshape(-1).contiguous()
    offsets = (torch.arange(n + 1, device="cuda", dtype=torch.int64) * 8).to(torch.int32)
    bitmap = torch.full(((n + 7) // 8,), 0xFF, dtype=torch.uint8, device="cuda")
    sa = StringArray(offsets, chars, bitmap, n)
    # big-endian value of the 8 bytes = byte-lexicographic order; computed per dictionary entry, looked up per row
    be = torch.from_numpy(pool.copy().view(">u8").astype(np.uint64).reshape(-1).view(np.int64)).cuda()
    flip = torch.tensor(-2**63, dtype=torch.int64, device="cuda")
    order = sa.argsort(True)
    sk = (be[ids[order]] ^ flip)
    assert bool((sk[1:] >= sk[:-1]).all())
    same = sk[1:] == sk[:-1]
    assert bool((order[1:][same] > order[:-1][same]).all())          # stable
    del sk, same, order
    val = torch.rand(n, device="cuda", dtype=torch.float64)
    rk, res = groupby_str_device(sa, [val], ("sum", "mean", "min", "max", "count"), True, 1, G)
    present = torch.unique(ids)
    assert rk.n == int(present.shape[0])
    assert int(res["count"][0].sum().item()) == n
    np.testing.assert_allclose(float(res["sum"][0].sum().item()), float(val.sum().item()), rtol=1e-9)
    got_keys = rk.to_pylist()
    assert got_keys == sorted(got_keys, key=lambda s_: s_.encode())
    # the counts per key against a device bincount over the dictionary ids, in string order
    cnt = torch.bincount(ids, minlength=G)[present]
    key_order = np.argsort(pool[present.cpu().numpy()].copy().view(">u8").reshape(-1), kind="stable")
    np.testing.assert_array_equal(res["count"][0].cpu().numpy(), cnt.cpu().numpy()[key_order])
    assert float(res["min"][0].min().item()) == float(val.min().item())
    assert float(res["max"][0].max().item()) == float(val.max().item())


@pytest.mark.parametrize("pool_kind", ["mixed", "single", "sample_lies"])
def test_argsort_few_distinct_strings_rank_path(pool_kind):
    """>= 2^20 rows with few distinct strings take the factorize -> rank -> radix-sort-of-ranks path of
    sdcb200_str_argsort: same order as the reference's stable sort in both directions (ties in input order), for
    strings shorter and longer than the 8-byte key word, the empty string, raw 0xFF bytes; a single distinct string;
    and a column whose first rows hold few strings but whose tail holds many (back to the LSD sort)."""
    from sdc_b200.str_arr import StringArray
    rng = np.random.default_rng(77)
    n = (1 << 20) + 4321
    if pool_kind == "mixed":
        pool = ["", "a", "ab", "ab\x00", "b", "abcdefgh", "abcdefghi", "abcdefghij" * 4, "zz", "été", "Zebra", "zebra"] \
            + _rand_strings(rng, 300, 12)
        idx = rng.integers(0, len(pool), n)
    elif pool_kind == "single":
        pool = ["only"]
        idx = np.zeros(n, dtype=np.int64)
    else:
        pool = _rand_strings(rng, 200_000, 10)
        idx = np.concatenate([rng.integers(0, 50, 1 << 18), rng.integers(0, len(pool), n - (1 << 18))])
    enc = [p.encode("utf-8") for p in pool]
    lens = np.array([len(e) for e in enc], dtype=np.int64)[idx]
    offsets = np.zeros(n + 1, dtype=np.uint32)
    np.cumsum(lens, out=offsets[1:], dtype=np.uint64)
    chars = np.frombuffer(b"".join(enc[i] for i in idx), dtype=np.uint8).copy()
    sa = StringArray(offsets, chars, np.full((n + 7) // 8, 0xFF, dtype=np.uint8), n)
    # oracle order through the pool: stable sort of the rows by their string's rank
    order_pool = sorted(range(len(pool)), key=lambda i: enc[i])
    rank = np.empty(len(pool), dtype=np.int64)
    r = -1
    prev = None
    for i in order_pool:                       # equal byte strings (none here, but the pool is random) share a rank
        if enc[i] != prev:
            r += 1
            prev = enc[i]
        rank[i] = r
    row_rank = rank[idx]
    for asc in (True, False):
        got = sa.argsort(asc).cpu().numpy()
        want = np.argsort(row_rank if asc else -row_rank, kind="stable")
        np.testing.assert_array_equal(got, want, err_msg="%s asc=%s" % (pool_kind, asc))
