"""GPU parity: csrc/join.cu (hash join) against the oracle's sort-merge restatement -- exact row sets."""
import numpy as np
import pandas as pd
import pytest

from oracle import join as ojoin

pytestmark = pytest.mark.gpu


def _device_pairs(l, r, how, ordered=True):
    import torch
    from sdc_b200.join import join_device
    li, ri = join_device(torch.from_numpy(np.ascontiguousarray(l)).cuda(), torch.from_numpy(np.ascontiguousarray(r)).cuda(), how,
                         ordered=ordered)
    li, ri = li.cpu().numpy(), ri.cpu().numpy()
    # inner (ordered=True) / left: pairs come out in left-row order (how='left' keeps the left order like pandas); right:
    # in right-row order; outer: the left join's pairs in left-row order, then the unmatched right rows in row order;
    # inner with ordered=False (SDCB200_JOIN_ANY_ORDER) is a row set
    if how == "inner" and not ordered:
        pass
    elif how in ("inner", "left"):
        assert np.all(np.diff(li) >= 0)
    elif how == "right":
        assert np.all(np.diff(ri) >= 0)
    else:
        tail = int((li < 0).sum())
        head = len(li) - tail
        assert np.all(li[:head] >= 0) and np.all(np.diff(li[:head]) >= 0)
        assert np.all(li[head:] == -1) and np.all(np.diff(ri[head:]) > 0)
    return ojoin.canonical(li, ri)


def _check(l, r, how):
    want = ojoin.join_indexers(l, r, how)
    for ordered in ((True, False) if how == "inner" else (True,)):
        got = _device_pairs(l, r, how, ordered)
        np.testing.assert_array_equal(got[0], want[0])
        np.testing.assert_array_equal(got[1], want[1])


@pytest.mark.parametrize("how", ["inner", "left", "right", "outer"])
def test_small_cases(how):
    i64 = np.int64
    sentinel = np.array([0x7ff8dead0000beef], dtype=np.uint64).view(np.int64)[0]
    cases = [
        (np.array([3, 5, 1, 0, 4], dtype=i64), np.array([5, 4, 2, 8, 0], dtype=i64)),
        (np.array([3, 1, 1, 3, 4], dtype=i64), np.array([2, 1, 3, 1, 9, 3], dtype=i64)),
        (np.array([0.5, np.nan, 2.0, -0.0, np.nan]), np.array([np.nan, 0.0, 2.0, 7.5])),
        (np.array([1, 2, 3], dtype=i64), np.array([], dtype=i64)),
        (np.array([], dtype=i64), np.array([1, 2, 3], dtype=i64)),
        (np.array([sentinel, 1, sentinel], dtype=i64), np.array([7, sentinel, sentinel], dtype=i64)),
        (np.array([-2**63, 2**63 - 1, 0], dtype=i64), np.array([2**63 - 1, -2**63, -2**63], dtype=i64)),
    ]
    for l, r in cases:
        _check(l, r, how)


@pytest.mark.parametrize("how", ["inner", "left", "right", "outer"])
@pytest.mark.parametrize("nl,nr,hi", [(1, 1, 2), (2047, 100, 50), (2048, 2049, 4000), (100_000, 10_000, 10_000),
                                      (300_000, 300_000, 50_000), (50_000, 500_000, 2_000_000), (200_000, 1000, 10)])
def test_random_sizes(nl, nr, hi, how):
    rng = np.random.default_rng(nl + nr)
    _check(rng.integers(-hi, hi, nl, dtype=np.int64), rng.integers(-hi, hi, nr, dtype=np.int64), how)


def test_heavy_duplicates_cross_product():
    l = np.zeros(3000, dtype=np.int64)
    r = np.zeros(500, dtype=np.int64)
    got = _device_pairs(l, r, "inner")
    assert len(got[0]) == 1_500_000
    want = ojoin.join_indexers(l, r, "inner")
    np.testing.assert_array_equal(got[1], want[1])


def test_float_keys_random():
    rng = np.random.default_rng(5)
    l = rng.integers(0, 2000, 100_000).astype(np.float64) / 4
    r = rng.integers(0, 2000, 30_000).astype(np.float64) / 4
    l[rng.integers(0, len(l), 100)] = np.nan
    r[rng.integers(0, len(r), 3)] = np.nan
    l[::1000] = -0.0
    r[5] = 0.0
    for how in ("inner", "left", "right", "outer"):
        _check(l, r, how)


def test_merge_frames_like_reference_tests():
    # sdc/tests/test_join.py:67-81 (inner, key1), :237-260 (left join, NaN fill), compared like :275-279
    from sdc_b200.join import merge
    df1 = pd.DataFrame({"key1": [2, 3, 5, 1, 2, 8], "A": np.array([4, 6, 3, 9, 9, -1], np.float64)})
    df2 = pd.DataFrame({"key1": [1, 2, 9, 3, 2], "B": np.array([1, 7, 2, 6, 5], np.float64)})
    for how in ("inner", "left", "right", "outer"):      # right / outer: sdc/tests/test_join.py:262-335
        got = merge(df1, df2, how=how, on="key1")
        want = df1.merge(df2, how=how, on="key1")
        key = ["key1", "A", "B"]
        pd.testing.assert_frame_equal(got.sort_values(key).reset_index(drop=True),
                                      want.sort_values(key).reset_index(drop=True))
    # how='left' preserves the left order exactly when the build keys are unique
    df3 = pd.DataFrame({"key1": [9, 2, 1], "C": [10, 20, 30]})
    pd.testing.assert_frame_equal(merge(df1, df3, how="left", on="key1"), df1.merge(df3, how="left", on="key1"))
    # left_on / right_on and overlapping payload names
    df4 = pd.DataFrame({"k": [2, 3], "A": [0.5, 1.5]})
    got = merge(df1, df4, how="inner", left_on="key1", right_on="k")
    want = df1.merge(df4, how="inner", left_on="key1", right_on="k")
    cols = list(want.columns)
    pd.testing.assert_frame_equal(got[cols].sort_values(cols).reset_index(drop=True),
                                  want.sort_values(cols).reset_index(drop=True))
    with pytest.raises(ValueError):
        merge(df1, df2, how="cross", on="key1")


@pytest.mark.parametrize("how", ["inner", "left", "right", "outer"])
def test_join_ids_replace_row_numbers(how):
    """sdcb200_join_ids: the caller's ids come out in place of local row numbers (-1 stays -1) on every path -- unique
    build keys (mapping folded into the probe's last kernel), duplicate build keys, keys with no dense range."""
    import torch
    from sdc_b200.join import join_device
    rng = np.random.default_rng(101)
    cases = [(rng.integers(0, 60_000, 200_001, dtype=np.int64), rng.permutation(80_000)[:50_000].astype(np.int64)),   # unique, dense
             (rng.integers(0, 5_000, 50_000, dtype=np.int64), rng.integers(0, 5_000, 20_000, dtype=np.int64)),        # duplicates
             (rng.integers(0, 60_000, 100_000, dtype=np.int64) * 1_000_003, rng.permutation(80_000)[:50_000].astype(np.int64) * 1_000_003)]
    for l, r in cases:
        lid = rng.permutation(len(l)).astype(np.int64) + 10**12
        rid = rng.permutation(len(r)).astype(np.int64) + 2 * 10**12
        dl, dr = torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda()
        for ordered in ((True, False) if how == "inner" else (True,)):
            li, ri = join_device(dl, dr, how, ordered=ordered, left_ids=torch.from_numpy(lid).cuda(), right_ids=torch.from_numpy(rid).cuda())
            pl, pr = join_device(dl, dr, how, ordered=ordered)
            pl, pr = pl.cpu().numpy(), pr.cpu().numpy()
            want_l = np.where(pl < 0, -1, lid[np.maximum(pl, 0)])
            want_r = np.where(pr < 0, -1, rid[np.maximum(pr, 0)])
            got = ojoin.canonical(li.cpu().numpy(), ri.cpu().numpy())
            want = ojoin.canonical(want_l, want_r)
            np.testing.assert_array_equal(got[0], want[0])
            np.testing.assert_array_equal(got[1], want[1])
        li, _ = join_device(dl, dr, how, ordered=True, left_ids=torch.from_numpy(lid).cuda())          # one side only
        np.testing.assert_array_equal(np.sort(li.cpu().numpy()), np.sort(want_l))


def test_c3_shape_row_count_and_properties():
    """BASELINE config C3 at 1/10 scale through the oracle, and at full size through invariants:
    unique build keys, every probe row matches once -> n_out == n_probe, ridx is the inverse permutation."""
    import torch
    from sdc_b200.join import join_device
    nb, npr = 10_000_000, 100_000_000
    perm = np.random.default_rng(2).permutation(nb).astype(np.int64)
    probe = np.random.default_rng(3).integers(0, nb, npr, dtype=np.int64)
    r, l = torch.from_numpy(perm).cuda(), torch.from_numpy(probe).cuda()
    li, ri = join_device(l, r, "inner", ordered=True)
    assert li.shape[0] == npr
    assert bool((li == torch.arange(npr, device="cuda")).all())
    assert bool((r[ri] == l).all())
    del li, ri
    # the row-set form (one kernel, a tile takes its output range with an atomic): the same pairs in some order
    li, ri = join_device(l, r, "inner")
    assert li.shape[0] == npr
    assert bool((r[ri] == l[li]).all())
    seen = torch.zeros(npr, dtype=torch.bool, device="cuda")
    seen[li] = True
    assert bool(seen.all())                                   # npr pairs covering every left row: each exactly once
    del li, ri, seen
    # 50 % hit rate, left join: every probe row appears once, misses carry -1
    probe2 = np.random.default_rng(3).integers(0, 2 * nb, npr, dtype=np.int64)
    l2 = torch.from_numpy(probe2).cuda()
    li, ri = join_device(l2, r, "left")
    assert li.shape[0] == npr
    miss = ri < 0
    assert bool((l2[miss] >= nb).all()) and bool((l2[~miss] < nb).all())
    assert bool((r[ri[~miss]] == l2[~miss]).all())
    li_in, _ = join_device(l2, r, "inner")
    assert li_in.shape[0] == int((~miss).sum().item())


def test_partitioned_inner_join_matches_oracle(monkeypatch):
    """Keys with no dense range and a table far larger than L2 take the partitioned path (both sides split by hash
    bucket, slots from the top hash bits).  SDCB200_JOIN_PART=2 forces it at test sizes; the pairs come out in bucket
    order, so only the row SET is compared -- the contract of an inner join."""
    import torch
    from sdc_b200.join import join_device
    monkeypatch.setenv("SDCB200_JOIN_PART", "2")
    rng = np.random.default_rng(77)
    sentinel = np.array([0x7ff8dead0000beef], dtype=np.uint64).view(np.int64)[0]
    cases = []
    r = rng.permutation(400_000)[:150_000].astype(np.int64) * 1_000_003              # unique, sparse
    cases.append((rng.integers(0, 400_000, 700_001, dtype=np.int64) * 1_000_003, r))
    cases.append((rng.integers(-3_000, 3_000, 50_000, dtype=np.int64) * 7_777_777, rng.integers(-3_000, 3_000, 20_000, dtype=np.int64) * 7_777_777))   # duplicates on both sides
    cases.append((np.array([sentinel, 5, sentinel, 9], dtype=np.int64), np.array([9, sentinel, 11], dtype=np.int64)))
    f = rng.standard_normal(30_000).round(2)
    f[::50] = np.nan
    cases.append((f, np.concatenate([f[:5_000] * 1.0, [np.nan, -0.0, 0.0]])))
    cases.append((np.array([1, 2, 3], dtype=np.int64), np.array([], dtype=np.int64)))
    # the table can be built and probed one group of buckets at a time: one group (the default), 256 groups
    for group_mb in ("-1", "0"):
        monkeypatch.setenv("SDCB200_JOIN_GROUP_MB", group_mb)
        for l, r in cases:
            li, ri = join_device(torch.from_numpy(np.ascontiguousarray(l)).cuda(), torch.from_numpy(np.ascontiguousarray(r)).cuda(), "inner")
            got = ojoin.canonical(li.cpu().numpy(), ri.cpu().numpy())
            want = ojoin.join_indexers(l, r, "inner")
            np.testing.assert_array_equal(got[0], want[0])
            np.testing.assert_array_equal(got[1], want[1])
    # keys whose hashes all fall into ONE bucket (top byte of (x ^ x >> 32) * 0x9e3779b97f4a7c15 == 0, common.cuh
    # hash_mix64): the bucket's slice cannot hold them and the build falls back to the table-wide path
    ginv = pow(0x9e3779b97f4a7c15, -1, 1 << 64)
    h = rng.integers(0, 1 << 56, 60_000, dtype=np.uint64)
    y = np.array([(int(v) * ginv) & ((1 << 64) - 1) for v in np.unique(h)], dtype=np.uint64)
    x = (y ^ (y >> np.uint64(32))).view(np.int64)                                  # x ^ (x >> 32) == y
    assert (((x.view(np.uint64) ^ (x.view(np.uint64) >> np.uint64(32))) * np.uint64(0x9e3779b97f4a7c15)) >> np.uint64(56)).max() == 0
    lx = np.concatenate([x[rng.integers(0, len(x), 300_000)], rng.integers(0, 1 << 40, 50_000, dtype=np.int64)])
    li, ri = join_device(torch.from_numpy(lx).cuda(), torch.from_numpy(x).cuda(), "inner")
    got = ojoin.canonical(li.cpu().numpy(), ri.cpu().numpy())
    want = ojoin.join_indexers(lx, x, "inner")
    np.testing.assert_array_equal(got[0], want[0])
    np.testing.assert_array_equal(got[1], want[1])
    # how='left' keeps left-row order whatever the switch says
    l, r = cases[0]
    li, _ = join_device(torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda(), "left")
    assert torch.equal(li, torch.arange(len(l), device="cuda"))
