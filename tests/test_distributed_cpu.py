"""N > 1 host logic under gloo (world_size 2, CPU): the shuffle protocol, ragged all-gather, halo exchange
and block partition of sdc_b200/distributed.py against the oracle's restatement of the reference protocol.
The device kernels are not involved (those run in tests/dist_check.py on >= 2 GPUs)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world_size, port, fn_name, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        ret[rank] = globals()[fn_name](rank, world_size)
    finally:
        dist.destroy_process_group()


def _run(fn_name, world_size=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world_size, _free_port(), fn_name, ret), nprocs=world_size, join=True)
    return [ret[r] for r in range(world_size)]


def _data(rank):
    rng = np.random.default_rng(100 + rank)
    n = 1000 + 37 * rank
    keys = rng.integers(0, 50, n)
    vals = rng.standard_normal(n)
    dest = (keys * 2654435761 % 7) % 2
    return keys, vals, dest


def _shuffle_case(rank, P):
    from sdc_b200 import distributed as D
    keys, vals, dest = _data(rank)
    perm = np.argsort(dest, kind="stable")                    # what sdcb200_hash_partition returns on a GPU
    counts = torch.from_numpy(np.bincount(dest, minlength=P).astype(np.int64))
    recv = D.exchange_counts(counts)
    k2 = D.exchange_rows(torch.from_numpy(keys[perm]), counts, recv)
    v2 = D.exchange_rows(torch.from_numpy(vals[perm]), counts, recv)
    return recv.numpy(), k2.numpy(), v2.numpy()


def test_shuffle_protocol_matches_reference_restatement():
    from oracle import shuffle as oshuffle
    got = _run("_shuffle_case")
    cols = [[_data(r)[0], _data(r)[1]] for r in range(2)]
    dests = [_data(r)[2] for r in range(2)]
    want, send_counts = oshuffle.shuffle(cols, dests)
    for r in range(2):
        np.testing.assert_array_equal(got[r][0], send_counts[:, r])
        np.testing.assert_array_equal(got[r][1], want[r][0])
        np.testing.assert_array_equal(got[r][2], want[r][1])
        # every key landed on the rank its destination names
        assert np.all((got[r][1] * 2654435761 % 7) % 2 == r)


def _ragged_case(rank, P):
    from sdc_b200 import distributed as D
    t = torch.arange(3 + 4 * rank, dtype=torch.float64) + 100 * rank
    cat, sizes = D.allgather_ragged(t)
    return cat.numpy(), sizes


def test_allgather_ragged():
    got = _run("_ragged_case")
    want = np.concatenate([np.arange(3) + 0.0, np.arange(7) + 100.0])
    for r in range(2):
        np.testing.assert_array_equal(got[r][0], want)
        assert got[r][1] == [3, 7]


def _halo_case(rank, P):
    from sdc_b200 import distributed as D
    x = np.arange(10.0) + 10 * rank
    h = D.halo_exchange(torch.from_numpy(x), 3)
    return None if h is None else h.numpy()


def test_halo_exchange_and_block_partition():
    from oracle import shuffle as oshuffle
    from sdc_b200 import distributed as D
    got = _run("_halo_case")
    assert got[0] is None
    np.testing.assert_array_equal(got[1], [7.0, 8.0, 9.0])
    for n, P in ((9, 4), (10, 3), (1, 2), (0, 2), (100, 8)):
        for r in range(P):
            assert D.block_range(n, r, P) == oshuffle.block_range(n, r, P)
    # sdc/_distributed.h:77-102: ceil(n / P) rows per rank, the tail ranks may be short or empty
    assert [D.block_range(9, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 9)]


def _halo_short_case(rank, P):
    """Shards shorter than the halo: the rows come from as many previous ranks as needed, oldest first."""
    from sdc_b200 import distributed as D
    lens = [5, 1, 0, 4]
    x = np.arange(float(lens[rank])) + 100 * rank
    h = D.halo_exchange(torch.from_numpy(x), 3)
    return None if h is None else h.numpy()


def test_halo_exchange_reaches_several_shards_back():
    got = _run("_halo_short_case", world_size=4)
    assert got[0] is None
    np.testing.assert_array_equal(got[1], [2.0, 3.0, 4.0])
    np.testing.assert_array_equal(got[2], [3.0, 4.0, 100.0])          # rank 1 holds one row only
    np.testing.assert_array_equal(got[3], [3.0, 4.0, 100.0])          # rank 2 is empty: same rows


def _rolling_short_shards_case(rank, P):
    from oracle import rolling as orolling
    from sdc_b200 import distributed as D
    x = np.random.default_rng(11).random(23)
    bounds = [(0, 9), (9, 11), (11, 11), (11, 23)][rank]
    loc = x[bounds[0]:bounds[1]]
    h = D.halo_exchange(torch.from_numpy(loc.copy()), 9)
    full = loc if h is None else np.concatenate([h.numpy(), loc])
    out = orolling.rolling(full, "sum", 10, 1, 1)
    return out[len(full) - len(loc):]


def test_row_sharded_rolling_with_short_and_empty_shards():
    from oracle import rolling as orolling
    got = np.concatenate(_run("_rolling_short_shards_case", world_size=4))
    want = orolling.rolling(np.random.default_rng(11).random(23), "sum", 10, 1, 1)
    np.testing.assert_allclose(got, want, rtol=1e-12)


def _rolling_halo_oracle_case(rank, P):
    """Row-sharded rolling with the halo exchange, the window arithmetic done by the oracle on each shard:
    shard + halo must reproduce the tail of the whole-series result."""
    from oracle import rolling as orolling
    from sdc_b200 import distributed as D
    x = np.random.default_rng(7).random(400)
    s, e = D.block_range(len(x), rank, P)
    h = D.halo_exchange(torch.from_numpy(x[s:e].copy()), 9)
    loc = x[s:e] if h is None else np.concatenate([h.numpy(), x[s:e]])
    out = orolling.rolling(loc, "mean", 10, 10, 1)
    return out[len(loc) - (e - s):]


def test_row_sharded_rolling_via_halo():
    from oracle import rolling as orolling
    got = np.concatenate(_run("_rolling_halo_oracle_case"))
    x = np.random.default_rng(7).random(400)
    want = orolling.rolling(x, "mean", 10, 10, 1)
    # rank 1's first rows see the halo, so only rank 0's own warm-up rows are NaN
    np.testing.assert_allclose(got[200:], want[200:], rtol=1e-12)
    np.testing.assert_allclose(got[:200], want[:200], rtol=1e-12, equal_nan=True)


def _string_wire_case(rank, P):
    """The string column's wire format over gloo: lengths and characters grouped by destination, one alltoallv each,
    offsets by a prefix sum on arrival (what distributed.str_shuffle does around its kernels)."""
    from sdc_b200 import distributed as D
    rng = np.random.default_rng(40 + rank)
    strs = [("s%d" % v).encode() * (1 + v % 3) for v in rng.integers(0, 30, 200 + 11 * rank)]
    dest = np.array([sum(b) % P for b in strs])
    perm = np.argsort(dest, kind="stable")
    packed = [strs[i] for i in perm]
    lens = torch.tensor([len(b) for b in packed], dtype=torch.int32)
    chars = torch.from_numpy(np.frombuffer(b"".join(packed), dtype=np.uint8).copy())
    counts = torch.from_numpy(np.bincount(dest, minlength=P).astype(np.int64))
    send_chars = torch.tensor([sum(len(strs[i]) for i in perm if dest[i] == d) for d in range(P)], dtype=torch.int64)
    both = D.exchange_counts(torch.stack([counts, send_chars], 1).reshape(-1)).view(P, 2)
    lens_r = D.exchange_rows(lens, counts, both[:, 0].contiguous())
    chars_r = D.exchange_rows(chars, send_chars, both[:, 1].contiguous())
    offsets = np.zeros(len(lens_r) + 1, dtype=np.uint32)
    offsets[1:] = np.cumsum(lens_r.numpy().astype(np.uint64))
    return offsets, chars_r.numpy()


def test_string_column_wire_format_matches_reference_restatement():
    from oracle import shuffle as oshuffle
    got = _run("_string_wire_case")
    strs, dests = [], []
    for rank in range(2):
        rng = np.random.default_rng(40 + rank)
        ss = [("s%d" % v).encode() * (1 + v % 3) for v in rng.integers(0, 30, 200 + 11 * rank)]
        strs.append(ss)
        dests.append([sum(b) % 2 for b in ss])
    want = oshuffle.shuffle_strings(strs, dests)
    for r in range(2):
        np.testing.assert_array_equal(got[r][0], want[r][0])
        np.testing.assert_array_equal(got[r][1], want[r][1])


def test_ordered_image_preserves_the_sort_order():
    """Splitters of the distributed sort_values travel as order-preserving uint64 images of the keys (host side of
    sdcb200_shuffle_plan_range): image order == numeric order, -0.0 == +0.0, int64 extremes included."""
    import numpy as np
    from sdc_b200.distributed import ordered_image
    rng = np.random.default_rng(12)
    f = np.concatenate([rng.standard_normal(1000) * 1e6, [0.0, -0.0, np.inf, -np.inf, 5e-324, -5e-324, 1.7e308, -1.7e308]])
    img = ordered_image(f)
    order = np.argsort(img, kind="stable")
    assert np.array_equal(f[order], np.sort(f, kind="stable"))
    assert img[f == 0.0].min() == img[f == 0.0].max()                 # both zeros share one image
    i = np.concatenate([rng.integers(-10**18, 10**18, 1000), [np.iinfo(np.int64).min, np.iinfo(np.int64).max, 0, -1]])
    assert np.array_equal(i[np.argsort(ordered_image(i), kind="stable")], np.sort(i, kind="stable"))
