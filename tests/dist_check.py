"""Multi-GPU parity check, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py

Every rank builds the SAME global synthetic table (seeded), takes its block of rows
(sdc/_distributed.h:77-102), runs the distributed groupby / join / rolling, and the gathered results are
compared with the oracle on the global table.  Exits non-zero on any mismatch.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    from oracle import groupby as ogroupby
    from oracle import join as ojoin
    from oracle import rolling as orolling
    from sdc_b200 import distributed as D

    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    P = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rng = np.random.default_rng(5)
    n = 400_003
    keys_lo = rng.integers(0, 1000, n)
    keys_hi = rng.integers(0, 150_000, n)
    vals = rng.standard_normal(n)
    vals[rng.integers(0, n, n // 100)] = np.nan
    ivals = rng.integers(-1000, 1000, n)
    s, e = D.block_range(n, rank, P)
    cols = [torch.from_numpy(vals[s:e].copy()).to(dev), torch.from_numpy(ivals[s:e].copy()).to(dev)]
    names = {"v": vals, "w": ivals}

    peer = D.PeerShuffle(2 * n // P + 4096, max_cols=3, narenas=2)      # NVLink peer-memory shuffle (CUDA IPC)

    def check_groupby(keys, aggs, mode, sort, hint, peer=None, route="auto"):          # peer=None: NCCL all-to-all
        k = torch.from_numpy(keys[s:e].copy()).to(dev)
        gk, res, layout = D.groupby_distributed(k, cols, aggs, sort, 1, hint, None, mode, peer, route)
        if layout == "sharded":
            gk_all, _ = D.allgather_ragged(gk)
            res = {a: [D.allgather_ragged(t)[0] for t in ts] for a, ts in res.items()}
            order = torch.argsort(gk_all)                       # test-side canonical order
            gk = gk_all[order]
            res = {a: [t[order] for t in ts] for a, ts in res.items()}
            sort_cmp = True
        else:
            sort_cmp = sort
        for a in aggs:
            wk, want = ogroupby.groupby_agg(keys, names, a, sort_cmp)
            np.testing.assert_array_equal(gk.cpu().numpy(), wk, err_msg="%s %s keys" % (mode, a))
            for i, c in enumerate(("v", "w")):
                g, w = res[a][i].cpu().numpy(), want[c]
                if a in ("min", "max", "count") or w.dtype.kind == "i":
                    np.testing.assert_array_equal(g, w, err_msg="%s %s %s" % (mode, a, c))
                else:
                    np.testing.assert_allclose(g, w, rtol=1e-9, atol=1e-12, equal_nan=True, err_msg="%s %s %s" % (mode, a, c))

    check_groupby(keys_lo, ("sum", "mean", "min", "max", "count"), "combine", True, 1000)
    check_groupby(keys_lo, ("sum", "count"), "combine", False, 1000)
    check_groupby(keys_lo, ("sum", "mean", "count"), "combine", True, 1000)        # every rank holds every key: row-wise add of the partial tables
    check_groupby(keys_hi, ("sum", "mean", "min", "max", "count"), "shuffle", True, 150_000)
    check_groupby(keys_hi, ("var", "std"), "shuffle", True, 150_000)
    check_groupby(keys_hi, ("sum", "mean", "min", "max", "count"), "shuffle", True, 150_000, peer)
    check_groupby(keys_hi, ("sum", "count"), "shuffle", True, 150_000, "auto")        # cached context of the group
    # the peer-memory shuffle delivers exactly what the NCCL all-to-all delivers: same rows, same order
    kh = torch.from_numpy(keys_hi[s:e].copy()).to(dev)
    for rep in range(3):                                   # arenas rotate: results of consecutive calls must not alias
        k_a, c_a = D.hash_shuffle(kh, cols)
        k_b, c_b = D.hash_shuffle(kh, cols, None, peer)
        assert torch.equal(k_a, k_b), "peer shuffle keys differ (rep %d)" % rep
        for x_a, x_b in zip(c_a, c_b):
            assert torch.equal(x_a.view(torch.int64), x_b.view(torch.int64)), "peer shuffle column differs (rep %d)" % rep

    # ROUTE_MOD (dense key ranges stay dense): same rows in the same order over both transports, keys travel as
    # floor(key / P) and every row lands on rank key mod P
    k_a, c_a = D.hash_shuffle(kh, cols, None, None, D.ROUTE_MOD)
    k_b, c_b = D.hash_shuffle(kh, cols, None, peer, D.ROUTE_MOD)
    assert torch.equal(k_a, k_b), "ROUTE_MOD: peer / NCCL keys differ"
    for x_a, x_b in zip(c_a, c_b):
        assert torch.equal(x_a.view(torch.int64), x_b.view(torch.int64)), "ROUTE_MOD: peer / NCCL column differs"
    back = (k_b * P + rank).cpu().numpy()
    assert np.all(np.mod(back, P) == rank), "ROUTE_MOD: a row landed on the wrong rank"
    all_back, _ = D.allgather_ragged(torch.from_numpy(back).to(dev))
    np.testing.assert_array_equal(np.sort(all_back.cpu().numpy()), np.sort(keys_hi), err_msg="ROUTE_MOD: key multiset changed")
    # negative keys: Euclidean mod / floor division
    kneg = torch.from_numpy((keys_hi[s:e] - 75_000).copy()).to(dev)
    k_n, _ = D.hash_shuffle(kneg, [], None, peer, D.ROUTE_MOD)
    backn = (k_n * P + rank).cpu().numpy()
    alln, _ = D.allgather_ragged(torch.from_numpy(backn).to(dev))
    np.testing.assert_array_equal(np.sort(alln.cpu().numpy()), np.sort(keys_hi - 75_000), err_msg="ROUTE_MOD: negative keys")
    # groupby through both routes gives the same groups
    check_groupby(keys_hi, ("sum", "count", "min"), "shuffle", True, 150_000, peer, D.ROUTE_HASH)
    check_groupby(keys_hi, ("sum", "count", "min"), "shuffle", True, 150_000, peer, D.ROUTE_MOD)
    check_groupby(keys_hi * 2, ("sum", "count"), "shuffle", True, 150_000, "auto", "auto")   # key mod 2 == 0 everywhere: auto falls back to hashing

    # join: probe rows sharded, build rows sharded; global row numbers come back
    nb = 20_011
    build = rng.permutation(60_000)[:nb].astype(np.int64)
    probe = rng.integers(0, 60_000, n)
    bs, be = D.block_range(nb, rank, P)
    def check_join(probe, build, mode, how, route="auto", pr=None):
        ps, pe = D.block_range(len(probe), rank, P)
        bs, be = D.block_range(len(build), rank, P)
        li, ri = D.join_distributed(torch.from_numpy(probe[ps:pe].copy()).to(dev), torch.from_numpy(build[bs:be].copy()).to(dev),
                                    how, None, "shuffle" if mode.startswith("shuffle") else mode, pr, route)
        li_all, _ = D.allgather_ragged(li)
        ri_all, _ = D.allgather_ragged(ri)
        got = ojoin.canonical(li_all.cpu().numpy(), ri_all.cpu().numpy())
        want = ojoin.join_indexers(probe, build, how)
        np.testing.assert_array_equal(got[0], want[0], err_msg="%s %s %s" % (mode, how, route))
        np.testing.assert_array_equal(got[1], want[1], err_msg="%s %s %s" % (mode, how, route))

    for mode, pr in (("broadcast", None), ("shuffle", None), ("shuffle_peer", peer), ("shuffle_auto", "auto")):
        for how in ("inner", "left"):
            check_join(probe, build, mode, how, "auto", pr)
    for route in (D.ROUTE_HASH, D.ROUTE_MOD):
        check_join(probe, build, "shuffle_peer", "inner", route, peer)
    check_join(probe, build, "shuffle_peer", "outer", "auto", peer)
    # duplicate build keys, negative keys, and a build side about twice the probe side (the cached context was sized
    # for smaller inputs: it has to grow BEFORE anything is sent, with nothing dangling)
    dupb = rng.integers(-5_000, 5_000, 30_000)
    check_join(rng.integers(-6_000, 6_000, 15_000), dupb, "shuffle_auto", "inner", "auto", "auto")
    big_build = rng.integers(0, 3_000_000, 2 * n)
    check_join(rng.integers(0, 3_000_000, n), big_build, "shuffle_auto", "left", "auto", "auto")

    # rolling: row-sharded with the halo from the previous rank -- through the all-gather of the shard tails and
    # through the peer-memory link inside the kernel (HaloLink); several calls in a row exercise the slot rotation
    x = rng.random(n)
    x[::977] = np.nan
    link = D.HaloLink()

    def check_rolling(xx, op, win, minp, lk, bounds=None):
        ss, ee = bounds if bounds is not None else D.block_range(len(xx), rank, P)
        out = D.rolling_distributed(torch.from_numpy(xx[ss:ee].copy()).to(dev), op, win, minp, 1, None, None, None, lk)
        full, _ = D.allgather_ragged(out)
        want = orolling.rolling(xx, op, win, minp, 1)
        g = full.cpu().numpy()
        assert np.array_equal(np.isnan(g), np.isnan(want)), (op, win, lk is not None)
        ok = ~np.isnan(want)
        np.testing.assert_allclose(g[ok], want[ok], rtol=1e-9, atol=1e-11, err_msg="%s %d" % (op, win))

    for lk in (None, link):
        for op in ("mean", "std", "count", "max", "sum", "min"):
            check_rolling(x, op, 100, 100, lk)
        check_rolling(x, "max", 5, 2, lk)               # short window: the segmented min / max kernel (generic link form)
        check_rolling(x, "mean", 3000, 10, lk)
        check_rolling(x, "sum", 1, 1, lk)
        # shards shorter than the window: rank r holds 7 + r rows, a window of 40 reaches several shards back
        tiny = rng.random(sum(7 + r for r in range(P)))
        lo = sum(7 + r for r in range(rank))
        check_rolling(tiny, "mean", 40, 1, lk, (lo, lo + 7 + rank))
        check_rolling(tiny, "min", 40, 3, lk, (lo, lo + 7 + rank))
    link.close()

    # string keys (str_arr offsets + chars): length array + characters travel, offsets rebuilt by a scan on arrival
    from sdc_b200.str_arr import StringArray
    words = np.array(["k%05d" % i if i % 7 else "key-%d-long-%s" % (i, "x" * (i % 13)) for i in range(3000)], dtype=object)
    skeys = words[rng.integers(0, 3000, n)]
    skeys[rng.integers(0, n, n // 200)] = None                       # NA keys are skipped by the aggregate
    sa = StringArray.from_pylist(list(skeys[s:e])).cuda()
    gk, res, layout = D.groupby_str_distributed(sa, cols, ("sum", "count", "max"), True, 1, 3000)
    assert layout == "sharded"
    mine = gk.to_pylist()
    assert mine == sorted(mine), "string groups of a rank come out in string order"
    import pandas as pd
    want = pd.DataFrame({"k": skeys, "v": vals, "w": ivals}).groupby("k")
    got_keys = [None] * P
    dist.all_gather_object(got_keys, mine)
    flat = [k for part in got_keys for k in part]
    assert sorted(flat) == sorted(want.sum().index.tolist()) and len(set(flat)) == len(flat), "string group keys"
    wsum, wcnt, wmax = want.sum(), want.count(), want.max()
    np.testing.assert_allclose(res["sum"][0].cpu().numpy(), wsum.loc[mine, "v"].to_numpy(), rtol=1e-9, atol=1e-12)
    np.testing.assert_array_equal(res["sum"][1].cpu().numpy(), wsum.loc[mine, "w"].to_numpy())
    np.testing.assert_array_equal(res["count"][0].cpu().numpy(), wcnt.loc[mine, "v"].to_numpy())
    np.testing.assert_array_equal(res["max"][1].cpu().numpy(), wmax.loc[mine, "w"].to_numpy())
    # the shuffle alone: every string lands on exactly one rank, with its payload row
    k2, (v2,) = D.str_shuffle(sa, [cols[1]])
    back = pd.DataFrame({"k": k2.to_pylist(), "w": v2.cpu().numpy()})
    allb = [None] * P
    dist.all_gather_object(allb, back)
    allb = pd.concat(allb)
    ref = pd.DataFrame({"k": skeys, "w": ivals})
    key = lambda d: d.fillna("<NA>").sort_values(["k", "w"]).reset_index(drop=True)
    pd.testing.assert_frame_equal(key(allb), key(ref))
    per_rank = [None] * P
    dist.all_gather_object(per_rank, set(x for x in k2.to_pylist() if x is not None))
    for a in range(P):
        for b in range(a + 1, P):
            assert not (per_rank[a] & per_rank[b]), "a string key landed on two ranks"

    # sort_values across GPUs (sample sort with range routing): the ranks' slices, concatenated in rank order, are the
    # single-GPU stable sort_values -- keys, original row numbers (ties in row order) and a payload column
    from oracle import sort as osort
    fk = rng.standard_normal(n).round(2)                       # many ties
    fk[rng.integers(0, n, n // 50)] = np.nan
    fk[rng.integers(0, n, 100)] = -0.0
    fk[rng.integers(0, n, 100)] = np.inf
    ik = rng.integers(-500, 500, n)
    pay = rng.integers(0, 1 << 40, n)
    for gkeys in (fk, ik, np.full(n, 7, dtype=np.int64), rng.standard_normal(n)):
        for asc in (True, False):
            for nap in ("last", "first"):
                dk = torch.from_numpy(gkeys[s:e].copy()).to(dev)
                sk, sid, (sp,) = D.sort_values_distributed(dk, asc, nap, payload=[torch.from_numpy(pay[s:e].copy()).to(dev)])
                all_k, _ = D.allgather_ragged(sk)
                all_i, _ = D.allgather_ragged(sid)
                all_p, _ = D.allgather_ragged(sp)
                want = osort.sort_values_indexer(gkeys, asc, nap)
                np.testing.assert_array_equal(all_i.cpu().numpy(), want, err_msg="sort_values rows asc=%s na=%s" % (asc, nap))
                np.testing.assert_array_equal(all_k.cpu().numpy(), gkeys[want])
                np.testing.assert_array_equal(all_p.cpu().numpy(), pay[want])

    peer.close()
    dist.barrier()
    if rank == 0:
        print("dist_check ok: world=%d" % P)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
