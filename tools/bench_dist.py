#!/usr/bin/env python
"""C5: census-shaped groupby + join hash-shuffled across the GPUs of one box, one process per GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \\
        tools/bench_dist.py [--rows-per-gpu 125000000] [--reps 3]

Weak scaling: every rank holds rows_per_gpu rows (default 1 B / 8).  Timing = CUDA events on each rank bracketed
by a barrier, MAX over ranks; one JSON line per case from rank 0.  Shuffle cases also report the bytes that left
each GPU and that volume / time as a fraction of the 900 GB/s NVLink direction bandwidth (upper bound on the
exchange alone: the timed region also holds the partition, pack and local aggregate / join kernels).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows-per-gpu", type=int, default=125_000_000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cases", default="groupby_lo,groupby_hi,join_bcast,join_shuffle,rolling")
    args = ap.parse_args()
    from sdc_b200 import distributed as D

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", rank))
    P = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if P > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.rows_per_gpu
    g = torch.Generator(device=dev).manual_seed(5 + rank)

    def timed(fn, warm=1):
        for _ in range(warm):
            fn()
        ts = []
        for _ in range(args.reps):
            if P > 1:
                dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
            if P > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ts.append(float(t.item()))
        return float(np.median(ts))

    def emit(case, ms, rows_total, **kw):
        if rank == 0:
            d = {"case": case, "n_gpus": P, "rows_per_gpu": n, "ms": ms, "rows_per_s": rows_total / (ms * 1e-3)}
            d.update(kw)
            print(json.dumps(d), flush=True)

    cases = args.cases.split(",")
    peer = D.PeerShuffle(int(n * 1.25) + 1024, max_cols=2, narenas=2) if P > 1 else None
    vals = torch.rand(n, device=dev, dtype=torch.float64, generator=g)
    if "groupby_lo" in cases:
        keys = torch.randint(0, 1000, (n,), device=dev, generator=g)
        ms = timed(lambda: D.groupby_distributed(keys, [vals], ("sum",), True, 1, 1000, None, "combine"))
        emit("groupby_sum K=1K (local aggregate + all-gather of partials + combine)", ms, n * P)
        del keys
    if "groupby_hi" in cases:
        K = 100_000_000
        keys = torch.randint(0, K, (n,), device=dev, generator=g)
        ms = timed(lambda: D.groupby_distributed(keys, [vals], ("sum",), True, 1, K, None, "shuffle", None))
        out_bytes = n * 16 * (P - 1) / P
        emit("groupby_sum K=100M (hash shuffle of key + value, local aggregate)", ms, n * P,
             shuffle_bytes_out_per_gpu=out_bytes, nvlink_frac_upper=(out_bytes / (ms * 1e-3)) / 900e9 if P > 1 else None)
        if peer is not None:
            ms = timed(lambda: D.groupby_distributed(keys, [vals], ("sum",), True, 1, K, None, "shuffle", peer))
            emit("groupby_sum K=100M, peer-memory shuffle (one pack-and-send kernel over NVLink)", ms, n * P,
                 shuffle_bytes_out_per_gpu=out_bytes, nvlink_frac_upper=(out_bytes / (ms * 1e-3)) / 900e9)
            ms_nccl = timed(lambda: D.hash_shuffle(keys, [vals]))
            ms_peer = timed(lambda: D.hash_shuffle(keys, [vals], None, peer))
            emit("shuffle alone: NCCL all-to-all per column", ms_nccl, n * P, shuffle_bytes_out_per_gpu=out_bytes,
                 nvlink_frac=(out_bytes / (ms_nccl * 1e-3)) / 900e9)
            emit("shuffle alone: peer-memory stores", ms_peer, n * P, shuffle_bytes_out_per_gpu=out_bytes,
                 nvlink_frac=(out_bytes / (ms_peer * 1e-3)) / 900e9)
        del keys
    nb = max(1, n // 10)
    if "join_bcast" in cases or "join_shuffle" in cases:
        build = (torch.randperm(nb * P, device=dev, generator=torch.Generator(device=dev).manual_seed(99))[rank * nb:(rank + 1) * nb]).contiguous()
        probe = torch.randint(0, nb * P, (n,), device=dev, generator=g)
        for mode in ("broadcast", "shuffle", "shuffle_peer"):
            if ("join_bcast" if mode == "broadcast" else "join_shuffle") not in cases:
                continue
            if mode == "shuffle_peer" and peer is None:
                continue
            res = {}

            def run():
                res.clear()
                res["o"] = D.join_distributed(probe, build, "inner", None, "shuffle" if mode == "shuffle_peer" else mode,
                                              peer if mode == "shuffle_peer" else None)
            ms = timed(run)
            kw = {}
            if mode != "broadcast" and P > 1:
                out_bytes = (n + nb) * 16 * (P - 1) / P
                kw = {"shuffle_bytes_out_per_gpu": out_bytes, "nvlink_frac_upper": (out_bytes / (ms * 1e-3)) / 900e9}
            emit("join_inner %s (probe %d x build %d rows per GPU)" % (mode, n, nb), ms, n * P, n_out_rank0=int(res["o"][0].shape[0]), **kw)
            res.clear()
        del build, probe
    if "rolling" in cases:
        out = torch.empty_like(vals)
        ms = timed(lambda: D.rolling_distributed(vals, "mean", 100, 100, 1, None, rank * n, out))
        emit("rolling_mean w=100 (row-sharded, win-1 halo rows from the previous rank)", ms, n * P)
    if P > 1:
        peer.close()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
