"""Multi-GPU pass: one process per GPU, NCCL over NVLink / NVSwitch through torch.distributed.

Replaces the reference's (dead) distributed layer for the hot path:
  * shuffle protocol ........ sdc/shuffle_utils.py:66-298: per-destination counts -> alltoall of the counts
    (:158) -> displacements (:161-162) -> one alltoallv per column (:268-270)
  * collectives ............. sdc/distributed_api.py:94-125, 439-466 over the single-process memcpy stub
    sdc/transport/hpat_transport_single_process.cpp:106-119 (rank 0 / size 1, :231-239)
  * block partition ......... sdc/_distributed.h:77-102
What runs where: destinations, counts and the stable grouping of rows by destination come from
`sdcb200_hash_partition` (csrc/shuffle.cu), packing from `sdcb200_gather`; the exchange is
`all_to_all_single` with int64 split sizes (the reference's counts are int32, distributed_api.py:122);
the local work is the same groupby / join / rolling kernels.  The functions that only move bytes
(`exchange_counts`, `exchange_rows`, `allgather_ragged`, `halo_exchange`) accept CPU tensors too, which is
how tests/test_distributed_cpu.py covers the N > 1 protocol under `gloo` without a GPU.
"""
import ctypes as ct

from . import native


def _torch():
    import torch
    return torch


def _dist():
    import torch.distributed as dist
    return dist


def world(group=None):
    dist = _dist()
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


# ------------------------------------------------------------------ byte movers (device-agnostic)
def exchange_counts(send_counts, group=None):
    """send_counts int64[P] (rows this rank sends to each rank) -> recv_counts int64[P].  shuffle_utils.py:158."""
    torch, dist = _torch(), _dist()
    recv = torch.empty_like(send_counts)
    dist.all_to_all_single(recv, send_counts, group=group)
    return recv


def exchange_rows(packed, send_counts, recv_counts, group=None):
    """`packed` holds this rank's rows grouped by destination (send_counts rows each); returns the rows
    received, grouped by source rank in rank order.  One alltoallv per column (shuffle_utils.py:268-270)."""
    torch, dist = _torch(), _dist()
    sc = [int(x) for x in send_counts.tolist()]
    rc = [int(x) for x in recv_counts.tolist()]
    out = torch.empty((sum(rc),) + tuple(packed.shape[1:]), dtype=packed.dtype, device=packed.device)
    dist.all_to_all_single(out, packed, rc, sc, group=group)
    return out


def allgather_ragged(t, group=None):
    """Concatenation of every rank's 1-D tensor in rank order, plus the per-rank lengths."""
    torch, dist = _torch(), _dist()
    _, P = world(group)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.empty_like(n) for _ in range(P)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes) if sizes else 0
    pad = torch.zeros(mx, dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    parts = [torch.empty_like(pad) for _ in range(P)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)]), sizes


def halo_exchange(col, halo_rows, group=None):
    """Row-sharded series: send my last `halo_rows` rows to rank + 1, receive rank - 1's (None on rank 0).
    The analogue of the per-chunk prelude of the reference's rolling loop
    (sdc/datatypes/hpat_pandas_series_rolling_functions.py:609-617)."""
    torch, dist = _torch(), _dist()
    rank, P = world(group)
    if P == 1 or halo_rows <= 0:
        return None
    # a shard shorter than the halo forwards what it has; windows then reach at most one shard back,
    # which is exact whenever every shard holds >= halo_rows rows (checked by the caller)
    send = col[max(0, col.shape[0] - halo_rows):].contiguous()
    recv = torch.empty(halo_rows, dtype=col.dtype, device=col.device) if rank > 0 else None
    ops = []
    if rank + 1 < P:
        ops.append(dist.P2POp(dist.isend, send, rank + 1, group))
    if rank > 0:
        ops.append(dist.P2POp(dist.irecv, recv, rank - 1, group))
    for w in dist.batch_isend_irecv(ops) if ops else []:
        w.wait()
    return recv


def block_range(n_total, rank, P):
    """Rows [start, end) of a 1-D block partition: ceil(n / P) rows per rank (sdc/_distributed.h:77-102)."""
    per = -(-n_total // P)
    start = min(n_total, rank * per)
    return start, min(n_total, start + per)


# ------------------------------------------------------------------ kernels + collectives
def hash_partition_device(keys, nparts):
    """-> (perm int64[n]: rows grouped by destination, stable; counts int64[nparts])."""
    torch = _torch()
    lib = native.load()
    name = str(keys.dtype).replace("torch.", "")
    perm = torch.empty(keys.shape[0], dtype=torch.int64, device=keys.device)
    counts = torch.empty(nparts, dtype=torch.int64, device=keys.device)
    with torch.cuda.device(keys.device):
        native.check(lib.sdcb200_hash_partition(native.DTYPE_CODES[name], keys.data_ptr(), keys.shape[0], nparts,
                                                perm.data_ptr(), counts.data_ptr(), native.current_stream_ptr()))
    return perm, counts


def hash_shuffle(keys, cols, group=None):
    """Route every row to rank hash(key) % P.  -> (keys', [cols'...]) now owned by this rank."""
    from .sort import gather_device
    _, P = world(group)
    if P == 1:
        return keys, list(cols)
    perm, counts = hash_partition_device(keys, P)
    recv_counts = exchange_counts(counts, group)
    out = []
    for c in [keys] + list(cols):
        out.append(exchange_rows(gather_device(c, perm), counts, recv_counts, group))
    return out[0], out[1:]


def combine_mean_device(sums, counts):
    torch = _torch()
    lib = native.load()
    out = torch.empty(sums.shape[0], dtype=torch.float64, device=sums.device)
    with torch.cuda.device(sums.device):
        native.check(lib.sdcb200_combine_mean(sums.data_ptr(), counts.data_ptr(), sums.shape[0], out.data_ptr(),
                                              native.current_stream_ptr()))
    return out


_PARTIALS = {"sum": ("sum",), "count": ("count",), "min": ("min",), "max": ("max",), "mean": ("sum", "count")}


def groupby_distributed(keys, cols, aggs, sort=True, ddof=1, expected_groups=0, group=None, mode="auto"):
    """df.groupby(key).<aggs>() over row-sharded columns.

    mode 'combine' (low cardinality): local hash-aggregate -> all-gather of the partial (key, sum, count,
      min, max) tables -> the same groupby kernel over the partials -> every rank holds the full result.
      This is the reference's per-chunk-dict-then-merge idea (hpat_pandas_dataframe_functions.py:3080-3102)
      with GPUs as the chunks.  sum / mean / min / max / count only.
    mode 'shuffle' (high cardinality, or var / std): rows go to rank hash(key) % P, each rank aggregates
      the keys it owns -> every rank holds a disjoint slice of the groups.
    -> (group_keys, {agg: [tensor per column]}, 'replicated' | 'sharded')"""
    from .groupby import groupby_device
    torch = _torch()
    aggs = tuple(aggs)
    _, P = world(group)
    if P == 1:
        k, r = groupby_device(keys, cols, aggs, sort, ddof, expected_groups)
        return k, r, "replicated"
    if mode == "auto":
        simple = all(a in _PARTIALS for a in aggs)
        mode = "combine" if simple and 0 < expected_groups <= (1 << 20) else "shuffle"
    if mode == "shuffle":
        k2, c2 = hash_shuffle(keys, cols, group)
        k, r = groupby_device(k2, c2, aggs, sort, ddof, expected_groups // P if expected_groups > 0 else 0)
        return k, r, "sharded"
    for a in aggs:
        if a not in _PARTIALS:
            raise ValueError("groupby_distributed: %r needs mode='shuffle'" % (a,))
    need = sorted({p for a in aggs for p in _PARTIALS[a]})
    lk, lres = groupby_device(keys, cols, need, False, ddof, expected_groups)
    gk, _ = allgather_ragged(lk, group)
    part = {p: [allgather_ragged(t, group)[0] for t in lres[p]] for p in need}
    final = {}
    # sums of sums and of counts, minima of minima, maxima of maxima: one pass of the same kernel each
    summed = [t for p in ("sum", "count") if p in part for t in part[p]]
    names = [(p, i) for p in ("sum", "count") if p in part for i in range(len(cols))]
    hint = max(int(expected_groups), 1)
    out_keys = None
    if summed:
        out_keys, r = groupby_device(gk, summed, ("sum",), sort, ddof, hint)
        for (p, i), t in zip(names, r["sum"]):
            final.setdefault(p, {})[i] = t
    for p in ("min", "max"):
        if p in part:
            out_keys, r = groupby_device(gk, part[p], (p,), sort, ddof, hint)
            final[p] = dict(enumerate(r[p]))
    if out_keys is None:      # no value columns
        out_keys, _ = groupby_device(gk, [], ("count",), sort, ddof, hint)
    res = {}
    for a in aggs:
        if a == "mean":
            res[a] = [combine_mean_device(final["sum"][i] if final["sum"][i].dtype == torch.float64
                                          else final["sum"][i].to(torch.float64), final["count"][i]) for i in range(len(cols))]
        else:
            res[a] = [final[a][i] for i in range(len(cols))]
    return out_keys, res, "replicated"


def join_distributed(left_keys, right_keys, how="inner", group=None, mode="auto"):
    """pd.merge indexers over row-sharded key columns.  Row numbers in the result are GLOBAL
    (rank-major: global = rows of lower ranks + local row), the result itself stays sharded.

    mode 'broadcast': the build (right) side is all-gathered and every rank joins its own probe rows
      against the whole build side -- cheaper than shuffling when build << probe (C3, C5).
    mode 'shuffle': both sides are hash-partitioned by key with their global row numbers as payload."""
    from .join import join_device
    from .sort import gather_device
    torch = _torch()
    rank, P = world(group)
    if P == 1:
        return join_device(left_keys, right_keys, how)
    dev = left_keys.device
    nl_all = allgather_ragged(torch.tensor([left_keys.shape[0]], dtype=torch.int64, device=dev), group)[0]
    loff = int(nl_all[:rank].sum().item())
    if mode == "auto":
        nr_all = allgather_ragged(torch.tensor([right_keys.shape[0]], dtype=torch.int64, device=dev), group)[0]
        mode = "broadcast" if int(nr_all.sum().item()) * 4 <= int(nl_all.sum().item()) else "shuffle"
    if mode == "broadcast":
        rk_all, _ = allgather_ragged(right_keys, group)          # position in the concatenation = global build row
        li, ri = join_device(left_keys, rk_all, how)
        return li + loff, ri
    nr_all = allgather_ragged(torch.tensor([right_keys.shape[0]], dtype=torch.int64, device=dev), group)[0]
    roff = int(nr_all[:rank].sum().item())
    lid = torch.arange(loff, loff + left_keys.shape[0], dtype=torch.int64, device=dev)
    rid = torch.arange(roff, roff + right_keys.shape[0], dtype=torch.int64, device=dev)
    lk2, (lid2,) = hash_shuffle(left_keys, [lid], group)
    rk2, (rid2,) = hash_shuffle(right_keys, [rid], group)
    li, ri = join_device(lk2, rk2, how)
    gl = gather_device(lid2, li)
    if how == "left":
        miss = ri < 0
        gr = gather_device(rid2, torch.where(miss, torch.zeros_like(ri), ri)) if rid2.shape[0] else torch.full_like(ri, -1)
        gr = torch.where(miss, torch.full_like(gr, -1), gr)
    else:
        gr = gather_device(rid2, ri)
    return gl, gr


def rolling_distributed(col, op, window, minp, ddof=1, group=None, row_offset=None, out=None):
    """Series.rolling over a row-sharded column: win-1 halo rows from the previous rank, then the local kernel."""
    from .rolling import rolling_device
    torch = _torch()
    rank, P = world(group)
    halo = halo_exchange(col, window - 1, group) if window > 1 else None
    if row_offset is None:
        n_all = allgather_ragged(torch.tensor([col.shape[0]], dtype=torch.int64, device=col.device), group)[0] if P > 1 else None
        row_offset = int(n_all[:rank].sum().item()) if P > 1 else 0
    return rolling_device(col, op, window, minp, ddof, out=out, halo=halo, row_offset=row_offset)
