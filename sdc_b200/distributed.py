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


class _ArenaView:
    """int64[n] window of a receive arena exposed through __cuda_array_interface__ (zero-copy into torch)."""

    def __init__(self, ptr, n, owner):
        self.owner = owner
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2, "strides": None}


class PeerShuffle:
    """Hash shuffle over NVLink peer memory (one node, one process per GPU, P <= 8).

    Every rank owns `narenas` receive arenas of `max_cols` columns x `capacity_rows` 8-byte rows (cudaMalloc,
    exported with CUDA IPC handles, opened by the peers once, here).  shuffle() then costs: one counting pass over
    the keys (`sdcb200_shuffle_plan`), ONE small all-gather of the per-destination counts, ONE fused partition-and-
    send kernel (`sdcb200_shuffle_send`: rows ranked per destination inside 4096-row tiles, staged in shared memory,
    stored into the peers' arenas over NVLink) and a barrier -- instead of a radix partition, a gather kernel and an
    NCCL all-to-all per column.  The returned tensors VIEW the arena used by this call; arenas rotate, so
    they stay valid for the next `narenas - 1` calls."""

    def __init__(self, capacity_rows, max_cols=4, narenas=2, group=None):
        torch, dist = _torch(), _dist()
        self.group = group
        self.rank, self.P = world(group)
        if self.P > 8:
            raise ValueError("PeerShuffle: at most 8 ranks (one NVSwitch node)")
        self.cap, self.max_cols, self.narenas = int(capacity_rows), int(max_cols), int(narenas)
        self.lib = native.load()
        self.dev = torch.device("cuda", torch.cuda.current_device())
        self.mine, self.peers, self.turn = [], [], 0
        nbytes = self.cap * self.max_cols * 8
        for _ in range(self.narenas):
            p = ct.c_void_p()
            native.check(self.lib.sdcb200_ipc_alloc(ct.byref(p), nbytes))
            h = (ct.c_uint8 * 64)()
            native.check(self.lib.sdcb200_ipc_get_handle(p, h))
            handles = [None] * self.P
            dist.all_gather_object(handles, bytes(h), group=group)
            bases = []
            for r, hb in enumerate(handles):
                if r == self.rank:
                    bases.append(p.value)
                else:
                    q = ct.c_void_p()
                    native.check(self.lib.sdcb200_ipc_open((ct.c_uint8 * 64).from_buffer_copy(hb), ct.byref(q)))
                    bases.append(q.value)
            self.mine.append(p.value)
            self.peers.append(bases)

    def close(self):
        torch = _torch()
        torch.cuda.synchronize()
        for a, bases in enumerate(self.peers):
            for r, b in enumerate(bases):
                if r != self.rank:
                    self.lib.sdcb200_ipc_close(b)
        _dist().barrier(group=self.group)          # nobody frees an arena a peer still maps
        for p in self.mine:
            self.lib.sdcb200_ipc_free(p)
        self.mine, self.peers = [], []

    def shuffle(self, keys, cols):
        """Route every row to rank hash(key) % P.  -> (keys', [cols'...]) views of this rank's arena."""
        torch, dist = _torch(), _dist()
        P, me = self.P, self.rank
        allc = [keys] + list(cols)
        if len(allc) > self.max_cols:
            raise ValueError("PeerShuffle: %d columns, arena holds %d" % (len(allc), self.max_cols))
        for c in allc:
            if c.element_size() != 8 or not c.is_contiguous():
                raise TypeError("PeerShuffle: contiguous 8-byte columns only")
        n = keys.shape[0]
        kname = str(keys.dtype).replace("torch.", "")
        if kname not in ("int64", "float64"):
            raise TypeError("PeerShuffle: key dtype %s" % kname)
        kcode = native.DTYPE_CODES[kname]
        counts = torch.empty(P, dtype=torch.int64, device=keys.device)
        tile_off = torch.empty(max(1, -(-n // 4096)) * 8, dtype=torch.int32, device=keys.device)
        with torch.cuda.device(keys.device):
            native.check(self.lib.sdcb200_shuffle_plan(kcode, keys.data_ptr(), n, P, tile_off.data_ptr(), counts.data_ptr(),
                                                       native.current_stream_ptr()))
        # counts[s][d] on every rank; entering the collective also orders this call after every rank's use of the
        # arena that is about to be overwritten (each rank's consumers run before it on the same stream)
        mat = torch.empty(P * P, dtype=torch.int64, device=keys.device)
        dist.all_gather_into_tensor(mat, counts, group=self.group)
        m = mat.cpu().view(P, P)
        recv_total = int(m[:, me].sum())
        if int(m.sum(0).max()) > self.cap:
            raise ArenaOverflow(int(m.sum(0).max()), self.cap)
        dst_off = [int(m[:me, d].sum()) for d in range(P)]       # my block in rank d's arena starts after lower ranks'
        a = self.turn
        self.turn = (self.turn + 1) % self.narenas
        cptr = (ct.c_void_p * len(allc))(*[c.data_ptr() for c in allc])
        offc = (ct.c_int64 * P)(*dst_off)
        basec = (ct.c_void_p * P)(*self.peers[a])
        with torch.cuda.device(keys.device):
            native.check(self.lib.sdcb200_shuffle_send(kcode, len(allc), cptr, n, P, tile_off.data_ptr(), offc, basec,
                                                       self.cap, native.current_stream_ptr()))
        # every rank's stores have landed once every rank's kernel is done: a collective on the same stream
        dist.barrier(group=self.group, device_ids=[self.dev.index])
        out = []
        for c, col in enumerate(allc):
            if recv_total == 0:
                out.append(torch.empty(0, dtype=col.dtype, device=col.device))
                continue
            v = torch.as_tensor(_ArenaView(self.mine[a] + c * self.cap * 8, recv_total, self), device=col.device)
            out.append(v.view(col.dtype))
        return out[0], out[1:]


_AUTO_PEERS = {}     # id(group) -> PeerShuffle | False (peer memory unavailable for this group)


class ArenaOverflow(native.NativeError):
    """Raised by PeerShuffle.shuffle on EVERY rank (the decision is taken on the all-gathered count matrix)."""

    def __init__(self, needed, cap):
        super().__init__("PeerShuffle: a rank would receive %d rows, arena holds %d" % (needed, cap))
        self.needed = needed


def auto_peer(group, n_rows, ncols, min_rows=0):
    """The cached PeerShuffle context of `group`, created (or re-created larger) collectively; None when the group is
    not one NCCL node of <= 8 GPUs or SDCB200_PEER_SHUFFLE=0.  Every decision depends only on values all ranks
    share (ncols, all-reduced sizes, the all-gathered count matrix), so all ranks take the same path."""
    import os
    import socket
    torch, dist = _torch(), _dist()
    key = id(group)
    ctx = _AUTO_PEERS.get(key)
    if ctx is False or os.environ.get("SDCB200_PEER_SHUFFLE", "1") == "0":
        return None
    _, P = world(group)
    if ctx is None:
        if P < 2 or P > 8 or dist.get_backend(group) != "nccl":
            _AUTO_PEERS[key] = False
            return None
        hosts = [None] * P
        dist.all_gather_object(hosts, socket.gethostname(), group=group)
        if len(set(hosts)) != 1:
            _AUTO_PEERS[key] = False
            return None
    if ctx is not None and ncols <= ctx.max_cols and min_rows <= ctx.cap:
        return ctx
    if ctx is None:
        need = torch.tensor([n_rows], dtype=torch.int64, device="cuda")
        dist.all_reduce(need, op=dist.ReduceOp.MAX, group=group)
        cap = max(int(need.item()) * 3 // 2 + 4096, min_rows)
        cols_cap = max(ncols, 2)
    else:
        cap = max(ctx.cap, min_rows * 5 // 4)
        cols_cap = max(ncols, ctx.max_cols)
        ctx.close()
    ctx = PeerShuffle(cap, max_cols=cols_cap, narenas=2, group=group)
    _AUTO_PEERS[key] = ctx
    return ctx


def hash_shuffle(keys, cols, group=None, peer=None):
    """Route every row to rank hash(key) % P.  -> (keys', [cols'...]) now owned by this rank.
    peer: a PeerShuffle context -> NVLink peer-memory stores instead of NCCL all-to-all per column."""
    from .sort import gather_device
    _, P = world(group)
    if P == 1:
        return keys, list(cols)
    if peer == "auto":
        peer = auto_peer(group, keys.shape[0], 1 + len(cols)) if keys.is_cuda else None
        if peer is not None:
            try:
                return peer.shuffle(keys, cols)
            except ArenaOverflow as e:           # skewed destinations: every rank is here, grow together and retry
                peer = auto_peer(group, keys.shape[0], 1 + len(cols), e.needed)
                return peer.shuffle(keys, cols)
    if peer is not None:
        return peer.shuffle(keys, cols)
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


def groupby_distributed(keys, cols, aggs, sort=True, ddof=1, expected_groups=0, group=None, mode="auto", peer="auto"):
    """df.groupby(key).<aggs>() over row-sharded columns.

    mode 'combine' (low cardinality): local hash-aggregate -> all-gather of the partial (key, sum, count,
      min, max) tables -> the same groupby kernel over the partials -> every rank holds the full result.
      This is the reference's per-chunk-dict-then-merge idea (hpat_pandas_dataframe_functions.py:3080-3102)
      with GPUs as the chunks.  sum / mean / min / max / count only.
    mode 'shuffle' (high cardinality, or var / std): rows go to rank hash(key) % P, each rank aggregates
      the keys it owns -> every rank holds a disjoint slice of the groups.  peer: 'auto' (default: the cached
      PeerShuffle context of the group -- NVLink peer-memory stores -- when the group is one node of <= 8 GPUs), a
      PeerShuffle context, or None for NCCL all-to-all.
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
        k2, c2 = hash_shuffle(keys, cols, group, peer)
        k, r = groupby_device(k2, c2, aggs, sort, ddof, expected_groups // P if expected_groups > 0 else 0)
        return k, r, "sharded"
    for a in aggs:
        if a not in _PARTIALS:
            raise ValueError("groupby_distributed: %r needs mode='shuffle'" % (a,))
    need = sorted({p for a in aggs for p in _PARTIALS[a]})
    # sort=False needs every partial table in local first-appearance order (the combine then sees the groups in
    # global first-appearance order, rank-major); sort=True does not care, and the sorted local path is cheaper
    lk, lres = groupby_device(keys, cols, need, bool(sort), ddof, expected_groups)
    gk = part = None
    cap = int(expected_groups)
    if 0 < cap <= (1 << 16) and lk.is_cuda:
        # ONE collective: every rank's partial table (keys + every partial of every column, padded to the hint, its
        # group count in the extra last column) travels in a single int64 matrix
        nrows = 1 + len(need) * len(cols)
        ng = lk.shape[0]
        buf = torch.zeros((nrows, cap + 1), dtype=torch.int64, device=lk.device)
        buf[0, cap] = ng
        if ng <= cap:
            buf[0, :ng] = lk.view(torch.int64)
            r = 1
            for p_ in need:
                for t in lres[p_]:
                    buf[r, :ng] = t.view(torch.int64)
                    r += 1
        allb = torch.empty((P, nrows, cap + 1), dtype=torch.int64, device=lk.device)
        _dist().all_gather_into_tensor(allb, buf, group=group)
        # the one host read of this path: every rank's group count, and whether all ranks hold the same key row
        same_keys = (allb[:, 0, :cap] == allb[0:1, 0, :cap]).all().to(torch.int64).reshape(1)
        host = torch.cat([allb[:, 0, cap], same_keys]).tolist()
        ngs, same = host[:P], bool(host[P])
        if same and all(g == ngs[0] for g in ngs) and 0 < ngs[0] <= cap and all(p_ in ("sum", "count") for p_ in need):
            # every rank found the same groups in the same order (the usual case at low cardinality): the partial
            # tables add up row by row -- no second aggregate, the result is already in the gathered matrix
            g0 = ngs[0]
            out_keys = allb[0, 0, :g0].view(lk.dtype)
            final, r = {}, 1
            for p_ in need:
                final[p_] = {}
                for i, t in enumerate(lres[p_]):
                    final[p_][i] = allb[:, r, :g0].view(t.dtype).sum(dim=0)
                    r += 1
            res = {}
            for a in aggs:
                if a == "mean":
                    res[a] = [combine_mean_device(final["sum"][i] if final["sum"][i].dtype == torch.float64
                                                  else final["sum"][i].to(torch.float64), final["count"][i]) for i in range(len(cols))]
                else:
                    res[a] = [final[a][i] for i in range(len(cols))]
            return out_keys, res, "replicated"
        if all(0 <= g <= cap for g in ngs):
            gk = torch.cat([allb[q, 0, :g] for q, g in enumerate(ngs)]).view(lk.dtype)
            part, r = {}, 1
            for p_ in need:
                part[p_] = []
                for t in lres[p_]:
                    part[p_].append(torch.cat([allb[q, r, :g] for q, g in enumerate(ngs)]).view(t.dtype))
                    r += 1
    if gk is None:            # no usable hint (or it was too small on some rank -- every rank sees that): ragged gathers
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


def join_distributed(left_keys, right_keys, how="inner", group=None, mode="auto", peer="auto"):
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
    lk2, (lid2,) = hash_shuffle(left_keys, [lid], group, peer)     # peer: two arenas, both results stay valid
    rk2, (rid2,) = hash_shuffle(right_keys, [rid], group, peer)
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
