"""Multi-GPU pass: one process per GPU, NCCL over NVLink / NVSwitch through torch.distributed.

Replaces the reference's (dead) distributed layer for the hot path:
  * shuffle protocol ........ sdc/shuffle_utils.py:66-298: per-destination counts -> alltoall of the counts
    (:158) -> displacements (:161-162) -> one alltoallv per column (:268-270)
  * collectives ............. sdc/distributed_api.py:94-125, 439-466 over the single-process memcpy stub
    sdc/transport/hpat_transport_single_process.cpp:106-119 (rank 0 / size 1, :231-239)
  * block partition ......... sdc/_distributed.h:77-102
What runs where: destinations, counts and the stable grouping of rows by destination come from
`sdcb200_hash_partition` / `sdcb200_shuffle_plan` (csrc/shuffle.cu); the exchange is either the reference's
protocol over NCCL (`all_to_all_single` with int64 split sizes; the reference's counts are int32,
distributed_api.py:122) or ONE partition-and-send kernel storing rows into the peers' receive arenas over NVLink
(`PeerShuffle`); the local work is the same groupby / join / rolling kernels.  Row-proportional work is done by the
library's kernels only (row ids are generated inside the send kernel, local -> global row numbers by
`sdcb200_gather_ids` / `sdcb200_affine_i64`); torch is plumbing (allocation, collectives on small tensors).
The functions that only move bytes (`exchange_counts`, `exchange_rows`, `allgather_ragged`, `halo_exchange`) accept
CPU tensors too, which is how tests/test_distributed_cpu.py covers the N > 1 protocol under `gloo` without a GPU.
"""
import ctypes as ct
import weakref

from . import native

ROUTE_HASH, ROUTE_MOD, ROUTE_RANGE = native.ROUTE_HASH, native.ROUTE_MOD, native.ROUTE_RANGE


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


def _global_rank(group, group_rank):
    """P2P ops address peers by GLOBAL rank; `group_rank` is a rank inside `group`."""
    dist = _dist()
    if group is None:
        return group_rank
    return dist.get_global_rank(group, group_rank)


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


def allgather_sizes(n, device, group=None):
    """Every rank's value of the host integer `n`, as a list (one small all-gather)."""
    torch, dist = _torch(), _dist()
    _, P = world(group)
    t = torch.tensor([int(n)], dtype=torch.int64, device=device)
    out = torch.empty(P, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, t, group=group)
    return [int(x) for x in out.tolist()]


def halo_exchange(col, halo_rows, group=None):
    """Row-sharded series: the `halo_rows` rows that precede this rank's shard in the global series (fewer at the
    head of the series; None on rank 0 / when nothing precedes).  The analogue of the per-chunk prelude of the
    reference's rolling loop (sdc/datatypes/hpat_pandas_series_rolling_functions.py:609-617).

    Every rank contributes its last min(len, halo_rows) rows to ONE ragged all-gather and takes the tails of as many
    previous ranks as it needs, so shards shorter than the halo (block_range leaves short or empty tail shards,
    sdc/_distributed.h:77-102) are exact: a window may reach several shards back."""
    torch = _torch()
    rank, P = world(group)
    if P == 1 or halo_rows <= 0:
        return None
    tail = col[max(0, col.shape[0] - halo_rows):].contiguous()
    cat, sizes = allgather_ragged(tail, group)
    if rank == 0:
        return None
    offs = [0]
    for s in sizes:
        offs.append(offs[-1] + s)
    need, parts = halo_rows, []
    for r in range(rank - 1, -1, -1):             # nearest shard first; each tail is already at most halo_rows long
        if need <= 0:
            break
        take = min(need, sizes[r])
        if take:
            parts.append(cat[offs[r + 1] - take:offs[r + 1]])
            need -= take
    if not parts:
        return None
    parts.reverse()
    return torch.cat(parts) if len(parts) > 1 else parts[0].contiguous()


def block_range(n_total, rank, P):
    """Rows [start, end) of a 1-D block partition: ceil(n / P) rows per rank (sdc/_distributed.h:77-102)."""
    per = -(-n_total // P)
    start = min(n_total, rank * per)
    return start, min(n_total, start + per)


# ------------------------------------------------------------------ small kernels of the pass
def affine_i64_(t, mul, add, keep_negative=False):
    """t[i] = t[i] * mul + add in place (library kernel); keep_negative leaves the -1 markers alone."""
    torch = _torch()
    if t.numel() == 0:
        return t
    assert t.is_cuda and t.dtype == torch.int64 and t.is_contiguous()
    lib = native.load()
    with torch.cuda.device(t.device):
        native.check(lib.sdcb200_affine_i64(t.data_ptr(), t.numel(), int(mul), int(add), 1 if keep_negative else 0,
                                            native.current_stream_ptr()))
    return t


def gather_ids_device(src, idx):
    """out[i] = idx[i] < 0 ? -1 : src[idx[i]] (library kernel): local join rows -> the ids that travelled with them."""
    torch = _torch()
    out = torch.empty(idx.shape[0], dtype=torch.int64, device=idx.device)
    if idx.numel() == 0:
        return out
    if src.numel() == 0:
        return out.fill_(-1)
    lib = native.load()
    with torch.cuda.device(idx.device):
        native.check(lib.sdcb200_gather_ids(src.data_ptr(), idx.data_ptr(), idx.numel(), out.data_ptr(),
                                            native.current_stream_ptr()))
    return out


def hash_partition_device(keys, nparts, route=ROUTE_HASH, want_shipped=False):
    """-> (perm int64[n]: rows grouped by destination, stable; counts int64[nparts][, the key column as it travels])."""
    torch = _torch()
    lib = native.load()
    name = str(keys.dtype).replace("torch.", "")
    perm = torch.empty(keys.shape[0], dtype=torch.int64, device=keys.device)
    counts = torch.empty(nparts, dtype=torch.int64, device=keys.device)
    shipped = torch.empty_like(keys) if want_shipped else None
    with torch.cuda.device(keys.device):
        native.check(lib.sdcb200_hash_partition(native.DTYPE_CODES[name], int(route), keys.data_ptr(), keys.shape[0], nparts,
                                                perm.data_ptr(), counts.data_ptr(),
                                                shipped.data_ptr() if want_shipped else None, native.current_stream_ptr()))
    return (perm, counts, shipped) if want_shipped else (perm, counts)


class _ArenaView:
    """int64[n] window of a receive arena exposed through __cuda_array_interface__ (zero-copy into torch).  The
    tensor made from it keeps this object alive; the context counts its live views (weakly) and never frees an
    arena while one exists."""

    def __init__(self, ptr, n, owner):
        self.owner = owner
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2, "strides": None}
        owner._views.add(self)


class ArenaOverflow(native.NativeError):
    """A rank would receive more rows than an arena holds.  Raised on EVERY rank (the decision is taken on the
    all-gathered count matrices) and BEFORE anything is sent, so the caller can take another transport."""

    def __init__(self, needed, cap):
        super().__init__("PeerShuffle: a rank would receive %d rows, arena holds %d" % (needed, cap))
        self.needed = needed


class _Plan:
    """Counting pass of one key column: per-tile offsets per destination (device) + per-destination totals."""
    __slots__ = ("keys", "kcode", "route", "tile_off", "counts", "matrix", "range")


class PeerShuffle:
    """Hash shuffle over NVLink peer memory (one node, one process per GPU, P <= 8).

    Every rank owns `narenas` receive arenas of `max_cols` columns x `capacity_rows` 8-byte rows (cudaMalloc,
    exported with CUDA IPC handles, opened by the peers once, here).  A shuffle then costs: one counting pass over
    the keys (`plan`), ONE small all-gather of the per-destination counts of every column that is about to travel
    (`exchange`: a join plans both sides first, so capacity is checked for both before anything moves), ONE fused
    partition-and-send kernel per key column (`send`: rows ranked per destination inside 4096-row tiles, staged in
    shared memory, stored into the peers' arenas over NVLink) and a barrier -- instead of a radix partition, a gather
    kernel and an NCCL all-to-all per column.  The returned tensors VIEW the arena used by the call; arenas rotate,
    so they stay valid for the next `narenas - 1` sends.  close() frees nothing while a view is alive."""

    def __init__(self, capacity_rows, max_cols=4, narenas=2, group=None):
        torch, dist = _torch(), _dist()
        self.group = group
        self.rank, self.P = world(group)
        if self.P > 8:
            raise ValueError("PeerShuffle: at most 8 ranks (one NVSwitch node)")
        self.cap, self.max_cols, self.narenas = (int(capacity_rows) + 1) & ~1, int(max_cols), int(narenas)   # even: bulk stores
        self.lib = native.load()
        self.dev = torch.device("cuda", torch.cuda.current_device())
        self.mine, self.peers, self.turn = [], [], 0
        self._tok = None
        self._views = weakref.WeakSet()
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

    def live_views(self):
        return len(self._views)

    def close(self):
        """Collective.  Unmaps the peers' arenas and frees this rank's -- unless a tensor returned by an earlier
        send still views an arena ON ANY RANK (the count is all-reduced), in which case nothing is freed and False
        is returned: freeing would turn those tensors into dangling pointers."""
        torch, dist = _torch(), _dist()
        if not self.mine:
            return True
        torch.cuda.synchronize()
        live = torch.tensor([self.live_views()], dtype=torch.int64, device=self.dev)
        dist.all_reduce(live, op=dist.ReduceOp.MAX, group=self.group)
        if int(live.item()) > 0:
            return False
        for bases in self.peers:
            for r, b in enumerate(bases):
                if r != self.rank:
                    self.lib.sdcb200_ipc_close(b)
        dist.barrier(group=self.group)          # nobody frees an arena a peer still maps
        for p in self.mine:
            self.lib.sdcb200_ipc_free(p)
        self.mine, self.peers = [], []
        return True

    # -- the three steps
    def plan(self, keys, route=ROUTE_HASH, rng=None):
        torch = _torch()
        kname = str(keys.dtype).replace("torch.", "")
        if kname not in ("int64", "float64"):
            raise TypeError("PeerShuffle: key dtype %s" % kname)
        if keys.element_size() != 8 or not keys.is_contiguous():
            raise TypeError("PeerShuffle: contiguous 8-byte columns only")
        if route == ROUTE_MOD and kname != "int64":
            raise TypeError("PeerShuffle: ROUTE_MOD needs int64 keys")
        n = keys.shape[0]
        pl = _Plan()
        pl.keys, pl.kcode, pl.route, pl.matrix = keys, native.DTYPE_CODES[kname], int(route), None
        pl.counts = torch.empty(self.P, dtype=torch.int64, device=keys.device)
        pl.tile_off = torch.empty(max(1, -(-n // 4096)) * 8, dtype=torch.int32, device=keys.device)
        pl.range = rng                      # ROUTE_RANGE: (splitters uint64[P - 1] (host), descending, nan_dest)
        with torch.cuda.device(keys.device):
            if route == ROUTE_RANGE:
                sp = (ct.c_uint64 * max(1, self.P - 1))(*[int(x) for x in rng[0]])
                native.check(self.lib.sdcb200_shuffle_plan_range(pl.kcode, sp, int(rng[1]), int(rng[2]), keys.data_ptr(), n, self.P,
                                                                 pl.tile_off.data_ptr(), pl.counts.data_ptr(),
                                                                 native.current_stream_ptr()))
            else:
                native.check(self.lib.sdcb200_shuffle_plan(pl.kcode, pl.route, keys.data_ptr(), n, self.P, pl.tile_off.data_ptr(),
                                                           pl.counts.data_ptr(), native.current_stream_ptr()))
        return pl

    def exchange(self, plans):
        """ONE all-gather for the count vectors of all `plans`: plan.matrix[s][d] = rows rank s sends to rank d, on
        every rank (host).  Entering the collective also orders this call after every rank's use of the arena that is
        about to be overwritten (each rank's consumers run before it on the same stream)."""
        torch, dist = _torch(), _dist()
        P, k = self.P, len(plans)
        mine = torch.cat([pl.counts for pl in plans]) if k > 1 else plans[0].counts
        mat = torch.empty(P * k * P, dtype=torch.int64, device=mine.device)
        dist.all_gather_into_tensor(mat, mine, group=self.group)
        m = mat.cpu().view(P, k, P)
        for i, pl in enumerate(plans):
            pl.matrix = m[:, i, :].contiguous()
        return plans

    def max_recv(self, plan):
        return int(plan.matrix.sum(0).max())

    def send(self, plan, cols, iota_col=-1, iota_base=0):
        """Partition-and-send of `plan.keys` + `cols` (payload columns; cols[iota_col - 1] may be None: that column is
        generated as iota_base + row).  -> (keys', [cols'...]) views of this rank's arena."""
        torch, dist = _torch(), _dist()
        P, me = self.P, self.rank
        allc = [plan.keys] + list(cols)
        if len(allc) > self.max_cols:
            raise ValueError("PeerShuffle: %d columns, arena holds %d" % (len(allc), self.max_cols))
        for i, c in enumerate(allc):
            if i == iota_col:
                continue
            if c.element_size() != 8 or not c.is_contiguous() or c.shape[0] != plan.keys.shape[0]:
                raise TypeError("PeerShuffle: contiguous 8-byte columns of the key column's length only")
        m = plan.matrix
        if self.max_recv(plan) > self.cap:
            raise ArenaOverflow(self.max_recv(plan), self.cap)
        n = plan.keys.shape[0]
        recv_total = int(m[:, me].sum())
        dst_off = [int(m[:me, d].sum()) for d in range(P)]       # my block in rank d's arena starts after lower ranks'
        a = self.turn
        self.turn = (self.turn + 1) % self.narenas
        cptr = (ct.c_void_p * len(allc))(*[None if i == iota_col else c.data_ptr() for i, c in enumerate(allc)])
        offc = (ct.c_int64 * P)(*dst_off)
        basec = (ct.c_void_p * P)(*self.peers[a])
        dev = plan.keys.device
        with torch.cuda.device(dev):
            if plan.route == ROUTE_RANGE:
                sp = (ct.c_uint64 * max(1, P - 1))(*[int(x) for x in plan.range[0]])
                native.check(self.lib.sdcb200_shuffle_send_range(plan.kcode, sp, int(plan.range[1]), int(plan.range[2]), len(allc), cptr,
                                                                 int(iota_col), int(iota_base), n, P, plan.tile_off.data_ptr(), offc,
                                                                 basec, self.cap, native.current_stream_ptr()))
            else:
                native.check(self.lib.sdcb200_shuffle_send(plan.kcode, plan.route, len(allc), cptr, int(iota_col), int(iota_base), n, P,
                                                           plan.tile_off.data_ptr(), offc, basec, self.cap,
                                                           native.current_stream_ptr()))
        # every rank's stores have landed once every rank's kernel is done: a collective on the same stream (a one-word
        # all-reduce, not dist.barrier -- the barrier also blocks the HOST until the stream has drained, and every
        # launch behind it then pays its latency in the open)
        if self._tok is None:
            self._tok = torch.zeros(1, dtype=torch.int32, device=self.dev)
        dist.all_reduce(self._tok, group=self.group)
        out = []
        for c, col in enumerate(allc):
            dt = torch.int64 if c == iota_col else col.dtype
            if recv_total == 0:
                out.append(torch.empty(0, dtype=dt, device=dev))
                continue
            v = torch.as_tensor(_ArenaView(self.mine[a] + c * self.cap * 8, recv_total, self), device=dev)
            out.append(v.view(dt))
        return out[0], out[1:]

    def shuffle(self, keys, cols, route=ROUTE_HASH):
        """plan + exchange + send of one key column and its payload columns."""
        pl = self.plan(keys, route)
        self.exchange([pl])
        return self.send(pl, cols)


_AUTO_PEERS = {}     # id(group) -> PeerShuffle | False (peer memory unavailable for this group)
_RETIRED = []        # contexts replaced by a larger one whose arenas were still viewed: freed once nobody looks


def auto_peer(group, n_rows, ncols, min_rows=0):
    """The cached PeerShuffle context of `group`, created (or re-created larger) collectively; None when the group is
    not one NCCL node of <= 8 GPUs or SDCB200_PEER_SHUFFLE=0.  Every decision depends only on values all ranks
    share (ncols, all-reduced sizes, all-gathered count matrices, all-reduced view counts), so all ranks take the
    same path.  A context that is too small is replaced; if tensors still view its arenas on some rank it is parked
    in a retired list (its memory stays valid) and freed by a later call once they are gone."""
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
    for old in list(_RETIRED):                   # collective: every rank holds the same list in the same order
        if old.group is group and old.close():
            _RETIRED.remove(old)
    if ctx is None:
        need = torch.tensor([n_rows], dtype=torch.int64, device="cuda")
        dist.all_reduce(need, op=dist.ReduceOp.MAX, group=group)
        cap = max(int(need.item()) * 3 // 2 + 4096, min_rows)
        cols_cap = max(ncols, 2)
    else:
        cap = max(ctx.cap, min_rows * 5 // 4)
        cols_cap = max(ncols, ctx.max_cols)
        if not ctx.close():
            _RETIRED.append(ctx)
    ctx = PeerShuffle(cap, max_cols=cols_cap, narenas=2, group=group)
    _AUTO_PEERS[key] = ctx
    return ctx


def _balanced(matrix, P):
    """Does the count matrix spread the rows evenly enough over the ranks to keep ROUTE_MOD?  (Keys that are all
    multiples of P, say, would all land on rank 0.)"""
    recv = matrix.sum(0)
    total = int(recv.sum())
    return int(recv.max()) <= total // P * 5 // 4 + 4096


def _nccl_shuffle(keys, cols, route, group, iota=None):
    """The reference's protocol, one to one: stable partition, counts all-to-all, one all-to-all(v) per column."""
    from .sort import gather_device
    torch = _torch()
    _, P = world(group)
    if route == ROUTE_MOD:
        perm, counts, shipped = hash_partition_device(keys, P, route, True)
    else:
        perm, counts = hash_partition_device(keys, P, route)
        shipped = keys
    recv_counts = exchange_counts(counts, group)
    out = [exchange_rows(gather_device(shipped, perm), counts, recv_counts, group)]
    for i, c in enumerate(cols):
        if iota is not None and i == iota[0]:
            packed = perm.clone()
            affine_i64_(packed, 1, iota[1])
        else:
            packed = gather_device(c, perm)
        out.append(exchange_rows(packed, counts, recv_counts, group))
    return out[0], out[1:]


def shuffle_many(sides, group=None, peer="auto", route="auto"):
    """Shuffle several row sets by key with ONE routing decision and one counts exchange (a join: both sides).

    sides : list of (keys, cols, iota) -- iota None, or (index into cols, base): that payload column is not read but
            generated as base + row inside the send kernel (global row ids).
    route : ROUTE_HASH, ROUTE_MOD, or 'auto': ROUTE_MOD for int64 keys unless the all-gathered counts show that
            key mod P leaves the ranks unbalanced (every rank sees the same matrix and decides alike).
    -> (route used, [(keys', [cols'...]) per side]).  With ROUTE_MOD keys' holds floor(key / P) (key = keys' * P + rank)."""
    torch = _torch()
    rank, P = world(group)
    int_keys = all(str(k.dtype) == "torch.int64" for k, _, _ in sides)
    if route == "auto":
        route_try = ROUTE_MOD if int_keys else ROUTE_HASH
    else:
        route_try = int(route)
        if route_try == ROUTE_MOD and not int_keys:
            raise TypeError("ROUTE_MOD needs int64 keys")
    if P == 1:
        return ROUTE_HASH, [(k, [torch.arange(io[1] or 0, (io[1] or 0) + k.shape[0], dtype=torch.int64, device=k.device)
                                 if io is not None and i == io[0] else c for i, c in enumerate(cols)]) for k, cols, io in sides]
    ctx = None
    if all(k.is_cuda for k, _, _ in sides):
        if peer == "auto":
            ctx = auto_peer(group, max(k.shape[0] for k, _, _ in sides), max(1 + len(c) for _, c, _ in sides))
        elif peer is not None:
            ctx = peer
    if ctx is None:
        if any(io is not None and io[1] is None for _, _, io in sides):
            fixed = []
            for k, cols, io in sides:
                if io is not None and io[1] is None:
                    io = (io[0], sum(allgather_sizes(k.shape[0], k.device, group)[:rank]))
                fixed.append((k, cols, io))
            sides = fixed
        if route == "auto" and route_try == ROUTE_MOD:
            # the balance test needs the counts of every rank: one partition pass + all-gather, only for the first side
            _, counts = hash_partition_device(sides[0][0], P, ROUTE_MOD)
            mat = torch.empty(P * P, dtype=torch.int64, device=counts.device)
            _dist().all_gather_into_tensor(mat, counts, group=group)
            if not _balanced(mat.cpu().view(P, P), P):
                route_try = ROUTE_HASH
        return route_try, [_nccl_shuffle(k, cols, route_try, group, io) for k, cols, io in sides]
    plans = ctx.exchange([ctx.plan(k, route_try) for k, _, _ in sides])
    if route == "auto" and route_try == ROUTE_MOD and not all(_balanced(pl.matrix, P) for pl in plans):
        route_try = ROUTE_HASH
        plans = ctx.exchange([ctx.plan(k, route_try) for k, _, _ in sides])
    need = max(ctx.max_recv(pl) for pl in plans)
    if need > ctx.cap:
        # skewed destinations.  Nothing has been sent yet: grow the cached context (collectively; arenas that are
        # still viewed are parked, not freed) -- or, for a caller-owned context, take the NCCL transport
        if peer == "auto":
            ctx = auto_peer(group, 0, max(1 + len(c) for _, c, _ in sides), need)
        else:
            return route_try, [_nccl_shuffle(k, cols, route_try, group, io) for k, cols, io in sides]
    out = []
    for pl, (k, cols, io) in zip(plans, sides):
        base = 0
        if io is not None:
            # base None: global row ids, rank-major -- the rows of lower ranks are the row sums of the count matrix
            base = int(pl.matrix[:rank].sum()) if io[1] is None else io[1]
        out.append(ctx.send(pl, cols, -1 if io is None else io[0] + 1, base))
    return route_try, out


def hash_shuffle(keys, cols, group=None, peer=None, route=ROUTE_HASH):
    """Route every row to its owner rank.  -> (keys', [cols'...]) now owned by this rank (route: see shuffle_many).
    peer: 'auto' | a PeerShuffle context (NVLink peer-memory stores) | None (NCCL all-to-all per column)."""
    _, P = world(group)
    if P == 1:
        return keys, list(cols)
    _, out = shuffle_many([(keys, list(cols), None)], group, peer, route)
    return out[0]


def combine_mean_device(sums, counts):
    torch = _torch()
    lib = native.load()
    out = torch.empty(sums.shape[0], dtype=torch.float64, device=sums.device)
    with torch.cuda.device(sums.device):
        native.check(lib.sdcb200_combine_mean(sums.data_ptr(), counts.data_ptr(), sums.shape[0], out.data_ptr(),
                                              native.current_stream_ptr()))
    return out


_PARTIALS = {"sum": ("sum",), "count": ("count",), "min": ("min",), "max": ("max",), "mean": ("sum", "count")}


def groupby_distributed(keys, cols, aggs, sort=True, ddof=1, expected_groups=0, group=None, mode="auto", peer="auto",
                        route="auto"):
    """df.groupby(key).<aggs>() over row-sharded columns.

    mode 'combine' (low cardinality): local hash-aggregate -> all-gather of the partial (key, sum, count,
      min, max) tables -> the same groupby kernel over the partials -> every rank holds the full result.
      This is the reference's per-chunk-dict-then-merge idea (hpat_pandas_dataframe_functions.py:3080-3102)
      with GPUs as the chunks.  sum / mean / min / max / count only.
    mode 'shuffle' (high cardinality, or var / std): rows go to their key's owner rank, each rank aggregates
      the keys it owns -> every rank holds a disjoint slice of the groups.  peer: 'auto' (default: the cached
      PeerShuffle context of the group -- NVLink peer-memory stores -- when the group is one node of <= 8 GPUs), a
      PeerShuffle context, or None for NCCL all-to-all.  route: 'auto' keeps dense int64 key ranges dense on every
      rank (ROUTE_MOD, see shuffle_many), so the local aggregate stays on its direct-addressed table.
    -> (group_keys, {agg: [tensor per column]}, 'replicated' | 'sharded')"""
    from .groupby import groupby_device
    torch = _torch()
    aggs = tuple(aggs)
    rank, P = world(group)
    if P == 1:
        k, r = groupby_device(keys, cols, aggs, sort, ddof, expected_groups)
        return k, r, "replicated"
    if mode == "auto":
        simple = all(a in _PARTIALS for a in aggs)
        mode = "combine" if simple and 0 < expected_groups <= (1 << 20) else "shuffle"
    if mode == "shuffle":
        used, out = shuffle_many([(keys, list(cols), None)], group, peer, route)
        k2, c2 = out[0]
        k, r = groupby_device(k2, c2, aggs, sort, ddof, expected_groups // P if expected_groups > 0 else 0)
        if used == ROUTE_MOD:
            affine_i64_(k, P, rank)               # floor(key / P) back to the key; ascending order is preserved
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
        lib = native.load()
        rows = [lk] + [t for p_ in need for t in lres[p_]]
        ops = [0]
        for p_ in need:
            for t in lres[p_]:
                f = t.dtype == torch.float64
                ops.append({"sum": 1 if f else 2, "count": 2, "min": 3 if f else 5, "max": 4 if f else 6}[p_])
        buf = torch.empty((nrows, cap + 1), dtype=torch.int64, device=lk.device)
        allb = torch.empty((P, nrows, cap + 1), dtype=torch.int64, device=lk.device)
        comb = torch.empty((nrows, cap + 1), dtype=torch.int64, device=lk.device)
        flag = torch.empty(2, dtype=torch.int64, device=lk.device)
        with torch.cuda.device(lk.device):
            ptrs = (ct.c_void_p * nrows)(*[t.data_ptr() if t.numel() else None for t in rows])
            native.check(lib.sdcb200_pack_partials(ptrs, nrows, ng, cap, buf.data_ptr(), native.current_stream_ptr()))
            _dist().all_gather_into_tensor(allb, buf, group=group)
            # ONE launch reduces the gathered tables over the ranks and tells whether that was legal: every rank found
            # the same groups in the same order (the usual case at low cardinality) -- then the partial tables add up
            # row by row and there is no second aggregate (P x groups values: table-sized, not row-sized, work)
            native.check(lib.sdcb200_combine_partials(allb.data_ptr(), P, nrows, (ct.c_int32 * nrows)(*ops), cap, comb.data_ptr(),
                                                      flag.data_ptr(), native.current_stream_ptr()))
        differ, g0 = flag.tolist()                 # the one host read of this path
        if not differ and 0 < g0 <= cap:
            out_keys = comb[0, :g0].view(lk.dtype)
            final, r = {}, 1
            for p_ in need:
                final[p_] = {}
                for i, t in enumerate(lres[p_]):
                    final[p_][i] = comb[r, :g0].view(t.dtype)
                    r += 1
            res = {}
            for a in aggs:
                if a == "mean":
                    res[a] = [combine_mean_device(final["sum"][i] if final["sum"][i].dtype == torch.float64
                                                  else final["sum"][i].to(torch.float64), final["count"][i]) for i in range(len(cols))]
                else:
                    res[a] = [final[a][i] for i in range(len(cols))]
            return out_keys, res, "replicated"
        ngs = allb[:, 0, cap].tolist()
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


def join_distributed(left_keys, right_keys, how="inner", group=None, mode="auto", peer="auto", route="auto"):
    """pd.merge indexers over row-sharded key columns.  Row numbers in the result are GLOBAL
    (rank-major: global = rows of lower ranks + local row), the result itself stays sharded.

    mode 'broadcast': the build (right) side is all-gathered and every rank joins its own probe rows
      against the whole build side -- cheaper than shuffling when build << probe and the table stays cache-sized.
    mode 'shuffle': both sides go to their keys' owner ranks with their global row numbers as payload -- the row
      ids are generated inside the send kernel, both sides are planned first and share ONE counts exchange and ONE
      routing decision.  With int64 keys the routing keeps dense key ranges dense (ROUTE_MOD), so the local join
      stays on its direct-addressed table; the local row numbers it returns are mapped to the global ids that
      travelled with the rows by `sdcb200_gather_ids`."""
    from .join import join_device
    torch = _torch()
    rank, P = world(group)
    if P == 1:
        return join_device(left_keys, right_keys, how)
    dev = left_keys.device
    if mode == "shuffle":
        # the global row ids' bases come out of the shuffle's own count matrices: no size exchange of its own
        _, out = shuffle_many([(left_keys, [None], (0, None)), (right_keys, [None], (0, None))], group, peer, route)
        (lk2, (lid2,)), (rk2, (rid2,)) = out
        return join_device(lk2, rk2, how, left_ids=lid2, right_ids=rid2)
    sizes = allgather_sizes(left_keys.shape[0] * (1 << 32) + right_keys.shape[0], dev, group)   # both lengths in one word
    nl_all, nr_all = [s >> 32 for s in sizes], [s & 0xffffffff for s in sizes]
    loff, roff = sum(nl_all[:rank]), sum(nr_all[:rank])
    if mode == "auto":
        # all-gathering the build side makes every rank's table P times larger: worth it only while that table is small
        # next to L2 (4 bytes per key of a dense build side) or the probe side is tiny
        tot_r, tot_l = sum(nr_all), sum(nl_all)
        mode = "broadcast" if (tot_r * 4 <= tot_l and tot_r <= (4 << 20) and how in ("inner", "left")) else "shuffle"
    if mode == "broadcast":
        if how in ("right", "outer"):
            raise ValueError("join_distributed: how=%r needs mode='shuffle'" % (how,))
        rk_all, _ = allgather_ragged(right_keys, group)          # position in the concatenation = global build row
        li, ri = join_device(left_keys, rk_all, how)
        return affine_i64_(li, 1, loff, keep_negative=True), ri
    _, out = shuffle_many([(left_keys, [None], (0, loff)), (right_keys, [None], (0, roff))], group, peer, route)
    (lk2, (lid2,)), (rk2, (rid2,)) = out
    # the global row ids that travelled with the rows replace the local row numbers inside the join's last kernel
    return join_device(lk2, rk2, how, left_ids=lid2, right_ids=rid2)


# ------------------------------------------------------------------ sort_values across GPUs (sample sort)
def ordered_image(values):
    """Order-preserving uint64 image of an int64 / float64 NumPy array (the library's canonical sort key: int64 ^ 2^63;
    float64 with -0.0 as +0.0, the sign bit set for non-negative values and every bit flipped for negative ones).
    NaN has no image here -- callers drop NaNs first."""
    import numpy as np
    v = np.ascontiguousarray(values)
    if v.dtype == np.int64:
        return v.view(np.uint64) ^ np.uint64(1 << 63)
    if v.dtype != np.float64:
        raise TypeError("ordered_image: int64 / float64 only")
    b = (v + 0.0).view(np.uint64)                    # -0.0 + 0.0 = +0.0
    neg = (b >> np.uint64(63)) != 0
    return np.where(neg, ~b, b | np.uint64(1 << 63))


def sort_values_distributed(keys, ascending=True, na_position="last", group=None, payload=(), samples_per_rank=512):
    """Series.sort_values over a row-sharded column (sample sort).  Every rank ends with one contiguous slice
    of the global result: concatenating the ranks' outputs in rank order gives exactly the single-GPU stable
    sort_values (sdc/datatypes/hpat_pandas_series_functions.py:3853-3959; NaN last in both directions unless
    na_position='first', ties in original row order).

      1. every rank contributes a strided sample of its non-NaN keys to ONE all-gather; the P - 1 splitters are picked
         from the sorted samples (host work on P x samples_per_rank values);
      2. rows travel to the rank that owns their key range through the peer-memory shuffle with range routing
         (sdcb200_shuffle_plan_range / _send_range), carrying their global row number (generated in the send kernel) and
         the payload columns; rows of one source arrive in source order and sources in rank order, so equal keys sit in
         global row order on arrival;
      3. a local STABLE argsort orders the slice (the library's onesweep), NaN block first or last.
    keys: int64 / float64 CUDA tensor; payload: 8-byte CUDA columns of the same length.
    -> (sorted keys, global row numbers of the sorted rows, [payload columns in sorted order])"""
    import numpy as np
    from .sort import gather_device, sort_values_indexer_device
    torch, dist = _torch(), _dist()
    rank, P = world(group)
    if na_position not in ("last", "first"):
        raise ValueError("Method sort_values(). Unsupported parameter. Given na_position != 'last', 'first'")
    kname = str(keys.dtype).replace("torch.", "")
    if kname not in ("int64", "float64"):
        raise TypeError("sort_values_distributed: int64 / float64 keys")
    n = keys.shape[0]
    if P == 1:
        idx = sort_values_indexer_device(keys, bool(ascending), na_position)
        return gather_device(keys, idx), idx, [gather_device(c, idx) for c in payload]
    # ---- 1. splitters
    S = int(samples_per_rank)
    take = min(S, n)
    if take:
        pos = torch.div(torch.arange(take, device=keys.device, dtype=torch.int64) * n, take, rounding_mode="floor")
        smp = gather_device(keys, pos)
    else:
        smp = keys[:0]
    buf = torch.zeros(S + 1, dtype=torch.int64, device=keys.device)
    buf[:take] = smp.view(torch.int64)
    buf[S] = take
    allb = torch.empty(P * (S + 1), dtype=torch.int64, device=keys.device)
    dist.all_gather_into_tensor(allb, buf, group=group)
    hb = allb.cpu().numpy().reshape(P, S + 1)
    vals = np.concatenate([hb[q, :hb[q, S]] for q in range(P)]).view(np.float64 if kname == "float64" else np.int64)
    if kname == "float64":
        vals = vals[~np.isnan(vals)]
    img = np.sort(ordered_image(vals))
    if img.size:
        splitters = [int(img[min(img.size - 1, (i * img.size) // P)]) for i in range(1, P)]
    else:
        splitters = [0] * (P - 1)
    nan_dest = 0 if na_position == "first" else P - 1
    # ---- 2. range shuffle of (key, global row id, payload...)
    ncols = 2 + len(payload)
    ctx = auto_peer(group, n, ncols)
    if ctx is None:
        raise RuntimeError("sort_values_distributed needs the peer-memory transport (one NVLink node, NCCL backend)")
    pl = ctx.plan(keys, ROUTE_RANGE, (splitters, 0 if ascending else 1, nan_dest))
    ctx.exchange([pl])
    need = ctx.max_recv(pl)
    if need > ctx.cap:
        ctx = auto_peer(group, 0, ncols, need)
        pl = ctx.plan(keys, ROUTE_RANGE, (splitters, 0 if ascending else 1, nan_dest))
        ctx.exchange([pl])
    base = int(pl.matrix[:rank].sum())
    k2, cols2 = ctx.send(pl, [None] + list(payload), 1, base)
    ids2, pay2 = cols2[0], cols2[1:]
    # ---- 3. local stable order
    idx = sort_values_indexer_device(k2, bool(ascending), na_position)
    return gather_device(k2, idx), gather_device(ids2, idx), [gather_device(c, idx) for c in pay2]


# ------------------------------------------------------------------ string columns (str_arr) in the shuffle
def str_hash_device(sa):
    """64-bit hash of every string of a device StringArray (NA rows share one value) -> int64 CUDA tensor: the routing
    key of a string column (the reference hashes the decoded unicode object, shuffle_utils.py:119-140)."""
    torch = _torch()
    lib = native.load()
    d = sa.cuda()
    out = torch.empty(d.n, dtype=torch.int64, device=d.offsets.device)
    with torch.cuda.device(d.offsets.device):
        native.check(lib.sdcb200_str_hash(d.chars.data_ptr(), d.offsets.data_ptr(), d.bitmap.data_ptr(), d.n, out.data_ptr(),
                                          native.current_stream_ptr()))
    return out


def str_shuffle(keys, cols, group=None):
    """Route every row of a string-keyed table to rank hash(string) % P.  The string column travels in the reference's
    wire format -- a length array plus the characters, offsets rebuilt by a scan on the receiving side
    (sdc/shuffle_utils.py:277-297, sdc_str_arr.hpp:242-252; the validity bit rides in bit 31 of the length) -- the
    numeric columns as in the numeric shuffle.  keys: device StringArray; cols: 8-byte CUDA columns.
    -> (StringArray of the rows this rank now owns, [cols'...]); rows of one source keep their order."""
    from .sort import gather_device
    from .str_arr import StringArray
    torch = _torch()
    lib = native.load()
    rank, P = world(group)
    d = keys.cuda()
    if P == 1:
        return d, list(cols)
    dev = d.offsets.device
    h = str_hash_device(d)
    perm, counts = hash_partition_device(h, P, ROUTE_HASH)
    packed = d.take(perm)                                     # strings grouped by destination (two-phase gather)
    lens = torch.empty(max(packed.n, 1), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        native.check(lib.sdcb200_str_pack_lens(packed.offsets.data_ptr(), packed.bitmap.data_ptr(), packed.n, lens.data_ptr(),
                                               native.current_stream_ptr()))
    lens = lens[:packed.n]
    # characters per destination: differences of the packed offsets at the destination boundaries (P + 1 values)
    bounds = torch.zeros(P + 1, dtype=torch.int64, device=dev)
    bounds[1:] = torch.cumsum(counts, 0)
    char_bounds = packed.offsets[bounds].to(torch.int64) & 0xffffffff        # P + 1 values (offsets are uint32 bits)
    send_chars = char_bounds[1:] - char_bounds[:-1]
    both = exchange_counts(torch.cat([counts, send_chars]).view(2, P).t().contiguous().view(-1), group).view(P, 2)
    recv_rows, recv_chars = both[:, 0].contiguous(), both[:, 1].contiguous()
    lens_r = exchange_rows(lens, counts, recv_rows, group)
    chars_r = exchange_rows(packed.chars[:int(char_bounds[-1])], send_chars, recv_chars, group)
    n_r = lens_r.shape[0]
    offsets = torch.empty(n_r + 1, dtype=torch.int32, device=dev)
    bitmap = torch.empty(max(1, (n_r + 7) // 8), dtype=torch.uint8, device=dev)
    total = ct.c_int64(0)
    with torch.cuda.device(dev):
        native.check(lib.sdcb200_str_unpack_lens(lens_r.data_ptr() if n_r else None, n_r, offsets.data_ptr(), bitmap.data_ptr(),
                                                 ct.byref(total), native.current_stream_ptr()))
    if chars_r.shape[0] == 0:
        chars_r = torch.zeros(1, dtype=torch.uint8, device=dev)[:0]
    out_keys = StringArray(offsets, chars_r if chars_r.shape[0] else torch.zeros(1, dtype=torch.uint8, device=dev), bitmap[:(n_r + 7) // 8] if n_r else bitmap[:0], n_r)
    out_cols = [exchange_rows(gather_device(c, perm), counts, recv_rows, group) for c in cols]
    return out_keys, out_cols


def groupby_str_distributed(keys, cols, aggs, sort=True, ddof=1, expected_groups=0, group=None):
    """df.groupby(<string column>).<aggs>() over row-sharded columns: rows go to rank hash(string) % P (str_shuffle), every
    rank aggregates the strings it owns (factorize + aggregate + key strings, str_arr.groupby_str_device).
    -> (StringArray of this rank's group keys, {agg: [tensor per column]}, 'sharded' | 'replicated')"""
    from .str_arr import groupby_str_device
    _, P = world(group)
    if P == 1:
        k, r = groupby_str_device(keys, cols, aggs, sort, ddof, expected_groups)
        return k, r, "replicated"
    k2, c2 = str_shuffle(keys, cols, group)
    k, r = groupby_str_device(k2, c2, aggs, sort, ddof, max(1, expected_groups // P) if expected_groups > 0 else 0)
    return k, r, "sharded"


class HaloLink:
    """Neighbour links for a row-sharded series: every rank owns a small mailbox in peer-mapped memory (CUDA IPC);
    the rolling kernel itself stores this rank's last win-1 rows into the next rank's mailbox over NVLink and waits
    (in the tile it processes last) for the previous rank's -- `sdcb200_rolling_linked`.  No collective, no extra
    launch per call.  One node, one process per GPU; collective constructor / close()."""

    def __init__(self, group=None):
        torch, dist = _torch(), _dist()
        self.group = group
        self.rank, self.P = world(group)
        self.lib = native.load()
        self.epoch = 0
        nbytes = int(self.lib.sdcb200_halo_mailbox_bytes())
        p = ct.c_void_p()
        native.check(self.lib.sdcb200_ipc_alloc(ct.byref(p), nbytes))
        self.own = p.value
        zero = torch.as_tensor(_RawView(self.own, nbytes // 8), device=torch.device("cuda", torch.cuda.current_device()))
        zero.zero_()
        torch.cuda.synchronize()
        h = (ct.c_uint8 * 64)()
        native.check(self.lib.sdcb200_ipc_get_handle(p, h))
        handles = [None] * self.P
        dist.all_gather_object(handles, bytes(h), group=group)      # also orders every rank's zeroing before any use
        self.next = self.prev = None
        for r in (self.rank - 1, self.rank + 1):
            if 0 <= r < self.P:
                q = ct.c_void_p()
                native.check(self.lib.sdcb200_ipc_open((ct.c_uint8 * 64).from_buffer_copy(handles[r]), ct.byref(q)))
                if r < self.rank:
                    self.prev = q.value
                else:
                    self.next = q.value
        dist.barrier(group=group)

    def close(self):
        torch, dist = _torch(), _dist()
        if self.own is None:
            return
        torch.cuda.synchronize()
        for q in (self.prev, self.next):
            if q is not None:
                self.lib.sdcb200_ipc_close(q)
        dist.barrier(group=self.group)
        self.lib.sdcb200_ipc_free(self.own)
        self.own = self.prev = self.next = None

    def rolling(self, col, op, window, minp, ddof=1, row_offset=0, out=None):
        torch = _torch()
        assert col.is_cuda and col.dim() == 1 and col.is_contiguous()
        name = {torch.float64: "float64", torch.int64: "int64"}[col.dtype]
        if out is None:
            out = torch.empty(col.shape[0], dtype=torch.float64, device=col.device)
        if window <= 1:                               # no row of another shard is ever needed: no exchange, no epoch
            from .rolling import rolling_device
            return rolling_device(col, op, window, minp, ddof, out=out, row_offset=row_offset)
        if (window - 1) * 8 * 2 > int(self.lib.sdcb200_halo_mailbox_bytes()) - 128:
            # the halo no longer fits a mailbox slot (windows beyond the tiled limit): collective exchange instead
            from .rolling import rolling_device
            halo = halo_exchange(col, window - 1, self.group)
            return rolling_device(col, op, window, minp, ddof, out=out, halo=halo, row_offset=row_offset)
        self.epoch += 1
        with torch.cuda.device(col.device):
            native.check(self.lib.sdcb200_rolling_linked(
                native.ROLL_OPS[op], native.DTYPE_CODES[name], col.data_ptr() if col.shape[0] else None, out.data_ptr() if col.shape[0] else None,
                col.shape[0], window, minp, ddof, row_offset, self.own, self.next, self.prev, self.epoch, native.current_stream_ptr()))
        return out


class _RawView:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2, "strides": None}


def rolling_distributed(col, op, window, minp, ddof=1, group=None, row_offset=None, out=None, link=None):
    """Series.rolling over a row-sharded column: the win-1 rows in front of the shard come from the previous
    rank(s), then the local kernel.  link: a HaloLink -> the exchange happens inside the kernel over peer memory;
    None -> a ragged all-gather of the shard tails in front of it (any backend, any shard lengths)."""
    from .rolling import rolling_device
    rank, P = world(group)
    if row_offset is None:
        row_offset = sum(allgather_sizes(col.shape[0], col.device, group)[:rank]) if P > 1 else 0
    if link is not None and P > 1:
        return link.rolling(col, op, window, minp, ddof, row_offset, out)
    halo = halo_exchange(col, window - 1, group) if window > 1 else None
    return rolling_device(col, op, window, minp, ddof, out=out, halo=halo, row_offset=row_offset)
