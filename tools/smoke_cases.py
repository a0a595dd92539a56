#!/usr/bin/env python
"""Small invocations of every kernel family and code path in one short process (every rolling op at windows on both
sides of the tiled limit, every groupby regime with the thresholds lowered, every join path with and without ids, the
shuffle's three routes with bulk and plain stores into eight local arenas): a quick fault check after a kernel change
(compute-sanitizer is closed on this pool, so this runs plain; parity is the tests' job).
Usage: python tools/sanitize_cases.py [rolling groupby join sort strings shuffle]"""
import ctypes as ct
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SDCB200_GB_SORT_ROWS", "1000")
os.environ.setdefault("SDCB200_GB_SORT_GROUPS", "1000")
import numpy as np
import torch
from sdc_b200 import native

fam = sys.argv[1:] or ["rolling", "groupby", "join", "sort", "strings", "shuffle"]
rng = np.random.default_rng(0)
dev = "cuda"


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


if "rolling" in fam:
    from sdc_b200.rolling import rolling_device
    x = rng.standard_normal(30_011)
    x[::97] = np.nan
    d = t(x)
    for op in ("sum", "mean", "var", "std", "count", "min", "max"):
        for win in (3, 100, 1025, 4097, 5000, 40_000):
            rolling_device(d, op, win, win // 2, 1)
    rolling_device(d[7:].clone(), "mean", 6000, 10, 1, halo=d[:7].clone(), row_offset=7)
    print("rolling ok", flush=True)
if "groupby" in fam:
    from sdc_b200.groupby import groupby_device
    n = 50_000
    v = t(rng.standard_normal(n))
    for keys, hint in ((rng.integers(0, 50, n), 50), (rng.integers(0, 3000, n), 3000), (rng.integers(-20_000, 20_000, n), 40_000),
                       (rng.integers(0, 20_000, n) * 1_000_003, 20_000)):
        for part in ("0", "1"):
            os.environ["SDCB200_GB_WIN_PART"] = part
            for aggs in (("sum",), ("sum", "mean", "min", "max", "count"), ("var",)):
                for sort in (True, False):
                    groupby_device(t(keys), [v], aggs, sort, 1, hint)
    print("groupby ok", flush=True)
if "join" in fam:
    from sdc_b200.join import join_device
    l = rng.integers(0, 9000, 40_001)
    for r in (rng.permutation(12_000)[:8000], rng.integers(0, 9000, 5000)):
        for mul in (1, 1_000_003):
            for how in ("inner", "left", "right", "outer"):
                for part in ("1", "2"):
                    os.environ["SDCB200_JOIN_PART"] = part
                    join_device(t(l * mul), t(r * mul), how, ordered=False)
                    join_device(t(l * mul), t(r * mul), how, ordered=True, left_ids=t(np.arange(len(l)) + 5), right_ids=t(np.arange(len(r)) + 9))
    print("join ok", flush=True)
if "sort" in fam:
    from sdc_b200.sort import argsort_device, sort_device
    for a in (rng.integers(-10**15, 10**15, 30_001), rng.standard_normal(30_001), rng.integers(0, 5, 10_000)):
        argsort_device(t(a), True)
        argsort_device(t(a), False)
        sort_device(t(a))
    print("sort ok", flush=True)
if "strings" in fam:
    from sdc_b200.str_arr import StringArray, groupby_str_device
    words = ["w%04d" % i if i % 5 else "a-long-word-%d" % i for i in range(700)]
    vals = [words[i] for i in rng.integers(0, 700, 20_000)]
    sa = StringArray.from_pylist(vals).cuda()
    sa.argsort(True)
    sa.argsort(False)
    groupby_str_device(sa, [t(rng.standard_normal(20_000))], ("sum", "count"), True, 1)
    print("strings ok", flush=True)
if "shuffle" in fam:
    lib = native.load()
    n, P = 50_001, 8
    keys = t(rng.integers(-1000, 100_000, n))
    vals = t(rng.standard_normal(n))
    cap = (n + 4096) & ~1
    arenas = [torch.empty(3 * cap, dtype=torch.int64, device=dev) for _ in range(P)]
    counts = torch.empty(P, dtype=torch.int64, device=dev)
    tile_off = torch.empty(-(-n // 4096) * 8, dtype=torch.int32, device=dev)
    s = native.current_stream_ptr()
    for route in (0, 1, 2):
        for bulk in ("1", "0"):
            os.environ["SDCB200_SHUFFLE_BULK"] = bulk           # read once per process: the second value only matters on a fresh run
            sp = (ct.c_uint64 * 7)(*[(1 << 63) + 12_000 * (i + 1) for i in range(7)])
            if route == 2:
                native.check(lib.sdcb200_shuffle_plan_range(6, sp, 0, 7, keys.data_ptr(), n, P, tile_off.data_ptr(), counts.data_ptr(), s))
            else:
                native.check(lib.sdcb200_shuffle_plan(6, route, keys.data_ptr(), n, P, tile_off.data_ptr(), counts.data_ptr(), s))
            cptr = (ct.c_void_p * 3)(keys.data_ptr(), None, vals.data_ptr())
            offc = (ct.c_int64 * P)(*([0] * P))
            basec = (ct.c_void_p * P)(*[a.data_ptr() for a in arenas])
            if route == 2:
                native.check(lib.sdcb200_shuffle_send_range(6, sp, 0, 7, 3, cptr, 1, 100, n, P, tile_off.data_ptr(), offc, basec, cap, s))
            else:
                native.check(lib.sdcb200_shuffle_send(6, route, 3, cptr, 1, 100, n, P, tile_off.data_ptr(), offc, basec, cap, s))
            torch.cuda.synchronize()
            assert int(counts.sum()) == n
    print("shuffle ok", flush=True)
torch.cuda.synchronize()
print("all done")
