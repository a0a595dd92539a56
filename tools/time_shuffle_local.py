#!/usr/bin/env python
"""The peer shuffle's two kernels on ONE GPU: 8 destination arenas that are all local buffers (no NVLink), so the kernels'
own cost is visible -- count pass + tile scan (sdcb200_shuffle_plan) and the partition-and-send kernel
(sdcb200_shuffle_send), 125 M rows of key + value (or key + generated row id)."""
import ctypes as ct
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sdc_b200 import native
lib = native.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 125_000_000
P = int(sys.argv[2]) if len(sys.argv) > 2 else 8
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
g = torch.Generator(device="cuda").manual_seed(5)
keys = torch.randint(0, 100_000_000, (n,), device="cuda", generator=g)
vals = torch.rand(n, device="cuda", dtype=torch.float64, generator=g)
cap = (n // P * 5 // 4 + 4096) & ~1
arenas = [torch.empty(2 * cap, dtype=torch.int64, device="cuda") for _ in range(P)]
counts = torch.empty(P, dtype=torch.int64, device="cuda")
tile_off = torch.empty(-(-n // 4096) * 8, dtype=torch.int32, device="cuda")
s = native.current_stream_ptr()
for route in (1, 0):
    for iota in (False, True):
        tp, ts = [], []
        for rep in range(reps + 1):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            native.check(lib.sdcb200_shuffle_plan(native.DTYPE_CODES["int64"], route, keys.data_ptr(), n, P, tile_off.data_ptr(),
                                                  counts.data_ptr(), s))
            e[1].record()
            cptr = (ct.c_void_p * 2)(keys.data_ptr(), None if iota else vals.data_ptr())
            offc = (ct.c_int64 * P)(*([0] * P))
            basec = (ct.c_void_p * P)(*[a.data_ptr() for a in arenas])
            native.check(lib.sdcb200_shuffle_send(native.DTYPE_CODES["int64"], route, 2, cptr, 1 if iota else -1, 7, n, P,
                                                  tile_off.data_ptr(), offc, basec, cap, s))
            e[2].record()
            torch.cuda.synchronize()
            if rep:
                tp.append(e[0].elapsed_time(e[1]))
                ts.append(e[1].elapsed_time(e[2]))
        c = counts.tolist()
        chk = sum(int(arenas[d][:c[d]].sum()) for d in range(P)) & 0xffffffffffff
        print("route", "mod" if route else "hash", "iota" if iota else "value", "plan ms %.3f send ms %.3f" % (float(np.median(tp)), float(np.median(ts))),
              "GB moved", 2 * n * 16 / 1e9, "check", chk, flush=True)
