#!/usr/bin/env python
"""CUDA-event time of the C3 inner join (100 M probes x 10 M unique build keys): left-row order (two passes) against the
row-set form (one kernel, atomic output ranges) with 2048- and 4096-row tiles; dense keys and keys * 1000003."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sdc_b200.join import join_device
r0 = torch.from_numpy(np.random.default_rng(2).permutation(10_000_000).astype(np.int64)).cuda()
l0 = torch.from_numpy(np.random.default_rng(3).integers(0, 10_000_000, 100_000_000, dtype=np.int64)).cuda()
for name, mul in (("dense", 1), ("sparse", 1000003)):
    l, r = l0 * mul, r0 * mul
    for ordered, items in ((True, "8"), (False, "8")):
        os.environ["SDCB200_JOIN_ATOMIC_ITEMS"] = items
        for _ in range(2):
            o = join_device(l, r, "inner", ordered=ordered)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            o = join_device(l, r, "inner", ordered=ordered)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ok = bool((r[o[1]] == l[o[0]]).all()) and o[0].shape[0] == l.shape[0]
        print(name, "ordered", ordered, "items", items, "ms", float(np.median(ts)), "ok", ok, flush=True)
        del o
