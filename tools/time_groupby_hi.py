#!/usr/bin/env python
"""groupby-sum over 125 M rows with K possible int64 keys (C5's local aggregate): dense-partitioned path on / off."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sdc_b200.groupby import groupby_device
n = 125_000_000
g = torch.Generator(device="cuda").manual_seed(5)
vals = torch.rand(n, device="cuda", dtype=torch.float64, generator=g)
for K in (12_500_000, 100_000_000):
    keys = torch.randint(0, K, (n,), device="cuda", generator=g)
    ref = None
    for dense, part in (("0", "0"), (str(256 << 20), "0"), (str(256 << 20), "1")):
        os.environ["SDCB200_GB_DENSE_GROUPS"] = dense
        os.environ["SDCB200_GB_WIN_PART"] = part
        for _ in range(2):
            gk, r = groupby_device(keys, [vals], ("sum",), True, 1, K)
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            gk, r = groupby_device(keys, [vals], ("sum",), True, 1, K)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        chk = (int(gk.shape[0]), int(gk.sum()), float(r["sum"][0].sum()))
        if ref is None:
            ref = (gk.clone(), r["sum"][0].clone())
        else:
            assert torch.equal(gk, ref[0]) and torch.allclose(r["sum"][0], ref[1], rtol=1e-9), "window path differs from the sort path"
        print("K", K, "dense_groups", dense, "part", part, "ms", float(np.median(ts)), chk, flush=True)
    del keys
