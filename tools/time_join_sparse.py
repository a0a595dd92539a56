#!/usr/bin/env python
"""Wall (CUDA event) time of the C3 join with keys that have no dense range, partitioned path on / off."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sdc_b200.join import join_device
r = torch.from_numpy(np.random.default_rng(2).permutation(10_000_000).astype(np.int64)).cuda() * 1000003
l = torch.from_numpy(np.random.default_rng(3).integers(0, 10_000_000, 100_000_000, dtype=np.int64)).cuda() * 1000003
for part in ("0", "1"):
    os.environ["SDCB200_JOIN_PART"] = part
    for _ in range(2):
        o = join_device(l, r, "inner")
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        o = join_device(l, r, "inner")
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print("part", part, "ms", float(np.median(ts)), "n_out", o[0].shape[0], "sum", int(o[0].sum()), int(o[1].sum()), flush=True)
