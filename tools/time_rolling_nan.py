#!/usr/bin/env python
"""Rolling mean / sum / std / min per 100 M rows, clean data and the 0.1 % NaN + a few inf variant, back to back."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sdc_b200.rolling import rolling_device
rng = np.random.default_rng(1)
xh = rng.random(100_000_000)
x = torch.from_numpy(xh).cuda()
bad = xh.copy()
bad[rng.integers(0, len(bad), len(bad) // 1000)] = np.nan
bad[rng.integers(0, len(bad), 8)] = np.inf
y = torch.from_numpy(bad).cuda()
for name, col in (("clean", x), ("nan", y)):
    for op in ("mean", "sum", "std", "min"):
        for _ in range(3):
            rolling_device(col, op, 100, 100)
        torch.cuda.synchronize()
        ts = []
        for _ in range(15):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            rolling_device(col, op, 100, 100)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        print(name, op, "ms median %.4f min %.4f" % (float(np.median(ts)), float(np.min(ts))), flush=True)
