#!/usr/bin/env python
"""C1 (10 M rows, 1 K groups, sum): blocking calls against pipelined begin / end calls (CUDA events around K calls)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sdc_b200.groupby import groupby_device, groupby_device_async
n = 10_000_000
rng = np.random.default_rng(0)
a = torch.from_numpy(rng.integers(0, 1000, n, dtype=np.int64)).cuda()
b = torch.from_numpy(rng.random(n)).cuda()
K = 20
for mode in ("blocking", "pipelined"):
    ts = []
    for rep in range(7):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if mode == "blocking":
            out = [groupby_device(a, [b], ("sum",), True, 1, 1000) for _ in range(K)]
        else:
            pend = [groupby_device_async(a, [b], ("sum",), True, 1, 1000) for _ in range(K)]
            e1.record()
            out = [p.result() for p in pend]
        if mode == "blocking":
            e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / K * 1e3)
    print(mode, "us per call: median %.1f min %.1f" % (float(np.median(ts)), min(ts)), "check", float(out[-1][1]["sum"][0].sum()))
