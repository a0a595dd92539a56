#!/usr/bin/env python
"""A/B of the partitioned (row-set) inner join's switches, one process per setting (the switches are read once):
SDCB200_JOIN_GROUP_MB (bytes of table built and probed at a time) x SDCB200_JOIN_LF (table load factor, percent).
Every setting prints its median CUDA-event time and order-independent checksums of the pairs."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    from sdc_b200.join import join_device
    r = torch.from_numpy(np.random.default_rng(2).permutation(10_000_000).astype(np.int64)).cuda() * 1000003
    hi = 20_000_000 if os.environ.get("AB_HALF") else 10_000_000
    l = torch.from_numpy(np.random.default_rng(3).integers(0, hi, 100_000_000, dtype=np.int64)).cuda() * 1000003
    for _ in range(2):
        o = join_device(l, r, "inner")
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        o = join_device(l, r, "inner")
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ok = bool((l[o[0]] == r[o[1]]).all())
    print({"group_mb": os.environ.get("SDCB200_JOIN_GROUP_MB"), "lf": os.environ.get("SDCB200_JOIN_LF"), "half": bool(os.environ.get("AB_HALF")),
           "ms": round(float(np.median(ts)), 4), "n_out": o[0].shape[0], "keys_equal": ok,
           "sum_l": int(o[0].sum()), "sum_r": int(o[1].sum())}, flush=True)
    sys.exit(0)
for half in ("", "1"):
    for gmb, lf in (("-1", "30"),):
        env = dict(os.environ, SDCB200_JOIN_GROUP_MB=gmb, SDCB200_JOIN_LF=lf)
        if half:
            env["AB_HALF"] = "1"
        subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, check=False)
