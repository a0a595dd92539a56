#!/usr/bin/env python
"""A/B timing of the onesweep variants (switches are read once per process, so every variant is a subprocess).

    python tools/sort_ab.py            # all variants, one JSON line each
    python tools/sort_ab.py --one      # the variant selected by the environment
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VARIANTS = [{"SDCB200_SORT_CFG": c, "SDCB200_SORT_TMA": t} for c in ("0", "1", "2", "3") for t in ("1", "0")]


def one():
    import numpy as np
    import torch
    from sdc_b200.sort import argsort_device, sort_device
    n = 100_000_000
    x = torch.from_numpy(np.random.default_rng(0).integers(-2**63, 2**63 - 1, n, dtype=np.int64)).cuda()
    order = argsort_device(x)
    ok = bool(torch.equal(x[order], torch.sort(x, stable=True)[0]))
    small = torch.from_numpy(np.random.default_rng(1).integers(0, 50, 3_000_001, dtype=np.int64)).cuda()
    ok = ok and bool(torch.equal(argsort_device(small), torch.argsort(small, stable=True)))
    ok = ok and bool(torch.equal(argsort_device(small, ascending=False), torch.argsort(small, stable=True, descending=True)))
    res = {"variant": {k: os.environ.get(k) for k in ("SDCB200_SORT_CFG", "SDCB200_SORT_TMA")}, "ok": ok}
    for name, fn in (("argsort_i64_100M", lambda: argsort_device(x)), ("sort_i64_100M", lambda: sort_device(x.clone()))):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        res[name + "_ms"] = float(np.median(ts))
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    if "--one" in sys.argv:
        one()
    else:
        for v in VARIANTS:
            env = dict(os.environ)
            env.update(v)
            subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=env)
