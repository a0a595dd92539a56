#!/usr/bin/env python
"""A/B of the lean groupby loop's two knobs inside one process (the library reads them on every call):
SDCB200_GB_RF = most lane-indexed copies of the direct window in the shared table (1 = off), SDCB200_GB_HOT = keep the
most frequent sampled keys' aggregates in registers.  Every setting is checked against setting (1, 0) on the same input.
  python tools/ab_lean.py [--reps N] [--settings rf:hot,rf:hot,...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
from bench_ops import timeit  # noqa: E402
from sdc_b200.groupby import groupby_device  # noqa: E402

reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 10
settings = [(1, 0), (2, 0), (4, 0), (32, 0), (1, 1), (32, 1)]
if "--settings" in sys.argv:
    settings = [tuple(int(x) for x in s.split(":")) for s in sys.argv[sys.argv.index("--settings") + 1].split(",")]
rng = np.random.default_rng(0)
n = 100_000_000
b = torch.from_numpy(rng.random(n)).cuda()
bn = b.clone()
bn[::97] = float("nan")
cases = {
    "dense1k": torch.from_numpy(rng.integers(0, 1000, n, dtype=np.int64)).cuda(),
    "zipf1.3": torch.from_numpy((np.minimum(rng.zipf(1.3, n), 1000) - 1).astype(np.int64)).cuda(),
    "dense64": torch.from_numpy(rng.integers(0, 64, n, dtype=np.int64)).cuda(),
    "dense10": torch.from_numpy(rng.integers(0, 10, n, dtype=np.int64)).cuda(),
}
for cname, keys in cases.items():
    for nn in ((n, 10_000_000) if cname == "dense1k" else (n,)):
        for aggs in (("sum",), ("sum", "mean", "min", "max", "count")):
            ref = None
            for hint in (1000, 0):
              for gs, m in settings:
                  os.environ["SDCB200_GB_RF"] = str(gs)
                  os.environ["SDCB200_GB_HOT"] = str(m)
                  k, v = keys[:nn], bn[:nn]
                  gk, res = groupby_device(k, [v], aggs, True, 1, hint)
                  ok = True
                  if ref is None:
                      ref = (gk, res)
                  else:
                      ok = bool(torch.equal(gk, ref[0]))
                      for a in aggs:
                          x, y = res[a][0], ref[1][a][0]
                          if a in ("sum", "mean"):
                              ok = ok and bool(torch.allclose(x, y, rtol=1e-9, atol=0, equal_nan=True))
                          else:
                              ok = ok and bool(torch.equal(x, y))
                  med, best = timeit(lambda: groupby_device(k, [b[:nn]], aggs, True, 1, hint), reps)
                  print(json.dumps({"case": cname, "n": nn, "aggs": "+".join(aggs), "hint": hint, "rf": gs, "hot": m, "ms_median": round(med, 4),
                                    "ms_best": round(best, 4), "Grows_per_s": round(nn / med / 1e6, 1), "same_as_first": ok}), flush=True)

# several float64 columns: the generic all-columns kernel (SDCB200_GB_SPLIT_COLS=0) against the lean loop per column
os.environ["SDCB200_GB_RF"] = "32"
os.environ["SDCB200_GB_HOT"] = "1"
b2 = torch.from_numpy(rng.standard_normal(n)).cuda()
b3 = b2 * 3.0
for cname in ("dense10", "dense64", "dense1k", "zipf1.3"):
    keys = cases[cname]
    for label, cols, aggs in (("2cols_sum", [b, b2], ("sum",)), ("3cols_mean", [b, b2, b3], ("mean",)),
                              ("2cols_5aggs", [b, b2], ("sum", "mean", "min", "max", "count"))):
        for split in (0, 1):
            os.environ["SDCB200_GB_SPLIT_COLS"] = str(split)
            med, best = timeit(lambda: groupby_device(keys, cols, aggs, True, 1, 1000), reps)
            print(json.dumps({"case": cname, "n": n, "aggs": label, "hint": 1000, "split_cols": split, "ms_median": round(med, 4),
                              "ms_best": round(best, 4), "Grows_per_s": round(n / med / 1e6, 1)}), flush=True)
