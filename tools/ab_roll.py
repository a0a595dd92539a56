#!/usr/bin/env python
"""Rolling kernel variants chosen by environment switches read once per process (SDCB200_ROLL_R = run length), one
process per setting:  SDCB200_ROLL_R=17 python tools/ab_roll.py [--reps N]
(The half-size-CTA experiment of DESIGN 4.7 used this script with a build that carried a 128-thread instantiation.)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
from bench_ops import timeit  # noqa: E402
from sdc_b200.rolling import rolling_device  # noqa: E402

reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 20
n = 100_000_000
x = torch.from_numpy(np.random.default_rng(1).random(n)).cuda()
out = torch.empty_like(x)
env = {k: v for k, v in os.environ.items() if k.startswith("SDCB200_")}
for op in ("sum", "mean", "std", "var", "count"):
    med, best = timeit(lambda: rolling_device(x, op, 100, 100, 1, out=out), reps)
    print(json.dumps({"case": "rolling_%s w=100 n=%d" % (op, n), "ms_median": round(med, 4), "ms_best": round(best, 4),
                      "frac_of_6553": round(16 * n / med / 1e6 / 6553, 3), "env": env}), flush=True)
