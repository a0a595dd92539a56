#!/usr/bin/env python
"""A/B timings of kernel variants selected by environment switches (read once per process, so each
variant runs in its own process):  python tools/ab_r1.py groupby|join [--reps N]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
from bench_ops import emit, timeit  # noqa: E402


def groupby(reps):
    from sdc_b200.groupby import groupby_device
    rng = np.random.default_rng(0)
    for n in (10_000_000, 100_000_000):
        a = torch.from_numpy(rng.integers(0, 1000, n, dtype=np.int64)).cuda()
        b = torch.from_numpy(rng.random(n)).cuda()
        variants = {"dense": a, "sparse": a * 1000003}
        if n == 100_000_000:
            z = torch.from_numpy(np.minimum(rng.zipf(1.3, n), 1000).astype(np.int64) - 1).cuda()
            variants["zipf1.3"] = z
        for vname, keys in variants.items():
            for aggs in (("sum",), ("sum", "mean", "min", "max", "count")):
                for hint in (1000, 0):
                    med, best = timeit(lambda: groupby_device(keys, [b], aggs, True, 1, hint), reps)
                    emit("groupby_%s %s n=%d hint=%d" % ("+".join(aggs), vname, n, hint), n, 16 * n, med, best,
                         env={k: v for k, v in os.environ.items() if k.startswith("SDCB200_")})


def join(reps):
    from sdc_b200.join import join_device
    nb, npr = 10_000_000, 100_000_000
    r0 = torch.from_numpy(np.random.default_rng(2).permutation(nb).astype(np.int64)).cuda()
    for name, hi in (("hit100", nb), ("hit50", 2 * nb), ("sparsekeys_hit100", nb)):
        l = torch.from_numpy(np.random.default_rng(3).integers(0, hi, npr, dtype=np.int64)).cuda()
        r = r0
        if name.startswith("sparse"):
            l = l * 1000003
            r = r0 * 1000003
        for how in ("inner", "left"):
            out = {}

            def run():
                out["r"] = join_device(l, r, how)
            med, best = timeit(run, reps)
            n_out = out["r"][0].shape[0]
            emit("join_%s_%s" % (how, name), npr, 8 * (npr + nb) + 16 * n_out, med, best, n_out=n_out,
                 env={k: v for k, v in os.environ.items() if k.startswith("SDCB200_")})
        del l


if __name__ == "__main__":
    reps = 10
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    if "--reps" in sys.argv:
        reps = int(sys.argv[sys.argv.index("--reps") + 1])
        args = [a for a in args if a != str(reps)]
    for w in args:
        {"groupby": groupby, "join": join}[w](reps)
