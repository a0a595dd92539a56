#!/usr/bin/env python
"""Experiment behind the fixed-width string-key path: C4's 8-character keys read as int64 words
through the numeric groupby / argsort, against the factorize path.  One JSON line per case."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from tools.bench_ops import timeit  # noqa: E402


def main(n=200_000_000, reps=5):
    from sdc_b200.groupby import groupby_device
    from sdc_b200.sort import argsort_device
    from sdc_b200.str_arr import StringArray, groupby_str_device
    alphabet = np.frombuffer(b"abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789", dtype=np.uint8)
    for groups in (1000, 1_000_000):
        rng = np.random.default_rng(4)
        dic = torch.from_numpy(alphabet[rng.integers(0, alphabet.size, (groups, 8))]).cuda()
        ids = torch.randint(0, groups, (n,), device="cuda", generator=torch.Generator(device="cuda").manual_seed(4))
        chars = dic[ids].reshape(-1).contiguous()
        del ids
        words = chars.view(torch.int64)
        vals = torch.rand(n, device="cuda", dtype=torch.float64)
        aggs = ("sum", "mean", "min", "max", "count")
        for hint in (groups, 0):
            med, best = timeit(lambda: groupby_device(words, [vals], aggs, True, 1, hint), reps, warm=1)
            print(json.dumps({"case": "words groupby 5 aggs G=%d hint=%d" % (groups, hint), "ms": med, "best": best}), flush=True)
        med, best = timeit(lambda: groupby_device(words, [vals], ("sum",), True, 1, groups), reps, warm=1)
        print(json.dumps({"case": "words groupby sum G=%d" % groups, "ms": med, "best": best}), flush=True)
        med, best = timeit(lambda: groupby_device(words.view(torch.float64), [vals], aggs, True, 1, groups), reps, warm=1)
        print(json.dumps({"case": "words-as-f64 groupby 5 aggs G=%d" % groups, "ms": med, "best": best}), flush=True)
        offsets = (torch.arange(n + 1, device="cuda", dtype=torch.int64) * 8).to(torch.int32)
        bitmap = torch.full(((n + 7) // 8,), 0xFF, dtype=torch.uint8, device="cuda")
        sa = StringArray(offsets, chars, bitmap, n)
        med, best = timeit(lambda: sa.factorize(groups), reps, warm=1)
        print(json.dumps({"case": "factorize G=%d" % groups, "ms": med, "best": best}), flush=True)
        med, best = timeit(lambda: groupby_str_device(sa, [vals], aggs, True, 1, groups, short_keys=False), reps, warm=1)
        print(json.dumps({"case": "str groupby 5 aggs G=%d (factorize)" % groups, "ms": med, "best": best}), flush=True)

        def words_path():
            sa._profile = None                       # the column profile is part of the measured call
            return groupby_str_device(sa, [vals], aggs, True, 1, groups)
        med, best = timeit(words_path, reps, warm=1)
        print(json.dumps({"case": "str groupby 5 aggs G=%d (words, profile inside)" % groups, "ms": med, "best": best}), flush=True)
        med, best = timeit(lambda: groupby_str_device(sa, [vals], aggs, True, 1, groups), reps, warm=1)
        print(json.dumps({"case": "str groupby 5 aggs G=%d (words, profile cached)" % groups, "ms": med, "best": best}), flush=True)
        med, best = timeit(lambda: sa.argsort(True), reps, warm=1)
        print(json.dumps({"case": "str argsort G=%d" % groups, "ms": med, "best": best}), flush=True)
        del sa, vals, chars, offsets, bitmap, words


if __name__ == "__main__":
    main()
