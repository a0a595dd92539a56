#!/usr/bin/env python
"""One short invocation of one operator, for `ncu` (launch list or --set full) under gpurun.

    python tools/prof_case.py groupby|rolling|sort|join [reps]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

what = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rng = np.random.default_rng(0)
if what == "groupby":
    from sdc_b200.groupby import groupby_device
    n = 10_000_000
    a = torch.from_numpy(rng.integers(0, 1000, n, dtype=np.int64)).cuda()
    b = torch.from_numpy(rng.random(n)).cuda()
    for _ in range(reps):
        groupby_device(a, [b], ("sum",), True, 1, 1000)
elif what in ("groupby_100m", "groupby_100m_5agg"):
    from sdc_b200.groupby import groupby_device
    n = 100_000_000
    a = torch.from_numpy(rng.integers(0, 1000, n, dtype=np.int64)).cuda()
    b = torch.from_numpy(rng.random(n)).cuda()
    aggs = ("sum",) if what == "groupby_100m" else ("sum", "mean", "min", "max", "count")
    for _ in range(reps):
        groupby_device(a, [b], aggs, True, 1, 1000)
elif what == "groupby_sparse_5agg":
    # 1 K groups under sparse int64 keys: the hashed lean loop of gb_smem_kernel (regime 1b)
    from sdc_b200.groupby import groupby_device
    n = 100_000_000
    a = torch.from_numpy(rng.integers(0, 1000, n, dtype=np.int64) * 1000003 + 17).cuda()
    b = torch.from_numpy(rng.random(n)).cuda()
    for _ in range(reps):
        groupby_device(a, [b], ("sum", "mean", "min", "max", "count"), True, 1, 1000)
elif what == "csv":
    import io
    import pyarrow as pa
    import pyarrow.csv as pacsv
    from sdc_b200.csv import read_csv_device
    table = pa.table({"c%d" % c: rng.standard_normal(500_000) * 10.0 ** (c % 7) for c in range(24)})
    sink = io.BytesIO()
    pacsv.write_csv(table, sink)
    raw = sink.getvalue()
    for _ in range(reps):
        read_csv_device(raw)
elif what == "groupby_small":
    from sdc_b200.groupby import groupby_device
    n = 200_000
    a = torch.from_numpy(rng.integers(0, 1000, n, dtype=np.int64)).cuda()
    b = torch.from_numpy(rng.random(n)).cuda()
    for _ in range(reps):
        groupby_device(a, [b], ("sum",), True, 1, 1000)
elif what == "groupby_hi":
    from sdc_b200.groupby import groupby_device
    n = 100_000_000
    a = torch.from_numpy(rng.integers(0, 1_000_000, n, dtype=np.int64)).cuda()
    b = torch.from_numpy(rng.random(n)).cuda()
    for _ in range(reps):
        groupby_device(a, [b], ("sum",), True, 1, 1_000_000)
elif what == "groupby_50m":
    from sdc_b200.groupby import groupby_device
    n = 100_000_000
    a = torch.from_numpy(rng.integers(0, 50_000_000, n, dtype=np.int64)).cuda()
    b = torch.from_numpy(rng.random(n)).cuda()
    for _ in range(reps):
        groupby_device(a, [b], ("sum",), True, 1, 50_000_000)
elif what == "rolling":
    from sdc_b200.rolling import rolling_device
    x = torch.from_numpy(np.random.default_rng(1).random(100_000_000)).cuda()
    out = torch.empty_like(x)
    for _ in range(reps):
        rolling_device(x, "mean", 100, 100, 1, out=out)
elif what in ("rolling_std", "rolling_min", "rolling_mean_nan", "rolling_var", "rolling_sum"):
    from sdc_b200.rolling import rolling_device
    x = torch.from_numpy(np.random.default_rng(1).random(100_000_000)).cuda()
    if what == "rolling_mean_nan":
        x[::1000] = float("nan")
    out = torch.empty_like(x)
    for _ in range(reps):
        rolling_device(x, {"rolling_std": "std", "rolling_min": "min", "rolling_mean_nan": "mean", "rolling_var": "var", "rolling_sum": "sum"}[what], 100, 100, 1, out=out)
elif what == "rolling_skew":
    from sdc_b200.rolling import rolling_moments_device
    x = torch.from_numpy(np.random.default_rng(1).random(100_000_000)).cuda()
    for _ in range(reps):
        rolling_moments_device(x, "skew", 100, 100)
elif what == "strings":
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_ops
    bench_ops.bench_strings(1, 50_000_000)
elif what == "sort":
    from sdc_b200.sort import argsort_device
    x = torch.from_numpy(rng.integers(-2**63, 2**63 - 1, 100_000_000, dtype=np.int64)).cuda()
    for _ in range(reps):
        argsort_device(x)
elif what == "join_left":
    from sdc_b200.join import join_device
    r = torch.from_numpy(np.random.default_rng(2).permutation(10_000_000).astype(np.int64)).cuda()
    l = torch.from_numpy(np.random.default_rng(3).integers(0, 10_000_000, 100_000_000, dtype=np.int64)).cuda()
    for _ in range(reps):
        join_device(l, r, "left")
elif what == "join_sparse":
    from sdc_b200.join import join_device
    r = torch.from_numpy(np.random.default_rng(2).permutation(10_000_000).astype(np.int64)).cuda() * 1000003
    l = torch.from_numpy(np.random.default_rng(3).integers(0, 10_000_000, 100_000_000, dtype=np.int64)).cuda() * 1000003
    for _ in range(reps):
        join_device(l, r, "left")
        join_device(l, r, "inner")
elif what == "join":
    from sdc_b200.join import join_device
    r = torch.from_numpy(np.random.default_rng(2).permutation(10_000_000).astype(np.int64)).cuda()
    l = torch.from_numpy(np.random.default_rng(3).integers(0, 10_000_000, 100_000_000, dtype=np.int64)).cuda()
    for _ in range(reps):
        join_device(l, r, "inner")
torch.cuda.synchronize()
print("ok")
