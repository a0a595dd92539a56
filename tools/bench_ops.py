#!/usr/bin/env python
"""Per-operator device timings (CUDA events, inputs resident in HBM) for the kernels that are not
bench.py's headline: groupby (C1), sort / argsort, join (C3).  One JSON line per case on stdout.

    python tools/bench_ops.py [groupby] [sort] [join] [--reps 10]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def timeit(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def emit(name, rows, alg_bytes, med, best, **kw):
    d = {"case": name, "rows": rows, "ms_median": med, "ms_best": best, "rows_per_s": rows / (med * 1e-3),
         "alg_GBps": alg_bytes / (med * 1e-3) / 1e9, "frac_of_measured_peak": alg_bytes / (med * 1e-3) / 1e9 / peak()}
    d.update(kw)
    print(json.dumps(d), flush=True)


def bench_groupby(reps):
    from sdc_b200.groupby import groupby_device
    for n, groups in ((10_000_000, 1000), (100_000_000, 1000), (100_000_000, 1_000_000), (100_000_000, 50_000_000)):
        rng = np.random.default_rng(0)
        a = torch.from_numpy(rng.integers(0, groups, n, dtype=np.int64)).cuda()
        b = torch.from_numpy(rng.random(n)).cuda()
        for aggs in (("sum",), ("sum", "mean", "min", "max", "count")):
            med, best = timeit(lambda: groupby_device(a, [b], aggs, True, 1, groups), reps)
            emit("groupby_%s n=%d groups=%d" % ("+".join(aggs), n, groups), n, 16 * n, med, best)
        if groups == 1_000_000:        # the same groups under non-dense keys: the hashed global table
            a2 = a * 1000003
            med, best = timeit(lambda: groupby_device(a2, [b], ("sum",), True, 1, groups), reps)
            emit("groupby_sum_sparsekeys n=%d groups=%d" % (n, groups), n, 16 * n, med, best)
            del a2
        del a, b
    # low cardinality beyond uniform keys: a handful of groups, Zipf-distributed keys, several value columns
    n = 100_000_000
    rng = np.random.default_rng(0)
    b = torch.from_numpy(rng.random(n)).cuda()
    b2 = torch.from_numpy(rng.standard_normal(n)).cuda()
    for kname, keys in (("10 groups", rng.integers(0, 10, n, dtype=np.int64)),
                        ("zipf1.3 over 1000", (np.minimum(rng.zipf(1.3, n), 1000) - 1).astype(np.int64)),
                        ("uniform 1000", rng.integers(0, 1000, n, dtype=np.int64))):
        a = torch.from_numpy(keys).cuda()
        for label, cols, aggs in (("sum", [b], ("sum",)), ("5aggs", [b], ("sum", "mean", "min", "max", "count")),
                                  ("sum 2 cols", [b, b2], ("sum",))):
            if kname == "uniform 1000" and len(cols) == 1:
                continue
            med, best = timeit(lambda: groupby_device(a, cols, aggs, True, 1, 1000), reps)
            emit("groupby_%s n=%d keys=%s" % (label, n, kname), n, (8 + 8 * len(cols)) * n, med, best)
        del a


def bench_sort(reps):
    from sdc_b200.sort import argsort_device, sort_device
    for n in (10_000_000, 100_000_000, 200_000_000):
        rng = np.random.default_rng(4)
        for name, arr in (("i64_full", rng.integers(-2**63, 2**63 - 1, n, dtype=np.int64)),
                          ("i64_1k", rng.integers(0, 1000, n, dtype=np.int64)),
                          ("f64", rng.standard_normal(n))):
            d = torch.from_numpy(arr).cuda()
            med, best = timeit(lambda: argsort_device(d), reps)
            emit("argsort_%s n=%d" % (name, n), n, 16 * n, med, best, note="alg bytes = key read + idx written")
            w = d.clone()
            med, best = timeit(lambda: sort_device(w.copy_(d)), reps)
            emit("sort_inplace+copy_%s n=%d" % (name, n), n, 16 * n, med, best)
            del d, w


def bench_join(reps):
    from sdc_b200.join import join_device
    nb, npr = 10_000_000, 100_000_000
    r = torch.from_numpy(np.random.default_rng(2).permutation(nb).astype(np.int64)).cuda()
    r0 = r
    for name, hi in (("hit100", nb), ("hit50", 2 * nb), ("sparsekeys_hit100", nb)):
        l = torch.from_numpy(np.random.default_rng(3).integers(0, hi, npr, dtype=np.int64)).cuda()
        r = r0
        if name.startswith("sparse"):      # keys spread over a wide range: the hash-table path instead of the dense one
            l = l * 1000003
            r = r0 * 1000003
        for how in ("inner", "left"):
            out = {}
            def run():
                out["r"] = join_device(l, r, how)
            med, best = timeit(run, reps)
            n_out = out["r"][0].shape[0]
            emit("join_%s_%s probe=%d build=%d" % (how, name, npr, nb), npr, 8 * (npr + nb) + 16 * n_out, med, best,
                 n_out=n_out)
        del l
    # build keys duplicated x4 (SURVEY 8d C3 variant): 25 M probe rows x 10 M build rows -> 100 M pairs
    rd = torch.from_numpy(np.tile(np.random.default_rng(2).permutation(nb // 4), 4).astype(np.int64)).cuda()
    ld = torch.from_numpy(np.random.default_rng(3).integers(0, nb // 4, npr // 4, dtype=np.int64)).cuda()
    for how in ("inner", "left"):
        out = {}

        def run_dup():
            out["r"] = join_device(ld, rd, how)
        med, best = timeit(run_dup, reps)
        n_out = out["r"][0].shape[0]
        emit("join_%s_dup4 probe=%d build=%d" % (how, npr // 4, nb), npr // 4, 8 * (npr // 4 + nb) + 16 * n_out, med, best, n_out=n_out)


def bench_rolling(reps, n=100_000_000):
    """C2: every Series.rolling(100) reduction on 100 M float64 rows, each as its own pass (16 B / row)."""
    from sdc_b200.rolling import rolling_device
    x = torch.from_numpy(np.random.default_rng(1).random(n)).cuda()
    out = torch.empty_like(x)
    for op in ("sum", "mean", "std", "var", "count", "min", "max"):
        med, best = timeit(lambda: rolling_device(x, op, 100, 100, 1, out=out), reps)
        emit("rolling_%s w=100 n=%d" % (op, n), n, 16 * n, med, best)
    from sdc_b200.rolling import rolling_moments_device
    y = torch.from_numpy(np.random.default_rng(2).random(n)).cuda()
    for op in ("skew", "kurt", "cov", "corr"):
        other = y if op in ("cov", "corr") else None
        med, best = timeit(lambda: rolling_moments_device(x, op, 100, 100, other), reps)
        emit("rolling_%s w=100 n=%d" % (op, n), n, (24 if other is not None else 16) * n, med, best)
    del y
    x[::1000] = float("nan")
    x[7::100_000] = float("inf")
    for op in ("mean", "std", "min"):
        med, best = timeit(lambda: rolling_device(x, op, 100, 100, 1, out=out), reps)
        emit("rolling_%s w=100 n=%d, 0.1%% NaN + a few inf" % (op, n), n, 16 * n, med, best)


def bench_strings(reps, n=200_000_000):
    """C4: sort_values + groupby multi-agg on 200 M rows with 8-char string keys (str_arr offsets + chars built
    on the device from a host dictionary of G distinct strings)."""
    from sdc_b200.str_arr import StringArray, groupby_str_device
    alphabet = np.frombuffer(b"abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789", dtype=np.uint8)
    for groups in (1000, 1_000_000):
        rng = np.random.default_rng(4)
        dic = torch.from_numpy(alphabet[rng.integers(0, alphabet.size, (groups, 8))]).cuda()       # [G, 8] uint8
        ids = torch.randint(0, groups, (n,), device="cuda", generator=torch.Generator(device="cuda").manual_seed(4))
        chars = dic[ids].reshape(-1).contiguous()
        del ids
        offsets = (torch.arange(n + 1, device="cuda", dtype=torch.int64) * 8).to(torch.int32)
        bitmap = torch.full(((n + 7) // 8,), 0xFF, dtype=torch.uint8, device="cuda")
        sa = StringArray(offsets, chars, bitmap, n)
        vals = torch.rand(n, device="cuda", dtype=torch.float64)
        med, best = timeit(lambda: sa.argsort(True), reps, warm=1)
        emit("str_argsort n=%d G=%d" % (n, groups), n, n * (12 + 8), med, best, note="alg bytes = offsets + chars read + idx written")
        aggs = ("sum", "mean", "min", "max", "count")
        med, best = timeit(lambda: groupby_str_device(sa, [vals], aggs, True, 1, groups), reps, warm=1)
        emit("str_groupby_multiagg n=%d G=%d" % (n, groups), n, n * (4 + 8 + 8), med, best,
             note="alg bytes = offsets + chars + value read once")
        del sa, vals, chars, offsets, bitmap


def bench_csv(reps, nrows=2_000_000, ncols=24):
    """Census-shaped CSV (24 float64 columns, SURVEY 8d C5 / benchmarks/census_benchmark.py:57-85): bytes resident in
    HBM -> columns (index + inference + conversion), the same from a pinned host buffer (H2D inside), and Arrow's
    multithreaded host reader (what the reference calls) on the box's cores."""
    import io
    import time
    import ctypes as ct
    import pyarrow as pa
    import pyarrow.csv as pacsv
    from sdc_b200 import native
    from sdc_b200.csv import read_csv_device
    rng = np.random.default_rng(6)
    table = pa.table({"c%d" % c: rng.standard_normal(nrows) * 10.0 ** (c % 7) for c in range(ncols)})
    sink = io.BytesIO()
    pacsv.write_csv(table, sink)
    raw = sink.getvalue()
    del table, sink
    nbytes = len(raw)
    med, best = timeit(lambda: read_csv_device(raw), max(3, reps // 2), warm=1)
    emit("csv_read_device %d rows x %d float64 cols, host bytes -> device columns (H2D + index + inference + conversion)" % (nrows, ncols),
         nrows, nbytes, med, best, csv_bytes=nbytes, csv_GBps=nbytes / (med * 1e-3) / 1e9)
    import tempfile
    with tempfile.NamedTemporaryFile(suffix=".csv", delete=False) as f:
        f.write(raw)
        path = f.name
    med, best = timeit(lambda: read_csv_device(path), max(3, reps // 2), warm=1)
    emit("csv_read_device %d rows x %d float64 cols, FILE (page cache) -> device columns" % (nrows, ncols),
         nrows, nbytes, med, best, csv_bytes=nbytes, csv_GBps=nbytes / (med * 1e-3) / 1e9)
    dtypes = {"c%d" % c: np.float64 for c in range(ncols)}
    med, best = timeit(lambda: read_csv_device(raw, dtype=dtypes), max(3, reps // 2), warm=1)
    emit("csv_read_device %d rows x %d float64 cols, dtypes given (no inference pass)" % (nrows, ncols), nrows, nbytes, med, best,
         csv_bytes=nbytes, csv_GBps=nbytes / (med * 1e-3) / 1e9)
    # the kernels alone: bytes already in HBM
    lib = native.load()
    data = torch.frombuffer(bytearray(raw), dtype=torch.uint8).cuda()
    first_nl = raw.index(b"\n") + 1
    outs = [torch.empty(nrows, dtype=torch.float64, device="cuda") for _ in range(ncols)]
    types = (ct.c_int32 * ncols)(*([2] * ncols))
    ptrs = (ct.c_void_p * ncols)(*[o.data_ptr() for o in outs])

    def kernels():
        h, n = ct.c_int64(0), ct.c_int64(0)
        native.check(lib.sdcb200_csv_open(data.data_ptr() + first_nl, nbytes - first_nl, ct.byref(h), ct.byref(n), native.current_stream_ptr()))
        native.check(lib.sdcb200_csv_parse(h.value, ncols, ord(","), types, ptrs))
        native.check(lib.sdcb200_csv_close(h.value))
    med, best = timeit(kernels, reps, warm=1)
    emit("csv_kernels %d rows x %d float64 cols, bytes resident in HBM (index + conversion)" % (nrows, ncols), nrows,
         nbytes + 8 * nrows * ncols, med, best, csv_bytes=nbytes, csv_GBps=nbytes / (med * 1e-3) / 1e9,
         note="alg bytes = text read once + columns written once")
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        pacsv.read_csv(path)
        ts.append(time.perf_counter() - t0)
    os.unlink(path)
    print(json.dumps({"case": "csv_read_arrow_host (the reference's reader, %d threads)" % pa.cpu_count(), "rows": nrows,
                      "ms_best": min(ts) * 1e3, "csv_GBps": nbytes / min(ts) / 1e9}), flush=True)


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    reps = 10
    if "--reps" in sys.argv:
        reps = int(sys.argv[sys.argv.index("--reps") + 1])
        args = [a for a in args if a != str(reps)]
    which = args or ["groupby", "sort"]
    for w in which:
        {"groupby": bench_groupby, "sort": bench_sort, "join": bench_join, "strings": bench_strings, "csv": bench_csv,
         "rolling": bench_rolling}[w](reps)
