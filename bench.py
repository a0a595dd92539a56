#!/usr/bin/env python
"""bench.py -- headline benchmark of the sdc_b200 hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Headline workload (BASELINE.json configs[1], "C2"): Series.rolling(window=100).mean() on a
100 M-row float64 column per GPU.  A "step" is one pass of the rolling kernel over the column.
Prints ONE JSON line (see the contract in DESIGN.md, "Measurement").  Under torchrun (N > 1) the same line also
carries C5 -- the census-shaped groupby and join hash-shuffled across the GPUs, 125 M rows per GPU -- under
`other_ops`, each case with its own in-run single-GPU time and an oracle parity check of the multi-GPU paths
(`parity`), so the scaling run certifies the shuffle as well as the rolling halo.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_ROWS = 100_000_000
WINDOW = 100
BYTES_PER_ROW = 16          # algorithmic: 8 read + 8 written
WORKLOAD = "C2: Series.rolling(100).mean(), 100M float64 rows per GPU, min_periods=None"
C5_ROWS = 125_000_000       # C5: 1 B rows over 8 GPUs, weak scaling
NVLINK_GBS = 900.0          # per direction per GPU (NVLink 5)


def config_dict():
    """The workload description -- identical in both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "rows_per_gpu": N_ROWS, "window": WINDOW, "min_periods": WINDOW,
            "l2": "inputs (800 MB in + 800 MB out per step) larger than L2",
            "multi_gpu": "row-sharded series, win-1 halo rows from the previous rank"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def make_column(rank=0):
    return np.random.default_rng(1 + rank).random(N_ROWS)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.thread = [], None, None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return

        def pump():
            for line in self.proc.stdout:
                self.rows.append(line.strip())
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------ CPU arms
def cpu_rolling_mean(x, repeats):
    """The oracle (restated reference algorithm, numba prange over all host threads)."""
    import numba
    from oracle import rolling as orolling
    orolling.rolling(x[:100_000], "mean", WINDOW)       # compile
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        orolling.rolling(x, "mean", WINDOW)
        best = min(best, time.perf_counter() - t0)
    return best, numba.config.NUMBA_NUM_THREADS, numba.threading_layer()


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the reference itself cannot be
    imported or compiled in this image) on the box's host cores, same config / metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numba
    from oracle import rolling as orolling
    x = make_column(0)
    orolling.rolling(x[:100_000], "mean", WINDOW)
    for _ in range(args.warmup):
        orolling.rolling(x, "mean", WINDOW)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orolling.rolling(x, "mean", WINDOW)
    dt = time.perf_counter() - t0
    value = N_ROWS * args.steps / dt
    cores = numba.config.NUMBA_NUM_THREADS
    sample = "full workload: %d rows per step, %d steps" % (N_ROWS, args.steps)
    print(json.dumps({
        "impl": "reference", "metric": "rolling_mean_rows_per_s", "value": value, "unit": "rows/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(),
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": cores, "kind": "port", "sample": sample,
                         "threading_layer": numba.threading_layer()},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the rolling kernel, from the committed
    `ncu --set full` summary (profiles/); None when no capture of the current kernel is on file."""
    for name in ("r2_rolling_mean_ncu_summary.json", "r1_rolling_mean_ncu_summary.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                d = json.load(f)
            return float(d["dram_bytes_read"]) + float(d["dram_bytes_write"]), "from profile (profiles/%s), not measured in this run" % name
        except Exception:
            continue
    return None, None


def _event_times(fn, reps, warm=3):
    import torch
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


def bench_other_ops(dev, peak, cpu_legs=True):
    """The other two operators BASELINE.json's metric names, at their configs, inputs resident in HBM:
    C1 groupby-sum (10 M rows, 1 K int64 groups, one f64 column) and C3 inner hash join (100 M probe x
    10 M build, int64 keys).  Whole-operator device time through the public Python API (CUDA events)."""
    import torch
    from sdc_b200.groupby import groupby_device, groupby_device_async
    from sdc_b200.join import join_device
    out = {}
    rng = np.random.default_rng(0)
    n = 10_000_000
    a_h = rng.integers(0, 1000, n, dtype=np.int64)
    b_h = rng.random(n)
    a, b = torch.from_numpy(a_h).to(dev), torch.from_numpy(b_h).to(dev)
    med, best = _event_times(lambda: groupby_device(a, [b], ("sum",), True, 1, 1000), 20)
    gb = {"workload": "C1: df.groupby('A').sum(), 10M rows, int64 key with 1K groups, one float64 column, sort=True",
          "rows_per_s": n / (med * 1e-3), "ms_median": med, "ms_best": best, "algorithmic_bytes": 16 * n,
          "achieved_GBps": 16 * n / (med * 1e-3) / 1e9, "frac_of_measured_peak": 16 * n / (med * 1e-3) / 1e9 / peak,
          "how": "one blocking call between two CUDA events (stream synchronisation and the Python mirror inside)"}
    # the same aggregate through the asynchronous entry points (sdcb200_groupby_begin / _end): K calls in flight, results
    # collected afterwards -- what a pipeline of small aggregates pays per call
    K, ts = 20, []
    want_k, want = groupby_device(a, [b], ("sum",), True, 1, 1000)
    for _ in range(7):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pend = [groupby_device_async(a, [b], ("sum",), True, 1, 1000) for _ in range(K)]
        e1.record()
        res = [p_.result() for p_ in pend]
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / K)
    ok = all(torch.equal(r[0], want_k) and torch.allclose(r[1]["sum"][0], want["sum"][0], rtol=1e-12) for r in res)
    pm = float(np.median(ts))
    gb["pipelined"] = {"ms_per_call": pm, "rows_per_s": n / (pm * 1e-3), "frac_of_measured_peak": 16 * n / (pm * 1e-3) / 1e9 / peak,
                       "calls_in_flight": K, "results_equal_blocking": bool(ok),
                       "api": "sdcb200_groupby_begin / sdcb200_groupby_end (no stream synchronisation per call)"}
    del res, pend
    n2 = 100_000_000
    a2 = torch.from_numpy(rng.integers(0, 1000, n2, dtype=np.int64)).to(dev)
    b2 = torch.from_numpy(rng.random(n2)).to(dev)
    med2, _ = _event_times(lambda: groupby_device(a2, [b2], ("sum",), True, 1, 1000), 10)
    gb["rows_per_s_100M"] = n2 / (med2 * 1e-3)
    gb["frac_of_measured_peak_100M"] = 16 * n2 / (med2 * 1e-3) / 1e9 / peak
    del a2, b2
    # "launch overhead matters at this size: also report 100 M and 1 B rows" -- 1 B rows = 16 GB in HBM,
    # generated on the device (same distributions: uniform keys in [0, 1000), uniform values in [0, 1))
    n3 = 1_000_000_000
    g3 = torch.Generator(device=dev).manual_seed(0)
    a3 = torch.randint(0, 1000, (n3,), device=dev, dtype=torch.int64, generator=g3)
    b3 = torch.rand(n3, device=dev, dtype=torch.float64, generator=g3)
    med3, _ = _event_times(lambda: groupby_device(a3, [b3], ("sum",), True, 1, 1000), 5, warm=2)
    k3, r3 = groupby_device(a3, [b3], ("sum",), True, 1, 1000)
    gb["rows_per_s_1B"] = n3 / (med3 * 1e-3)
    gb["ms_1B"] = med3
    gb["frac_of_measured_peak_1B"] = 16 * n3 / (med3 * 1e-3) / 1e9 / peak
    gb["check_1B"] = {"groups": int(k3.shape[0]), "sum_of_sums_rel_err": abs(float(r3["sum"][0].sum()) - float(b3.sum())) / float(b3.sum())}
    del a3, b3, k3, r3
    # C5's census shape on one GPU (24 float64 columns, 125 M rows per GPU = 24 GB + 1 GB of keys): every
    # column summed per key, K = 1 K -- the local aggregate of the combiner path
    nc, ncols = 125_000_000, 24
    a4 = torch.randint(0, 1000, (nc,), device=dev, dtype=torch.int64, generator=g3)
    cols4 = [torch.rand(nc, device=dev, dtype=torch.float64, generator=g3) for _ in range(ncols)]
    med4, _ = _event_times(lambda: groupby_device(a4, cols4, ("sum",), True, 1, 1000), 3, warm=1)
    k4, r4 = groupby_device(a4, cols4, ("sum",), True, 1, 1000)
    bytes4 = nc * 8 * (ncols + 1)
    out["groupby_sum_census24"] = {
        "workload": "C5 census shape on one GPU: df.groupby('A').sum() over 24 float64 columns, 125M rows, 1K int64 groups",
        "ms_median": med4, "rows_per_s": nc / (med4 * 1e-3), "algorithmic_bytes": bytes4,
        "frac_of_measured_peak": bytes4 / (med4 * 1e-3) / 1e9 / peak,
        "check": {"groups": int(k4.shape[0]),
                  "max_rel_err_of_column_totals": max(abs(float(r4["sum"][i].sum()) - float(cols4[i].sum())) / float(cols4[i].sum())
                                                      for i in range(ncols))}}
    del a4, cols4, k4, r4
    torch.cuda.empty_cache()
    if cpu_legs:
        import numba
        import pandas as pd
        from oracle import groupby as ogroupby
        ns = 2_000_000          # the dict-of-row-lists algorithm is ~5 Mrows/s: bounded sample
        ogroupby.groupby_agg(a_h[:10000], {"B": b_h[:10000]}, "sum")
        t0 = time.perf_counter()
        ogroupby.groupby_agg(a_h[:ns], {"B": b_h[:ns]}, "sum")
        dt = time.perf_counter() - t0
        df = pd.DataFrame({"A": a_h, "B": b_h})
        t0 = time.perf_counter()
        df.groupby("A").sum()
        dtp = time.perf_counter() - t0
        gb["cpu_baseline"] = {"value": ns / dt, "unit": "rows/s", "cores": 1, "kind": "port",
                              "sample": "first 2M rows of the workload, one call after a compile call; the port builds the "
                                        "per-chunk dicts in a SERIAL loop (the reference uses prange), so it understates the "
                                        "reference on a many-core host -- pandas on the full 10M rows is reported beside it",
                              "numba_threads_available": numba.config.NUMBA_NUM_THREADS, "pandas_rows_per_s": n / dtp}
    # end to end through the C ABI with HOST buffers (pinned): H2D of both columns, aggregate, D2H of keys + sums
    import ctypes as ct
    from sdc_b200 import native
    lib = native.load()
    a_p, b_p = torch.from_numpy(a_h).pin_memory(), torch.from_numpy(b_h).pin_memory()
    ptrs = (ct.c_void_p * 1)(b_p.data_ptr())
    dts = (ct.c_int32 * 1)(native.DTYPE_CODES["float64"])
    rk, rv = np.empty(4096, np.int64), np.empty(4096, np.float64)

    def host_call():
        h, ng = ct.c_int64(0), ct.c_int64(0)
        native.check(lib.sdcb200_host_groupby(native.DTYPE_CODES["int64"], a_p.data_ptr(), n, 1, ptrs, dts, native.AGG_BITS["sum"],
                                              1, 1, 1000, ct.byref(h), ct.byref(ng)))
        native.check(lib.sdcb200_groupby_keys(h.value, rk.ctypes.data, 1, None))
        native.check(lib.sdcb200_groupby_values(h.value, 0, native.AGG_BITS["sum"], rv.ctypes.data, 1, None))
        lib.sdcb200_groupby_free(h.value)
        return ng.value
    host_call()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        ng = host_call()
        ts.append(time.perf_counter() - t0)
    gb["e2e"] = {"value": n / float(np.median(ts)), "unit": "rows/s", "h2d_bytes_per_step": 16 * n, "d2h_bytes_per_step": 16 * ng,
                 "api": "sdcb200_host_groupby + sdcb200_groupby_keys / _values (pinned host buffers)", "ms_median": float(np.median(ts)) * 1e3}
    del a_p, b_p
    out["groupby_sum_c1"] = gb
    del a, b

    nb, npr = 10_000_000, 100_000_000
    r_h = np.random.default_rng(2).permutation(nb).astype(np.int64)
    l_h = np.random.default_rng(3).integers(0, nb, npr, dtype=np.int64)
    r, l = torch.from_numpy(r_h).to(dev), torch.from_numpy(l_h).to(dev)
    res = {}

    def run():
        res["o"] = join_device(l, r, "inner")
    med, best = _event_times(run, 5, warm=2)
    n_out = int(res["o"][0].shape[0])
    alg = 8 * (npr + nb) + 16 * n_out
    jn = {"workload": "C3: pd.merge inner, 100M-row probe x 10M-row build (unique), int64 keys -> int64 indexers",
          "rows_per_s": npr / (med * 1e-3), "ms_median": med, "ms_best": best, "n_out": n_out, "algorithmic_bytes": alg,
          "achieved_GBps": alg / (med * 1e-3) / 1e9, "frac_of_measured_peak": alg / (med * 1e-3) / 1e9 / peak}
    res.clear()

    # the same join with keys that are NOT a dense range (every key scaled by 1000003): the hashed / partitioned path
    r_s, l_s = r * 1000003, l * 1000003

    def run_sparse():
        res["o"] = join_device(l_s, r_s, "inner")
    meds, _ = _event_times(run_sparse, 3, warm=1)
    jn["sparse_keys"] = {"workload": "same join, keys * 1000003 (no dense range: partitioned hash join, pairs as a row set)",
                         "ms_median": meds, "rows_per_s": npr / (meds * 1e-3), "frac_of_measured_peak": alg / (meds * 1e-3) / 1e9 / peak}
    res.clear()
    del r_s, l_s
    if cpu_legs:
        import pandas as pd
        ns = 10_000_000
        L = pd.DataFrame({"k": l_h[:ns], "li": np.arange(ns)})
        R = pd.DataFrame({"k": r_h, "ri": np.arange(nb)})
        t0 = time.perf_counter()
        L.merge(R, on="k", how="inner")
        dt = time.perf_counter() - t0
        jn["cpu_baseline"] = {"value": ns / dt, "unit": "probe rows/s", "cores": 1, "kind": "pandas.merge (the reference ships no merge)",
                              "sample": "first 10M probe rows x the full 10M-row build side"}
    l_p, r_p = torch.from_numpy(l_h).pin_memory(), torch.from_numpy(r_h).pin_memory()
    li_p = torch.empty(npr, dtype=torch.int64).pin_memory()
    ri_p = torch.empty(npr, dtype=torch.int64).pin_memory()

    def host_join():
        h, no = ct.c_int64(0), ct.c_int64(0)
        native.check(lib.sdcb200_host_join(native.JOIN_HOW["inner"], native.DTYPE_CODES["int64"], l_p.data_ptr(), npr, r_p.data_ptr(), nb,
                                           ct.byref(h), ct.byref(no)))
        native.check(lib.sdcb200_join_indexers(h.value, li_p.data_ptr(), ri_p.data_ptr(), 1, None))
        lib.sdcb200_join_free(h.value, None)
        return no.value
    host_join()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        no = host_join()
        ts.append(time.perf_counter() - t0)
    jn["e2e"] = {"value": npr / float(np.median(ts)), "unit": "probe rows/s", "h2d_bytes_per_step": 8 * (npr + nb),
                 "d2h_bytes_per_step": 16 * no, "api": "sdcb200_host_join + sdcb200_join_indexers (pinned host buffers)",
                 "ms_median": float(np.median(ts)) * 1e3}
    out["join_inner_c3"] = jn
    del l, r, l_p, r_p, li_p, ri_p

    # sort_values feeding the path: stable int64 argsort, 100 M full-range keys (1.6 GB algorithmic: keys in, int64 rows out)
    from sdc_b200.sort import argsort_device
    ks = torch.from_numpy(np.random.default_rng(6).integers(-2**62, 2**62, 100_000_000, dtype=np.int64)).to(dev)
    meda, _ = _event_times(lambda: argsort_device(ks), 3, warm=1)
    out["argsort_i64_100M"] = {"workload": "stable argsort, 100M full-range int64 keys", "ms_median": meda, "rows_per_s": 1e8 / (meda * 1e-3),
                               "frac_of_measured_peak": 1.6e9 / (meda * 1e-3) / 1e9 / peak,
                               "frac_per_pass": 2.4e9 / (meda * 1e-3 / 8) / 1e9 / peak}
    del ks

    # C4: 200 M rows with 8-character string keys drawn from 1 K distinct strings (str_arr offsets + chars built on the
    # device): sort_values' argsort and the five-aggregate groupby; checked through totals (tests/test_strings_gpu.py
    # holds the full parity run).  alg bytes: argsort = offsets + chars in, int64 rows out; groupby = offsets + chars + value
    from sdc_b200.str_arr import StringArray, groupby_str_device
    n4, g4 = 200_000_000, 1000
    alphabet = np.frombuffer(b"abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789", dtype=np.uint8)
    dic = torch.from_numpy(alphabet[np.random.default_rng(4).integers(0, alphabet.size, (g4, 8))]).to(dev)
    ids = torch.randint(0, g4, (n4,), device=dev, generator=torch.Generator(device=dev).manual_seed(4))
    chars = dic[ids].reshape(-1).contiguous()
    del ids
    offsets = (torch.arange(n4 + 1, device=dev, dtype=torch.int64) * 8).to(torch.int32)
    bitmap = torch.full(((n4 + 7) // 8,), 0xFF, dtype=torch.uint8, device=dev)
    sa = StringArray(offsets, chars, bitmap, n4)
    vals = torch.rand(n4, device=dev, dtype=torch.float64)
    aggs = ("sum", "mean", "min", "max", "count")

    def c4_groupby():
        sa._profile = None                   # the column's length profile is part of the measured call
        return groupby_str_device(sa, [vals], aggs, True, 1, g4)
    medg, _ = _event_times(c4_groupby, 3, warm=1)
    rk, res = c4_groupby()
    ok = rk.n <= g4 and int(res["count"][0].sum().item()) == n4 and \
        abs(float(res["sum"][0].sum().item()) - float(vals.sum().item())) <= 1e-9 * n4
    keys_sorted = rk.to_pylist()
    ok = ok and keys_sorted == sorted(keys_sorted, key=lambda t: t.encode())
    meds, _ = _event_times(lambda: sa.argsort(True), 3, warm=1)
    out["strings_c4"] = {
        "workload": "C4: 200M rows, 8-char string keys from 1K distinct strings: five-aggregate groupby (sum, mean, min, max, count), stable argsort",
        "groupby_ms": medg, "groupby_rows_per_s": n4 / (medg * 1e-3), "groupby_frac_of_measured_peak": n4 * 20 / (medg * 1e-3) / 1e9 / peak,
        "argsort_ms": meds, "argsort_rows_per_s": n4 / (meds * 1e-3), "argsort_frac_of_measured_peak": n4 * 20 / (meds * 1e-3) / 1e9 / peak,
        "totals_ok": bool(ok)}
    del sa, vals, chars, offsets, bitmap, rk, res
    return out


# ------------------------------------------------------------------------------ C5 (multi-GPU groupby + join)
def c5_parity(dev, rank, world, link):
    """The multi-GPU paths against the oracle on a 400 K-row global table every rank regenerates from one seed
    (tests/dist_check.py is the long form).  -> {case: bool}; ANDed over ranks by the caller."""
    import torch
    from oracle import groupby as ogroupby
    from oracle import join as ojoin
    from oracle import rolling as orolling
    from sdc_b200 import distributed as D
    res = {}
    rng = np.random.default_rng(5)
    n = 400_003
    keys_lo, keys_hi = rng.integers(0, 1000, n), rng.integers(0, 150_000, n)
    vals = rng.standard_normal(n)
    vals[rng.integers(0, n, n // 100)] = np.nan
    s, e = D.block_range(n, rank, world)
    v = torch.from_numpy(vals[s:e].copy()).to(dev)

    def gb(keys, mode, hint):
        k = torch.from_numpy(keys[s:e].copy()).to(dev)
        gk, r, layout = D.groupby_distributed(k, [v], ("sum", "count"), True, 1, hint, None, mode)
        sm, ct_ = r["sum"][0], r["count"][0]
        if layout == "sharded":
            gk, sm, ct_ = D.allgather_ragged(gk)[0], D.allgather_ragged(sm)[0], D.allgather_ragged(ct_)[0]
            o = torch.argsort(gk)
            gk, sm, ct_ = gk[o], sm[o], ct_[o]
        wk, want = ogroupby.groupby_agg(keys, {"v": vals}, "sum", True)
        _, wcnt = ogroupby.groupby_agg(keys, {"v": vals}, "count", True)
        return (np.array_equal(gk.cpu().numpy(), wk) and np.allclose(sm.cpu().numpy(), want["v"], rtol=1e-9, atol=1e-12)
                and np.array_equal(ct_.cpu().numpy(), wcnt["v"]))
    res["groupby_combine_1k"] = gb(keys_lo, "combine", 1000)
    res["groupby_shuffle_150k"] = gb(keys_hi, "shuffle", 150_000)
    nb = 20_011
    build = rng.permutation(60_000)[:nb].astype(np.int64)
    probe = rng.integers(0, 60_000, n)
    bs, be = D.block_range(nb, rank, world)
    for mode in ("shuffle", "broadcast"):
        for how in ("inner", "left"):
            li, ri = D.join_distributed(torch.from_numpy(probe[s:e].copy()).to(dev), torch.from_numpy(build[bs:be].copy()).to(dev), how, None, mode)
            got = ojoin.canonical(D.allgather_ragged(li)[0].cpu().numpy(), D.allgather_ragged(ri)[0].cpu().numpy())
            want = ojoin.join_indexers(probe, build, how)
            res["join_%s_%s" % (mode, how)] = bool(np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]))
    if world > 1:
        # sort_values across GPUs (sample sort): the ranks' slices in rank order = the stable single-GPU order
        from oracle import sort as osort
        sk = rng.standard_normal(n).round(2)
        sk[rng.integers(0, n, n // 50)] = np.nan
        ok_all = True
        for asc, nap in ((True, "last"), (False, "first")):
            _, sid, _ = D.sort_values_distributed(torch.from_numpy(sk[s:e].copy()).to(dev), asc, nap)
            ok_all = ok_all and np.array_equal(D.allgather_ragged(sid)[0].cpu().numpy(), osort.sort_values_indexer(sk, asc, nap))
        res["sort_values_sample_sort"] = bool(ok_all)
    x = rng.random(n)
    x[::977] = np.nan
    for op in ("mean", "std", "max"):
        o = D.rolling_distributed(torch.from_numpy(x[s:e].copy()).to(dev), op, WINDOW, WINDOW, 1, None, None, None, link)
        g = D.allgather_ragged(o)[0].cpu().numpy()
        want = orolling.rolling(x, op, WINDOW, WINDOW, 1)
        ok = ~np.isnan(want)
        res["rolling_link_%s" % op] = bool(np.array_equal(np.isnan(g), np.isnan(want)) and np.allclose(g[ok], want[ok], rtol=1e-9, atol=1e-11))
    return res


def bench_c5(dev, rank, world, peak, reps=3):
    """C5 at `world` GPUs, 125 M rows per GPU (weak scaling): groupby-sum with K = 1 K (local aggregate + combine of
    the partial tables) and K = 100 M (hash shuffle + local aggregate), inner join 125 M probe x 12.5 M build rows per
    GPU (build side all-gathered / both sides shuffled).  Device time, MAX over ranks, median of `reps`.  Every case
    also times the single-GPU operator on this rank's own shard in the same run (`n1_ms`: what one GPU needs for its
    125 M rows when nothing has to move), so `efficiency` = n1_ms / ms is a same-run weak-scaling figure."""
    import torch
    import torch.distributed as dist
    from sdc_b200 import distributed as D
    from sdc_b200.groupby import groupby_device
    from sdc_b200.join import join_device
    n = C5_ROWS
    g = torch.Generator(device=dev).manual_seed(5 + rank)

    def timed(fn, warm=1):
        for _ in range(warm):
            fn()
        ts = []
        for _ in range(reps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ts.append(float(t.item()))
        return float(np.median(ts))

    out = {}
    vals = torch.rand(n, device=dev, dtype=torch.float64, generator=g)

    def case(name, workload, dist_fn, local_fn, shuffle_bytes=None, alg_bytes=None):
        d = {"workload": workload, "rows_per_gpu": n}
        n1 = timed(local_fn)
        d["n1_ms"] = n1
        if world > 1:
            ms = timed(dist_fn)
            d.update({"ms": ms, "rows_per_s": n * world / (ms * 1e-3), "efficiency": n1 / ms})
            if shuffle_bytes:
                d["shuffle_bytes_out_per_gpu"] = shuffle_bytes
                d["nvlink_frac"] = shuffle_bytes / (ms * 1e-3) / 1e9 / NVLINK_GBS
        else:
            d.update({"ms": n1, "rows_per_s": n / (n1 * 1e-3), "efficiency": 1.0})
        if alg_bytes:
            d["hbm_frac_n1"] = alg_bytes / (n1 * 1e-3) / 1e9 / peak
        out[name] = d

    keys = torch.randint(0, 1000, (n,), device=dev, generator=g)
    case("groupby_sum_k1k", "C5 groupby-sum, K = 1K groups: local aggregate, all-gather of the partial tables, combine",
         lambda: D.groupby_distributed(keys, [vals], ("sum",), True, 1, 1000, None, "combine"),
         lambda: groupby_device(keys, [vals], ("sum",), True, 1, 1000), None, 16 * n)
    del keys
    K = 100_000_000
    keys = torch.randint(0, K, (n,), device=dev, generator=g)
    case("groupby_sum_k100m", "C5 groupby-sum, K = 100M possible keys: shuffle of key + value to the key's owner (peer-memory "
         "stores over NVLink, dense-preserving routing), local aggregate",
         lambda: D.groupby_distributed(keys, [vals], ("sum",), True, 1, K, None, "shuffle"),
         lambda: groupby_device(keys, [vals], ("sum",), True, 1, K), n * 16 * (world - 1) / world, 16 * n)
    if world > 1:
        ms = timed(lambda: D.hash_shuffle(keys, [vals], None, "auto", "auto"))
        b = n * 16 * (world - 1) / world
        out["shuffle_key_value"] = {"workload": "the shuffle alone: 125M rows of key + value per GPU (count pass, counts all-gather, "
                                                "partition-and-send kernel, barrier)", "ms": ms, "shuffle_bytes_out_per_gpu": b,
                                    "nvlink_frac": b / (ms * 1e-3) / 1e9 / NVLINK_GBS}
    del keys
    # sort_values of a row-sharded float64 column (sample sort: splitters from one all-gather, range-routed shuffle of key +
    # global row id, local stable argsort + gathers) against the single-GPU sort_values of one shard
    from sdc_b200.sort import gather_device, sort_values_indexer_device

    def local_sort():
        idx = sort_values_indexer_device(vals, True, "last")
        res_s["o"] = (gather_device(vals, idx), idx)
    res_s = {}

    def dist_sort():
        res_s["o"] = D.sort_values_distributed(vals, True, "last")
    case("sort_values_f64", "Series.sort_values over 125M float64 rows per GPU (stable, NaN last): sample sort -- range-routed "
         "shuffle of key + global row id over NVLink, local onesweep argsort", dist_sort, local_sort,
         n * 16 * (world - 1) / world, 16 * n)
    res_s.clear()
    del vals
    nb = n // 10
    build = torch.randperm(nb * world, device=dev, generator=torch.Generator(device=dev).manual_seed(99))[rank * nb:(rank + 1) * nb].contiguous()
    probe = torch.randint(0, nb * world, (n,), device=dev, generator=g)
    lb = torch.randperm(nb, device=dev, generator=torch.Generator(device=dev).manual_seed(98))
    lp = torch.randint(0, nb, (n,), device=dev, generator=g)
    res = {}

    def dj(mode):
        def f():
            res.clear()
            res["o"] = D.join_distributed(probe, build, "inner", None, mode)
        return f

    def lj():
        res.clear()
        res["o"] = join_device(lp, lb, "inner")
    alg = 8 * (n + nb) + 16 * n
    case("join_inner_shuffle", "C5 inner join, 125M probe x 12.5M build rows per GPU, both sides shuffled to the key's owner with "
         "their global row ids (generated in the send kernel), local join, ids gathered",
         dj("shuffle"), lj, (n + nb) * 16 * (world - 1) / world, alg)
    if world > 1:
        out["join_inner_shuffle"]["n_out_rank0"] = int(res["o"][0].shape[0])
        ms = timed(dj("broadcast"))
        out["join_inner_broadcast"] = {"workload": "same join, build side all-gathered instead (every rank's table holds all %d build rows)" % (nb * world),
                                       "ms": ms, "rows_per_s": n * world / (ms * 1e-3), "n1_ms": out["join_inner_shuffle"]["n1_ms"],
                                       "efficiency": out["join_inner_shuffle"]["n1_ms"] / ms}
    res.clear()
    return out


# ------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from sdc_b200 import native
    from sdc_b200.rolling import rolling_device

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    link = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        from sdc_b200 import distributed as D
        link = D.HaloLink()
    lib = native.load()

    x_host = make_column(rank)
    col = torch.from_numpy(x_host).to(dev)
    out = torch.empty(N_ROWS, dtype=torch.float64, device=dev)

    def step():
        # row-sharded series: rank r's first windows need rank r-1's last win-1 rows.  The exchange is part
        # of the kernel: CTA 0 stores this rank's tail into the next rank's mailbox over NVLink, the tile that needs the
        # previous rank's rows runs last and finds them there (sdcb200_rolling_linked) -- one launch per step, no collective
        if link is not None:
            link.rolling(col, "mean", WINDOW, WINDOW, 1, rank * N_ROWS, out)
        else:
            rolling_device(col, "mean", WINDOW, WINDOW, 1, out=out, halo=None, row_offset=0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = native.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    for i in range(args.steps):
        step()
        ev[i + 1].record()
    barrier()
    launches = native.launch_count() - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    per_step = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * N_ROWS * args.steps / (total_ms * 1e-3)

    # the result of the timed steps against the oracle: this rank's first rows (they see the previous rank's tail
    # through the link) and a stretch from the middle of the shard
    from oracle import rolling as orolling
    head = 4096
    prev_tail = make_column(rank - 1)[N_ROWS - (WINDOW - 1):] if rank > 0 else np.empty(0)
    want_head = orolling.rolling(np.concatenate([prev_tail, x_host[:head]]), "mean", WINDOW, WINDOW, 1)[len(prev_tail):]
    got_head = out[:head].cpu().numpy()
    mid = N_ROWS // 2
    want_mid = orolling.rolling(x_host[mid - WINDOW + 1:mid + head], "mean", WINDOW, WINDOW, 1)[WINDOW - 1:]
    got_mid = out[mid:mid + head].cpu().numpy()
    ok_roll = bool(np.allclose(got_head, want_head, rtol=1e-9, atol=0, equal_nan=True) and np.allclose(got_mid, want_mid, rtol=1e-9, atol=0))

    # kernel-only duration for the roofline: back-to-back launches of the local kernel between two events (the launch
    # path overlaps the previous kernel, so this is the kernel's duration, not the call's)
    reps = 10
    rolling_device(col, "mean", WINDOW, WINDOW, 1, out=out, halo=None, row_offset=0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        rolling_device(col, "mean", WINDOW, WINDOW, 1, out=out, halo=None, row_offset=0)
    b.record()
    torch.cuda.synchronize()
    kern = a.elapsed_time(b) / reps
    peak, peak_src = measured_peak()
    achieved = BYTES_PER_ROW * N_ROWS / (kern * 1e-3) / 1e9

    # ---- e2e: host buffers through the C-ABI host entry point (pinned), copies inside the timed region
    import ctypes as ct
    nbytes = N_ROWS * 8
    pin_in, pin_out = ct.c_void_p(), ct.c_void_p()
    native.check(lib.sdcb200_host_alloc(ct.byref(pin_in), nbytes))
    native.check(lib.sdcb200_host_alloc(ct.byref(pin_out), nbytes))
    ct.memmove(pin_in.value, x_host.ctypes.data, nbytes)

    def e2e_step():
        native.check(lib.sdcb200_host_rolling(1, 9, pin_in.value, pin_out.value, N_ROWS, WINDOW, WINDOW, 1))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    res = np.ctypeslib.as_array(ct.cast(pin_out.value, ct.POINTER(ct.c_double)), shape=(N_ROWS,))
    check = float(np.nanmean(res[::1000]))
    e2e = {"value": world * N_ROWS * e2e_steps / float(tt.item()), "unit": "rows/s",
           "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes, "steps": e2e_steps,
           "api": "sdcb200_host_rolling (pinned host buffers, 3-stream chunk pipeline)",
           "result_check_mean": check}
    lib.sdcb200_host_free(pin_in)
    lib.sdcb200_host_free(pin_out)
    del col, out

    ops, parity = None, {"rolling_timed_steps": ok_roll}
    if not args.no_ops:
        if world > 1:
            parity.update(c5_parity(dev, rank, world, link))
        ops = {}
        if rank == 0 and world == 1:
            ops.update(bench_other_ops(dev, peak, cpu_legs=not args.no_cpu))
        if not args.no_c5:
            ops.update(bench_c5(dev, rank, world, peak))
    # a parity flag counts only when it holds on every rank
    names = sorted(parity)
    flags = torch.tensor([1 if parity[k] else 0 for k in names], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    parity = {k: bool(f) for k, f in zip(names, flags.tolist())}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        best, cores, layer = cpu_rolling_mean(x_host, 5)
        cpu = {"value": N_ROWS / best, "unit": "rows/s", "cores": cores, "kind": "port",
               "sample": "full workload (100M rows), best of 5", "threading_layer": layer}

    if rank == 0:
        traffic, traffic_src = ncu_traffic_bytes()
        line = {
            "metric": "rolling_mean_rows_per_s", "value": value, "unit": "rows/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": config_dict(),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "kernel": "rolling_run_kernel<MEAN,f64,R=17>", "kernel_ms": kern,
                         "kernel_ms_how": "10 back-to-back launches between two CUDA events / 10",
                         "algorithmic_bytes_per_launch": BYTES_PER_ROW * N_ROWS},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "ms_per_step_each": per_step,
            "halo": None if world == 1 else "peer-memory mailbox inside the rolling kernel (sdcb200_rolling_linked), no collective per step",
            "parity_ok": all(parity.values()), "parity": parity,
            "other_ops": ops,
        }
        print(json.dumps(line))
    if world > 1:
        link.close()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-ops", action="store_true", help="skip the groupby / join side measurements and the C5 parity run")
    ap.add_argument("--no-c5", action="store_true", help="skip the C5-shaped (125M rows per GPU) groupby / join cases")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
