#!/usr/bin/env python
"""bench.py -- headline benchmark of the sdc_b200 hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Headline workload (BASELINE.json configs[1], "C2"): Series.rolling(window=100).mean() on a
100 M-row float64 column per GPU.  A "step" is one pass of the rolling kernel over the column.
Prints ONE JSON line (see the contract in DESIGN.md, "Measurement").
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
        "config": {"workload": WORKLOAD, "rows": N_ROWS, "window": WINDOW},
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": cores, "kind": "port", "sample": sample,
                         "threading_layer": numba.threading_layer()},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the rolling kernel, from the committed
    `ncu --set full` summary (profiles/); None when no capture of the current kernel is on file."""
    p = os.path.join(ROOT, "profiles", "r1_rolling_mean_ncu_summary.json")
    try:
        with open(p) as f:
            d = json.load(f)
        return float(d["dram_bytes_read"]) + float(d["dram_bytes_write"])
    except Exception:
        return None


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
    from sdc_b200.groupby import groupby_device
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
          "achieved_GBps": 16 * n / (med * 1e-3) / 1e9, "frac_of_measured_peak": 16 * n / (med * 1e-3) / 1e9 / peak}
    n2 = 100_000_000
    a2 = torch.from_numpy(rng.integers(0, 1000, n2, dtype=np.int64)).to(dev)
    b2 = torch.from_numpy(rng.random(n2)).to(dev)
    med2, _ = _event_times(lambda: groupby_device(a2, [b2], ("sum",), True, 1, 1000), 10)
    gb["rows_per_s_100M"] = n2 / (med2 * 1e-3)
    gb["frac_of_measured_peak_100M"] = 16 * n2 / (med2 * 1e-3) / 1e9 / peak
    del a2, b2
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
        gb["cpu_baseline"] = {"value": ns / dt, "unit": "rows/s", "cores": numba.config.NUMBA_NUM_THREADS, "kind": "port",
                              "sample": "first 2M rows of the workload, one call after a compile call",
                              "pandas_rows_per_s": n / dtp}
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
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = native.load()

    x_host = make_column(rank)
    col = torch.from_numpy(x_host).to(dev)
    out = torch.empty(N_ROWS, dtype=torch.float64, device=dev)
    halo_n = WINDOW - 1
    halo = torch.empty(halo_n, dtype=torch.float64, device=dev) if world > 1 else None

    def step():
        # row-sharded series: rank r's window needs rank r-1's last win-1 rows
        h = None
        if world > 1:
            ops = []
            if rank + 1 < world:
                ops.append(dist.P2POp(dist.isend, col[N_ROWS - halo_n:], rank + 1))
            if rank > 0:
                ops.append(dist.P2POp(dist.irecv, halo, rank - 1))
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            h = halo if rank > 0 else None
        rolling_device(col, "mean", WINDOW, WINDOW, 1, out=out, halo=h, row_offset=rank * N_ROWS)

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

    # kernel-only duration for the roofline: events directly around the launch (N=1: == step)
    kern_ms = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rolling_device(col, "mean", WINDOW, WINDOW, 1, out=out, halo=None, row_offset=0)
        b.record()
        torch.cuda.synchronize()
        kern_ms.append(a.elapsed_time(b))
    kern = float(np.mean(kern_ms))
    peak, peak_src = measured_peak()
    achieved = BYTES_PER_ROW * N_ROWS / (kern * 1e-3) / 1e9

    # ---- e2e: host buffers through the C-ABI host entry point (pinned), copies inside the timed region
    e2e = None
    if True:
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

    ops = None
    if rank == 0 and world == 1 and not args.no_ops:
        ops = bench_other_ops(dev, peak, cpu_legs=not args.no_cpu)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        best, cores, layer = cpu_rolling_mean(x_host, 5)
        cpu = {"value": N_ROWS / best, "unit": "rows/s", "cores": cores, "kind": "port",
               "sample": "full workload (100M rows), best of 5", "threading_layer": layer}

    if rank == 0:
        line = {
            "metric": "rolling_mean_rows_per_s", "value": value, "unit": "rows/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "rows_per_gpu": N_ROWS, "window": WINDOW,
                       "l2": "inputs (800 MB in + 800 MB out per step) larger than L2",
                       "multi_gpu": "row-sharded series, win-1 halo rows from the previous rank (NCCL send/recv)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic_bytes(), "peak_source": peak_src,
                         "kernel": "rolling_run_kernel<MEAN,f64,R=17>", "kernel_ms": kern,
                         "algorithmic_bytes_per_launch": BYTES_PER_ROW * N_ROWS},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "ms_per_step_each": per_step,
            "other_ops": ops,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-ops", action="store_true", help="skip the groupby (C1) / join (C3) side measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
