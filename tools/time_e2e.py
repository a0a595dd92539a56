#!/usr/bin/env python
"""End-to-end rolling mean through sdcb200_host_rolling with pinned host buffers: zero-copy (kernel streams host
memory) vs the chunked H2D -> kernel -> D2H pipeline (SDCB200_HOST_ZEROCOPY=0).  Run once per setting."""
import ctypes as ct
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from sdc_b200 import native  # noqa: E402

n = 100_000_000
lib = native.load()
torch.cuda.init()
x = np.random.default_rng(1).random(n)
pin_in, pin_out = ct.c_void_p(), ct.c_void_p()
native.check(lib.sdcb200_host_alloc(ct.byref(pin_in), n * 8))
native.check(lib.sdcb200_host_alloc(ct.byref(pin_out), n * 8))
ct.memmove(pin_in.value, x.ctypes.data, n * 8)
for op, name in ((1, "mean"), (3, "std"), (5, "min")):
    native.check(lib.sdcb200_host_rolling(op, 9, pin_in.value, pin_out.value, n, 100, 100, 1))
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        native.check(lib.sdcb200_host_rolling(op, 9, pin_in.value, pin_out.value, n, 100, 100, 1))
        ts.append(time.perf_counter() - t0)
    res = np.ctypeslib.as_array(ct.cast(pin_out.value, ct.POINTER(ct.c_double)), shape=(n,))
    print("zerocopy=%s %s: best %.2f ms median %.2f ms -> %.2f Grows/s, checksum %.12f" % (
        os.environ.get("SDCB200_HOST_ZEROCOPY", "1"), name, min(ts) * 1e3, float(np.median(ts)) * 1e3, n / min(ts) / 1e9,
        float(np.nansum(res[::997]))))
