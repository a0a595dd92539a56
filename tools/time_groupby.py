#!/usr/bin/env python
"""Where does a C1 groupby call spend its time?  wall clock per call (host, with sync), CUDA-event time,
and the bare C-ABI call (no Python result wrapping)."""
import ctypes as ct
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from sdc_b200 import native  # noqa: E402
from sdc_b200.groupby import groupby_device  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
rng = np.random.default_rng(0)
a = torch.from_numpy(rng.integers(0, 1000, n, dtype=np.int64)).cuda()
b = torch.from_numpy(rng.random(n)).cuda()
lib = native.load()
for _ in range(5):
    groupby_device(a, [b], ("sum",), True, 1, 1000)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200):
    groupby_device(a, [b], ("sum",), True, 1, 1000)
torch.cuda.synchronize()
print("python api wall us/call:", (time.perf_counter() - t0) / 200 * 1e6)

ptrs = (ct.c_void_p * 1)(b.data_ptr())
dts = (ct.c_int32 * 1)(native.DTYPE_CODES["float64"])
h, ng = ct.c_int64(0), ct.c_int64(0)
stream = native.current_stream_ptr()


def ccall():
    native.check(lib.sdcb200_groupby(native.DTYPE_CODES["int64"], a.data_ptr(), n, 1, ptrs, dts, 1, 1, 1, 1000, stream,
                                     ct.byref(h), ct.byref(ng)))
    lib.sdcb200_groupby_free(h.value)


for _ in range(5):
    ccall()
t0 = time.perf_counter()
for _ in range(200):
    ccall()
print("C call wall us/call:", (time.perf_counter() - t0) / 200 * 1e6, "groups", ng.value)
ts = []
for _ in range(50):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ccall()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print("C call event us median:", float(np.median(ts)), "min", min(ts))

# the wrapper's own share: handle -> tensors (two torch.as_tensor over __cuda_array_interface__ views)
import sdc_b200.groupby as G  # noqa: E402
kp, vp, stride = ct.c_void_p(), ct.c_void_p(), ct.c_int64(0)
native.check(lib.sdcb200_groupby(native.DTYPE_CODES["int64"], a.data_ptr(), n, 1, ptrs, dts, 1, 1, 1, 1000, stream,
                                 ct.byref(h), ct.byref(ng)))
native.check(lib.sdcb200_groupby_result(h.value, ct.byref(kp), ct.byref(vp), ct.byref(stride)))
owner = object()
t0 = time.perf_counter()
for _ in range(200):
    x = torch.as_tensor(G._DeviceView(kp.value, (ng.value,), "<i8", owner), device=a.device)
print("torch.as_tensor(cuda array interface) us:", (time.perf_counter() - t0) / 200 * 1e6)
lib.sdcb200_groupby_free(h.value)
