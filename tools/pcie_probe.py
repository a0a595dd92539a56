#!/usr/bin/env python
"""Host<->device copy bandwidth of the box (pinned memory), one direction and both at once: the
ceiling of bench.py's e2e leg."""
import json
import torch

n = 800_000_000
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def t(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        torch.cuda.synchronize()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def h2d():
    d_in.copy_(h_in, non_blocking=True)


def d2h():
    h_out.copy_(d_out, non_blocking=True)


def both():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def chunks(k):
    c = n // k
    def f():
        for i in range(k):
            with torch.cuda.stream(s1):
                d_in[i * c:(i + 1) * c].copy_(h_in[i * c:(i + 1) * c], non_blocking=True)
            with torch.cuda.stream(s2):
                h_out[i * c:(i + 1) * c].copy_(d_out[i * c:(i + 1) * c], non_blocking=True)
    return f


r = {"h2d_GBps": n / t(h2d) / 1e6, "d2h_GBps": n / t(d2h) / 1e6, "duplex_ms_800MB_each": t(both),
     "duplex_16chunks_ms": t(chunks(16)), "duplex_64chunks_ms": t(chunks(64))}
print(json.dumps(r))
