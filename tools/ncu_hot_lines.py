#!/usr/bin/env python
"""Hottest source lines of a kernel in an .ncu-rep (warp-stall samples per CUDA source line).

    python tools/ncu_hot_lines.py gpurun_out/prof.ncu-rep [top=30]
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur_file = ""
hdr = None
lines = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].strip().isdigit():
        continue
    d = dict(zip(hdr[4:], r[4:]))
    try:
        samples = int(d.get("# Samples", "0") or 0)
    except ValueError:
        samples = 0
    if samples:
        stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "(Not Issued)" not in k and v.isdigit() and int(v)}
        lines.append((samples, cur_file, int(r[0]), r[1].strip()[:90], int(d.get("Instructions Executed", "0") or 0), stalls))
tot = sum(x[0] for x in lines) or 1
lines.sort(key=lambda x: x[0], reverse=True)
print("total samples", tot)
for s, f, ln, src, inst, stalls in lines[:top]:
    st = ",".join("%s=%d" % kv for kv in sorted(stalls.items(), key=lambda kv: -kv[1])[:3])
    print("%5.1f%% %-12s:%-4d inst=%-9d %-40s | %s" % (100.0 * s / tot, f, ln, inst, st, src))
