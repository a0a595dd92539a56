#!/usr/bin/env python
"""Condense an `ncu --metrics ... --csv` launch list into a per-kernel table (markdown) for profiles/.

    python tools/launch_summary.py gpurun_out/launches_x.csv [profiles/r1_x_launches.md] [title]
"""
import csv
import sys
from collections import OrderedDict


def load(path):
    with open(path) as fh:
        lines = [l for l in fh if l.startswith('"')]
    per = OrderedDict()
    for row in csv.DictReader(lines):
        k = (row["ID"], row["Kernel Name"], row["Grid Size"], row["Block Size"])
        per.setdefault(k, {})[row["Metric Name"]] = float(row["Metric Value"].replace(",", ""))
    return per


def main():
    per = load(sys.argv[1])
    agg = OrderedDict()
    for (_id, name, grid, block), m in per.items():
        a = agg.setdefault((name, grid, block), {"n": 0, "ns": 0.0, "rd": 0.0, "wr": 0.0, "has_bytes": False})
        a["n"] += 1
        a["ns"] += m.get("gpu__time_duration.sum", 0.0)
        if "dram__bytes_read.sum" in m:
            a["has_bytes"] = True
            a["rd"] += m["dram__bytes_read.sum"]
            a["wr"] += m.get("dram__bytes_write.sum", 0.0)
    total = sum(a["ns"] for a in agg.values()) or 1.0
    out = []
    title = sys.argv[3] if len(sys.argv) > 3 else sys.argv[1]
    out.append("# %s\n" % title)
    out.append("ncu `--metrics gpu__time_duration.sum[,dram__bytes_*] --clock-control none` launch list; durations are "
               "cold-cache and serialised, so read the SHARE column, not the absolute.\n")
    out.append("| kernel | grid | block | launches | avg us | share | DRAM rd MB/launch | DRAM wr MB/launch | DRAM GB/s |")
    out.append("|---|---|---|---|---|---|---|---|---|")
    for (name, grid, block), a in agg.items():
        avg = a["ns"] / a["n"]
        if a["has_bytes"]:
            rd, wr = a["rd"] / a["n"] / 1e6, a["wr"] / a["n"] / 1e6
            bw = "%.0f" % ((a["rd"] + a["wr"]) / a["ns"])
            rd, wr = "%.1f" % rd, "%.1f" % wr
        else:
            rd = wr = bw = "-"
        short = name.replace("unnamed>::", "").replace("void ", "")
        short = short[:short.find("(")] if "(" in short else short
        out.append("| `%s` | %s | %s | %d | %.1f | %.1f%% | %s | %s | %s |" % (
            short, grid.strip("()").split(",")[0], block.strip("()").split(",")[0], a["n"], avg / 1e3,
            100 * a["ns"] / total, rd, wr, bw))
    text = "\n".join(out) + "\n"
    if len(sys.argv) > 2:
        with open(sys.argv[2], "w") as f:
            f.write(text)
    print(text)


if __name__ == "__main__":
    main()
