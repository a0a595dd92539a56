#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a small JSON for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_x_ncu_summary.json [kernel-regex]
"""
import csv
import io
import json
import re
import subprocess
import sys

KEEP = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid_size",
    "launch__block_size": "block_size",
    "launch__shared_mem_per_block_dynamic": "dynamic_smem_per_block",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio": "stall_mio_throttle",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio": "stall_lg_throttle",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
}
UNIT = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    pat = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    kernels = []
    for r in rows[2:]:
        if pat and not pat.search(r[kn]):
            continue
        d = {"kernel": r[kn]}
        for m, name in KEEP.items():
            if m in hdr:
                i = hdr.index(m)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                u = units[i]
                if name in ("duration", "dram_bytes_read", "dram_bytes_write") and u in UNIT:
                    v *= UNIT[u]
                d[name] = v
        if "duration" in d and "dram_bytes_read" in d:
            d["dram_GBps"] = (d["dram_bytes_read"] + d["dram_bytes_write"]) / d["duration"] / 1e9
        kernels.append(d)
    res = kernels[0] if len(kernels) == 1 else {"kernels": kernels}
    res["source"] = rep
    res["note"] = "ncu --set full --clock-control none; durations under ncu are cold-cache and serialised"
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res)[:600])


if __name__ == "__main__":
    main()
