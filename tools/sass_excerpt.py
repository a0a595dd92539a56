#!/usr/bin/env python
"""SASS excerpts of the dominant kernels from the built objects (cuobjdump -sass): per kernel the instruction count, a
histogram of the mnemonics that matter for the design claims (bulk / TMA copies, mbarrier waits, shared and global
atomics / reductions, votes, 128-bit accesses) and the first lines of the listing.
Usage: python tools/sass_excerpt.py  -> profiles/r2_sass_<name>.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "sdc_b200", "lib", "obj")
WANT = [  # (object, regex on the demangled kernel name, output name)
    ("rolling.o", r"rolling_run_kernel<\(int\)1, double, \(int\)17>", "rolling_run_mean"),
    ("groupby.o", r"gb_smem_kernel<\(bool\)0, \(int\)1, \(bool\)1>", "groupby_smem_lean_sum"),
    ("groupby.o", r"gb_window_kernel", "groupby_window"),
    ("join.o", r"probe_unique_row1_kernel<\(bool\)0>", "join_probe_row1"),
    ("sort.o", r"onesweep_kernel<unsigned long, unsigned int, \(bool\)1, .*SortDigit<unsigned long, \(int\)1, \(bool\)0>, \(int\)256, \(int\)16, \(int\)3, \(bool\)0>", "sort_onesweep_i64_u32"),
    ("shuffle.o", r"shuffle_send_kernel<\(bool\)0, \(bool\)1>", "shuffle_send_mod"),
]
KEYS = ["UBLKCP", "SYNCS", "LDG.E.128", "LDG.E.64", "STG.E.128", "STG.E.64", "LDS.128", "LDS.64", "STS.128", "STS.64", "ATOMS", "ATOMG", "REDG", "REDS", "VOTE",
        "MATCH", "SHFL", "DADD", "DMUL", "DFMA", "MUFU", "BAR.SYNC", "POPC", "NANOSLEEP"]


def main():
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    for obj, pat, name in WANT:
        txt = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, obj)], capture_output=True, text=True).stdout
        dem = subprocess.run(["cu++filt"], input=txt, capture_output=True, text=True).stdout
        blocks = re.split(r"\n\s*Function : ", dem)
        hit = [b for b in blocks[1:] if re.search(pat, b.split("\n", 1)[0])]
        if not hit:
            print("no kernel matches", pat, file=sys.stderr)
            continue
        b = hit[0]
        lines = [l for l in b.split("\n") if re.search(r"/\*[0-9a-f]{4}\*/", l)]
        ops = collections.Counter()
        for l in lines:
            m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
            if m:
                for k in KEYS:
                    if m.group(1).startswith(k):
                        ops[k] += 1
        with open(os.path.join(ROOT, "profiles", "r2_sass_%s.txt" % name), "w") as f:
            f.write("# cuobjdump -sass %s | cu++filt, kernel: %s\n" % (obj, b.split("\n", 1)[0][:300]))
            f.write("# %d SASS instructions; mnemonic counts (static): %s\n" % (len(lines), ", ".join("%s %d" % (k, ops[k]) for k in KEYS if ops[k])))
            f.write("# first 160 instructions follow\n")
            f.write("\n".join(l.rstrip() for l in lines[:160]) + "\n")
        print(name, len(lines), dict(ops))


if __name__ == "__main__":
    main()
