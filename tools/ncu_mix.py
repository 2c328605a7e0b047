#!/usr/bin/env python3
"""Instruction mix and warp-stall shares of one kernel from `ncu --page source --csv` output.

    ncu -i prof.ncu-rep --page source --csv --kernel-name regex:k2_replay_lin > src.csv
    python tools/ncu_mix.py src.csv
"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if r and r[0] == "Address")
H = {h: i for i, h in enumerate(hdr)}
ops, samp, stall = collections.Counter(), collections.Counter(), collections.Counter()
stallcols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = 0
launches = sum(1 for r in rows if r and r[0] == "Kernel Name")
for r in rows:
    if len(r) < len(hdr) or not r[0].startswith("0x"):
        continue
    src = r[H["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2).split(".")[0] if m else src
    n = int(r[H["Instructions Executed"]])
    ops[op] += n
    tot += n
    samp[op] += int(r[H["# Samples"]])
    for c in stallcols:
        stall[c] += int(r[H[c]])
print(f"launches in file: {launches}; warp instructions executed (all launches): {tot}")
for op, n in ops.most_common(28):
    print(f"{op:12s} {n:12d} {100*n/tot:5.1f}%  samples {samp[op]}")
ts = sum(stall.values())
for c, n in stall.most_common(12):
    print(f"{c:28s} {n:8d} {100*n/ts:5.1f}%")
