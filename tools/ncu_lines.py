#!/usr/bin/env python3
"""Per-source-line instruction / stall-sample shares of one kernel from an ncu report.

Joins `ncu --page source --csv` (per-SASS metrics, no line numbers) with `nvdisasm -g` line
info of the same cubin by instruction offset.  Usage:

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep k2_replay_lin '_Z13k2_replay_linILi2ELi1ELi4E' [min_pct]

The shared library must be the one the profile was taken with.
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter, defaultdict

rep, kregex, mangled_prefix = sys.argv[1], sys.argv[2], sys.argv[3]
min_pct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.7
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "extrack_b200", "libxtrack_b200.so")

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
# one cubin per translation unit: take the one that holds the wanted function
sass = ""
for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    if ".text." + mangled_prefix in txt:
        sass = txt
        break

# offset -> (file, line) for the wanted function
off2line = {}
infn = False
cur = ("?", 0)
for ln in sass.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m:
        infn = m.group(1).startswith(mangled_prefix)
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        off2line[int(m.group(1), 16)] = cur

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kregex], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = his[0]
end = his[1] - 1 if len(his) > 1 else len(rows)  # first captured launch only
hdr = rows[hi]
ie, iss, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
data = [r for r in rows[hi + 1 : end] if len(r) > ie and r[ie].isdigit()]
base = int(data[0][0], 16)
inst = defaultdict(int)
samp = defaultdict(int)
ops = defaultdict(Counter)
for r in data:
    off = int(r[0], 16) - base
    key = off2line.get(off, ("?", 0))
    inst[key] += int(r[ie])
    samp[key] += int(r[iss]) if r[iss].isdigit() else 0
    toks = r[isrc].split()
    op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
    ops[key][op] += int(r[ie])
ti, ts = sum(inst.values()), max(1, sum(samp.values()))
print(f"kernel {kregex}: {ti} warp-instructions, {ts} samples, {len(data)} SASS lines")
srcs = {}
for key in sorted(inst, key=lambda k: (k[0], k[1])):
    pi, ps = 100 * inst[key] / ti, 100 * samp[key] / ts
    if pi < min_pct and ps < min_pct:
        continue
    f, l = key
    if f not in srcs:
        for dp, _, fs in os.walk(root):
            if f in fs:
                srcs[f] = open(os.path.join(dp, f)).read().splitlines()
                break
        else:
            srcs[f] = []
    text = srcs[f][l - 1].strip() if 0 < l <= len(srcs[f]) else ""
    top = ",".join(f"{o}:{100*c/inst[key]:.0f}" for o, c in ops[key].most_common(4))
    print(f"{f[:20]:20s}:{l:4d} inst {pi:5.1f}% samp {ps:5.1f}%  [{top}]  {text[:80]}")
