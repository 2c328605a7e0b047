#!/usr/bin/env python3
"""profiles/k2_traffic.json (read by bench.py for roofline.traffic) from an `ncu --set full` capture of the bench:
DRAM bytes read + written by one whole-data-set launch of the FP64 replay kernel.

    python tools/make_k2_traffic.py gpurun_out/r12/prof_k12.ncu-rep 1000000 profiles/r12/ncu_full_k1_k2.json
"""
import csv
import io
import json
import os
import subprocess
import sys

rep, tracks, source = sys.argv[1], int(sys.argv[2]), sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
for r in rows[2:]:
    name = r[h.index("Kernel Name")]
    if "k2_replay_fused" not in name:
        continue
    def val(k):
        i = h.index(k)
        v, u = float(r[i]), units[i].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    out = {"kernel": name, "tracks": tracks, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
           "source": source, "how": "dram__bytes_read.sum + dram__bytes_write.sum of one whole-data-set launch, ncu --set full --clock-control none"}
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "k2_traffic.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out))
    break
