#!/usr/bin/env python3
"""In-process multi-GPU strong scaling (xt_multi_*): ONE Python process drives 1, 2, 4, 8 GPUs of the box on the
same 10^6-track config-2 data set; wall time per objective evaluation through TrackSet.sum_logp (parameters change at
every call, BFGS finite-difference pattern) and the bitwise comparison with the one-GPU value.

    python tools/multi_probe.py [n_tracks]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

import bench  # noqa: E402
from extrack_b200 import _native, tracking as xt  # noqa: E402
from extrack_b200.simulate import sim_tracks  # noqa: E402

n_tracks = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
tracks = sim_tracks(n_tracks, seed=0, device="cuda:0", **bench.SIM_KW)
st, _ = xt._sorted_buckets(tracks)
pvar = bench.param_variants(st[0].shape[1])
ndev = _native.device_count()
steps = sum((a.shape[1] - 1) * a.shape[0] for a in st)
base = None
t1 = None
out = []
for G in (1, 2, 4, 8):
    if G > ndev:
        break
    ts = xt.TrackSet(st, devices=list(range(G)) if G > 1 else None, device=0)
    vals = [ts.sum_logp(p) for p in pvar]  # first round: plans are built
    if base is None:
        base = vals
    for _ in range(2):
        for p in pvar:
            ts.sum_logp(p)
    reps = 10
    t = time.perf_counter()
    for _ in range(reps):
        for p in pvar:
            v = ts.sum_logp(p)
    wall = (time.perf_counter() - t) / (reps * len(pvar)) * 1e3
    if t1 is None:
        t1 = wall
    load = ts.engine.device_load() if G > 1 else [(0, len(ts.chunks), steps)]
    row = {"gpus": G, "ms_per_eval": wall, "track_steps_per_s": steps / (wall * 1e-3), "efficiency_vs_one_gpu": t1 / (G * wall),
           "bitwise_equal_to_one_gpu": [ts.sum_logp(p) for p in pvar] == base, "chunks_per_device": [n for _, n, _ in load],
           "stats": {k: ts.engine.stats()[k] for k in ("plan_verified", "replanned")}}
    print(json.dumps(row), flush=True)
    out.append(row)
    ts.close()
