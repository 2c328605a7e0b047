#!/usr/bin/env python3
"""Strong-scaling probe on ONE GPU: time the shard that rank 0 of a world of N = 1, 2, 4, 8 would own
(the same 10^6-track config-2 data set split by `shard_chunks`), two-phase (plan / replay kernel
times) and in the default evaluation mode.  The all-reduce of the real N-GPU run is not included.

    python tools/strong_probe.py [n_tracks] [worlds, e.g. 1,2,4,8]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

from extrack_b200 import tracking as xt  # noqa: E402
from extrack_b200.simulate import sim_tracks  # noqa: E402
from helpers import engine_params, make_model  # noqa: E402

n_tracks = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
worlds = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,2,4,8").split(",")]
tracks = sim_tracks(n_tracks, seed=0, device="cuda", max_track_len=30, min_track_len=10, LocErr=0.02, Ds=[0, 0.25], nb_dims=2,
                    initial_fractions=[0.6, 0.4], TrMat=[[0.9, 0.1], [0.1, 0.9]], dt=0.02, pBL=0.05, cell_dims=[1, None, None])
st, _ = xt._sorted_buckets(tracks)
model = make_model(nS=2, nsub=1, frame_len=8, min_len=st[0].shape[1], Ds=[1e-5, 0.25], Fs=[0.6, 0.4])
p = engine_params(model, 2)
print("lib", os.environ.get("XT_LIB_PATH", "default"))
t1 = None
for W in worlds:
    ts = xt.TrackSet(st, rank=0, world_size=W)
    eng = ts.engine
    for name, val in [kv.split("=") for kv in os.environ.get("XT_OPTS", "").split(",") if kv]:
        eng.set_option(name, int(val))
    eng.set_option("pipeline", 0)
    pl, rp = [], []
    for _ in range(6):
        v = ts.engine.sum_logp(p)
        s = eng.stats()
        pl.append(s["ms_plan"])
        rp.append(s["ms_replay"])
    eng.set_option("pipeline", 1)
    for _ in range(5):
        v2 = eng.sum_logp(p)
    reps = 50
    t = time.perf_counter()
    for _ in range(reps):
        v2 = eng.sum_logp(p)
    wall = (time.perf_counter() - t) / reps * 1e3
    if W == 1:
        t1 = wall
    eff = t1 / (W * wall) if t1 else float("nan")
    print(f"world {W}: chunks {ts.n_local_chunks:4d} track-steps {s['track_steps']:9d}  plan {np.median(pl[1:]):.4f} ms  replay {np.median(rp[1:]):.4f} ms"
          f"  default-mode wall {wall:.4f} ms/eval  strong-scaling efficiency (compute only) {eff:.3f}  sum {v!r} {v2!r}")
    ts.close()
