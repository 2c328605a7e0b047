#!/usr/bin/env python3
"""Per-phase cycle counts of the plan kernel (debug build: nvcc -DXT_K1_PROF).  Run on the GPU box:

    python tools/k1_phase_prof.py [n_tracks]
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from extrack_b200 import _native, tracking as xt  # noqa: E402
from extrack_b200.simulate import sim_tracks  # noqa: E402
from helpers import engine_params, make_model  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
cfg = os.environ.get("CFG", "2")
if cfg == "5":  # BASELINE config 5: long 3-D tracks, 3 states (fewer chunks than SMs, hundreds of steps each)
    import bench

    c5 = bench.SECONDARY["5"]
    tracks = sim_tracks(n, seed=4242, device="cuda", **c5["sim"])
    st, _ = xt._sorted_buckets(tracks)
    LocErr, ds, Fs, TrMat, pBL = xt.extract_params(bench._params_for(c5["sim"]), c5["sim"]["dt"], 3, 1)
    ev = c5["ev"]
    p = xt.build_tables(LocErr, ds, Fs, TrMat, pBL, [c5["sim"]["cell_dims"][0]], 1, ev["frame_len"], st[0].shape[1], ev["threshold"],
                        ev["max_nb_states"], 3)
else:
    tracks = sim_tracks(n, seed=0, device="cuda", max_track_len=30, min_track_len=10, LocErr=0.02, Ds=[0, 0.25], nb_dims=2,
                        initial_fractions=[0.6, 0.4], TrMat=[[0.9, 0.1], [0.1, 0.9]], dt=0.02, pBL=0.05, cell_dims=[1, None, None])
    st, _ = xt._sorted_buckets(tracks)
    model = make_model(nS=2, nsub=1, frame_len=8, min_len=st[0].shape[1], Ds=[1e-5, 0.25], Fs=[0.6, 0.4])
    p = engine_params(model, 2)
ts = xt.TrackSet(st, rank=0, world_size=int(os.environ.get("WORLD", "1")))
for name, val in [kv.split("=") for kv in os.environ.get("XT_OPTS", "").split(",") if kv]:
    ts.engine.set_option(name, int(val))
ts.engine.set_option("pipeline", int(os.environ.get("PIPE", "0")))
for _ in range(3):
    ts.engine.sum_logp(p)
print(ts.engine.stats())
lib = _native.load_library()
nch = ts.n_local_chunks
buf = np.zeros((nch, 12), dtype=np.int64)
assert lib.xt_debug_k1_prof(buf.ctypes.data_as(C.c_void_p), nch) == 0
names = ["update+codes", "batch rows", "batch resolve", "copy+zero+barriers", "history (thread 0)", "records (thread 0)", "merge (warp 0)", "end barrier"]
Ls = np.array([st[ts.chunks[i][0]].shape[1] for i in ts.my_chunks])
for L in ((100, 150, 200) if cfg == "5" else (10, 20, 30)):
    sel = buf[Ls == L]
    if len(sel) == 0:
        continue
    m = sel.mean(0)
    print(f"L={L}: {len(sel)} chunks, total {m[:8].sum():.0f} cycles = {m[:8].sum()/1.965e3:.1f} us; per fused step {m[:8].sum()/(L-3):.0f}")
    tot = m[:8].sum()
    for nm, v in zip(names, m[:8]):
        print(f"   {nm:22s} {v:10.0f}  {100*v/tot:5.1f}%  per step {v/(L-3):8.0f}")
    print(f"   history: slowest thread {m[8]/(L-3):.0f} cycles/step; member visits {m[10]/(L-3):.0f}/step of which slow path {m[9]/(L-3):.0f}; items {m[11]/(L-3):.0f}/step")
