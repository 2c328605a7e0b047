#!/usr/bin/env python3
"""Where the wall time of one predict_Bs call goes (host side), on 1/8 of config 4:  python tools/predict_api_prof.py [n_tracks]"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from extrack_b200 import tracking as xt  # noqa: E402
from extrack_b200.simulate import sim_tracks  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 125_000
tracks = sim_tracks(n, seed=99, device="cuda:0", **bench.SIM_KW)
params = bench.eval_params()
kw = dict(cell_dims=bench.CELL, nb_states=2, frame_len=bench.FRAME_LEN, gather=False)
xt.predict_Bs({k: v[:64] for k, v in tracks.items()}, bench.DT, params, **kw)
for _ in range(3):
    t = time.perf_counter()
    xt.predict_Bs(tracks, bench.DT, params, **kw)
    print(f"predict_Bs({n} tracks): {(time.perf_counter() - t) * 1e3:.1f} ms")
pr = cProfile.Profile()
pr.enable()
xt.predict_Bs(tracks, bench.DT, params, **kw)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
