"""Throughput of the segment-length histogram on the config-2 track distribution.

    python tools/bench_seglen.py [n_tracks] [max_nb_states] [cpu_sample_tracks]

GPU: extrack_b200.histograms.len_hist (upload + k4_seglen + read-back), kernel time from CUDA events.
CPU: the numpy oracle port (oracle/seglen_oracle.py, one process) on a bounded sample of the same
tracks, and a parity check of the GPU result on that sample.  Prints one JSON line.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from extrack_b200 import histograms as xh  # noqa: E402
from extrack_b200 import tracking as xt  # noqa: E402
from extrack_b200.simulate import sim_tracks  # noqa: E402

n_tracks = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
max_nb = int(sys.argv[2]) if len(sys.argv) > 2 else 500
n_cpu = int(sys.argv[3]) if len(sys.argv) > 3 else 300
tracks = sim_tracks(n_tracks, seed=0, device="cuda", max_track_len=30, min_track_len=10, LocErr=0.02, Ds=[0, 0.25], nb_dims=2,
                    initial_fractions=[0.6, 0.4], TrMat=[[0.9, 0.1], [0.1, 0.9]], dt=0.02, pBL=0.05, cell_dims=[1, None, None])
st, keys = xt._sorted_buckets(tracks)
tr = {str(a.shape[1]): a for a in st}
params = xt.generate_params(nb_states=2, LocErr_type=1, nb_dims=2, LocErr_bounds=[0.005, 0.1], D_max=10, Fractions_bounds=[0.001, 0.99],
                            estimated_LocErr=[0.02], estimated_Ds=[1e-5, 0.25], estimated_Fs=[0.6, 0.4], estimated_transition_rates=0.1)
params["pBL"].value = 0.05
import contextlib, io  # noqa: E402
res = {}
for rep in range(3):
    tm = {}
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        h = xh.len_hist(tr, params, 0.02, cell_dims=[1.0], nb_states=2, max_nb_states=max_nb, _timing=tm)
    wall = time.perf_counter() - t0
    res = {"wall_s": wall, "kernel_ms": tm["kernel_ms"]}
locs = int(sum(a.shape[0] * a.shape[1] for a in st))
steps = int(sum(a.shape[0] * (a.shape[1] - 1) for a in st))
out = {"what": "len_hist (segment-length histogram), 2-state 2-D sim_FOV tracks of 10-30 localisations, max_nb_states=%d" % max_nb,
       "tracks": int(sum(len(a) for a in st)), "track_steps": steps, "kernel_ms": res["kernel_ms"], "wall_s_incl_upload": res["wall_s"],
       "tracks_per_s_kernel": sum(len(a) for a in st) / (res["kernel_ms"] * 1e-3),
       "track_steps_per_s_kernel": steps / (res["kernel_ms"] * 1e-3), "hist_sum": float(h.sum())}
if n_cpu > 0:
    from oracle import extrack_oracle as orc
    from oracle import seglen_oracle as so

    per = max(1, n_cpu // len(st))
    sample = [a[:per] for a in st]
    LocErr, ds, Fs, TrMat, pBL = xt.extract_params(params, 0.02, 2, 1)
    m = orc.Model(np.asarray(LocErr).reshape(-1), ds, Fs, TrMat, pBL, [1.0], 1, 6, int(st[0].shape[1]), 0.2, max_nb)
    t0 = time.perf_counter()
    ref = so.len_hist(sample, m, max_nb)
    cpu_s = time.perf_counter() - t0
    with contextlib.redirect_stdout(io.StringIO()):
        got = xh.len_hist({str(a.shape[1]): a for a in sample}, params, 0.02, cell_dims=[1.0], nb_states=2, max_nb_states=max_nb)
    csteps = int(sum(a.shape[0] * (a.shape[1] - 1) for a in sample))
    out["cpu_baseline"] = {"kind": "port", "cores": 1, "sample": "%d tracks (%d per length bucket), numpy oracle" % (per * len(st), per),
                           "secs": cpu_s, "track_steps_per_s": csteps / cpu_s}
    out["parity_max_abs_diff_on_sample"] = float(np.abs(got - ref).max())
    out["sample_hist_sum"] = float(ref.sum())
print(json.dumps(out))
