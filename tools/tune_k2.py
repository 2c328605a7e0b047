"""A/B timing of replay-kernel configurations on the config-2 data set (GPU box only).

    python tools/tune_k2.py [n_tracks]            # current build
    XT_LIB_PATH=... python tools/tune_k2.py       # another build of the engine

Prints CUDA-event times of the two-phase evaluation (plan, replay) for warps-per-tile /
tracks-per-thread variants, and the wall time of the pipelined evaluation per group count.
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
tracks = sim_tracks(n_tracks, seed=0, device="cuda", max_track_len=30, min_track_len=10, LocErr=0.02, Ds=[0, 0.25], nb_dims=2,
                    initial_fractions=[0.6, 0.4], TrMat=[[0.9, 0.1], [0.1, 0.9]], dt=0.02, pBL=0.05, cell_dims=[1, None, None])
st, _ = xt._sorted_buckets(tracks)
model = make_model(nS=2, nsub=1, frame_len=8, min_len=st[0].shape[1], Ds=[1e-5, 0.25], Fs=[0.6, 0.4])
p = engine_params(model, 2)
ts = xt.TrackSet(st)
eng = ts.engine
print("lib", os.environ.get("XT_LIB_PATH", "default"), "chunks", len(ts.chunks))


def two_phase(tag, reps=6):
    eng.set_option("pipeline", 0)
    pl, rp = [], []
    v = None
    for _ in range(reps):
        v = ts.sum_logp(p)
        s = eng.stats()
        pl.append(s["ms_plan"])
        rp.append(s["ms_replay"])
    print(f"{tag:28s} plan {np.median(pl[1:]):.4f} ms  replay {np.median(rp[1:]):.4f} ms  sum {v!r}")
    return v


def pipelined(tag, reps=30):
    eng.set_option("pipeline", 1)
    for _ in range(3):
        v = ts.sum_logp(p)
    t = time.perf_counter()
    for _ in range(reps):
        v = ts.sum_logp(p)
    dtm = (time.perf_counter() - t) / reps
    print(f"{tag:28s} wall {dtm*1e3:.4f} ms/eval -> {eng.stats()['track_steps']/dtm/1e9:.3f} G track-steps/s  sum {v!r}")
    return v


ref = two_phase("wpc=4 tpt=1")
if os.environ.get("TUNE_GST"):
    eng.set_option("k2_gst_below_ctas", 99)
    v = two_phase("wpc=4 tpt=1 state in global memory")
    print("   rel diff vs shared-memory state", abs(v - ref) / abs(ref))
    eng.set_option("k2_gst_below_ctas", 3)
if os.environ.get("TUNE_FP32"):
    eng.set_option("fp32_replay", 1)
    v = two_phase("fp32 replay (FP64 plan)")
    print("   fp32 kernel used:", eng.stats()["fp32"], " rel diff vs FP64", abs(v - ref) / abs(ref))
    eng.set_option("n_groups", 6)
    pipelined("fp32 pipelined G=6")
    eng.set_option("fp32_replay", 0)
if os.environ.get("TUNE_VARIANTS"):
    for wpc, tpt in ((4, 2), (8, 1), (8, 2), (2, 1), (2, 2)):
        eng.set_option("k2_wpc", wpc)
        eng.set_option("k2_tpt", tpt)
        v = two_phase(f"wpc={wpc} tpt={tpt}")
        if abs(v - ref) > 1e-11 * abs(ref):
            print("   MISMATCH", v, ref)
    eng.set_option("k2_wpc", 4)
    eng.set_option("k2_tpt", 1)
for ns in ((8,) if os.environ.get("TUNE_QUICK") else (4, 8, 16)):
    try:
        eng.set_option("n_streams", ns)
    except Exception:
        if ns != 4:
            continue
    for g in ((6,) if os.environ.get("TUNE_QUICK") else (2, 3, 4, 6, 8, 12, 16)):
        if g > ns * 2:
            continue
        eng.set_option("n_groups", g)
        pipelined(f"pipelined streams={ns} G={g}")
ts.close()
