#!/usr/bin/env python3
"""Secondary measurements of BASELINE.json configs 3, 4 and 5 on ONE GPU (bounded sizes).

Not the bench line (bench.py measures configs[1]); these numbers go to profiles/ to show where
the other regimes of the same path stand: many live sequences (config 3), state annotation
(config 4), long 3-D tracks (config 5).  Prints one JSON line per config.

    python tools/bench_configs.py [--only 3,4,5] [--scale 1.0]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from extrack_b200 import _native  # noqa: E402
from extrack_b200 import tracking as xt  # noqa: E402
from extrack_b200._lmfit_compat import Parameters  # noqa: E402
from extrack_b200.simulate import sim_tracks  # noqa: E402


def params_from(LocErr, Ds, Fs, Tr, pBL):
    p = Parameters()
    p.add("LocErr", value=LocErr)
    nS = len(Ds)
    for i, D in enumerate(Ds):
        p.add(f"D{i}", value=max(D, 1e-5))
    for i, F in enumerate(Fs[:-1]):
        p.add(f"F{i}", value=F)
    p.add(f"F{nS-1}", expr="1-" + "-".join(f"F{i}" for i in range(nS - 1)))
    for i in range(nS):
        for j in range(nS):
            if i != j:
                p.add(f"p{i}{j}", value=max(Tr[i][j], 1e-4))
    p.add("pBL", value=pBL)
    return p


def time_eval(eng, p, reps):
    if os.environ.get("K2_GST_BELOW"):
        eng.set_option("k2_gst_below_ctas", int(os.environ["K2_GST_BELOW"]))
    eng.sum_logp(p)  # two-phase (sizes the launches)
    eng.sum_logp(p)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        v = eng.sum_logp(p)
    wall = (time.perf_counter() - t) / reps
    st = eng.stats()
    eng.set_option("pipeline", 0)
    eng.sum_logp(p)
    s2 = eng.stats()
    eng.set_option("pipeline", 1)
    return v, wall, st, s2


def likelihood_config(name, n_tracks, sim_kw, eval_kw, reps=5):
    t0 = time.perf_counter()
    tracks = sim_tracks(n_tracks, seed=4242, device="cuda:0", **sim_kw)
    st, _ = xt._sorted_buckets(tracks)
    gen = time.perf_counter() - t0
    d = st[0].shape[2]
    nS = len(sim_kw["Ds"])
    pr = params_from(sim_kw["LocErr"], sim_kw["Ds"], sim_kw["initial_fractions"], sim_kw["TrMat"], sim_kw["pBL"])
    LocErr, ds, Fs, TrMat, pBL = xt.extract_params(pr, sim_kw["dt"], nS, eval_kw["nb_substeps"])
    p = xt.build_tables(LocErr, ds, Fs, TrMat, pBL, [sim_kw["cell_dims"][0]], eval_kw["nb_substeps"], eval_kw["frame_len"],
                        st[0].shape[1], eval_kw["threshold"], eval_kw["max_nb_states"], d)
    ts = xt.TrackSet(st, rank=0, world_size=1, device=0)
    v, wall, s1, s2 = time_eval(ts.engine, p, reps)
    steps = s1["track_steps"]
    flops = s1["seq_updates"] * (25 + 9 * d) + s1["seq_groups"] * (3 + d)
    out = {"config": name, "tracks": int(s1["n_tracks"]), "track_steps": int(steps), "chunks": int(s1["n_chunks"]),
           "max_live_sequences": int(s1["max_nB_in"]), "sum_logp": v, "ms_per_eval": wall * 1e3,
           "track_steps_per_s": steps / wall, "pipelined": int(s1["pipelined"]),
           "two_phase_ms": {"plan": s2["ms_plan"], "replay": s2["ms_replay"]},
           "seq_updates_per_track_step": s1["seq_updates"] / steps, "algorithmic_flops": flops,
           "replay_tflops": flops / (s2["ms_replay"] * 1e-3) / 1e12, "generator_s": round(gen, 1), **eval_kw}
    ts.close()
    print(json.dumps(out), flush=True)


def predict_config(n_tracks, reps=3):
    sim_kw = dict(max_track_len=30, min_track_len=10, LocErr=0.02, Ds=[0, 0.25], nb_dims=2, initial_fractions=[0.6, 0.4],
                  TrMat=[[0.9, 0.1], [0.1, 0.9]], dt=0.02, pBL=0.05, cell_dims=[1, None, None])
    tracks = sim_tracks(n_tracks, seed=99, device="cuda:0", **sim_kw)
    st, _ = xt._sorted_buckets(tracks)
    pr = params_from(0.02, [1e-5, 0.25], [0.6, 0.4], sim_kw["TrMat"], 0.05)
    LocErr, ds, Fs, TrMat, pBL = xt.extract_params(pr, 0.02, 2, 1)
    p = xt.build_tables(LocErr, ds, Fs, TrMat, pBL, [1], 1, 8, st[0].shape[1], 0.1, 200, 2)
    eng = _native.Engine(0)
    eng.upload(st, [0 if a.shape[1] == st[-1].shape[1] else 1 for a in st], xt.MAX_TRACKS_PER_CHUNK)
    locs = sum(a.shape[0] * a.shape[1] for a in st)
    for kv in filter(None, os.environ.get("XT_OPTS", "").split(",")):
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    eng.predict(p, 2)
    if os.environ.get("K3_SWEEP"):  # kernel time vs resident CTAs per SM (per-warp scratch vs L2 capacity)
        for c in (1, 2, 3, 4, 6, 8):
            eng.set_option("k3_ctas_per_sm", c)
            eng.predict(p, 2)
            print(f"k3_ctas_per_sm={c}: kernel {eng.stats().get('ms_predict'):.1f} ms", file=sys.stderr, flush=True)
        eng.set_option("k3_ctas_per_sm", int(os.environ["K3_SWEEP"]))
    t = time.perf_counter()
    for _ in range(reps):
        out = eng.predict(p, 2)
    wall = (time.perf_counter() - t) / reps
    s = eng.stats()
    res = {"config": "4: predict_Bs 2-state fl=8 th=0.1 max_nb_states=200 nb_max=1", "tracks": n_tracks, "localisations": locs,
           "ms_per_call_incl_d2h": wall * 1e3, "localisations_per_s": locs / wall,
           "kernel_ms": s.get("ms_predict"), "kernel_launches": s.get("k3_launches"), "sequence_capacity": s.get("k3_cap"),
           "d2h_bytes": locs * 2 * 8,
           "mean_posterior_state0": float(np.mean([o[..., 0].mean() for o in out]))}
    eng.close()
    print(json.dumps(res), flush=True)


def fit_config_1():
    """BASELINE config 1: param_fitting on the tutorial tracks (golden fixture with the reference's fit)."""
    import contextlib
    import io

    z = np.load(os.path.join(ROOT, "tests", "golden", "fit_tracks_csv.npz"))
    tracks = {str(k): z["C" + k] for k in z["keys"]}
    want = dict(zip([str(n) for n in z["names"]], z["fitted"]))
    out = {}
    for rep in range(2):  # the second run has the library and the CUDA context warm
        params = xt.generate_params(nb_states=2, LocErr_type=1, nb_dims=2, LocErr_bounds=[0.005, 0.1], D_max=10,
                                    Fractions_bounds=[0.001, 0.99])
        buf = io.StringIO()
        t = time.perf_counter()
        with contextlib.redirect_stdout(buf):
            fit = xt.param_fitting(tracks, 0.02, params=params, nb_states=2, nb_substeps=1, frame_len=6, verbose=0, method="bfgs",
                                   cell_dims=[1], threshold=0.2, max_nb_states=120)
        wall = time.perf_counter() - t
        out = {"config": "1: 2-state param_fitting on Tutorials/tracks.csv (613 tracks, 33 buckets, frame_len 6, bfgs)",
               "fit_seconds": wall, "objective_evaluations": buf.getvalue().count(".") + buf.getvalue().count("x"),
               "neglogl": float(fit.residual[0]), "neglogl_reference_fit": float(z["neglogl"]),
               "max_rel_param_diff_vs_reference_fit": max(abs(fit.params[k].value - w) / max(abs(w), 1e-3) for k, w in want.items()),
               "reference_fit_seconds_single_core_build_container": 772.4}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="3,4,5")
    ap.add_argument("--scale", type=float, default=1.0)
    a = ap.parse_args()
    only = set(a.only.split(","))
    if "1" in only:
        fit_config_1()
    if "3" in only:
        likelihood_config(
            "3: 3-state 2D nb_substeps=2 frame_len=6 max_nb_states=500", int(100_000 * a.scale),
            dict(max_track_len=30, min_track_len=10, LocErr=0.02, Ds=[0, 0.04, 0.25], nb_dims=2,
                 initial_fractions=[0.33, 0.33, 0.34], TrMat=[[0.9, 0.1, 0.0], [0.05, 0.91, 0.04], [0.01, 0.06, 0.93]],
                 dt=0.02, pBL=0.05, cell_dims=[1, None, None]),
            dict(nb_substeps=2, frame_len=6, threshold=0.2, max_nb_states=500))
    if "3a" in only:
        likelihood_config(
            "3a: 3-state 2D nb_substeps=1 frame_len=6 max_nb_states=120", int(200_000 * a.scale),
            dict(max_track_len=30, min_track_len=10, LocErr=0.02, Ds=[0, 0.04, 0.25], nb_dims=2,
                 initial_fractions=[0.33, 0.33, 0.34], TrMat=[[0.9, 0.1, 0.0], [0.05, 0.91, 0.04], [0.01, 0.06, 0.93]],
                 dt=0.02, pBL=0.05, cell_dims=[1, None, None]),
            dict(nb_substeps=1, frame_len=6, threshold=0.2, max_nb_states=120))
    if "4" in only:
        predict_config(int(1_000_000 * a.scale))
    if "5" in only:
        likelihood_config(
            "5: 3-state 3D long tracks (100-200) frame_len=10 max_nb_states=120", int(20_000 * a.scale),
            dict(max_track_len=200, min_track_len=100, LocErr=0.02, Ds=[0, 0.05, 0.25], nb_dims=3,
                 initial_fractions=[0.3, 0.3, 0.4], TrMat=[[0.9, 0.05, 0.05], [0.05, 0.9, 0.05], [0.05, 0.05, 0.9]],
                 dt=0.02, pBL=0.002, cell_dims=[10, None, None]),
            dict(nb_substeps=1, frame_len=10, threshold=0.2, max_nb_states=120), reps=3)


if __name__ == "__main__":
    main()
