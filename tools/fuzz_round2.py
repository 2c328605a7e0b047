"""Randomised parity sweep of the round-2 features on the GPU box (one-off confidence check next to the pytest suite):

* predict_Bs with nb_max > 1 (plans shared by the tracks of a chunk) against the numpy oracle,
* position refinement against its numpy oracle,
* plan verification: an engine that verifies the resident plan against an engine that plans every evaluation from
  scratch, bit for bit, along random walks through parameter space (small moves, line-search moves, jumps).

    python tools/fuzz_round2.py [seconds] [seed]
"""
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from extrack_b200 import refined_localization as rl  # noqa: E402
from extrack_b200 import tracking as xt  # noqa: E402
from helpers import engine_params, make_model, random_walk_tracks  # noqa: E402
from oracle import extrack_oracle as orc  # noqa: E402
from oracle import refine_oracle  # noqa: E402

KINDS = tuple(os.environ.get("FUZZ_KINDS", "predict,refine,verify,seglen,multi").split(","))
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 240.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
t_end = time.time() + budget
n = {"predict": 0, "refine": 0, "verify": 0, "seglen": 0, "multi": 0}
bad = 0
limit_hits = 0
worst = {"predict": 0.0, "refine": 0.0, "seglen": 0.0}
served = {"verified": 0, "replanned_chunks": 0, "scratch": 0}


def finding(kind, desc, what):
    global bad
    bad += 1
    print("FINDING", kind, desc, what, flush=True)


while time.time() < t_end:
    kind = KINDS[int(rng.integers(0, len(KINDS)))]
    nS = int(rng.choice([2, 2, 3]))
    d = int(rng.choice([1, 2, 2, 3]))
    fl = int(rng.integers(3, 9))
    Ds = np.sort(np.r_[1e-5, rng.random(nS - 1) * 0.4 + 0.02])
    th = float(rng.choice([0.05, 0.1, 0.2]))
    try:
        if kind == "predict":
            nb_max = int(rng.choice([2, 7, 30, 31, 64, 250]))
            tracks = {str(L): random_walk_tracks(int(rng.integers(1, 3 * nb_max + 2)), L, d, rng, Ds=Ds)
                      for L in sorted(set(int(x) for x in rng.integers(2, 26, size=3)))}
            desc = f"nS={nS} d={d} fl={fl} th={th} nb_max={nb_max} buckets={[(k, len(v)) for k, v in tracks.items()]}"
            m = make_model(nS=nS, frame_len=fl, threshold=th, max_nb_states=200, Ds=Ds, rates=float(rng.random() * 0.3 + 0.02),
                           pBL=float(rng.random() * 0.2 + 0.01), min_len=min(int(k) for k in tracks))
            params = xt.generate_params(nb_states=nS, LocErr_type=1, nb_dims=d, estimated_LocErr=[0.02], estimated_Ds=list(Ds),
                                        estimated_Fs=list(m.Fs[:-1]), estimated_transition_rates=0.1)
            LocErr, ds, Fs, TrMat, pBL = xt.extract_params(params, 0.02, nS, 1)
            keys = sorted(tracks, key=int)
            st = [tracks[k] for k in keys]
            mo = orc.Model(np.asarray(LocErr[0]).reshape(-1), ds, Fs, TrMat, pBL, [1], 1, fl, st[0].shape[1], th, 200)
            try:
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    want = orc.predict_states(st, mo, nb_max=nb_max)
            except ValueError:
                continue  # the reference's own grouping error path
            got = xt.predict_Bs(tracks, 0.02, params, cell_dims=[1], nb_states=nS, frame_len=fl, threshold=th, nb_max=nb_max)
            err = max(float(np.max(np.abs(got[k] - w))) for k, w in zip(keys, want))
            worst["predict"] = max(worst["predict"], err)
            if not err < 1e-6:
                finding(kind, desc, err)
        elif kind == "refine":
            tracks = {str(L): random_walk_tracks(int(rng.integers(1, 120)), L, d, rng, Ds=Ds)
                      for L in sorted(set(int(x) for x in rng.integers(2, 22, size=2)))}
            desc = f"nS={nS} d={d} fl={fl} th={th} buckets={[(k, len(v)) for k, v in tracks.items()]}"
            m = make_model(nS=nS, frame_len=fl, Ds=Ds, rates=float(rng.random() * 0.3 + 0.02))
            loc = 0.02 + 0.01 * float(rng.random())
            try:
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    want_mu, want_sig = refine_oracle.position_refinement(tracks, loc, m.ds, m.Fs, m.TrMat, fl, th, 1000)
            except ValueError:
                continue
            got_mu, got_sig = rl.position_refinement(tracks, loc, m.ds, m.Fs, m.TrMat, frame_len=fl, threshold=th, max_nb_states=1000)
            err = max(max(float(np.max(np.abs(got_mu[k] - want_mu[k]))), float(np.max(np.abs(got_sig[k] - want_sig[k])))) for k in tracks)
            worst["refine"] = max(worst["refine"], err)
            if not err < 1e-8:
                finding(kind, desc, err)
        elif kind == "seglen":
            from extrack_b200 import histograms as xh
            from oracle import seglen_oracle as so

            L, nT, mx = int(rng.integers(2, 31)), int(rng.integers(1, 51)), int(rng.choice([12, 40, 200, 500]))
            isBL, min_l = int(rng.integers(0, 2)), int(rng.integers(2, 8))
            per_dim = d > 1 and rng.integers(0, 3) == 0
            loc = tuple(0.02 + 0.01 * rng.random(d)) if per_dim else (0.02,)
            desc = f"nS={nS} d={d} L={L} nT={nT} mx={mx} isBL={isBL} min_l={min_l} loc={loc}"
            m = make_model(nS=nS, loc_err=loc, Ds=Ds, rates=float(rng.random() * 0.3 + 0.02), pBL=float(rng.random() * 0.2 + 0.01))
            C = random_walk_tracks(nT, L, d, rng, Ds=Ds)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                LP0, h0 = so.segment_len_chunk(C, m, isBL, mx, min_l)
            LP, Bs, hist = xh.P_segment_len(C, np.asarray(m.loc_err)[None, None], m.ds, m.Fs, m.TrMat, min_l=min_l, pBL=m.pBL, isBL=isBL,
                                            cell_dims=[1.0], nb_substeps=1, max_nb_states=mx)
            fin = np.isfinite(LP0)
            e1 = float(np.max(np.abs(LP[fin] - LP0[fin]) / np.maximum(np.abs(LP0[fin]), 1e-300))) if fin.any() else 0.0
            e2 = float(np.max(np.abs(hist - h0))) if h0.size else 0.0
            worst["seglen"] = max(worst["seglen"], e2)
            if not (e1 < 1e-9 and e2 < 1e-6 and LP.shape == LP0.shape):
                finding(kind, desc, (e1, e2))
        elif kind == "multi":
            st = [random_walk_tracks(int(rng.integers(1, 4200)), L, d, rng, Ds=Ds)
                  for L in sorted(set(int(x) for x in rng.integers(3, 30, size=int(rng.integers(1, 6)))))]
            nd = int(rng.integers(2, 6))
            desc = f"nS={nS} d={d} fl={fl} th={th} logical devices={nd} buckets={[a.shape[:2] for a in st]}"
            a, b = xt.TrackSet(st, 2000, devices=[0] * nd), xt.TrackSet(st, 2000)
            try:
                cur = dict(nS=nS, frame_len=fl, threshold=th, min_len=st[0].shape[1], Ds=Ds, loc_err=(0.02,),
                           rates=float(rng.random() * 0.3 + 0.02), pBL=float(rng.random() * 0.2 + 0.01))
                for step in range(int(rng.integers(3, 8))):
                    scale = float(rng.choice([1.5e-8, 1e-4, 3e-3, 0.2]))
                    cur = dict(cur)
                    cur["Ds"] = np.sort(np.asarray(cur["Ds"]) * (1 + scale * rng.standard_normal(nS)))
                    cur["loc_err"] = (cur["loc_err"][0] * (1 + scale * float(rng.standard_normal())),)
                    p = engine_params(make_model(**cur), d)
                    va, vb = a.sum_logp(p), b.sum_logp(p)
                    if not (va == vb or (np.isnan(va) and np.isnan(vb))):
                        finding(kind, desc, f"step {step} scale {scale}: {va!r} != {vb!r}")
                        break
            finally:
                a.close()
                b.close()
        else:
            st = [random_walk_tracks(int(rng.integers(1, 2600)), L, d, rng, Ds=Ds)
                  for L in sorted(set(int(x) for x in rng.integers(3, 30, size=int(rng.integers(1, 5)))))]
            desc = f"nS={nS} d={d} fl={fl} th={th} buckets={[a.shape[:2] for a in st]}"
            a, b = xt.TrackSet(st, 2000), xt.TrackSet(st, 2000)
            b.engine.set_option("plan_verify", 0)
            try:
                base = dict(nS=nS, frame_len=fl, threshold=th, min_len=st[0].shape[1], Ds=Ds, loc_err=(0.02,),
                            rates=float(rng.random() * 0.3 + 0.02), pBL=float(rng.random() * 0.2 + 0.01))
                cur = dict(base)
                for step in range(int(rng.integers(4, 12))):
                    scale = float(rng.choice([1.5e-8, 1.5e-8, 1e-4, 3e-3, 0.2]))
                    cur = dict(cur)
                    cur["Ds"] = np.sort(np.asarray(cur["Ds"]) * (1 + scale * rng.standard_normal(nS)))
                    cur["loc_err"] = (cur["loc_err"][0] * (1 + scale * float(rng.standard_normal())),)
                    cur["pBL"] = min(0.9, abs(cur["pBL"] * (1 + scale * float(rng.standard_normal()))))
                    p = engine_params(make_model(**cur), d)
                    va, vb = a.sum_logp(p), b.sum_logp(p)
                    s = a.engine.stats()
                    served["verified" if s["plan_verified"] else "scratch"] += 1
                    served["replanned_chunks"] += s["replanned"]
                    if not (va == vb or (np.isnan(va) and np.isnan(vb))):
                        finding(kind, desc, f"step {step} scale {scale}: {va!r} != {vb!r}")
                        break
                    c = int(rng.integers(0, len(a.chunks)))
                    _, c0, c1, _ = a.chunks[c]
                    if not np.array_equal(a.engine.chunk_logp(c, c1 - c0, p), b.engine.chunk_logp(c, c1 - c0, p), equal_nan=True):
                        finding(kind, desc, f"step {step}: per-track values of chunk {c} differ")
                        break
            finally:
                a.close()
                b.close()
        n[kind] += 1
    except ValueError as e:
        if "live state sequences" in str(e):  # the documented capacity limit of the engine, not a parity error
            limit_hits += 1
        else:
            finding(kind, "?", repr(e)[:300])
    except Exception as e:  # engine errors are findings too
        finding(kind, "?", repr(e)[:300])
print(f"fuzz round 2: cases {n}, {bad} findings, worst |posterior error| {worst['predict']:.2e}, worst |refined position / sigma error| "
      f"{worst['refine']:.2e} um, worst |segment-length histogram error| {worst['seglen']:.2e}, verification steps served {served}; cases that hit the documented sequence-capacity limit: {limit_hits}")
sys.exit(1 if bad else 0)
