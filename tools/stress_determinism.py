#!/usr/bin/env python3
"""Stress test (GPU box): the objective must have the same bits
  * on an engine that verifies its resident plan (default) and on one that plans every evaluation from scratch,
  * on repeated calls, and
  * on several contexts (logical shards on one GPU),
along a random walk of parameter sets (tiny BFGS-like perturbations, moderate moves, jumps, changes of frame_len).
Prints every mismatch and how the evaluations were served (verified / chunks planned again / planned from scratch)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from extrack_b200 import tracking as xt  # noqa: E402
from helpers import engine_params, make_model, random_walk_tracks  # noqa: E402

rng = np.random.default_rng(5)
st = [random_walk_tracks(n, L, 2, rng) for L, n in ((7, 2300), (11, 4100), (16, 2600), (24, 900))]
one = xt.TrackSet(st, 2000)
ref = xt.TrackSet(st, 2000)
ref.engine.set_option("plan_verify", 0)
many = xt.TrackSet(st, 2000, devices=[0, 0, 0])
n_eval = int(sys.argv[1]) if len(sys.argv) > 1 else 300
bad = 0
served = {"verified": 0, "verified, some chunks planned again": 0, "planned from scratch": 0}
replanned = 0
par = dict(D0=1e-4, D1=0.2, le=0.02, F0=0.5, rate=0.1, pBL=0.1, fl=6)
for it in range(n_eval):
    u = rng.random()
    if u < 0.6:      # finite-difference-like perturbation of one parameter
        k = ["D0", "D1", "le", "F0", "rate", "pBL"][int(rng.integers(0, 6))]
        par[k] *= 1.0 + float(rng.choice([1.5e-8, -1.5e-8, 1e-6, 1e-4]))
    elif u < 0.85:   # line-search-like move of all parameters
        for k in ("D0", "D1", "le", "F0", "rate", "pBL"):
            par[k] *= float(np.exp(rng.normal(0, 0.03)))
    elif u < 0.95:   # jump
        par.update(D0=float(10 ** rng.uniform(-6, -2)), D1=float(10 ** rng.uniform(-2.5, 0.5)), le=float(rng.uniform(0.005, 0.1)),
                   F0=float(rng.uniform(0.05, 0.95)), rate=float(10 ** rng.uniform(-3, -0.3)), pBL=float(rng.uniform(0.01, 0.3)))
    else:
        par["fl"] = int(rng.integers(3, 8))
    par["F0"] = min(max(par["F0"], 0.02), 0.98)
    par["pBL"] = min(max(par["pBL"], 0.005), 0.5)
    par["D0"] = min(par["D0"], par["D1"])
    m = make_model(frame_len=par["fl"], min_len=7, Ds=[par["D0"], par["D1"]], Fs=[par["F0"], 1 - par["F0"]],
                   loc_err=(par["le"],), rates=par["rate"], pBL=par["pBL"])
    p = engine_params(m, 2)
    a = one.sum_logp(p)
    s = one.engine.stats()
    if s["plan_verified"]:
        served["verified, some chunks planned again" if s["replanned"] else "verified"] += 1
        replanned += s["replanned"]
    else:
        served["planned from scratch"] += 1
    a2 = one.sum_logp(p)
    r = ref.sum_logp(p)
    b = many.sum_logp(p)
    if not (a == a2 == r == b):
        bad += 1
        print(f"MISMATCH it={it} {par}: verify {a!r} again {a2!r} scratch {r!r} multi {b!r} (stats {s})", flush=True)
        for c in range(len(one.chunks)):
            bb, aa, zz, _ = one.chunks[c]
            x = one.engine.chunk_logp(c, zz - aa, p)
            y = ref.engine.chunk_logp(c, zz - aa, p)
            if not np.array_equal(x, y):
                d = np.flatnonzero(x != y)
                print(f"   chunk {c}: {len(d)} tracks differ, first {d[:5]}, verify {x[d[:3]]} scratch {y[d[:3]]}", flush=True)
print(f"{n_eval} parameter sets, {bad} mismatches; served: {served}; chunks planned again: {replanned}")
