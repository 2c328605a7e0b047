#!/usr/bin/env python3
"""Stress test (GPU box): the objective must have the same bits on one context, on repeated calls, and on several
contexts (logical shards on one GPU) for many parameter sets.  Prints every mismatch."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from extrack_b200 import tracking as xt  # noqa: E402
from helpers import engine_params, make_model, random_walk_tracks  # noqa: E402

rng = np.random.default_rng(5)
st = [random_walk_tracks(n, L, 2, rng) for L, n in ((7, 2300), (11, 2100), (16, 600))]
one = xt.TrackSet(st, 2000)
many = xt.TrackSet(st, 2000, devices=[0, 0, 0])
n_eval = int(sys.argv[1]) if len(sys.argv) > 1 else 300
bad = 0
for it in range(n_eval):
    D1 = float(10 ** rng.uniform(-2.5, 0.5))
    le = float(rng.uniform(0.005, 0.1))
    F0 = float(rng.uniform(0.05, 0.95))
    rate = float(10 ** rng.uniform(-3, -0.3))
    m = make_model(frame_len=int(rng.integers(3, 8)), min_len=7, Ds=[float(10 ** rng.uniform(-6, -2)), D1], Fs=[F0, 1 - F0],
                   loc_err=(le,), rates=rate, pBL=float(rng.uniform(0.01, 0.3)))
    p = engine_params(m, 2)
    print(f'it={it} D1={D1:.4g} le={le:.4g} fl={m.frame_len} rate={rate:.3g}', flush=True)
    a = one.sum_logp(p)
    a2 = one.sum_logp(p)
    b = many.sum_logp(p)
    b2 = many.sum_logp(p)
    if not (a == a2 == b == b2):
        bad += 1
        print(f"MISMATCH it={it} D1={D1:.4g} le={le:.4g} fl={m.frame_len}: one {a!r} {a2!r} many {b!r} {b2!r}", flush=True)
        for c in range(len(one.chunks)):
            bb, aa, zz, _ = one.chunks[c]
            x = one.engine.chunk_logp(c, zz - aa, p)
            y = many.engine.chunk_logp(c, zz - aa, p)
            x2 = one.engine.chunk_logp(c, zz - aa, p)
            if not (np.array_equal(x, y) and np.array_equal(x, x2)):
                d = np.flatnonzero((x != y) | (x != x2))
                print(f"   chunk {c}: {len(d)} tracks differ, first {d[:5]}, one {x[d[:3]]} again {x2[d[:3]]} many {y[d[:3]]}", flush=True)
print(f"{n_eval} parameter sets, {bad} mismatches")
