#!/usr/bin/env python3
"""Wall time of xt_predict per call (kernel + read-back) on config 4, for several piece counts, with fresh and with
reused output arrays.  Run on the GPU box:  python tools/k3_probe.py [n_tracks]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from extrack_b200 import _native, tracking as xt  # noqa: E402
from extrack_b200.simulate import sim_tracks  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
tracks = sim_tracks(n, seed=99, device="cuda:0", **bench.SIM_KW)
st, _ = xt._sorted_buckets(tracks)
LocErr, ds, Fs, TrMat, pBL = xt.extract_params(bench.eval_params(), bench.DT, 2, 1)
p = xt.build_tables(LocErr, ds, Fs, TrMat, pBL, [1], 1, 8, st[0].shape[1], 0.1, 200, 2)
eng = _native.Engine(0)
eng.upload(st, [0 if a.shape[1] == st[-1].shape[1] else 1 for a in st], xt.MAX_TRACKS_PER_CHUNK)
locs = sum(a.shape[0] * a.shape[1] for a in st)
eng.predict(p, 2)
ref = eng.predict(p, 2)
for pieces in (1, 4, 6, 10, 16):
    eng.set_option("k3_pieces", pieces)
    for reuse in (0, 1):
        outs = [np.empty((nn, L, 2)) for (L, nn) in eng.segments]
        ts = []
        for _ in range(6):
            if not reuse:
                outs = [np.empty((nn, L, 2)) for (L, nn) in eng.segments]
            ptrs = (C.c_void_p * len(outs))(*[o.ctypes.data for o in outs])
            t = time.perf_counter()
            eng._check(eng._lib.xt_predict(eng._h, C.byref(p), ptrs))
            ts.append((time.perf_counter() - t) * 1e3)
        same = all(np.array_equal(a, b) for a, b in zip(outs, ref))
        print(f"pieces {pieces:2d} {'reused' if reuse else 'fresh '} outputs: ms per call {' '.join(f'{x:6.1f}' for x in ts)}  kernel {eng.stats()['ms_predict']:.1f} ms"
              f"  launches {eng.stats()['k3_launches']}  identical {same}  best {locs / min(ts) / 1e3:.3g} loc/s", flush=True)
eng.close()
