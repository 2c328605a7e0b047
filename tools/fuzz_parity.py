"""Randomised parity sweep on the GPU box: engine vs oracle on random small chunks (per-track log P,
plan equality, state posteriors), including peak-wise LocErr / per-track dt and the plan-kernel
variants.  One-off confidence check (the pytest suite holds the fixed cases).

    python tools/fuzz_parity.py [seconds] [seed]
"""
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from extrack_b200 import _native  # noqa: E402
from extrack_b200 import tracking as xt  # noqa: E402
from helpers import gid_from_groups, make_model, random_walk_tracks  # noqa: E402
from oracle import extrack_oracle as orc  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 300.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
t_end = time.time() + budget
n_case = n_bad = 0
worst = 0.0
while time.time() < t_end:
    nS = int(rng.choice([2, 2, 2, 3, 3, 4]))
    nsub = int(rng.choice([1, 1, 1, 2])) if nS <= 3 else 1
    d = int(rng.choice([1, 2, 2, 2, 3]))
    fl = int(rng.integers(2, 11))
    L = int(rng.integers(2, 36))
    nT = int(rng.choice([1, 2, 7, 29, 30, 31, 64, 200]))
    isBL = int(rng.integers(0, 2))
    th = float(rng.choice([0.05, 0.1, 0.2, 0.2, 0.4]))
    mx = int(rng.choice([12, 40, 120, 500]))
    kloc = int(rng.choice([0, 0, 1, d]))
    var_dt = bool(rng.integers(0, 3) == 0)
    per_dim = kloc == 0 and d > 1 and rng.integers(0, 4) == 0
    loc = tuple(0.02 + 0.01 * rng.random(d)) if per_dim else (0.02,)
    Ds = np.sort(np.r_[1e-5, rng.random(nS - 1) * 0.5 + 0.01])
    m = make_model(nS=nS, nsub=nsub, loc_err=loc, frame_len=fl, min_len=int(rng.integers(2, 8)), threshold=th,
                   max_nb_states=mx, pBL=float(rng.random() * 0.2 + 0.01), Ds=Ds, rates=float(rng.random() * 0.3 + 0.02))
    C = random_walk_tracks(nT, L, d, rng, Ds=Ds)
    sig = 0.02 * (1 + 0.5 * rng.random((nT, L, kloc))) if kloc else None
    dts = 0.02 * (1 + 0.5 * rng.random((nT, L))) if var_dt else None
    ds3 = np.sqrt(2 * Ds[None, None] * dts[:, :, None]) if var_dt else None
    desc = f"nS={nS} nsub={nsub} d={d} fl={fl} L={L} nT={nT} isBL={isBL} th={th} max={mx} kloc={kloc} var_dt={var_dt} per_dim={per_dim}"
    try:
        plan = []
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = orc.chunk_logp(C, m, isBL, plan_out=plan, sig=sig, ds3=ds3)
    except ValueError:
        continue  # grouping failure in the oracle (reference error path)
    if max((r["nB_in"] for r in plan), default=0) > 2000:
        continue
    p = xt.build_tables(m.loc_err, m.ds, m.Fs, m.TrMat, m.pBL, m.cell_dims, nsub, fl, m.min_len, th, mx, d,
                        var_loc_k=kloc, var_dt=var_dt, Ds=Ds)
    eng = _native.Engine(0)
    try:
        eng.upload([C], [isBL], nT)
        if kloc or var_dt:
            eng.upload_aux([sig] if kloc else None, [dts] if var_dt else None)
        if var_dt:
            eng.set_stay_tables(False, *xt.stay_tables(Ds, np.array([xt._mid2(dts[:, 0])]), m.TrMat, m.pBL, m.cell_dims, nsub))
        for opts in (dict(), dict(k1_threads=256), dict(k1_threads=256, force_global_replay=1)):
            for k, v in opts.items():
                eng.set_option(k, v)
            got = eng.chunk_logp(0, nT, p)
            eng.sum_logp(p)  # second evaluation: shared-memory scratch / pipelined path
            got2 = eng.chunk_logp(0, nT, p)
            for g in (got, got2):
                err = float(np.max(np.abs(g - ref) / np.maximum(np.abs(ref), 1e-300))) if np.isfinite(ref).all() else 0.0
                worst = max(worst, err)
                if not err < 1e-9:
                    n_bad += 1
                    print("LOGP MISMATCH", desc, opts, err, flush=True)
            for rec in plan:
                nB, nG, gid, _ = eng.plan_dump(0, rec["step"])
                if nB != rec["nB_in"] or nG != len(rec["groups"]) or not np.array_equal(gid, gid_from_groups(rec["groups"], nB)):
                    n_bad += 1
                    print("PLAN MISMATCH", desc, opts, "step", rec["step"], flush=True)
                    break
        if nsub == 1 and nT <= 31:
            m2 = make_model(nS=nS, nsub=1, loc_err=loc, frame_len=fl, min_len=m.min_len, threshold=0.1, max_nb_states=200,
                            pBL=m.pBL, Ds=Ds, rates=0.1)
            m2.TrMat = m.TrMat
            try:
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    want = np.concatenate([orc.chunk_recursion(C[i:i + 1], m2, isBL, 1, None, None, None if sig is None else sig[i:i + 1],
                                                               None if ds3 is None else ds3[i:i + 1])[2] for i in range(nT)])
                p2 = xt.build_tables(m2.loc_err, m2.ds, m2.Fs, m2.TrMat, m2.pBL, m2.cell_dims, 1, fl, m2.min_len, 0.1, 200, d,
                                     var_loc_k=kloc, var_dt=var_dt, Ds=Ds)
                if var_dt:
                    eng.set_stay_tables(True, *xt.stay_tables(Ds, np.stack([dts[:, 0], dts[:, 0]], 1), m2.TrMat, m2.pBL, m2.cell_dims, 1))
                pr = eng.predict(p2, nS)[0]
                perr = float(np.max(np.abs(pr - want)))
                if not perr < 1e-6:
                    n_bad += 1
                    print("PREDICT MISMATCH", desc, perr, flush=True)
            except ValueError:
                pass
    except Exception as e:  # engine errors are findings too
        n_bad += 1
        print("ENGINE ERROR", desc, repr(e)[:300], flush=True)
    finally:
        eng.close()
    n_case += 1
print(f"fuzz: {n_case} cases, {n_bad} findings, worst relative log P error {worst:.2e}")
sys.exit(1 if n_bad else 0)
