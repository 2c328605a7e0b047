"""Ad-hoc GPU bring-up script (not a pytest): engine vs oracle on a handful of chunks, with timing.

    python tests/gpu_quick.py
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from extrack_b200 import _native  # noqa: E402
from helpers import engine_params, gid_from_groups, make_model, random_walk_tracks  # noqa: E402
from oracle import extrack_oracle as orc  # noqa: E402


def one(nS, nsub, d, fl, L, nT, isBL, seed=0, **kw):
    rng = np.random.default_rng(seed)
    model = make_model(nS=nS, nsub=nsub, frame_len=fl, **kw)
    C = random_walk_tracks(nT, L, d, rng, Ds=(model.ds**2 / (2 * 0.02)))
    plan = []
    t = time.time()
    ref = orc.chunk_logp(C, model, isBL, plan_out=plan)
    t_or = time.time() - t
    p = engine_params(model, d)
    eng = _native.Engine(0)
    eng.upload([C], [isBL], max(nT, 1))
    t = time.time()
    got = eng.chunk_logp(0, nT, p)
    t_en = time.time() - t
    err = np.max(np.abs(got - ref) / np.abs(ref))
    eng.set_option("force_global_replay", 1)
    got2 = eng.chunk_logp(0, nT, p)
    err = max(err, np.max(np.abs(got2 - ref) / np.abs(ref)))
    eng.set_option("force_global_replay", 0)
    eng.set_option("k2_variant", 1)
    got3 = eng.chunk_logp(0, nT, p)
    err3 = np.max(np.abs(got3 - ref) / np.abs(ref))
    if err3 > 1e-9:
        print("   first-generation linear kernel relerr", err3)
    err = max(err, err3)
    eng.set_option("k2_variant", 0)
    eng.set_option("k2_tpt", 2)
    got4 = eng.chunk_logp(0, nT, p)
    err4 = np.max(np.abs(got4 - ref) / np.abs(ref))
    if err4 > 1e-9:
        print("   fused kernel, two tracks per thread: relerr", err4)
    err = max(err, err4)
    eng.set_option("k2_tpt", 1)
    eng.chunk_logp(0, nT, p)
    mism = 0
    for rec in plan:
        nB, nG, gid, th = eng.plan_dump(0, rec["step"])
        want = gid_from_groups(rec["groups"], rec["nB_in"])
        if nB != rec["nB_in"] or nG != len(rec["groups"]) or not np.array_equal(gid, want):
            mism += 1
    st = eng.stats()
    print(f"nS={nS} nsub={nsub} d={d} fl={fl} L={L} nT={nT} isBL={isBL} kw={kw}: relerr={err:.2e} sum_rel={abs(got.sum()-ref.sum())/abs(ref.sum()):.2e} "
          f"plan_mismatch={mism}/{len(plan)} maxnB={st['max_nB_in']} oracle={t_or:.3f}s engine={t_en*1e3:.2f}ms (plan {st['ms_plan']:.3f} replay {st['ms_replay']:.3f})")
    eng.close()
    return err, mism


if __name__ == "__main__":
    bad = 0
    cfgs = [] if os.environ.get("QUICK_SKIP") else [
        dict(nS=2, nsub=1, d=2, fl=8, L=20, nT=300, isBL=1),
        dict(nS=2, nsub=1, d=2, fl=8, L=12, nT=25, isBL=0),
        dict(nS=2, nsub=1, d=2, fl=4, L=3, nT=50, isBL=1),
        dict(nS=2, nsub=1, d=2, fl=4, L=2, nT=50, isBL=1),
        dict(nS=2, nsub=1, d=2, fl=4, L=4, nT=1, isBL=0),
        dict(nS=2, nsub=1, d=2, fl=6, L=30, nT=2000, isBL=1),
        dict(nS=3, nsub=2, d=2, fl=6, L=15, nT=100, isBL=1, max_nb_states=500),
        dict(nS=3, nsub=1, d=3, fl=6, L=14, nT=100, isBL=0, max_nb_states=60),
        dict(nS=3, nsub=1, d=3, fl=5, L=14, nT=100, isBL=1, loc_err=(0.02, 0.02, 0.03)),
        dict(nS=2, nsub=1, d=2, fl=6, L=14, nT=100, isBL=1, min_len=10),
        dict(nS=2, nsub=2, d=2, fl=6, L=14, nT=64, isBL=1),
        dict(nS=4, nsub=1, d=2, fl=5, L=12, nT=64, isBL=1, max_nb_states=200),
        dict(nS=3, nsub=2, d=2, fl=6, L=15, nT=100, isBL=1, max_nb_states=500, int8_wrap=False),
        dict(nS=3, nsub=2, d=3, fl=4, L=12, nT=40, isBL=0, max_nb_states=120),
    ]
    for cfg in cfgs:
        try:
            err, mism = one(**cfg)
            if not (err < 1e-9) or mism:
                bad += 1
        except Exception as e:  # keep going: show every failure in one GPU call
            import traceback

            traceback.print_exc()
            bad += 1
    eng = _native.Engine(0)
    print("fp64 peak TFLOP/s:", eng.fp64_peak_tflops())
    eng.close()
    # multi-bucket data set: objective vs oracle + repeated timing
    from extrack_b200 import tracking as xt
    from extrack_b200.simulate import sim_tracks

    n_tracks = int(os.environ.get("QUICK_TRACKS", "100000"))
    t = time.time()
    tracks = sim_tracks(n_tracks, seed=0, device="cuda", max_track_len=30, min_track_len=10, LocErr=0.02, Ds=[0, 0.25],
                        nb_dims=2, initial_fractions=[0.6, 0.4], TrMat=[[0.9, 0.1], [0.1, 0.9]], dt=0.02, pBL=0.05,
                        cell_dims=[1, None, None])
    print("generated", sum(len(v) for v in tracks.values()), "tracks in", round(time.time() - t, 2), "s")
    sorted_tracks, _ = xt._sorted_buckets(tracks)
    model = make_model(nS=2, nsub=1, frame_len=8, min_len=sorted_tracks[0].shape[1], Ds=[1e-5, 0.25], Fs=[0.6, 0.4])
    p = engine_params(model, 2)
    t = time.time()
    ts = xt.TrackSet(sorted_tracks)
    print("upload", round(time.time() - t, 3), "s; chunks", len(ts.chunks))
    def timed(tag, n=4):
        for it in range(n):
            t = time.time()
            v = ts.sum_logp(p)
            dtm = time.time() - t
            st = ts.engine.stats()
            print(f"{tag} eval {it}: sum_logp={v!r} wall={dtm*1e3:.3f}ms plan={st['ms_plan']:.3f}ms replay={st['ms_replay']:.3f}ms "
                  f"pipelined={st['pipelined']} launches={st['k1_launches']}+{st['k2_launches']} -> {st['track_steps']/dtm/1e9:.3f} G track-steps/s")
        return v

    ts.engine.set_option("pipeline", 0)
    v = timed("two-phase")
    ts.engine.set_option("pipeline", 1)
    for g in (1, 2, 3, 4, 6, 8, 21):
        ts.engine.set_option("n_groups", g)
        v2 = timed(f"pipelined G={g}", 3)
        if v2 != v:
            print("   MISMATCH pipelined vs two-phase", v2, v)
            bad += 1
    ts.engine.set_option("n_groups", 4)
    # host-buffer objective: upload overlapped with the kernels
    pinned = []
    for a in sorted_tracks:
        b = _native.pinned_empty(a.shape)
        b[...] = a
        pinned.append(b)
    bl = [0 if a.shape[1] == sorted_tracks[-1].shape[1] else 1 for a in sorted_tracks]
    e2 = _native.Engine(0)
    for it in range(5):
        t = time.time()
        vh = e2.sum_logp_host(pinned, bl, 2000, p)
        dtm = time.time() - t
        st = e2.stats()
        print(f"host eval {it}: sum_logp={vh!r} wall={dtm*1e3:.3f}ms pipelined={st['pipelined']} -> {st['track_steps']/dtm/1e9:.3f} G track-steps/s")
    if vh != v:
        print("   MISMATCH host vs resident", vh, v)
        bad += 1
    e2.set_option("pipeline", 0)
    for it in range(2):
        t = time.time()
        vh = e2.sum_logp_host(pinned, bl, 2000, p)
        dtm = time.time() - t
        print(f"host eval two-phase {it}: sum_logp={vh!r} wall={dtm*1e3:.3f}ms")
    e2.close()
    if n_tracks <= 200000:
        t = time.time()
        ref = -orc.neg_log_likelihood(sorted_tracks, model, workers=8)
        print("oracle", ref, "in", round(time.time() - t, 2), "s; rel err", abs(ref - v) / abs(ref))
        if not abs(ref - v) / abs(ref) < 1e-9:
            bad += 1
    print("FAILED" if bad else "ALL OK", bad)
    sys.exit(1 if bad else 0)
