"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Each case stores the inputs (tracks, model quantities exactly as `extract_params` produced them,
configuration) and the reference outputs of `Proba_Cs` (per-track log P) and, for nb_substeps=1,
of `P_Cs_inter_bound_stats_th(do_preds=1)` (state posteriors).  One multi-bucket case stores the
value of `cum_Proba_Cs`.  Seeds are fixed; numpy/scipy versions are recorded.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

from helpers import make_model, random_walk_tracks  # noqa: E402
from oracle import ref_loader  # noqa: E402

CASES = [
    dict(name="s2_fl8_L20", nS=2, nsub=1, d=2, fl=8, L=20, nT=120, isBL=1, seed=1),
    dict(name="s2_fl6_L30_noBL", nS=2, nsub=1, d=2, fl=6, L=30, nT=64, isBL=0, seed=2),
    dict(name="s2_fl4_L3", nS=2, nsub=1, d=2, fl=4, L=3, nT=40, isBL=1, seed=3),
    dict(name="s2_fl4_L2", nS=2, nsub=1, d=2, fl=4, L=2, nT=40, isBL=1, seed=4),
    dict(name="s2_single_track", nS=2, nsub=1, d=2, fl=5, L=9, nT=1, isBL=1, seed=5),
    dict(name="s3_nsub2_wrap", nS=3, nsub=2, d=2, fl=6, L=13, nT=48, isBL=1, seed=6, max_nb_states=500),
    dict(name="s3_3d", nS=3, nsub=1, d=3, fl=6, L=14, nT=64, isBL=0, seed=7, max_nb_states=60),
    dict(name="s3_3d_locerr_per_dim", nS=3, nsub=1, d=3, fl=5, L=12, nT=48, isBL=1, seed=8, loc_err=(0.02, 0.02, 0.03)),
    dict(name="s2_minlen10", nS=2, nsub=1, d=2, fl=6, L=14, nT=64, isBL=1, seed=9, min_len=10),
    dict(name="s2_nsub2", nS=2, nsub=2, d=2, fl=6, L=12, nT=48, isBL=1, seed=10),
    dict(name="s4", nS=4, nsub=1, d=2, fl=5, L=11, nT=48, isBL=1, seed=11, max_nb_states=200),
    dict(name="s2_escalate", nS=2, nsub=1, d=2, fl=10, L=16, nT=48, isBL=1, seed=12, max_nb_states=12, threshold=0.05),
]


def main():
    trk = ref_loader.load_tracking()
    import scipy

    meta = dict(numpy=np.__version__, scipy=scipy.__version__)
    for c in CASES:
        kw = {k: c[k] for k in ("loc_err", "min_len", "max_nb_states", "threshold") if k in c}
        model = make_model(nS=c["nS"], nsub=c["nsub"], frame_len=c["fl"], **kw)
        rng = np.random.default_rng(c["seed"])
        C = random_walk_tracks(c["nT"], c["L"], c["d"], rng, Ds=model.ds**2 / (2 * 0.02))
        LocErr = np.asarray(model.loc_err)[None, None]
        with contextlib.redirect_stdout(io.StringIO()):
            logp = trk.Proba_Cs(C, LocErr, model.ds, model.Fs, model.TrMat, model.pBL, c["isBL"], model.cell_dims,
                                model.nb_substeps, model.frame_len, model.min_len, model.threshold, model.max_nb_states)
            preds = np.zeros(0)
            if c["nsub"] == 1:  # predict_Bs forces nb_substeps = 1; default nb_max = 1 => one track per call
                preds = np.concatenate([
                    trk.P_Cs_inter_bound_stats_th(C[i : i + 1], LocErr, model.ds, model.Fs, model.TrMat, model.pBL, c["isBL"],
                                                  model.cell_dims, 1, model.frame_len, 1, model.min_len, 0.1, 200)[2]
                    for i in range(min(len(C), 24))
                ])
        np.savez_compressed(
            os.path.join(HERE, c["name"] + ".npz"), C=C, loc_err=model.loc_err, ds=model.ds, Fs=model.Fs, TrMat=model.TrMat,
            pBL=model.pBL, cell_dims=np.asarray(model.cell_dims), nsub=model.nb_substeps, frame_len=model.frame_len,
            min_len=model.min_len, threshold=model.threshold, max_nb_states=model.max_nb_states, isBL=c["isBL"],
            ref_logp=logp, ref_preds=preds, meta=str(meta))
        print(c["name"], float(np.sum(logp)))

    # multi-bucket objective through the reference's own parameter path (lmfit-style Parameters)
    from lmfit import Parameters

    rng = np.random.default_rng(100)
    buckets = {str(L): random_walk_tracks(n, L, 2, rng) for L, n in ((5, 30), (8, 70), (12, 2100), (17, 45))}
    p = Parameters()
    for k, v in dict(D0=1e-5, D1=0.25, LocErr=0.02, F0=0.6, p01=0.1, p10=0.12, pBL=0.05).items():
        p.add(k, value=v)
    p.add("F1", expr="1-F0")
    keys = sorted(buckets, key=int)
    st = [buckets[k] for k in keys]
    vals = {}
    for fl in (4, 7):
        with contextlib.redirect_stdout(io.StringIO()):
            vals[fl] = float(trk.cum_Proba_Cs(p, st, 0.02, [1], None, 2, 1, fl, 0, 1, 1, 0.2, 120))
    np.savez_compressed(os.path.join(HERE, "objective_multibucket.npz"), keys=np.array(keys), fl=np.array(list(vals)),
                        neglogl=np.array(list(vals.values())), meta=str(meta), **{"C" + k: buckets[k] for k in keys})
    print("objective", vals)
    make_var_cases(trk, meta)
    make_fit_case(trk, meta)


# peak-wise input_LocErr / per-track dt: objective and predictions through the
# reference's own public functions (cum_Proba_Cs with lists, predict_Bs with dicts)
VAR_CASES = [
    dict(name="var_loc_k1", nS=2, nsub=1, d=2, fl=6, kloc=1, var_dt=False, slope=False, seed=21),
    dict(name="var_loc_kd", nS=2, nsub=1, d=2, fl=6, kloc=2, var_dt=False, slope=False, seed=22),
    dict(name="var_dt", nS=2, nsub=1, d=2, fl=6, kloc=0, var_dt=True, slope=False, seed=23),
    dict(name="var_slope_dt", nS=2, nsub=1, d=2, fl=5, kloc=1, var_dt=True, slope=True, seed=24),
    dict(name="var_dt_s3_nsub2", nS=3, nsub=2, d=2, fl=4, kloc=0, var_dt=True, slope=False, seed=25, chunk=30),
    dict(name="var_all_s3_3d", nS=3, nsub=1, d=3, fl=5, kloc=3, var_dt=True, slope=False, seed=26, chunk=20),
]


def make_var_cases(trk, meta):
    for c in VAR_CASES:
        rng = np.random.default_rng(c["seed"])
        nS, d = c["nS"], c["d"]
        Ds = [1e-5, 0.25] if nS == 2 else [1e-5, 0.04, 0.25]
        st = [random_walk_tracks(n, L, d, rng, Ds=Ds) for L, n in ((5, 40), (8, 50), (12, 35))]
        kw = dict(nb_states=nS, nb_dims=d, estimated_Ds=Ds, estimated_Fs=[1 / nS] * (nS - 1), estimated_transition_rates=0.1)
        if c["slope"]:
            params = trk.generate_params(LocErr_type=4, slope_offsets_estimates=[1.1, 0.002], **kw)
        else:
            params = trk.generate_params(LocErr_type=1, estimated_LocErr=[0.02], **kw)
        il = None if c["kloc"] == 0 else [0.02 * (1 + 0.5 * rng.random(a.shape[:2] + (c["kloc"],))) for a in st]
        dts = 0.02 if not c["var_dt"] else [0.02 * (1 + 0.5 * rng.random(a.shape[:2])) for a in st]
        chunk = c.get("chunk", 2000)
        with contextlib.redirect_stdout(io.StringIO()):
            val = float(trk.cum_Proba_Cs(params, st, dts, [1], il, nS, c["nsub"], c["fl"], 0, 1, 1, 0.2, 120, chunk))
            preds = {}
            if c["nsub"] == 1:
                tr = {str(a.shape[1]): a for a in st}
                ild = None if il is None else {str(a.shape[1]): x for a, x in zip(st, il)}
                dtd = dts if not c["var_dt"] else {str(a.shape[1]): x for a, x in zip(st, dts)}
                preds = trk.predict_Bs(tr, dtd, params, cell_dims=[1], nb_states=nS, frame_len=c["fl"], input_LocErr=ild)
        pv = {k: float(params[k].value) for k in params}
        arrays = {"C%d" % i: a for i, a in enumerate(st)}
        if il is not None:
            arrays.update({"S%d" % i: a for i, a in enumerate(il)})
        if c["var_dt"]:
            arrays.update({"T%d" % i: a for i, a in enumerate(dts)})
        arrays.update({"P%d" % i: preds[str(a.shape[1])] for i, a in enumerate(st) if preds})
        np.savez_compressed(os.path.join(HERE, c["name"] + ".npz"), nS=nS, nsub=c["nsub"], fl=c["fl"], kloc=c["kloc"],
                            var_dt=int(c["var_dt"]), slope=int(c["slope"]), chunk=chunk, neglogl=val,
                            param_names=np.array(list(pv)), param_values=np.array(list(pv.values())), meta=str(meta), **arrays)
        print(c["name"], val)


def make_fit_case(trk, meta):
    """BASELINE config 1: 2-state param_fitting on Tutorials/tracks.csv (frame_len 6, bfgs) by the
    UNMODIFIED reference, driven by the same lmfit stand-in the repo uses where lmfit is absent.
    Stores the tracks as read by the reference's own reader, the start values, the fitted parameters
    and the final objective."""
    rd = ref_loader.load_readers()
    with contextlib.redirect_stdout(io.StringIO()):
        tracks, _, _ = rd.read_table(os.path.join(ref_loader.REFERENCE_ROOT, "Tutorials", "tracks.csv"), lengths=np.arange(5, 50),
                                     dist_th=0.3, frames_boundaries=[0, 10000], fmt="csv", colnames=["X", "Y", "frame", "track_ID"],
                                     opt_colnames=[], remove_no_disp=True)
        params = trk.generate_params(nb_states=2, LocErr_type=1, nb_dims=2, LocErr_bounds=[0.005, 0.1], D_max=10,
                                     Fractions_bounds=[0.001, 0.99])
        start = {k: float(params[k].value) for k in params}
        fit = trk.param_fitting(tracks, 0.02, params=params, nb_states=2, nb_substeps=1, frame_len=6, verbose=0, workers=1,
                                method="bfgs", cell_dims=[1], threshold=0.2, max_nb_states=120)
    fitted = {k: float(fit.params[k].value) for k in fit.params}
    keys = sorted((k for k in tracks if len(tracks[k])), key=int)
    np.savez_compressed(os.path.join(HERE, "fit_tracks_csv.npz"), keys=np.array(keys), names=np.array(list(fitted)),
                        start=np.array([start[k] for k in fitted]), fitted=np.array(list(fitted.values())),
                        neglogl=float(fit.residual[0]), meta=str(meta), **{"C" + k: np.asarray(tracks[k]) for k in keys})
    print("fit", fitted, float(fit.residual[0]))


if __name__ == "__main__":
    main()
