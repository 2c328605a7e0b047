"""Golden vectors for the window-only mode (`P_Cs_inter_bound_stats`, the pure sliding
window the north star names, = the threshold engine in the limit threshold -> 0+).

Runs the UNMODIFIED reference function `extrack/tracking.py:109-318` (dead in the live path; it needs
`np.product`, removed in numpy 2, restored here as an alias of `np.prod`) on seeded chunks and stores the
per-sequence log-weights reduced per track (log-sum-exp, what `Proba_Cs` :781-786 does with them).

    python tests/golden/make_golden_window.py      # build container only (needs /root/reference)
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
    dict(name="window_s2_fl4_L12", nS=2, nsub=1, d=2, fl=4, L=12, nT=40, isBL=1, seed=41),
    dict(name="window_s2_fl6_L15_noBL", nS=2, nsub=1, d=2, fl=6, L=15, nT=33, isBL=0, seed=42),
    dict(name="window_s3_fl3_L10", nS=3, nsub=1, d=2, fl=3, L=10, nT=24, isBL=1, seed=43),
    dict(name="window_s2_fl3_L9_3d", nS=2, nsub=1, d=3, fl=3, L=9, nT=31, isBL=1, seed=44),
    dict(name="window_s2_nsub2_fl4_L8", nS=2, nsub=2, d=2, fl=4, L=8, nT=20, isBL=1, seed=45),
]


def logsumexp_rows(LP):
    mx = LP.max(1, keepdims=True)
    return np.log(np.exp(LP - mx).sum(1, keepdims=True))[:, 0] + mx[:, 0]


def main():
    trk = ref_loader.load_tracking()
    if not hasattr(np, "product"):
        np.product = np.prod  # tracking.py:417-421 (fuse_tracks_general) predates numpy 2
    import scipy

    meta = dict(numpy=np.__version__, scipy=scipy.__version__)
    for c in CASES:
        model = make_model(nS=c["nS"], nsub=c["nsub"], frame_len=c["fl"])
        rng = np.random.default_rng(c["seed"])
        C = random_walk_tracks(c["nT"], c["L"], c["d"], rng, Ds=model.ds**2 / (2 * 0.02))
        LocErr = np.asarray(model.loc_err)[None, None]
        with contextlib.redirect_stdout(io.StringIO()):
            LP, _, _ = trk.P_Cs_inter_bound_stats(C, LocErr, model.ds, model.Fs, model.TrMat, model.pBL, c["isBL"],
                                                  model.cell_dims, model.nb_substeps, model.frame_len, 0, model.min_len)
            # the live threshold function in the limit threshold -> 0+ (SURVEY.md §0 [probe]: same value)
            LPth, _, _ = trk.P_Cs_inter_bound_stats_th(C, LocErr, model.ds, model.Fs, model.TrMat, model.pBL, c["isBL"],
                                                       model.cell_dims, model.nb_substeps, model.frame_len, 0, model.min_len,
                                                       1e-12, 10**9)
        logp = logsumexp_rows(np.asarray(LP))
        logp_th = logsumexp_rows(np.asarray(LPth))
        rel = float(np.max(np.abs(logp - logp_th) / np.abs(logp)))
        assert rel < 1e-12, (c["name"], rel)
        np.savez_compressed(
            os.path.join(HERE, c["name"] + ".npz"), C=C, loc_err=model.loc_err, ds=model.ds, Fs=model.Fs, TrMat=model.TrMat,
            pBL=model.pBL, cell_dims=np.asarray(model.cell_dims), nsub=model.nb_substeps, frame_len=model.frame_len,
            min_len=model.min_len, isBL=c["isBL"], ref_logp=logp, ref_logp_th=logp_th, n_seq=np.asarray(LP).shape[1], meta=str(meta))
        print(c["name"], "sequences", np.asarray(LP).shape[1], "sum", float(logp.sum()), "window vs threshold->0 rel", rel)


if __name__ == "__main__":
    main()
