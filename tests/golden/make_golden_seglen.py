"""Golden vectors of the segment-length histogram from the *unmodified* reference
(``extrack/histograms.py:P_segment_len`` / ``len_hist``), run in the build container:

    python tests/golden/make_golden_seglen.py

Each case is kept only if the oracle restatement reproduces it (the reference's order among equal sort
keys is unspecified, see oracle/seglen_oracle.py), so the committed vectors pin both.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from helpers import random_walk_tracks  # noqa: E402
from oracle import extrack_oracle as orc  # noqa: E402
from oracle import ref_loader, seglen_oracle as so  # noqa: E402


def model(nS, le, pBL=0.07, cell=1.0, D0=1e-5):
    Ds = np.array([D0, 0.04, 0.25, 0.6])[:nS]
    ds = np.sqrt(2 * Ds * 0.02)
    Fs = np.array([0.45, 0.3, 0.15, 0.1])[:nS]
    Fs = Fs / Fs.sum()
    Tr = np.array([[0, 0.06, 0.03, 0.01], [0.05, 0, 0.04, 0.02], [0.01, 0.07, 0, 0.03], [0.02, 0.03, 0.08, 0]])[:nS, :nS]
    Tr = Tr + np.diag(1 - Tr.sum(1))
    return orc.Model(np.asarray(le, float), ds, Fs, Tr, pBL, [cell], 1, 6, 3, 0.2, 120)


CASES = [  # name, nS, L, d, max_nb_states, isBL, LocErr, min_l, nT
    ("seglen_s2_L8", 2, 8, 2, 16, 1, (0.02,), 3, 50),
    ("seglen_s2_L20_noBL", 2, 20, 2, 128, 0, (0.02,), 8, 37),
    ("seglen_s2_L10_noprune", 2, 10, 2, 100000, 1, (0.02,), 5, 9),
    ("seglen_s2_L2", 2, 2, 2, 10, 1, (0.02,), 2, 5),
    ("seglen_s2_L3", 2, 3, 2, 10, 0, (0.02,), 2, 5),
    ("seglen_s3_L9", 3, 9, 2, 30, 1, (0.02,), 3, 50),
    ("seglen_s3_3d_locerr_per_dim", 3, 8, 3, 100, 0, (0.02, 0.03, 0.04), 4, 21),
    ("seglen_s4_L6", 4, 6, 2, 64, 1, (0.02,), 3, 13),
    ("seglen_s2_1d", 2, 10, 1, 32, 1, (0.03,), 2, 17),
    ("seglen_s2_L260_rescale", 2, 260, 2, 24, 1, (0.02,), 100, 6),   # final LP > 600: per-column rescale (:243-244)
    ("seglen_s3_3d_L120_rescale", 3, 120, 3, 30, 0, (0.02,), 50, 4),
]


def main():
    H = ref_loader.load_histograms()
    rng = np.random.default_rng(20261017)
    for name, nS, L, d, mx, isBL, le, minl, nT in CASES:
        m = model(nS, le)
        C = random_walk_tracks(nT, L, d, rng, Ds=m.ds**2 / 0.04)
        with contextlib.redirect_stdout(io.StringIO()):
            LP, Bs, hist = H.P_segment_len(C, np.asarray(le)[None, None], m.ds, m.Fs, m.TrMat, min_l=minl, pBL=m.pBL, isBL=isBL,
                                           cell_dims=[1.0], nb_substeps=1, max_nb_states=mx)
        LP1, h1, H1 = so.segment_len_chunk(C, m, isBL, mx, minl, want_histories=True)
        ok = np.abs(LP - LP1).max() <= 1e-12 * np.abs(LP).max() and np.abs(hist - h1).max() <= 1e-12 and (np.asarray(Bs) == H1).all()
        print(name, "oracle == reference:", ok, "hist sum", hist.sum())
        if not ok:
            raise SystemExit("oracle and reference disagree on " + name)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), C=C, loc_err=np.asarray(le), ds=m.ds, Fs=m.Fs, TrMat=m.TrMat,
                            pBL=m.pBL, cell_dims=np.asarray([1.0]), isBL=isBL, min_l=minl, max_nb_states=mx,
                            ref_LP=np.asarray(LP), ref_Bs=np.asarray(Bs).astype(np.int8), ref_hist=hist)
    # len_hist over several length buckets (chunks of 50 tracks, longest bucket without the leave term)
    trk = ref_loader.load_tracking()
    m = model(2, (0.02,))
    tracks = {str(L): random_walk_tracks(n, L, 2, rng, Ds=m.ds**2 / 0.04) for L, n in ((6, 120), (9, 70), (14, 55))}
    params = trk.generate_params(nb_states=2, LocErr_type=1, nb_dims=2, LocErr_bounds=[0.005, 0.1], D_max=10,
                                 Fractions_bounds=[0.001, 0.99], estimated_LocErr=[0.02], estimated_Ds=[1e-5, 0.25],
                                 estimated_Fs=[0.6, 0.4], estimated_transition_rates=0.1)
    with contextlib.redirect_stdout(io.StringIO()):
        ref = H.len_hist(tracks, params, 0.02, cell_dims=[1.0], nb_states=2, max_nb_states=64, workers=1, nb_substeps=1)
    vals = {k: float(params[k].value) for k in params}
    np.savez_compressed(os.path.join(HERE, "seglen_len_hist.npz"), keys=np.array(list(tracks)), ref_hist=ref,
                        names=np.array(list(vals)), values=np.array(list(vals.values())),
                        **{"C" + k: v for k, v in tracks.items()})
    print("len_hist", ref.sum())


if __name__ == "__main__":
    main()
