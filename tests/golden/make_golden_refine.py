"""Golden vectors for the position refinement: the UNMODIFIED reference
`extrack/refined_localization.py:position_refinement` on seeded buckets; a case is kept only if the numpy restatement
(`oracle/refine_oracle.py`) reproduces it.

    python tests/golden/make_golden_refine.py      # build container only (needs /root/reference)
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
from oracle import ref_loader, refine_oracle  # noqa: E402

CASES = [
    dict(name="refine_s2", nS=2, d=2, fl=6, th=0.1, seed=71, buckets=((5, 40), (9, 55), (14, 35))),
    dict(name="refine_s2_short", nS=2, d=2, fl=4, th=0.2, seed=72, buckets=((2, 12), (3, 15), (4, 33))),
    dict(name="refine_s3", nS=3, d=2, fl=4, th=0.1, seed=73, buckets=((6, 45), (11, 31))),
    dict(name="refine_s2_3d", nS=2, d=3, fl=5, th=0.1, seed=74, buckets=((8, 32), (12, 20))),
]


def main():
    rl = ref_loader.load_refined_localization()
    import scipy

    meta = dict(numpy=np.__version__, scipy=scipy.__version__)
    for c in CASES:
        m = make_model(nS=c["nS"], frame_len=c["fl"])
        rng = np.random.default_rng(c["seed"])
        tracks = {str(L): random_walk_tracks(n, L, c["d"], rng, Ds=m.ds**2 / (2 * 0.02)) for L, n in c["buckets"]}
        le = float(m.loc_err[0])
        with contextlib.redirect_stdout(io.StringIO()):
            mus, sig = rl.position_refinement(tracks, le, m.ds, m.Fs, m.TrMat, frame_len=c["fl"], threshold=c["th"], max_nb_states=1000)
        omus, osig = refine_oracle.position_refinement(tracks, le, m.ds, m.Fs, m.TrMat, c["fl"], c["th"], 1000)
        worst = max(max(float(np.max(np.abs(mus[k] - omus[k]))), float(np.max(np.abs(sig[k] - osig[k])))) for k in tracks)
        assert worst < 1e-10, (c["name"], worst)
        keys = sorted(tracks, key=int)
        np.savez_compressed(os.path.join(HERE, c["name"] + ".npz"), keys=np.array(keys), loc_err=le, ds=m.ds, Fs=m.Fs, TrMat=m.TrMat,
                            frame_len=c["fl"], threshold=c["th"], max_nb_states=1000, meta=str(meta),
                            **{"C" + k: tracks[k] for k in keys}, **{"M" + k: mus[k] for k in keys}, **{"S" + k: sig[k] for k in keys})
        print(c["name"], "oracle vs reference", worst, "mean shift", float(np.mean(np.abs(mus[keys[-1]] - tracks[keys[-1]]))))


if __name__ == "__main__":
    main()
