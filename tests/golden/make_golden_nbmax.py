"""Golden vectors for predict_Bs with nb_max > 1 (plan shared by the nb_max tracks of a chunk, decided from its first
30 tracks with per-track weighted histories: fuse_tracks_th(do_preds=1), tracking.py:652-743, called from :860-896).

Runs the UNMODIFIED reference `predict_Bs(..., nb_max=...)` on seeded buckets.

    python tests/golden/make_golden_nbmax.py      # build container only (needs /root/reference)
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

from helpers import random_walk_tracks  # noqa: E402
from oracle import ref_loader  # noqa: E402

CASES = [
    dict(name="nbmax_s2_50", nS=2, d=2, fl=6, nb_max=50, seed=61, buckets=((6, 70), (11, 130), (17, 64))),
    dict(name="nbmax_s2_2000", nS=2, d=2, fl=8, nb_max=2000, seed=62, buckets=((7, 90), (14, 75), (25, 40))),
    dict(name="nbmax_s3_40", nS=3, d=2, fl=4, nb_max=40, seed=63, buckets=((5, 45), (9, 85), (13, 33))),
    dict(name="nbmax_s2_3d_7", nS=2, d=3, fl=5, nb_max=7, seed=64, buckets=((8, 30), (12, 20))),
]


def main():
    trk = ref_loader.load_tracking()
    import scipy

    meta = dict(numpy=np.__version__, scipy=scipy.__version__)
    for c in CASES:
        rng = np.random.default_rng(c["seed"])
        nS, d = c["nS"], c["d"]
        Ds = [1e-5, 0.25] if nS == 2 else [1e-5, 0.04, 0.25]
        st = {str(L): random_walk_tracks(n, L, d, rng, Ds=Ds) for L, n in c["buckets"]}
        params = trk.generate_params(nb_states=nS, LocErr_type=1, nb_dims=d, estimated_LocErr=[0.02], estimated_Ds=Ds,
                                     estimated_Fs=[1 / nS] * (nS - 1), estimated_transition_rates=0.1)
        with contextlib.redirect_stdout(io.StringIO()):
            preds = trk.predict_Bs(st, 0.02, params, cell_dims=[1], nb_states=nS, frame_len=c["fl"], nb_max=c["nb_max"])
        pv = {k: float(params[k].value) for k in params}
        keys = sorted(st, key=int)
        assert all(np.all(np.isfinite(preds[k])) for k in keys)
        np.savez_compressed(os.path.join(HERE, c["name"] + ".npz"), nS=nS, fl=c["fl"], nb_max=c["nb_max"], keys=np.array(keys),
                            param_names=np.array(list(pv)), param_values=np.array(list(pv.values())), meta=str(meta),
                            **{"C" + k: st[k] for k in keys}, **{"P" + k: preds[k] for k in keys})
        print(c["name"], {k: preds[k].shape for k in keys}, float(np.mean(preds[keys[0]][..., 0])))


if __name__ == "__main__":
    main()
