"""Drop-in mirror of ``extrack.refined_localization.position_refinement`` (refined_localization.py:304-338) on the
sm_100a CUDA engine: refined positions of every localisation (weighted mean of the per-sequence Gaussian products of
the two passes of the recursion, ``get_LC_Km_Ks`` :48-204 and ``get_pos_PDF`` :207-298) and their standard deviations.

Same arguments, outputs and prints as the reference; every length bucket is handed to the engine as one chunk, as the
reference hands whole buckets to ``get_LC_Km_Ks`` (the grouping plan comes from the bucket's first 30 tracks).
Peak-wise localisation errors (``LocErr`` as a dict) are not provided by the engine (``NotImplementedError``).
There is no CPU fallback.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

from . import _native
from . import tracking as _trk


def position_refinement(all_tracks, LocErr, ds, Fs, TrMat, frame_len=7, threshold=0.1, max_nb_states=1000) -> Tuple[Dict, Dict]:
    """``({l: mu [n, l, d]}, {l: sigma [n, l]})`` for ``all_tracks = {l: float64 [n, l, d]}``."""
    if type(LocErr) == float or type(LocErr) == np.float64 or type(LocErr) == np.float32:
        loc = np.array([float(LocErr)])
        LocErr_type = "array"
    elif type(LocErr) == np.ndarray:
        LocErr_type = "array"
        loc = np.asarray(LocErr, dtype=np.float64).reshape(-1)
    elif type(LocErr) == dict:
        LocErr_type = "dict"
    else:
        LocErr_type = "other"
    print("LocErr_type", LocErr_type)
    if LocErr_type == "dict":
        raise NotImplementedError("position_refinement on the CUDA engine takes one localisation error (or one per dimension); "
                                  "peak-wise errors (a dict) are not implemented")
    if LocErr_type == "other":
        raise ValueError("LocErr must be a float, an array or a dict")
    keys = [l for l in all_tracks.keys()]
    all_mus = {l: np.zeros((len(all_tracks[l]), int(l), np.asarray(all_tracks[l]).shape[2])) for l in keys}
    all_sigmas = {l: np.zeros((len(all_tracks[l]), int(l))) for l in keys}
    live = [l for l in keys if len(all_tracks[l]) > 0]
    if not live:
        return all_mus, all_sigmas
    nb_dims = np.asarray(all_tracks[live[0]]).shape[2]
    if len(loc) not in (1, nb_dims):
        raise ValueError("Localization error is not specified correctly: a float, or one value per dimension")
    for l in live:
        if int(l) < 2:
            raise ValueError("minimal track length = 2, here track length = %s" % l)
    ds = np.asarray(ds, dtype=np.float64)
    Fs = np.asarray(Fs, dtype=np.float64)
    TrMat = np.asarray(TrMat, dtype=np.float64)
    nS = len(ds)
    kw = dict(pBL=0.1, cell_dims=[1.0], nb_substeps=1, frame_len=int(frame_len), min_len=2, threshold=float(threshold),
              max_nb_states=int(max_nb_states), nb_dims=nb_dims)
    # pass 1 consumes a track from its last to its first localisation (get_pos_PDF's first call); pass 2 runs in forward
    # time with the transposed transition matrix and neutral fractions (refined_localization.py:214-218)
    p_rev = _trk.build_tables(loc, ds, Fs, TrMat, **kw)
    p_fwd = _trk.build_tables(loc, ds, np.ones(nS) / nS, TrMat.T.copy(), **kw)
    eng = _native.Engine(_trk._default_device())
    try:
        segs = [np.ascontiguousarray(all_tracks[l], dtype=np.float64) for l in live]
        eng.upload(segs, [0] * len(segs), max(len(s) for s in segs))  # one chunk per bucket
        mus, sigmas = eng.refine_positions(p_rev, p_fwd)
    finally:
        eng.close()
    for l, m, s in zip(live, mus, sigmas):
        all_mus[l], all_sigmas[l] = m, s
    return all_mus, all_sigmas
