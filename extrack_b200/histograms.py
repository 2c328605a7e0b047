"""Segment-length histograms on the GPU — host mirror of ``extrack/histograms.py``.

Same call signatures as the reference (``P_segment_len`` :26, ``len_hist`` :265); the arithmetic runs
in the CUDA engine (``csrc/xt_seglen.cuh`` through ``xt_seglen_hist``), there is no CPU path.

Differences that are documented rather than hidden:

* ``nb_substeps`` must be 1 and ``input_LocErr`` (peak-wise localisation errors) is not supported:
  ``NotImplementedError`` / ``ValueError`` from the engine;
* equal sort keys in the top-``max_nb_states`` pruning (histograms.py:192-193 uses numpy's unstable
  default ``argsort``, so the reference's order among them is unspecified) are ordered by descending
  index; keys whose ``exp`` underflows to zero in the reference keep the order of their keys here.

The reference's rescale of final log-probabilities above 600 (per column over the tracks of a chunk,
:243-244) is reproduced on the device (one CTA per chunk).
"""
from __future__ import annotations

from typing import Sequence

import numpy as np

from . import _native
from . import tracking as _trk

NB_MAX = 50  # histograms.py:319: tracks per chunk


def _tables(LocErr, ds, Fs, TrMat, pBL, cell_dims, nb_substeps, min_l, max_nb_states, nb_dims):
    if int(nb_substeps) != 1:
        raise NotImplementedError("segment-length histogram: nb_substeps must be 1 on the GPU path")
    cd = [c for c in np.asarray(cell_dims, dtype=object).reshape(-1) if c is not None]  # histograms.py:60-61
    ds = np.asarray(ds, dtype=float)
    nS = len(ds)
    p = _trk.build_tables(np.asarray(LocErr, dtype=float).reshape(-1), ds, Fs, TrMat, pBL, cd, 1, 1, int(min_l), 0.2,
                          int(max_nb_states), nb_dims)
    # leave term per head = newest + nS * previous (histograms.py:224-232): p_stay[s] only if both are s, else p_stay[0]
    p_stay = _trk._p_stay(ds, nS, 1, cd)
    new, prev = np.arange(nS * nS) % nS, np.arange(nS * nS) // nS
    e = np.where(new == prev, p_stay[prev], p_stay[0])
    leave = np.log(pBL + (1 - e) - pBL * (1 - e))
    return p, leave, nS


def P_segment_len(Cs, LocErr, ds, Fs, TrMat, min_l=3, pBL=0.1, isBL=1, cell_dims=[0.5], nb_substeps=1, max_nb_states=1000):
    """One chunk ``Cs[nT, L, d]`` -> ``(LP[nT, nB], cur_Bs[nT, nB, L], seg_len_hist[L-1, nS])`` (histograms.py:26-258)."""
    Cs = np.ascontiguousarray(Cs, dtype=np.float64)
    if Cs.ndim != 3 or Cs.shape[1] < 2:
        raise ValueError("Cs must have shape [nb_tracks, nb_locs >= 2, nb_dims]")
    LocErr = np.asarray(LocErr, dtype=float)
    if LocErr.ndim == 3 and (LocErr.shape[0] != 1 or LocErr.shape[1] != 1):
        raise NotImplementedError("segment-length histogram: peak-wise localisation errors are not supported on the GPU path")
    nT, L, d = Cs.shape
    p, leave, nS = _tables(LocErr, ds, Fs, TrMat, pBL, cell_dims, nb_substeps, min_l, max_nb_states, d)
    eng = _native.Engine(_trk._default_device())
    try:
        eng.upload([Cs], [int(bool(isBL))], nT)
        hist, LP, Bs = eng.seglen_hist(p, leave, L, nS, 1, dbg_chunk=0, dbg_shape=(nT, L))
    finally:
        eng.close()
    return LP, Bs.astype(int), hist[0, :L - 1]


def len_hist(all_tracks, params, dt, cell_dims=[0.5, None, None], nb_states=2, max_nb_states=500, workers=1, nb_substeps=1,
             input_LocErr=None, _timing: dict = None):
    """Sum of the per-chunk histograms over all length buckets -> ``[L_longest, nb_states]`` (histograms.py:265-373).

    ``workers`` is accepted and ignored (the reference's process-pool width)."""
    if input_LocErr is not None:
        raise NotImplementedError("len_hist: input_LocErr (peak-wise localisation errors) is not supported on the GPU path")
    if isinstance(all_tracks, dict):
        min_l = int(np.min(np.array(list(all_tracks.keys())).astype(int)))
        tracks: Sequence[np.ndarray] = [all_tracks[k] for k in all_tracks]   # dict order, like the reference
    else:
        tracks = list(all_tracks)
        min_l = int(min(a.shape[1] for a in tracks))
    tracks = [np.ascontiguousarray(a, dtype=np.float64) for a in tracks]
    keep = [i for i, a in enumerate(tracks) if len(a)]
    LocErr, ds, Fs, TrMat, pBL = _trk.extract_params(params, dt, nb_states, nb_substeps, None)
    nb_dims = tracks[0].shape[2]
    p, leave, nS = _tables(np.asarray(LocErr).reshape(-1), ds, Fs, TrMat, pBL, cell_dims, nb_substeps, min_l, max_nb_states, nb_dims)
    Lout = tracks[-1].shape[1]
    segs = [tracks[i] for i in keep]
    bl = [0 if i == len(tracks) - 1 else 1 for i in keep]   # the last bucket did not disappear (histograms.py:312-315)
    if max(a.shape[1] for a in segs) > Lout:
        raise ValueError("len_hist: the last length bucket must be the longest (the histogram has its number of rows)")
    n_chunks = int(sum(-(-len(a) // NB_MAX) for a in segs))
    print("number of chunks:", n_chunks)
    eng = _native.Engine(_trk._default_device())
    try:
        eng.upload(segs, bl, NB_MAX)
        hist = eng.seglen_hist(p, leave, Lout, nS, n_chunks)
        if _timing is not None:
            _timing["kernel_ms"] = eng.seglen_last_ms()
    finally:
        eng.close()
    print("")
    return hist.sum(axis=0)
