"""Table reader producing the length-bucketed ``all_tracks`` dictionary the fit / predict API takes.

Mirror of ``extrack.readers.read_table`` (``readers.py:101-221``) for
tabular files (csv, any separator, pickled DataFrames): same arguments, same bucketing rules and
the same ``(tracks, frames, opt_metrics)`` return value, so that a script can switch packages
without touching its data loading.  The reference walks a pandas ``groupby`` track by track in
Python; here the table is sorted once and every rule is evaluated on whole arrays (segment
reductions), which keeps 10^6-track tables in the seconds range.

Rules restated from the reference (in its order):

* peaks are grouped by track ID (groups in sorted ID order) and sorted by frame inside a track
  (``:173-176``);
* ``remove_no_disp``: a track is dropped when more than 5 % of its per-dimension displacements are
  exactly zero (``:178-180``);
* a track is kept only if its first frame lies inside ``frames_boundaries`` (``:182``) and no step is
  longer than ``dist_th`` (``:183``);
* length in ``lengths`` -> that bucket (``:185-190``); longer than ``max(lengths)`` -> truncated into
  the largest bucket (``:192-197``); between ``min`` and ``max`` but not listed -> truncated to
  ``lengths[argmin(floor(len / lengths)) - 1]`` (``:199-203``; optional metrics are not collected for
  these tracks in the reference either);
* buckets are ``float64[n, l, d]`` arrays keyed by ``str(l)`` in the order of ``lengths``, empty
  buckets are removed, and the non-empty lengths are printed (``:207-218``).

Peaks without a track ID (``'None'`` / NaN) are dropped (the reference turns them into one-peak
tracks when the IDs are integers, ``:156-160``; such tracks only matter if ``1`` is in ``lengths``).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np


def _load(path, fmt):
    import pandas as pd

    if fmt == "csv":
        return pd.read_csv(path, sep=",")
    if fmt == "pkl":
        return pd.read_pickle(path)
    return pd.read_csv(path, sep=fmt)


def read_table(paths, lengths=np.arange(5, 40), dist_th=np.inf, frames_boundaries=[-np.inf, np.inf], fmt="csv",
               colnames=["POSITION_X", "POSITION_Y", "FRAME", "TRACK_ID"], opt_colnames=[], remove_no_disp=True
               ) -> Tuple[Dict[str, np.ndarray], Dict[str, np.ndarray], Dict[str, Dict[str, np.ndarray]]]:
    """Read one or several tables of localisations into ``{str(length): float64[n, length, d]}``.

    ``colnames`` = coordinate columns, frame column, track-ID column (or a list of columns that
    together identify a track); ``opt_colnames`` = additional per-peak columns to collect.
    Returns ``(tracks, frames, opt_metrics)`` like the reference.
    """
    import pandas as pd

    if isinstance(paths, (str, np.str_)):
        paths = [paths]
    colnames = list(colnames)
    lengths = np.asarray(lengths)
    nb_dims = len(colnames) - 2
    per_len: Dict[int, List[np.ndarray]] = {int(l): [] for l in lengths}
    per_len_fr: Dict[int, List[np.ndarray]] = {int(l): [] for l in lengths}
    per_len_opt: Dict[str, Dict[int, List[np.ndarray]]] = {m: {int(l): [] for l in lengths} for m in opt_colnames}
    lmax, lmin = int(np.max(lengths)), int(np.min(lengths))
    for path in paths:
        data = _load(path, fmt)
        for col, what in zip(colnames[:2] + [colnames[nb_dims]], ("x", "y", "frame")):
            if not pd.api.types.is_numeric_dtype(data.dtypes[col]):
                raise ValueError("The %s values are not numerical. Verify the presence of non numerical values in the file. "
                                 "In particular, verify that the file contains only one row of headers" % what)
        idcol = colnames[-1]
        if not isinstance(idcol, (str, np.str_)):  # several columns identify a track
            bad = np.zeros(len(data), dtype=bool)
            for c in idcol:
                bad |= (data[c].astype(str) == "None").values | pd.isna(data[c]).values
            data = data[~bad]
            key = data[idcol[0]].astype(str)
            for c in idcol[1:]:
                key = key + "_" + data[c].astype(str)
            ids = key.values
        else:
            bad = (data[idcol].astype(str) == "None").values | pd.isna(data[idcol]).values
            data = data[~bad]
            ids = data[idcol].values
        if len(data) == 0:
            continue
        codes, _ = pd.factorize(ids, sort=True)  # group order of DataFrame.groupby (sorted keys)
        xyz = data[colnames[:nb_dims]].values.astype("float64")
        fr = data[colnames[nb_dims]].values.astype("float64")
        order = np.lexsort((fr, codes))  # stable: by track, then by frame
        codes, xyz, fr = codes[order], xyz[order], fr[order]
        opt = {m: data[m].values[order] for m in opt_colnames}
        starts = np.flatnonzero(np.r_[True, codes[1:] != codes[:-1]])
        lens = np.diff(np.r_[starts, len(codes)])
        # per-track statistics of the displacements (pairs inside a track only)
        d2 = (xyz[1:] - xyz[:-1]) ** 2
        inside = codes[1:] == codes[:-1]
        pair_track = np.cumsum(np.r_[True, codes[1:] != codes[:-1]])[1:] - 1  # track index of the pair's second peak
        n_tracks = len(starts)
        zeros = np.bincount(pair_track[inside], weights=(d2[inside] == 0).sum(1), minlength=n_tracks)
        with np.errstate(invalid="ignore", divide="ignore"):
            zero_frac = zeros / ((lens - 1) * nb_dims)  # nan for one-peak tracks, as np.mean of an empty array
        dist = np.sum(d2, axis=1) ** 0.5
        too_far = np.zeros(n_tracks, dtype=bool)
        np.logical_or.at(too_far, pair_track[inside], dist[inside] > dist_th)
        keep = (fr[starts] >= frames_boundaries[0]) & (fr[starts] <= frames_boundaries[1]) & ~too_far
        if remove_no_disp:
            keep &= ~(zero_frac > 0.05)
        # target bucket of every track (0 = none)
        target = np.zeros(n_tracks, dtype=np.int64)
        exact = np.isin(lens, lengths)
        target[exact] = lens[exact]
        longer = ~exact & (lens > lmax)
        target[longer] = lmax
        between = ~exact & ~longer & (lens < lmax) & (lens > lmin)
        if between.any():
            l_idx = np.argmin(np.floor(lens[between][:, None] / lengths[None, :]), axis=1) - 1
            target[between] = lengths[l_idx]
        target[~keep] = 0
        for l in np.unique(target[target > 0]):
            sel = target == l
            idx = starts[sel][:, None] + np.arange(int(l))[None, :]
            per_len[int(l)].append(xyz[idx])
            per_len_fr[int(l)].append(fr[idx])
            has_opt = sel & ~between  # the reference collects no optional metrics for the in-between tracks
            if opt_colnames and has_opt.any():
                idx_o = starts[has_opt][:, None] + np.arange(int(l))[None, :]
                for m in opt_colnames:
                    per_len_opt[m][int(l)].append(opt[m][idx_o])
    tracks: Dict[str, np.ndarray] = {}
    frames: Dict[str, np.ndarray] = {}
    opt_metrics: Dict[str, Dict[str, np.ndarray]] = {m: {} for m in opt_colnames}
    for l in lengths:
        l = int(l)
        if per_len[l]:
            print(l)
            tracks[str(l)] = np.concatenate(per_len[l])
            frames[str(l)] = np.concatenate(per_len_fr[l])
            for m in opt_colnames:
                opt_metrics[m][str(l)] = np.concatenate(per_len_opt[m][l]) if per_len_opt[m][l] else np.array([])
    return tracks, frames, opt_metrics


def read_trackmate_xml(paths, lengths=np.arange(5, 40), dist_th=0.5, frames_boundaries=[-np.inf, np.inf], remove_no_disp=True,
                       opt_metrics_names=["t", "x"], opt_metrics_types=[int, "float64"]
                       ) -> Tuple[Dict[str, np.ndarray], Dict[str, np.ndarray], Dict[str, Dict[str, np.ndarray]]]:
    """TrackMate "export tracks to XML" files -> ``(tracks, frames, opt_metrics)`` with the rules of
    ``extrack.readers.read_trackmate_xml`` (``readers.py:5-98``; the reference parses with ``xmltodict``, absent from this
    image: ``xml.etree`` here).

    Per ``<particle>`` (in file order): the ``x`` / ``y`` / ``t`` attributes of its ``<detection>`` elements (``:53``);
    ``remove_no_disp`` drops the track if any x-displacement or any y-displacement is exactly zero (``:60-63``: the
    product of the minimal squared displacements); it is kept if its first frame lies inside ``frames_boundaries`` and
    every step is shorter than ``dist_th`` (``:65-66``); length in ``lengths`` -> that bucket, longer than
    ``max(lengths)`` -> truncated into the largest bucket (``:68-79``), other lengths are dropped.  ``frames`` holds
    the frame numbers as floats (they pass through the reference's float track array), ``opt_metrics[name][str(l)]``
    the detection attribute ``name`` cast to ``opt_metrics_types`` (``:84-93``).  Empty buckets are removed.  A particle
    the reference cannot walk (fewer than two detections) raises the same ``ValueError`` (``:80-81``)."""
    import xml.etree.ElementTree as ET

    if isinstance(paths, (str, np.str_)):
        paths = [paths]
    lengths = np.asarray(lengths)
    names = list(opt_metrics_names)
    types = ["float64"] * len(names) if opt_metrics_types is None else list(opt_metrics_types)
    keys = [str(l) for l in lengths]
    traces: Dict[str, list] = {k: [] for k in keys}
    frames: Dict[str, list] = {k: [] for k in keys}
    opt: Dict[str, Dict[str, list]] = {m: {k: [] for k in keys} for m in names}
    lmax = int(np.max(lengths))
    for path in paths:
        root = ET.parse(path).getroot()
        framerate = float(root.attrib["frameInterval"]) / 1000.0  # (kept for parity of the failure modes: the attribute must exist)
        del framerate
        for particle in root.iter("particle"):
            dets = particle.findall("detection")
            try:
                if len(dets) < 2:
                    raise ValueError("a particle needs at least two detections")
                xy = np.array([(float(d.attrib["x"]), float(d.attrib["y"])) for d in dets])
                fr = np.array([float(int(d.attrib["t"])) for d in dets])
                met = np.empty((int(particle.attrib["nSpots"]), len(names)), dtype=object)
                for k, d in enumerate(dets):
                    for j, m in enumerate(names):
                        met[k, j] = d.attrib[m]
            except Exception:
                raise ValueError("problem with data on path: " + path)
            step = xy[1:] - xy[:-1]
            if remove_no_disp and np.min(step[:, 0] ** 2) * np.min(step[:, 1] ** 2) == 0:
                continue
            dists = np.sum(step**2, axis=1) ** 0.5
            if not (fr[0] >= frames_boundaries[0] and fr[0] <= frames_boundaries[1] and np.all(dists < dist_th)):
                continue
            l = len(xy)
            if np.any(lengths == l):
                key, cut = str(l), l
            elif l > lmax:
                key, cut = str(lmax), lmax
            else:
                continue
            traces[key].append(xy[:cut])
            frames[key].append(fr[:cut])
            for j, m in enumerate(names):
                opt[m][key].append(met[:cut, j])
    for k in keys:
        if len(traces[k]) > 0:
            traces[k] = np.array(traces[k])
            frames[k] = np.array(frames[k])
            for j, m in enumerate(names):
                cur = np.array(opt[m][k])
                try:
                    cur = cur.astype(types[j])
                except Exception:
                    print("Error of type with the optional metric:", m)
                opt[m][k] = cur
        else:
            del traces[k], frames[k]
            for m in names:
                del opt[m][k]
    return traces, frames, opt
