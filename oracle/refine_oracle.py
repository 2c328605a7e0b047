"""TEST INFRASTRUCTURE ONLY — numpy restatement of ExTrack's position refinement
(``extrack/refined_localization.py``).  Only ``tests/`` and golden generators may import it.

Pinned: ``tests/golden/make_golden_refine.py`` runs the unmodified reference (``oracle.ref_loader``) and keeps a case
only if this restatement reproduces it (tests/test_oracle.py).

What the reference computes (file:line = refined_localization.py):
  * ``get_LC_Km_Ks`` (:48-204): the recursion of ``P_Cs_inter_bound_stats_th(do_preds=1)`` over a whole length bucket
    as ONE chunk (plan from its first 30 tracks, per-track weighted histories), consuming the localisations from the
    LAST to the first, without field-of-view / bleaching terms, with the initial-fraction term added at the end for the
    newest state; it keeps, for every step, the mean ``Km``, the standard deviation ``Ks`` and the log-weight ``LP``
    of every surviving sequence and the sequence's newest state.  (``all_LP[-1]`` aliases the array the end-of-track
    term is added to in place, :188-194: the last entry carries that term.)
  * ``get_pos_PDF`` (:207-298): one such pass over the track and one over the time-reversed track (transposed
    transition matrix, neutral fractions); for every localisation k the pairs (sequence of pass 1 that has consumed
    localisations k+1.., sequence of pass 2 that has consumed ..k-1) with the same state at k are combined with the
    localisation itself by a product of three Gaussians (:33-43).
  * ``position_refinement`` (:304-338): weighted mean position and standard deviation per localisation.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

from . import extrack_oracle as orc


def lc_km_ks(C: np.ndarray, loc_err, ds, Fs, TrMat, frame_len: int, threshold: float, max_nb_states: int, int8_wrap: bool = True):
    """One pass of ``get_LC_Km_Ks`` (:48-204) with nb_substeps = 1.  C: [nT, L, d].  Returns lists over the steps of
    (Km [nT, nB, d], Ks [nT, nB, k], LP [nT, nB], newest state [nB])."""
    C = np.asarray(C, dtype=np.float64)
    nT, L, d = C.shape
    model = orc.Model(np.asarray(loc_err, dtype=float).reshape(-1), np.asarray(ds, dtype=float), np.asarray(Fs, dtype=float),
                      np.asarray(TrMat, dtype=float), 0.1, [1.0], 1, int(frame_len), 3, float(threshold), int(max_nb_states), int8_wrap)
    tb = orc.HeadTables(model)
    nS, K = model.nS, tb.K
    l2 = (model.loc_err**2)[None, None, :]
    th = float(threshold)
    nB = K * nS
    head = np.arange(nB)
    cur = tb.digits[head, 0].copy()
    hist = (tb.digits[head][None, :, :, None] == np.arange(nS)[None, None, None, :]).astype(np.float64)
    LP = np.repeat(tb.LT[head][None], nT, axis=0)  # no initial-fraction term here (:90-93)
    s2 = l2 + tb.dd[head][None, :, None]
    s2 = np.repeat(s2, nT, axis=0) if s2.shape[0] == 1 else s2  # (:112-113)
    m = np.repeat(C[:, None, L - 1, :], nB, axis=1)
    all_m, all_s, all_LP, all_cur = [m], [s2**0.5], [LP], [cur]
    for step in range(2, L):  # consumes localisation L - step
        n = len(cur) * nS
        lab = orc._label(np.arange(n), nS, int8_wrap)
        new_row = np.repeat((lab[None, :, None, None] == np.arange(nS)[None, None, None, :]).astype(np.float64), hist.shape[0], 0)
        hist = np.concatenate((new_row, np.repeat(hist, nS, 1)), -2)
        child = np.arange(n)
        head = (child % K) + K * cur[child // K]
        cur = child % nS
        dd = tb.dd[head][None, :, None]
        m = np.repeat(m, K, axis=1)
        s2 = np.repeat(s2, K, axis=1)
        LP = np.repeat(LP, K, axis=1)
        c = C[:, None, L - step, :]
        q = l2 + s2
        new_m = (m * l2 + c * s2) / (l2 + s2)
        new_s2 = (dd * l2 + dd * s2 + l2 * s2) / q
        if s2.shape[2] == 1:
            LC = d * -0.5 * np.log(2 * np.pi * q[:, :, 0]) - np.sum((c - m) ** 2 / (2 * q), axis=2)
        else:
            LC = np.sum(-0.5 * np.log(2 * np.pi * q), 2) - np.sum((c - m) ** 2 / (2 * q), axis=2)
        m, s2 = new_m, new_s2
        LP = LP + (tb.LT[head][None] + LC)
        if len(cur) > max_nb_states:
            th = th * 1.2
        if step < L - 1:
            groups = orc._plan_groups(m, s2, hist, int(frame_len), th)
            m, s2, LP, cur, hist = orc._merge(m, s2, LP, cur, hist, groups, nT, int(frame_len), 1)
        all_m.append(m)
        all_s.append(s2**0.5)
        all_LP.append(LP)
        all_cur.append(cur)
    q = s2 + l2
    term = np.sum(-0.5 * np.log(2 * np.pi * q) - (C[:, None, 0, :] - m) ** 2 / (2 * q), axis=2)
    LP = LP + (term + np.log(model.Fs[cur])[None])
    all_LP[-1] = LP  # (:188-194: in the reference the last stored array is the one updated in place)
    return all_m, all_s, all_LP, all_cur


def _prod2(s1, s2, mu1, mu2):
    """refined_localization.py:33-37"""
    v = s1**2 + s2**2
    sigma = ((s1**2 * s2**2) / v) ** 0.5
    mu = (mu1 * s2**2 + mu2 * s1**2) / v
    LK = np.sum(-0.5 * np.log(2 * np.pi * v) - (mu1 - mu2) ** 2 / (2 * v), -1)
    return sigma, mu, LK


def _prod3(s1, s2, s3, mu1, mu2, mu3):
    """refined_localization.py:39-43"""
    sigma, mu, LK = _prod2(s1, s2, mu1, mu2)
    sigma, mu, LK2 = _prod2(sigma, s3, mu, mu3)
    return sigma, mu, LK + LK2


def pos_pdf(C: np.ndarray, loc_err, ds, Fs, TrMat, frame_len: int, threshold: float, max_nb_states: int):
    """``get_pos_PDF`` (:207-298) for scalar / per-dimension LocErr: per localisation (means [nT, n, d],
    stds [nT, n, k], log-weights [nT, n])."""
    C = np.asarray(C, dtype=np.float64)
    nT, L, d = C.shape
    le = np.asarray(loc_err, dtype=float).reshape(1, 1, -1)
    TrMat = np.asarray(TrMat, dtype=float)
    nS = TrMat.shape[0]
    m1, s1, LP1, c1 = lc_km_ks(C, loc_err, ds, Fs, TrMat, frame_len, threshold, max_nb_states)
    m2, s2, LP2, c2 = lc_km_ks(C[:, ::-1], loc_err, ds, np.ones(nS) / nS, TrMat.T.copy(), frame_len, threshold, max_nb_states)
    out = []
    sig, mu, LC = _prod2(le, s1[-1], C[:, None, 0], m1[-1])
    out.append((mu, sig, LP1[-1] + LC))
    for k in range(1, L - 1):
        A_LP, A_m, A_s, A_c = LP1[-1 - k], m1[-1 - k], s1[-1 - k], c1[-1 - k]
        B_LP, B_m, B_s, B_c = LP2[k - 1], m2[k - 1], s2[k - 1], c2[k - 1]
        mus, sigs, lps = [], [], []
        for state in range(nS):
            ia, ib = np.where(A_c == state)[0], np.where(B_c == state)[0]
            sub_sig, sub_mu, sub_LC = _prod3(A_s[:, ia][:, :, None], le[:, None], B_s[:, ib][:, None], A_m[:, ia][:, :, None],
                                             C[:, None, None, k], B_m[:, ib][:, None])
            sub_LP = A_LP[:, ia][:, :, None] + B_LP[:, ib][:, None] + sub_LC
            n = len(ia) * len(ib)
            sigs.append(sub_sig.reshape(nT, n, -1))
            mus.append(sub_mu.reshape(nT, n, d))
            lps.append(sub_LP.reshape(nT, n))
        out.append((np.concatenate(mus, 1), np.concatenate(sigs, 1), np.concatenate(lps, 1)))
    sig, mu, LC = _prod2(le, s2[-1], C[:, None, -1], m2[-1])
    out.append((mu, sig, LP2[-1] + LC))
    return out


def position_refinement(all_tracks: Dict[str, np.ndarray], loc_err, ds, Fs, TrMat, frame_len: int = 7, threshold: float = 0.1,
                        max_nb_states: int = 1000) -> Tuple[Dict[str, np.ndarray], Dict[str, np.ndarray]]:
    """``position_refinement`` (:304-338): ``({l: mu [n, L, d]}, {l: sigma [n, L]})``."""
    all_mus, all_sigmas = {}, {}
    for l, Cs in all_tracks.items():
        Cs = np.asarray(Cs, dtype=np.float64)
        mus = np.zeros((Cs.shape[0], int(l), Cs.shape[2]))
        sigmas = np.zeros((Cs.shape[0], int(l)))
        for k, (mean, std, w) in enumerate(pos_pdf(Cs, loc_err, ds, Fs, TrMat, frame_len, threshold, max_nb_states)):
            P = np.exp(w - np.max(w, 1, keepdims=True))
            mus[:, k] = np.sum(P[:, :, None] * mean, 1) / np.sum(P, 1)[:, None]
            sigmas[:, k] = (np.sum(P * std[:, :, 0] ** 2, 1) / np.sum(P, 1)) ** 0.5
        all_mus[l], all_sigmas[l] = mus, sigmas
    return all_mus, all_sigmas
