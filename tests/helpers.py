"""Shared test helpers: seeded synthetic chunks, model construction for oracle and engine."""
from __future__ import annotations

import numpy as np

from oracle import extrack_oracle as orc


def random_walk_tracks(n, L, d, rng, Ds=(1e-5, 0.25), loc_err=0.02, dt=0.02, p_switch=0.1):
    """Cheap multi-state random walks with localisation noise (for parity tests)."""
    Ds = np.asarray(Ds, dtype=float)
    nS = len(Ds)
    states = np.empty((n, L), dtype=int)
    states[:, 0] = rng.integers(0, nS, size=n)
    flip = rng.random((n, L)) < p_switch
    new = rng.integers(0, nS, size=(n, L))
    for k in range(1, L):
        states[:, k] = np.where(flip[:, k], new[:, k], states[:, k - 1])
    steps = rng.normal(size=(n, L, d)) * np.sqrt(2 * Ds[states] * dt)[:, :, None]
    pos = np.cumsum(steps, 1) + rng.random((n, 1, d))
    return pos + rng.normal(size=(n, L, d)) * loc_err


def make_model(nS=2, nsub=1, loc_err=(0.02,), frame_len=6, min_len=3, threshold=0.2, max_nb_states=120,
               pBL=0.05, cell_dims=(1.0,), dt=0.02, Ds=None, Fs=None, rates=0.1, int8_wrap=True):
    """(LocErr, ds, Fs, TrMat, pBL) the way extract_params produces them + the oracle Model."""
    if Ds is None:
        Ds = [1e-5, 0.25] if nS == 2 else list(np.linspace(1e-5, 0.25, nS) ** 1.0)
        if nS == 3:
            Ds = [1e-5, 0.04, 0.25]
    Ds = np.asarray(Ds, dtype=float)
    if Fs is None:
        Fs = np.full(nS, 1.0 / nS) if nS != 2 else np.array([0.6, 0.4])
    ds = np.sqrt(2 * Ds * dt)
    R = np.full((nS, nS), float(rates)) / nsub
    Tr = 1 - np.exp(-R)
    Tr[np.arange(nS), np.arange(nS)] = 0
    Tr[np.arange(nS), np.arange(nS)] = 1 - Tr.sum(1)
    model = orc.Model(np.asarray(loc_err, dtype=float), ds, np.asarray(Fs, dtype=float), Tr, pBL, list(cell_dims), nsub,
                      frame_len, min_len, threshold, max_nb_states, int8_wrap)
    return model


def engine_params(model, nb_dims):
    from extrack_b200 import tracking as xt

    return xt.build_tables(model.loc_err, model.ds, model.Fs, model.TrMat, model.pBL, model.cell_dims, model.nb_substeps,
                           model.frame_len, model.min_len, model.threshold, model.max_nb_states, nb_dims, model.int8_wrap)


def gid_from_groups(groups, nB):
    gid = np.full(nB, -1, dtype=np.int32)
    for g, mem in enumerate(groups):
        gid[mem] = g
    return gid


def load_var_case(path):
    """Golden case with peak-wise LocErr / per-track dt (tests/golden/make_golden.py:make_var_cases):
    tracks, optional sigma / dt lists, lmfit-style parameters, configuration and reference outputs."""
    from extrack_b200._lmfit_compat import Parameters

    z = np.load(path, allow_pickle=False)
    nb = len([k for k in z.files if k.startswith("C")])
    st = [z["C%d" % i] for i in range(nb)]
    il = [z["S%d" % i] for i in range(nb)] if int(z["kloc"]) else None
    dts = [z["T%d" % i] for i in range(nb)] if int(z["var_dt"]) else None
    preds = [z["P%d" % i] for i in range(nb)] if "P0" in z.files else None
    params = Parameters()
    for k, v in zip(z["param_names"], z["param_values"]):
        params.add(str(k), value=float(v))
    cfg = dict(nS=int(z["nS"]), nsub=int(z["nsub"]), fl=int(z["fl"]), chunk=int(z["chunk"]), slope=int(z["slope"]),
               neglogl=float(z["neglogl"]))
    return st, il, dts, params, preds, cfg


def var_oracle_inputs(st, il, dts, params, cfg, threshold=0.2, max_nb_states=120, nsub=None):
    """Oracle model + per-bucket sigma / ds arrays from the same parameters (restating extract_params)."""
    nS = cfg["nS"]
    nsub = cfg["nsub"] if nsub is None else nsub
    Ds = np.array([params["D%d" % i].value for i in range(nS)])
    Fs = np.array([params["F%d" % i].value for i in range(nS)])
    R = np.zeros((nS, nS))
    for i in range(nS):
        for j in range(nS):
            if i != j:
                R[i, j] = params["p%d%d" % (i, j)].value
    Tr = 1 - np.exp(-R / nsub)
    Tr[np.arange(nS), np.arange(nS)] = 0
    Tr[np.arange(nS), np.arange(nS)] = 1 - Tr.sum(1)
    sigs = None
    if il is not None:
        sigs = il
        if cfg["slope"]:
            sigs = [np.clip(a * params["slope_LocErr"].value + params["offset_LocErr"].value, 0.000001, np.inf) for a in il]
    loc = np.array([params["LocErr"].value]) if "LocErr" in params.keys() else np.array([0.02])
    if dts is not None:
        ds_list = [np.sqrt(2 * Ds[None, None] * t[:, :, None]) for t in dts]
        ds = np.median(ds_list[0], axis=(0, 1))
    else:
        ds_list = None
        ds = np.sqrt(2 * Ds * 0.02)
    model = orc.Model(loc, ds, Fs, Tr, params["pBL"].value, [1], nsub, cfg["fl"], st[0].shape[1], threshold, max_nb_states)
    return model, sigs, ds_list
