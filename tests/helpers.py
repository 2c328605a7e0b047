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
