"""Vectorised synthetic-track generator with the statistics of ``extrack.simulate_tracks.sim_FOV``.

The reference generator (``simulate_tracks.py:123-244``) loops over trajectories in Python
(~1.3 k kept tracks/s/core) and lives in the reference tree, which does not exist on the GPU
box.  This module re-implements the same stochastic process with tensor operations (torch, on
CPU or on a CUDA device) so that benchmark inputs of 10^6-10^7 tracks are generated in seconds.
It is input tooling, not part of the likelihood path.

Process (same as the reference): a Markov chain over diffusive states at ``nb_sub_steps``
sub-steps per frame with off-diagonal rates ``TrMat/nb_sub_steps``; Brownian increments with
variance ``2 D[state] dt/nb_sub_steps`` per sub-step; start positions uniform in
``[-cell, cell]``; positions sampled at the first sub-step of each frame; a track is a maximal
run of frames inside the field of view ``0 < x_i < cell_dims[i]``; each in-view localisation
bleaches with probability ``pBL`` (the run is cut after it and the trajectory is dropped);
Gaussian localisation error of standard deviation ``LocErr`` is added; runs with
``min_track_len <= length <= max_track_len`` are kept and bucketed by length.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch


def sim_FOV(nb_tracks=10000, max_track_len=40, min_track_len=2, LocErr=0.02, Ds=(0, 0.05), nb_dims=2,
            initial_fractions=(0.6, 0.4), TrMat=((0.9, 0.1), (0.1, 0.9)), LocErr_std=0, dt=0.02, pBL=0.1,
            cell_dims=(0.5, None, None), seed: Optional[int] = 0, device: str = "cpu", nb_sub_steps: int = 20,
            return_states: bool = False) -> Tuple[Dict[str, np.ndarray], Dict[str, np.ndarray]]:
    """Returns ``(all_tracks, all_states)`` keyed by ``str(length)``: ``float64[n, L, nb_dims]`` / ``int[n, L]``.

    ``LocErr_std != 0`` (simulate_tracks.py:131,207-209): every localisation gets its own error standard deviation
    ``sigma = LocErr * chi2(k) / k`` with ``k = 2 / LocErr_std**2`` (mean ``LocErr``, relative spread ``LocErr_std``),
    the noise is drawn with it, and a third dictionary ``{str(length): float64[n, L, nb_dims]}`` of these sigmas is
    returned (what the reference returns as its third value and ``param_fitting(input_LocErr=...)`` consumes)."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    if seed is not None:
        gen.manual_seed(int(seed))
    f64 = torch.float64
    Ds_t = torch.tensor(list(Ds), dtype=f64, device=dev)
    Tr = torch.tensor(np.array(TrMat, dtype=float), dtype=f64, device=dev)
    nS = Tr.shape[0]
    sub = Tr / nb_sub_steps
    sub[torch.arange(nS), torch.arange(nS)] = 0
    sub[torch.arange(nS), torch.arange(nS)] = 1 - sub.sum(1)
    cum = torch.cumsum(sub, 1)
    cumF = torch.cumsum(torch.tensor(list(initial_fractions), dtype=f64, device=dev), 0)
    cell = [c for c in cell_dims]
    n_lim = sum(c is not None for c in cell)
    N = (2**n_lim) * int(nb_tracks)
    cell0 = torch.tensor([1.0 if c is None else float(c) for c in cell], dtype=f64, device=dev)
    sub_dt = dt / nb_sub_steps

    # Markov chain at sub-step resolution; accumulate per-frame displacement variance on the fly
    u = torch.rand(N, generator=gen, device=dev, dtype=f64)
    state = (u[:, None] > cumF[None, : nS - 1]).sum(1)
    frame_state = torch.empty((N, max_track_len), dtype=torch.int64, device=dev)
    var = torch.zeros((N, max_track_len), dtype=f64, device=dev)  # variance of increment INTO frame k
    for k in range(max_track_len):
        frame_state[:, k] = state
        if k == max_track_len - 1:
            break
        acc = torch.zeros(N, dtype=f64, device=dev)
        for _ in range(nb_sub_steps):
            acc += 2 * Ds_t[state] * sub_dt
            u = torch.rand(N, generator=gen, device=dev, dtype=f64)
            state = (u[:, None] > cum[state][:, : nS - 1]).sum(1)
        var[:, k + 1] = acc
    start = (2 * torch.rand((N, 1, 3), generator=gen, device=dev, dtype=f64) - 1) * cell0
    incr = torch.randn((N, max_track_len, 3), generator=gen, device=dev, dtype=f64) * var.sqrt()[:, :, None]
    pos = start + torch.cumsum(incr, 1)

    inside = torch.ones((N, max_track_len), dtype=torch.bool, device=dev)
    for i, c in enumerate(cell):
        if c is not None:
            inside &= (pos[:, :, i] < c) & (pos[:, :, i] > 0)

    # run extraction, frame by frame (vector state machine over trajectories)
    alive = torch.ones(N, dtype=torch.bool, device=dev)
    in_run = torch.zeros(N, dtype=torch.bool, device=dev)
    run_start = torch.zeros(N, dtype=torch.int64, device=dev)
    rec_traj, rec_start, rec_len = [], [], []
    ar = torch.arange(N, device=dev)

    def emit(mask, end_excl):
        if mask.any():
            rec_traj.append(ar[mask])
            rec_start.append(run_start[mask])
            rec_len.append(end_excl - run_start[mask])

    for k in range(max_track_len):
        vis = inside[:, k] & alive
        ended = in_run & ~vis & alive
        emit(ended, torch.full((N,), k, dtype=torch.int64, device=dev)[ended])
        in_run = in_run & ~ended
        starting = vis & ~in_run
        run_start = torch.where(starting, torch.full_like(run_start, k), run_start)
        in_run = in_run | starting
        bleach = vis & (torch.rand(N, generator=gen, device=dev, dtype=f64) < pBL)
        emit(bleach, torch.full((N,), k + 1, dtype=torch.int64, device=dev)[bleach])
        in_run = in_run & ~bleach
        alive = alive & ~bleach
    emit(in_run & alive, torch.full((N,), max_track_len, dtype=torch.int64, device=dev)[in_run & alive])

    all_tracks: Dict[str, np.ndarray] = {}
    all_states: Dict[str, np.ndarray] = {}
    all_sigmas: Dict[str, np.ndarray] = {}
    if not rec_traj:
        return (all_tracks, all_states, all_sigmas) if LocErr_std != 0 else (all_tracks, all_states)
    traj = torch.cat(rec_traj)
    st = torch.cat(rec_start)
    ln = torch.cat(rec_len)
    order = torch.argsort(traj * max_track_len + st)  # trajectory-major, like the reference
    traj, st, ln = traj[order], st[order], ln[order]
    for Lk in range(min_track_len, max_track_len + 1):
        sel = ln == Lk
        n = int(sel.sum())
        if n == 0:
            continue
        idx = st[sel][:, None] + torch.arange(Lk, device=dev)[None, :]
        p = pos[traj[sel][:, None], idx]  # [n, L, 3]
        if LocErr_std != 0:
            kk = 2.0 / (float(LocErr_std) ** 2 + 1e-20)  # chi2(k) / k: mean 1, standard deviation LocErr_std
            loc = torch.as_tensor(np.asarray(LocErr, dtype=float).reshape(-1), dtype=f64, device=dev)  # scalar or per dimension
            if loc.numel() not in (1, 3):
                loc = torch.cat([loc, loc[-1:].expand(3 - loc.numel())])
            gam = torch._standard_gamma(torch.full(p.shape, kk / 2.0, dtype=f64, device=dev), generator=gen)
            sig = (2.0 * gam / kk) * loc
            p = p + sig * torch.randn(p.shape, generator=gen, device=dev, dtype=f64)
            all_sigmas[str(Lk)] = sig[:, :, :nb_dims].contiguous().cpu().numpy()
        else:
            p = p + LocErr * torch.randn(p.shape, generator=gen, device=dev, dtype=f64)
        all_tracks[str(Lk)] = p[:, :, :nb_dims].contiguous().cpu().numpy()
        if return_states:
            all_states[str(Lk)] = frame_state[traj[sel][:, None], idx].cpu().numpy()
    if LocErr_std != 0:
        return all_tracks, all_states, all_sigmas
    return all_tracks, all_states


def sim_tracks(n_target: int, block: int = 20000, seed: int = 0, device: str = "cpu", **kw) -> Dict[str, np.ndarray]:
    """Concatenate ``sim_FOV`` blocks (seeds seed, seed+1, ...) per bucket until ``n_target`` tracks exist,
    then trim the last block so the total is exactly ``n_target`` (SURVEY.md §8d config 2 protocol)."""
    parts: Dict[str, list] = {}
    total = 0
    s = seed
    while total < n_target:
        tr, _ = sim_FOV(nb_tracks=block, seed=s, device=device, **kw)
        got = sum(len(v) for v in tr.values())
        if got == 0:
            raise RuntimeError("generator produced no tracks with these settings")
        if total + got > n_target:  # trim proportionally, bucket by bucket
            keep = n_target - total
            frac = keep / got
            trimmed, acc = {}, 0
            keys = sorted(tr, key=int)
            for i, k in enumerate(keys):
                m = int(round(len(tr[k]) * frac)) if i < len(keys) - 1 else keep - acc
                m = max(0, min(m, len(tr[k]), keep - acc))
                acc += m
                if m:
                    trimmed[k] = tr[k][:m]
            tr, got = trimmed, acc
        for k, v in tr.items():
            parts.setdefault(k, []).append(v)
        total += got
        s += 1
    return {k: np.concatenate(v) for k, v in sorted(parts.items(), key=lambda kv: int(kv[0]))}
