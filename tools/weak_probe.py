#!/usr/bin/env python3
"""Weak-scaling step of bench.py in isolation (torchrun): every rank evaluates its own 10^6-track field of view and the
partial sums are all-reduced per step.  A/B of engine options given as XT_AB="name=v0,v1".

    python -m torch.distributed.run --nproc-per-node N tools/weak_probe.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from extrack_b200 import tracking as xt  # noqa: E402
from extrack_b200.simulate import sim_tracks  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
tracks = sim_tracks(int(os.environ.get("N_TRACKS", 1_000_000)), seed=rank, device=f"cuda:{local}", **bench.SIM_KW)
st, _ = xt._sorted_buckets(tracks)
ts = xt.TrackSet(st, rank=0, world_size=1, device=local)
eng = ts.engine
pvar = bench.param_variants(st[0].shape[1])
buf = torch.zeros(1, dtype=torch.float64, device=f"cuda:{local}")
stream = torch.cuda.current_stream().cuda_stream
name, vals = os.environ.get("XT_AB", "verify_fork_k2=0,1").split("=")
for rep in range(2):
    for v in vals.split(","):
        eng.set_option(name, int(v))
        for mode in ("allreduce", "no collective"):
            n = 0
            for i in range(6):
                eng.sum_logp_async(pvar[n % len(pvar)], buf.data_ptr(), stream); n += 1
                if world > 1 and mode == "allreduce":
                    dist.all_reduce(buf)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(60):
                eng.sum_logp_async(pvar[n % len(pvar)], buf.data_ptr(), stream); n += 1
                if world > 1 and mode == "allreduce":
                    dist.all_reduce(buf)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / 60], dtype=torch.float64, device=f"cuda:{local}")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rank == 0:
                print(f"{name}={v} {mode:13s}: {float(t.item()):.4f} ms per step (max over {world} ranks), verified {eng.stats()['plan_verified']}", flush=True)
if world > 1:
    dist.destroy_process_group()
