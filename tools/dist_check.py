"""Two-or-more-rank check on real GPUs (NCCL): sharded objective == single-GPU objective, sharded
predict_Bs (all-gather of the slices) == single-GPU predict_Bs.

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from extrack_b200 import tracking as xt  # noqa: E402
from extrack_b200._lmfit_compat import Parameters  # noqa: E402
from helpers import random_walk_tracks  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world)
rng = np.random.default_rng(5)  # identical data on every rank
tracks = {str(L): random_walk_tracks(n, L, 2, rng) for L, n in ((6, 900), (11, 5100), (19, 2300), (30, 4100))}
prm = Parameters()
for k, v in dict(D0=1e-5, D1=0.25, LocErr=0.02, F0=0.6, F1=0.4, p01=0.1, p10=0.12, pBL=0.05).items():
    prm.add(k, value=v)
st, _ = xt._sorted_buckets(tracks)
sharded = xt.TrackSet(st)  # picks rank / world from torch.distributed
v_sh = xt.cum_Proba_Cs(prm, st, 0.02, [1], None, 2, 1, 7, 0, _trackset=sharded)
single = xt.TrackSet(st, rank=0, world_size=1, device=local)
v_1 = xt.cum_Proba_Cs(prm, st, 0.02, [1], None, 2, 1, 7, 0, _trackset=single)
ok = abs(v_sh - v_1) <= 1e-12 * abs(v_1)
print(f"rank {rank}: sharded {v_sh!r} single {v_1!r} chunks {len(sharded.my_chunks)}/{len(sharded.chunks)} ok={ok}", flush=True)
small = {k: v[:257] for k, v in tracks.items()}
p_sh = xt.predict_Bs(small, 0.02, prm, cell_dims=[1], nb_states=2, frame_len=6)
dist.destroy_process_group()
p_1 = xt.predict_Bs(small, 0.02, prm, cell_dims=[1], nb_states=2, frame_len=6)  # world size 1 now
err = max(float(np.max(np.abs(p_sh[k] - p_1[k]))) for k in small)
print(f"rank {rank}: predict_Bs gathered vs single max abs diff {err:.2e}", flush=True)
sys.exit(0 if (ok and err == 0.0) else 1)
