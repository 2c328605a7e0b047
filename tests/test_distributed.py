"""N > 1 host logic on CPU: chunk sharding + one all-reduce per objective call, world_size 2, gloo.

There is no GPU in the build container, so the CUDA engine is replaced *in this test only* by a
stand-in that evaluates a rank's chunks with the oracle; what is under test is `TrackSet`'s
partition of the reference chunk list over ranks and the reduction of the partial sums."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    from extrack_b200 import _native
    from extrack_b200 import tracking as xt
    from helpers import make_model, random_walk_tracks
    from oracle import extrack_oracle as orc

    model = make_model(frame_len=5, min_len=6)

    class OracleEngine:  # same surface as _native.Engine, oracle inside (test double)
        def __init__(self, device=0):
            self.segs = []

        def upload(self, segments, isBL, chunk_size):
            self.segs = [(np.asarray(s), int(b)) for s, b in zip(segments, isBL)]
            assert all(len(s) <= chunk_size for s, _ in self.segs)

        def sum_logp(self, p):
            return float(sum(orc.chunk_logp(s, model, b).sum() for s, b in self.segs))

        def predict(self, p, nS):
            pm = make_model(frame_len=5, min_len=6, threshold=0.1, max_nb_states=200)
            return [np.concatenate([orc.chunk_recursion(s[i : i + 1], pm, b, 1)[2] for i in range(len(s))]) for s, b in self.segs]

        def set_option(self, name, value):
            assert name == "predict_shared_plans" and value == 0  # nb_max = 1 here: per-track plans

        def close(self):
            pass

    _native.Engine = OracleEngine
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)  # identical data on every rank
    st = [random_walk_tracks(n, L, 2, rng) for L, n in ((6, 700), (9, 1500), (14, 1200))]
    ts = xt.TrackSet(st, chunk=500)
    total = ts.sum_logp(None)
    # state annotation: every rank annotates its slice of each bucket, the dict is reassembled everywhere
    from extrack_b200._lmfit_compat import Parameters

    prm = Parameters()
    for k, v in dict(D0=1e-5, D1=0.25, LocErr=0.02, F0=0.6, F1=0.4, p01=0.1, p10=0.1, pBL=0.05).items():
        prm.add(k, value=v)
    small = {"6": st[0][:7], "9": st[1][:1], "14": st[2][:10]}  # a one-track bucket: rank 0 gets no row of it
    pred = xt.predict_Bs(small, 0.02, prm, cell_dims=[1], nb_states=2, frame_len=5)
    part = xt.predict_Bs(small, 0.02, prm, cell_dims=[1], nb_states=2, frame_len=5, gather=False)
    q.put((rank, total, ts.my_chunks, len(ts.chunks), {k: v.tolist() for k, v in pred.items()}, {k: len(v) for k, v in part.items()}))
    dist.destroy_process_group()


def test_two_rank_objective_equals_single_process():
    import torch.multiprocessing as mp

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import make_model, random_walk_tracks
    from oracle import extrack_oracle as orc

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    st = [random_walk_tracks(n, L, 2, rng) for L, n in ((6, 700), (9, 1500), (14, 1200))]
    want = -orc.neg_log_likelihood(st, make_model(frame_len=5, min_len=6), chunk=500)
    (r0, t0, c0, n0, p0, m0), (r1, t1, c1, n1, p1, m1) = res
    assert t0 == t1  # every rank sees the same all-reduced value
    assert abs(t0 - want) <= 1e-12 * abs(want)
    assert sorted(c0 + c1) == list(range(n0)) and c0 and c1  # disjoint cover of the reference chunk list
    # predict_Bs: identical full dictionaries on both ranks, equal to the single-process oracle result
    pm = make_model(frame_len=5, min_len=6, threshold=0.1, max_nb_states=200)
    small = [st[0][:7], st[1][:1], st[2][:10]]
    want_p = orc.predict_states(small, pm)
    assert p0 == p1
    for a, w in zip(small, want_p):
        np.testing.assert_allclose(np.array(p0[str(a.shape[1])]), w, rtol=0, atol=1e-12)
    assert m0 == {"6": 3, "9": 0, "14": 5} and m1 == {"6": 4, "9": 1, "14": 5}
