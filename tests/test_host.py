"""Host-side logic and the C-ABI surface (CPU only; no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

from extrack_b200 import _native
from extrack_b200 import tracking as xt
from extrack_b200._lmfit_compat import Parameters, minimize
from helpers import make_model, random_walk_tracks
from oracle import extrack_oracle as orc
from oracle import ref_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def two_state_params():
    p = Parameters()
    for k, v in dict(D0=1e-5, D1=0.25, LocErr=0.02, F0=0.6, p01=0.1, p10=0.12, pBL=0.05).items():
        p.add(k, value=v, min=0, max=10)
    p.add("F1", expr="1-F0")
    return p


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "xtrack.h")).read()
    declared = set(re.findall(r"\b(xt_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_native.EXPORTS)
    lib = _native.load_library()
    for name in declared:
        assert hasattr(lib, name), name


def test_params_struct_layout_matches_header():
    assert ctypes.sizeof(_native.XtParams) == 8 * 4 + 8 + 8 * 3 + 5 * 8 * _native.XT_MAX_HEADS + 8 * _native.XT_MAX_STATES + 2 * 8
    assert ctypes.sizeof(_native.XtStats) == 4 * 8 + 4 * 4 + 2 * 4 + 2 * 4 + 2 * 4 + 4 + 4 + 2 * 4  # (+ fp32, tail padding, plan_verified, replanned)


def test_engine_fails_loudly_without_gpu(gpu_available):
    if gpu_available:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_native.EngineError, match="no CPU fallback"):
        _native.Engine(0)
    with pytest.raises(_native.EngineError):
        xt.Proba_Cs(np.zeros((2, 4, 2)), np.array([[[0.02]]]), np.array([0.01, 0.1]), np.array([0.5, 0.5]),
                    np.array([[0.9, 0.1], [0.1, 0.9]]), 0.1, 1, [1], 1, 4, 3, 0.2, 120)


def test_extract_params_conventions():
    LocErr, ds, Fs, TrMat, pBL = xt.extract_params(two_state_params(), 0.02, 2, 1)
    assert LocErr[0].shape == (1, 1, 1) and LocErr[0][0, 0, 0] == 0.02
    np.testing.assert_allclose(ds, np.sqrt(2 * np.array([1e-5, 0.25]) * 0.02))
    np.testing.assert_allclose(Fs, [0.6, 0.4])
    assert pBL == 0.05
    np.testing.assert_allclose(TrMat, [[np.exp(-0.1), 1 - np.exp(-0.1)], [1 - np.exp(-0.12), np.exp(-0.12)]])
    _, _, _, T2, _ = xt.extract_params(two_state_params(), 0.02, 2, 2)
    np.testing.assert_allclose(T2[0, 1], 1 - np.exp(-0.05))
    _, _, _, T0, _ = xt.extract_params(two_state_params(), 0.02, 2, 1, Matrix_type=0)
    np.testing.assert_allclose(T0, [[0.9, 0.1], [0.12, 0.88]])
    for mt in (2, 3, 4):
        _, _, _, Tm, _ = xt.extract_params(two_state_params(), 0.02, 2, 1, Matrix_type=mt)
        np.testing.assert_allclose(Tm.sum(1), 1.0, atol=1e-3)


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not present")
@pytest.mark.parametrize("nsub,mt", [(1, 1), (2, 1), (1, 0), (1, 2), (1, 3), (1, 4)])
def test_extract_params_equals_reference(nsub, mt):
    trk = ref_loader.load_tracking()
    from lmfit import Parameters as RP

    rp = RP()
    for k, v in dict(D0=1e-5, D1=0.04, D2=0.3, LocErr=0.02, F0=0.3, F1=0.3, p01=0.1, p02=0.03, p10=0.12, p12=0.05, p20=0.02,
                     p21=0.07, pBL=0.05).items():
        rp.add(k, value=v)
    rp.add("F2", expr="1-F0-F1")
    a = trk.extract_params(rp, 0.02, 3, nsub, None, mt)
    b = xt.extract_params(rp, 0.02, 3, nsub, None, mt)
    for x, y in zip(a, b):
        np.testing.assert_array_equal(np.asarray(x[0] if isinstance(x, list) else x), np.asarray(y[0] if isinstance(y, list) else y))


def test_per_dim_locerr_param_names():
    p = xt.generate_params(nb_states=2, LocErr_type=2, nb_dims=3)
    LocErr, *_ = xt.extract_params(p, 0.02, 2, 1)
    assert LocErr[0].shape == (1, 1, 3)
    p3 = xt.generate_params(nb_states=2, LocErr_type=3)
    assert p3["LocErr1"].value == p3["LocErr0"].value


def test_extract_params_peakwise_locerr_and_dt_lists():
    """input_LocErr / dt lists (tracking.py:926-932,:979-982): pass-through, slope/offset clip, ds per track."""
    rng = np.random.default_rng(0)
    il = [0.02 + 0.01 * rng.random((4, 5, 1)), 0.02 + 0.01 * rng.random((3, 7, 1))]
    dts = [0.02 * (1 + rng.random((4, 5))), 0.02 * (1 + rng.random((3, 7)))]
    LocErr, ds, Fs, TrMat, pBL = xt.extract_params(two_state_params(), dts, 2, 1, input_LocErr=il)
    assert LocErr is il
    assert [a.shape for a in ds] == [(4, 5, 2), (3, 7, 2)]
    np.testing.assert_array_equal(ds[1], np.sqrt(2 * np.array([1e-5, 0.25])[None, None] * dts[1][:, :, None]))
    p4 = xt.generate_params(nb_states=2, LocErr_type=4, slope_offsets_estimates=[1.5, -0.04])
    LocErr, *_ = xt.extract_params(p4, 0.02, 2, 1, input_LocErr=il)
    np.testing.assert_array_equal(LocErr[0], np.clip(il[0] * 1.5 - 0.04, 0.000001, np.inf))
    assert LocErr[0].min() == 0.000001  # some peaks are clipped


def test_stay_tables_match_oracle_per_chunk_tables():
    """Per-chunk field-of-view tables from the middle order statistics of dt[:, 0] == the oracle's
    HeadTables built from np.median(ds3[:, 0], axis=0) (tracking.py:501-506), odd and even counts."""
    rng = np.random.default_rng(1)
    Ds = np.array([1e-5, 0.04, 0.25])
    m = make_model(nS=3, nsub=2)
    rows, want_stay, want_leave = [], [], []
    for n in (1, 2, 7, 10):
        t = 0.02 * (1 + rng.random((n, 6)))
        ds3 = np.sqrt(2 * Ds[None, None] * t[:, :, None])
        tb = orc.HeadTables(m, np.median(ds3[:, 0], axis=0))
        rows.append(xt._mid2(t[:, 0]))
        want_stay.append(tb.Lp_stay)
        want_leave.append(tb.L_leave)
    Lp, Ll = xt.stay_tables(Ds, np.array(rows), m.TrMat, m.pBL, m.cell_dims, 2)
    np.testing.assert_array_equal(Lp, np.array(want_stay))
    np.testing.assert_array_equal(Ll, np.array(want_leave))


@pytest.mark.parametrize("kw", [dict(nS=2, nsub=1), dict(nS=3, nsub=2), dict(nS=3, nsub=1, loc_err=(0.02, 0.02, 0.03)), dict(nS=4, nsub=1)])
def test_build_tables_match_oracle_tables(kw):
    m = make_model(**kw)
    d = len(m.loc_err) if len(m.loc_err) > 1 else 2
    p = xt.build_tables(m.loc_err, m.ds, m.Fs, m.TrMat, m.pBL, m.cell_dims, m.nb_substeps, m.frame_len, m.min_len, m.threshold,
                        m.max_nb_states, d)
    tb = orc.HeadTables(m)
    nH = tb.K * m.nS
    for name in ("dd", "LT", "LF", "L_leave"):
        np.testing.assert_array_equal(np.array(getattr(p, name)[:nH]), getattr(tb, name))
    np.testing.assert_array_equal(np.array(p.Lp_stay[: tb.K]), tb.Lp_stay)
    assert (p.nS, p.nsub, p.d, p.n_loc) == (m.nS, m.nb_substeps, d, len(m.loc_err))
    assert p.flags & _native.XT_FLAG_INT8_WRAP


def test_build_tables_rejects_bad_locerr_shape():
    m = make_model()
    with pytest.raises(ValueError, match="Localization error"):
        xt.build_tables(np.array([0.02, 0.03]), m.ds, m.Fs, m.TrMat, m.pBL, m.cell_dims, 1, 6, 3, 0.2, 120, 3)


def test_chunk_table_matches_reference_order():
    rng = np.random.default_rng(0)
    st = [np.zeros((n, L, 2)) for L, n in ((5, 10), (7, 4100), (9, 2000))]
    ch = xt.chunk_table(st, 2000)
    assert ch == orc.make_chunks(st, 2000)
    assert ch[0] == (2, 0, 2000, 0) and ch[-1] == (0, 0, 10, 1)
    assert [c[3] for c in ch] == [0, 1, 1, 1, 1]


def test_shard_chunks_is_balanced_partition():
    st = [np.zeros((n, L, 2)) for L, n in ((10, 6000), (20, 9000), (30, 15000))]
    ch = xt.chunk_table(st, 2000)
    for ws in (1, 2, 3, 8):
        own = xt.shard_chunks(ch, st, ws)
        flat = sorted(i for o in own for i in o)
        assert flat == list(range(len(ch)))
        cost = [sum((ch[i][2] - ch[i][1]) * (st[ch[i][0]].shape[1] - 1) for i in o) for o in own]
        assert max(cost) - min(cost) <= 2000 * 29


def test_sorted_buckets_numeric_order_and_empty_dropped():
    d = {"10": np.zeros((1, 10, 2)), "9": np.zeros((2, 9, 2)), "100": np.zeros((3, 100, 2)), "11": np.zeros((0, 11, 2))}
    st, keys = xt._sorted_buckets(d)
    assert [a.shape[1] for a in st] == [9, 10, 100]
    assert keys == ["9", "10", "11", "100"]


def test_generate_params_names_and_exprs():
    p = xt.generate_params(nb_states=3, LocErr_type=1)
    names = list(p.keys())
    assert names[:3] == ["D0", "D1", "D2"] and "LocErr" in names and names[-1] == "pBL"
    assert {"p01", "p02", "p10", "p12", "p20", "p21"} <= set(names)
    assert abs(p["F2"].value - (1 - p["F0"].value - p["F1"].value)) < 1e-15
    p["F0"].value = 0.5
    assert abs(p["F2"].value - (0.5 - p["F1"].value)) < 1e-15


def test_get_params_generic_branch():
    p = xt.get_params()
    assert list(p.keys()) == ["LocErr", "D0", "D1_minus_D0", "D1", "F0", "F1", "p01", "p10", "pBL"]
    assert abs(p["D1"].value - 0.05) < 1e-11 and abs(p["F1"].value - 0.55) < 1e-15  # D0 is clipped to its lower bound


def test_lmfit_standin_bounds_and_bfgs():
    p = Parameters()
    p.add("a", value=0.5, min=0, max=2)
    p.add("b", value=3.0, min=1)
    p.add("c", expr="a + b")

    def f(pp):
        return (pp["a"].value - 1.25) ** 2 + (pp["b"].value - 2.0) ** 2 + 0 * pp["c"].value

    res = minimize(f, p, method="bfgs", nan_policy="propagate")
    assert abs(res.params["a"].value - 1.25) < 1e-5 and abs(res.params["b"].value - 2.0) < 1e-5
    assert abs(res.params["c"].value - 3.25) < 1e-4
    assert res.residual.shape == (1,) and res.residual[0] < 1e-9
    res2 = minimize(f, p, method="powell")
    assert abs(res2.params["a"].value - 1.25) < 1e-4


def test_simulator_statistics():
    from extrack_b200.simulate import sim_FOV, sim_tracks

    tr, st = sim_FOV(nb_tracks=3000, max_track_len=20, min_track_len=5, LocErr=0.02, Ds=[0, 0.25], nb_dims=2,
                     initial_fractions=[0.6, 0.4], TrMat=[[0.9, 0.1], [0.1, 0.9]], dt=0.02, pBL=0.05, cell_dims=[1, None, None],
                     seed=3, return_states=True)
    assert all(v.shape[1] == int(k) and v.shape[2] == 2 for k, v in tr.items())
    assert set(tr) <= {str(i) for i in range(5, 21)} and len(tr["20"]) > len(tr["19"])
    d = np.concatenate([(v[:, 1:] - v[:, :-1]).reshape(-1, 2) for v in tr.values()])
    s = np.concatenate([v[:, :-1].reshape(-1) for v in st.values()])
    assert 8e-4 < np.mean(d[s == 0] ** 2) < 1.6e-3  # immobile at frame start: 2 sigma^2 (+ switches inside the frame)
    assert 0.006 < np.mean(d[s == 1] ** 2) < 0.012  # 2 D dt + 2 sigma^2 ~ 0.0108 (minus state switching)
    same = sim_FOV(nb_tracks=200, max_track_len=12, min_track_len=5, seed=5)[0]
    again = sim_FOV(nb_tracks=200, max_track_len=12, min_track_len=5, seed=5)[0]
    assert all(np.array_equal(same[k], again[k]) for k in same)
    many = sim_tracks(5000, block=3000, seed=1, max_track_len=12, min_track_len=5)
    assert sum(len(v) for v in many.values()) == 5000


# ---- table reader ----
def _synthetic_table(tmp_path, rng, n_tracks=300):
    import pandas as pd

    rows = []
    for tid in rng.permutation(n_tracks):
        L = int(rng.integers(1, 30))
        pos = np.cumsum(rng.normal(size=(L, 2)) * 0.05, 0)
        if tid % 17 == 0 and L > 3:
            pos[3:] = pos[2]          # no displacement: removed by remove_no_disp
        if tid % 23 == 0 and L > 4:
            pos[4:] += 5.0            # one long step: removed by dist_th
        f0 = int(rng.integers(0, 120))
        fr = np.arange(f0, f0 + L)
        perm = rng.permutation(L)     # rows of a track are not in frame order in the file
        for k in perm:
            rows.append((pos[k, 0], pos[k, 1], fr[k], int(tid), float(rng.random())))
    df = pd.DataFrame(rows, columns=["X", "Y", "frame", "track_ID", "QUALITY"])
    path = str(tmp_path / "table.csv")
    df.to_csv(path, index=False)
    return path


def test_read_table_bucketing_rules(tmp_path):
    """Own expectation (loop restating readers.py:173-203) on a synthetic table: exact lengths,
    truncation of longer tracks, the in-between rule, the three filters, frame sorting."""
    import contextlib
    import io

    import pandas as pd

    from extrack_b200 import readers

    rng = np.random.default_rng(0)
    path = _synthetic_table(tmp_path, rng)
    lengths = np.array([4, 5, 6, 9, 12])
    with contextlib.redirect_stdout(io.StringIO()) as out:
        tracks, frames, opt = readers.read_table(path, lengths=lengths, dist_th=1.0, frames_boundaries=[5, 100], fmt="csv",
                                                 colnames=["X", "Y", "frame", "track_ID"], opt_colnames=["QUALITY"])
    df = pd.read_csv(path)
    want = {int(l): [] for l in lengths}
    for tid, g in df.groupby("track_ID"):
        g = g.sort_values("frame")
        m = g[["X", "Y", "frame"]].values.astype(float)
        d2 = (m[1:, :2] - m[:-1, :2]) ** 2
        if len(m) > 1 and np.mean(d2 == 0) > 0.05:
            continue
        if not (5 <= m[0, 2] <= 100) or np.any(np.sum(d2, 1) ** 0.5 > 1.0):
            continue
        n = len(m)
        if n in lengths:
            want[n].append(m[:, :2])
        elif n > 12:
            want[12].append(m[:12, :2])
        elif 4 < n < 12:
            l = int(lengths[np.argmin(np.floor(n / lengths)) - 1])
            want[l].append(m[:l, :2])
    assert list(tracks) == [str(l) for l in lengths if want[int(l)]]
    assert [int(x) for x in out.getvalue().split()] == [int(k) for k in tracks]
    for k, arr in tracks.items():
        np.testing.assert_array_equal(arr, np.array(want[int(k)]))
        assert frames[k].shape == arr.shape[:2] and np.all(np.diff(frames[k], axis=1) == 1)
    assert opt["QUALITY"]["12"].shape[1] == 12


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not present on this box")
def test_read_table_matches_reference_reader(tmp_path):
    import contextlib
    import io

    from extrack_b200 import readers

    rd = ref_loader.load_readers()
    cases = [(os.path.join(ref_loader.REFERENCE_ROOT, "Tutorials", "tracks.csv"),
              dict(lengths=np.arange(5, 50), dist_th=0.3, frames_boundaries=[0, 10000], fmt="csv",
                   colnames=["X", "Y", "frame", "track_ID"], opt_colnames=[], remove_no_disp=True)),
             (_synthetic_table(tmp_path, np.random.default_rng(1)),
              dict(lengths=np.array([4, 5, 6, 9, 12]), dist_th=1.0, frames_boundaries=[5, 100], fmt="csv",
                   colnames=["X", "Y", "frame", "track_ID"], opt_colnames=["QUALITY"], remove_no_disp=True)),
             (_synthetic_table(tmp_path, np.random.default_rng(2)),
              dict(lengths=np.arange(3, 20), dist_th=np.inf, frames_boundaries=[-np.inf, np.inf], fmt=",",
                   colnames=["X", "Y", "frame", "track_ID"], opt_colnames=[], remove_no_disp=False))]
    for path, kw in cases:
        with contextlib.redirect_stdout(io.StringIO()) as o1:
            t1, f1, m1 = rd.read_table(path, **{k: (list(v) if isinstance(v, list) else v) for k, v in kw.items()})
        with contextlib.redirect_stdout(io.StringIO()) as o2:
            t2, f2, m2 = readers.read_table(path, **kw)
        assert list(t1) == list(t2) and o1.getvalue() == o2.getvalue()
        for k in t1:
            np.testing.assert_array_equal(t1[k], t2[k])
            np.testing.assert_array_equal(f1[k], f2[k])
    assert sum(len(v) for v in t2.values()) > 0


# ---- TrackMate xml reader (readers.py:5-98) and per-peak localisation errors of the generator ----
def _synthetic_trackmate_xml(tmp_path, rng, n_tracks=120, name="tracks.xml"):
    """A TrackMate "export tracks" file: <Tracks frameInterval=..><particle nSpots=..><detection t= x= y= z=/>..."""
    lines = ['<?xml version="1.0" encoding="UTF-8"?>',
             '<Tracks nTracks="%d" spaceUnits="um" frameInterval="20.0" timeUnits="ms" from="TrackMate v4.0.1">' % n_tracks]
    spec = []
    for tid in range(n_tracks):
        L = int(rng.integers(2, 30))
        pos = np.cumsum(rng.normal(size=(L, 2)) * 0.05, 0) + 10.0
        if tid % 11 == 0 and L > 3:
            pos[3, 0] = pos[2, 0]     # one zero x-displacement: removed by remove_no_disp
        if tid % 13 == 0 and L > 4:
            pos[4:] += 5.0            # one long step: removed by dist_th
        f0 = int(rng.integers(0, 60))
        lines.append('  <particle nSpots="%d">' % L)
        for k in range(L):
            lines.append('    <detection t="%d" x="%r" y="%r" z="0.0" />' % (f0 + k, float(pos[k, 0]), float(pos[k, 1])))
        lines.append("  </particle>")
        spec.append((pos, f0))
    lines.append("</Tracks>")
    path = str(tmp_path / name)
    with open(path, "w") as fh:
        fh.write("\n".join(lines))
    return path, spec


def test_read_trackmate_xml_rules(tmp_path):
    """Own expectation (loop restating readers.py:51-79) on a synthetic file."""
    from extrack_b200 import readers

    path, spec = _synthetic_trackmate_xml(tmp_path, np.random.default_rng(3))
    lengths = np.array([4, 5, 6, 9, 12])
    tracks, frames, opt = readers.read_trackmate_xml(path, lengths=lengths, dist_th=1.0, frames_boundaries=[5, 50])
    want, want_f = {}, {}
    for pos, f0 in spec:
        st = pos[1:] - pos[:-1]
        if np.min(st[:, 0] ** 2) * np.min(st[:, 1] ** 2) == 0 or not (5 <= f0 <= 50) or not np.all(np.sum(st**2, 1) ** 0.5 < 1.0):
            continue
        L = len(pos)
        key = L if L in lengths else (12 if L > 12 else None)
        if key is None:
            continue
        want.setdefault(key, []).append(pos[:key])
        want_f.setdefault(key, []).append(np.arange(f0, f0 + key, dtype=float))
    assert sorted(int(k) for k in tracks) == sorted(want) and len(want) >= 3
    for k, arr in tracks.items():
        np.testing.assert_array_equal(arr, np.array(want[int(k)]))
        np.testing.assert_array_equal(frames[k], np.array(want_f[int(k)]))
        assert opt["t"][k].dtype.kind == "i" and np.array_equal(opt["t"][k], frames[k].astype(int))
        np.testing.assert_array_equal(opt["x"][k], arr[:, :, 0])
    # a particle with a single detection is the reference's error path (:80-81)
    bad = str(tmp_path / "bad.xml")
    with open(bad, "w") as fh:
        fh.write('<Tracks nTracks="1" frameInterval="20.0"><particle nSpots="1"><detection t="1" x="0.5" y="0.5" z="0"/></particle></Tracks>')
    with pytest.raises(ValueError, match="problem with data"):
        readers.read_trackmate_xml(bad)


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not present on this box")
def test_read_trackmate_xml_matches_reference_reader(tmp_path):
    """Against the unmodified reference reader (its xmltodict import served by the ElementTree stand-in of
    oracle.ref_loader) on the tutorial file and a synthetic one: same buckets, same arrays, same dtypes."""
    from extrack_b200 import readers

    rd = ref_loader.load_readers()
    syn, _ = _synthetic_trackmate_xml(tmp_path, np.random.default_rng(4))
    tut = os.path.join(ref_loader.REFERENCE_ROOT, "Tutorials", "example_tracks.xml")
    cases = [(tut, dict()),
             (tut, dict(lengths=np.arange(3, 12), dist_th=0.3, frames_boundaries=[0, 40], remove_no_disp=False,
                        opt_metrics_names=["t", "x", "z"], opt_metrics_types=[int, "float64", "float64"])),
             ([tut, syn], dict(lengths=np.array([4, 7, 10]), dist_th=np.inf, opt_metrics_names=[], opt_metrics_types=None)),
             (syn, dict(lengths=np.arange(2, 15), dist_th=1.0, frames_boundaries=[5, 50]))]
    n = 0
    for path, kw in cases:
        t1, f1, m1 = rd.read_trackmate_xml(path, **kw)
        t2, f2, m2 = readers.read_trackmate_xml(path, **kw)
        assert list(t1) == list(t2) and list(m1) == list(m2)
        for k in t1:
            np.testing.assert_array_equal(t1[k], t2[k])
            np.testing.assert_array_equal(f1[k], f2[k])
            assert t1[k].dtype == t2[k].dtype and f1[k].dtype == f2[k].dtype
            for m in m1:
                np.testing.assert_array_equal(m1[m][k], m2[m][k])
                assert m1[m][k].dtype == m2[m][k].dtype
            n += len(t1[k])
    assert n > 100


def test_sim_FOV_per_peak_localisation_errors():
    """LocErr_std != 0 (simulate_tracks.py:207-209): sigma = LocErr * chi2(k) / k with k = 2 / LocErr_std^2, i.e. mean
    LocErr and relative spread LocErr_std; the noise added to a localisation has that localisation's sigma."""
    from extrack_b200.simulate import sim_FOV

    kw = dict(nb_tracks=4000, max_track_len=20, min_track_len=5, LocErr=0.03, Ds=[0.0, 0.0], nb_dims=2, pBL=0.02, seed=11)
    tr, _, sg = sim_FOV(LocErr_std=0.25, **kw)
    assert list(tr) == list(sg) and all(tr[k].shape == sg[k].shape for k in tr)
    s = np.concatenate([v.reshape(-1) for v in sg.values()])
    assert abs(s.mean() / 0.03 - 1) < 0.01 and abs(s.std() / s.mean() / 0.25 - 1) < 0.03 and s.min() > 0
    # immobile particles (D = 0): a track's scatter around its mean is the localisation noise itself
    z = np.concatenate([((tr[k] - tr[k].mean(1, keepdims=True)) / sg[k]).reshape(-1) for k in tr if int(k) >= 10])
    assert 0.85 < z.std() < 1.05
    tr0 = sim_FOV(**kw)[0]  # LocErr_std = 0 keeps the two-value return and the same selection of runs
    assert {k: v.shape for k, v in tr0.items()} == {k: v.shape for k, v in tr.items()}


def test_p_stay_columns_reproduce_the_array_formula():
    """tracking._p_stay memoises one column per distinct diffusion length (sequential sum over the 1000 grid points) - it
    must have the bits of the reference's array formula (tracking.py:515-523: np.mean over axis 0 of a [1000, 1, K] array)."""
    from scipy.special import ndtr

    from extrack_b200 import tracking as xt

    rng = np.random.default_rng(5)
    for it in range(120):
        nS = int(rng.choice([2, 3, 4]))
        nsub = int(rng.choice([1, 2])) if nS < 4 else 1
        cells = [float(c) for c in rng.choice([0.5, 1.0, 3.0, 10.0], size=int(rng.integers(1, 3)), replace=False)]
        ds = np.sort(rng.random(nS) * 0.2 * 10.0 ** float(rng.integers(-2, 1)) + 1e-4)
        K = nS**nsub
        tup = np.arange(K)[:, None] // nS ** np.arange(nsub)[None, :] % nS
        sub_ds = np.mean(ds[tup][None] ** 2, axis=2) ** 0.5
        want = np.ones(sub_ds.shape[-1])
        for cell_len in cells:
            xs = np.linspace(0 + cell_len / 2000, cell_len - cell_len / 2000, 1000)
            cur = np.mean(ndtr((cell_len - xs[:, None, None]) / (sub_ds + 1e-200)) - ndtr(-xs[:, None, None] / (sub_ds + 1e-200)), 0)
            want = want * cur
        got = xt._p_stay(ds, nS, nsub, cells)
        np.testing.assert_array_equal(got, want[0])
        ds2 = ds.copy()
        ds2[int(rng.integers(0, nS))] *= 1 + 1.5e-8   # a finite-difference step: the other columns come from the memo
        ds2 = np.sort(ds2)
        sub2 = np.mean(ds2[tup][None] ** 2, axis=2) ** 0.5
        want2 = np.ones(K)
        for cell_len in cells:
            xs = np.linspace(0 + cell_len / 2000, cell_len - cell_len / 2000, 1000)
            want2 = want2 * np.mean(ndtr((cell_len - xs[:, None, None]) / (sub2 + 1e-200)) - ndtr(-xs[:, None, None] / (sub2 + 1e-200)), 0)[0]
        np.testing.assert_array_equal(xt._p_stay(ds2, nS, nsub, cells), want2)
