"""Committed golden vectors (tests/golden/tls_small.npz, made by tools/make_golden.py from the CPU oracle --
the reference has none).  CPU: the oracle still reproduces them (regression pin of the
checker).  GPU: the CUDA path reproduces them through the C ABI without running the oracle."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLOUDS = ("edge", "sphere", "planar", "ground")


@pytest.fixture(scope="module")
def gold():
    g = np.load(os.path.join(ROOT, "tests", "golden", "tls_small.npz"))
    mp = [g["origin"] + g[f"map_{n}"].astype(np.float64) for n in CLOUDS]
    scan = [g[f"scan_{n}"] for n in CLOUDS]
    return g, mp, scan


def pose_err(A, B):
    d = np.linalg.inv(A) @ B
    return np.linalg.norm(d[:3, 3]), np.arccos(np.clip((np.trace(d[:3, :3]) - 1) / 2, -1, 1))


def test_oracle_reproduces_golden(oracle, gold):
    g, mp, scan = gold
    o = oracle.Oracle()
    o.set_input_target(mp)
    o.set_input_source(scan)
    rc, T, st = o.scan_matching(g["predict"])
    assert rc == 0 and st.n_outer == int(g["n_outer"])
    assert np.allclose(T, g["pose"], atol=1e-12)
    assert np.array_equal([list(st.outer[i].n_factors) for i in range(st.n_outer)], g["n_factors"])
    assert np.allclose([st.outer[i].initial_cost for i in range(st.n_outer)], g["initial_cost"], rtol=1e-12)
    dt, _ = pose_err(T, g["T_gt"])
    assert dt < 0.05                                     # ~800 noisy features with 10 % outliers
    for i in range(8):
        r, J, c = oracle.eval_point_to_line(g["f_x"], g["f_p"][i], g["f_a"][i], g["f_b"][i], g["f_w"][i])
        assert np.allclose(r, g["pl_r"][i], atol=1e-13) and np.allclose(J, g["pl_J"][i], atol=1e-12)


@pytest.mark.gpu
def test_gpu_reproduces_golden(gold):
    import tloam_b200
    g, mp, scan = gold
    reg = tloam_b200.LocalRegistration()
    reg.set_input_target(mp)
    reg.set_input_source(scan)
    T, st = reg.scan_matching(g["predict"], want_stats=True)
    dt, dr = pose_err(T, g["pose"])
    assert dt < 1e-8 and dr < 1e-9, (dt, dr)              # far inside 1e-4 m / 1e-5 rad
    assert st.n_outer == int(g["n_outer"])
    for i in range(st.n_outer):
        assert list(st.outer[i].n_factors) == list(g["n_factors"][i])
        assert st.outer[i].n_inner == g["n_inner"][i] and st.outer[i].termination == g["termination"][i]
        assert np.isclose(st.outer[i].initial_cost, g["initial_cost"][i], rtol=1e-9)
        assert np.allclose(np.array(st.outer[i].H0).reshape(6, 6), g["H0"][i], rtol=1e-9, atol=1e-9 * np.abs(g["H0"][i]).max())
        assert np.allclose(list(st.outer[i].x_end), g["x_end"][i], atol=1e-8)
        acc = [st.outer[i].inner[k].accepted if k < st.outer[i].n_inner else -9 for k in range(8)]
        assert acc == list(g["accepted"][i])
    x = reg.se3_log(g["predict"])
    for c, name in enumerate(CLOUDS):
        v, p = reg.build_factors(c, x)
        assert np.array_equal(v, g[f"valid_{name}"])
        if c == 0:
            d = np.minimum(np.abs(p - g["prim_edge"]).max(1), np.abs(p - g["prim_edge"][:, [3, 4, 5, 0, 1, 2]]).max(1))
            assert np.all(d < 1e-7)
        else:
            assert np.allclose(p, g[f"prim_{name}"], atol=1e-9)
    r, J, c = reg.eval_point_to_point(g["f_x"], g["f_p"], g["f_q"], g["f_w"])
    assert np.allclose(r, g["pp_r"], atol=1e-12) and np.allclose(J, g["pp_J"], atol=1e-12) and np.allclose(c, g["pp_c"], rtol=1e-12)
    r, J, c = reg.eval_point_to_line(g["f_x"], g["f_p"], g["f_a"], g["f_b"], g["f_w"])
    assert np.allclose(r, g["pl_r"], atol=1e-10) and np.allclose(J, g["pl_J"], atol=1e-10)
    r, J, c = reg.eval_point_to_plane(g["f_x"], g["f_p"], g["f_n"], g["f_d"], g["f_w"])
    assert np.allclose(r, g["pn_r"], atol=1e-12) and np.allclose(J, g["pn_J"], atol=1e-12) and np.allclose(c, g["pn_c"], rtol=1e-12)
    reg.close()


# ---- (f)-2: PCA feature extraction (tests/golden/feature_small.npz, made by tools/make_golden.py) ----
FE_LISTS = ("planar_scan", "planar_submap", "sphere_scan", "sphere_submap", "sphere_candidates")


@pytest.fixture(scope="module")
def fgold():
    return np.load(os.path.join(ROOT, "tests", "golden", "feature_small.npz"))


def check_feature(fgold, info, lists):
    for k in ("cvr", "flatness", "sphericity", "normal", "num_sum"):
        assert np.array_equal(info[k], fgold[k], equal_nan=True), k
    assert np.array_equal(info["neigh"], fgold["neigh"].astype(np.int32))
    for name, lst in zip(FE_LISTS, lists):
        assert np.array_equal(np.asarray(lst, dtype=np.int64), fgold[name].astype(np.int64)), name
    assert len(fgold["planar_submap"]) > 100 and len(fgold["sphere_submap"]) > 3


def test_oracle_reproduces_feature_golden(oracle, fgold):
    check_feature(fgold, oracle.pca_info(fgold["points"]), oracle.extract_planar_sphere(fgold["points"]))


@pytest.mark.gpu
def test_gpu_reproduces_feature_golden(fgold):
    import tloam_b200
    reg = tloam_b200.LocalRegistration()
    check_feature(fgold, reg.pca_info(fgold["points"]), reg.extract_planar_sphere(fgold["points"]))
    reg.close()
