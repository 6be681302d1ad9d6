"""Size-independent properties of the registration path, checked at BASELINE's full size (F = 40k, M = 500k)
where no test wants to wait for many oracle runs: rigid equivariance, insensitivity to the order of the map
points, idempotence of a converged pose, independence from the handle's history."""
import numpy as np
import pytest

from tloam_b200 import synth

pytestmark = pytest.mark.gpu
BIG = 10 ** 9
CAPS = dict(edge_maxnum=BIG, sphere_maxnum=BIG, planar_maxnum=BIG, ground_maxnum=BIG)


@pytest.fixture(scope="module")
def full():
    return synth.config1()


def pose_err(A, B):
    d = np.linalg.inv(A) @ B
    return np.linalg.norm(d[:3, 3]), np.arccos(np.clip((np.trace(d[:3, :3]) - 1) / 2, -1, 1))


def run(mp, scan, predict, **cfg):
    import tloam_b200
    r = tloam_b200.LocalRegistration(**CAPS, **cfg)
    r.set_input_target(mp)
    r.set_input_source(scan)
    T = r.scan_matching(predict)
    r.close()
    return T


def test_rigid_equivariance(full):
    """Moving the map AND the prediction by G moves the result by G (the scan stays in the sensor frame).
    Not bit-exact: the map origin and the FP32 quantisation move with G."""
    G = synth.se3_exp([30.0, -12.0, 0.4, 0.0, 0.0, 0.5])
    T = run(full["map"], full["scan"], full["predict"])
    mapG = [c @ G[:3, :3].T + G[:3, 3] for c in full["map"]]
    TG = run(mapG, full["scan"], G @ full["predict"])
    dt, dr = pose_err(G @ T, TG)
    assert dt < 1e-4 and dr < 1e-5, (dt, dr)


def test_map_point_order_does_not_matter(full):
    rng = np.random.default_rng(0)
    T = run(full["map"], full["scan"], full["predict"])
    shuffled = [c[rng.permutation(len(c))] for c in full["map"]]
    T2 = run(shuffled, full["scan"], full["predict"])
    assert np.array_equal(T, T2)       # kNN order is (d2, index): only exact distance ties could differ


def test_converged_pose_is_a_fixed_point(full):
    T1 = run(full["map"], full["scan"], full["predict"])
    T2 = run(full["map"], full["scan"], T1)
    dt, dr = pose_err(T1, T2)
    assert dt < 2e-3 and dr < 2e-4, (dt, dr)      # re-solving from the solution stays at the solution (noise level)
    gt = pose_err(T2, full["T_gt"])[0]
    assert gt < 5e-3


def test_result_does_not_depend_on_handle_history(full):
    import tloam_b200
    r = tloam_b200.LocalRegistration(**CAPS)
    other = synth.config1(synth.SceneConfig(seed=77, n_map=(30000, 6000, 60000, 50000), n_feat=(3000, 600, 6000, 5000)))
    r.set_input_target(other["map"])
    r.set_input_source(other["scan"])
    r.scan_matching(other["predict"])
    r.set_input_target(full["map"])
    r.set_input_source(full["scan"])
    T = r.scan_matching(full["predict"])
    r.close()
    assert np.array_equal(T, run(full["map"], full["scan"], full["predict"]))


def test_feature_caps_monotone(full):
    """With the reference's default caps (1200/200/2500/2000 of 40k features) the solve still lands near the
    ground truth, and the factor counts respect the caps."""
    import tloam_b200
    r = tloam_b200.LocalRegistration()
    r.set_input_target(full["map"])
    r.set_input_source(full["scan"])
    T, st = r.scan_matching(full["predict"], want_stats=True)
    r.close()
    for i in range(st.n_outer):
        nf = list(st.outer[i].n_factors)
        assert nf[0] <= 1200 and nf[1] <= 200 and nf[2] <= 2500 and nf[3] <= 2000
    assert pose_err(T, full["T_gt"])[0] < 2e-2
