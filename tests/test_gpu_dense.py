"""Dense-map correspondence path (dense_search.cuh: queries binned by map cell, neighbourhood tiles staged into shared
memory by TMA bulk copies, fine-grid nearest-first search).  It must return the SAME exact radius-truncated kNN as the
lane-pair path -- so poses, factor counts and normal equations are bit-identical whichever path runs -- and stay within
parity of the CPU oracle.  BASELINE config 3 (dense indoor, planar-only) at a reduced size is the parity case."""
import os

import numpy as np
import pytest

from tloam_b200 import synth

pytestmark = pytest.mark.gpu
BIG = 10 ** 9
CAPS = dict(edge_maxnum=BIG, sphere_maxnum=BIG, planar_maxnum=BIG, ground_maxnum=BIG)


def pose_err(A, B):
    d = np.linalg.inv(A) @ B
    return np.linalg.norm(d[:3, 3]), np.arccos(np.clip((np.trace(d[:3, :3]) - 1) / 2, -1, 1))


def quantized(sc):
    """Map coordinates exactly representable as origin + float32 (what the device stores), so that GPU and oracle see the
    same points (origin rule: integer-rounded centre of the bounding box of the first non-empty cloud)."""
    first = next(np.asarray(c).reshape(-1, 3) for c in sc["map"] if len(c))
    origin = np.rint(0.5 * (first.min(0) + first.max(0)))
    out = dict(sc)
    out["map"] = [origin + (c - origin).astype(np.float32).astype(np.float64) for c in sc["map"]]
    return out


def make_reg(dense, **cfg):
    import tloam_b200
    os.environ["TLOAM_B200_DENSE"] = "1" if dense else "0"
    os.environ["TLOAM_B200_FINE"] = "0"              # the reference side of every comparison is the plain lane-pair search
    os.environ["TLOAM_B200_DENSE_CHECK"] = "1"       # every dense query is re-searched by the plain path on the device
    try:
        return tloam_b200.LocalRegistration(**cfg)
    finally:
        for k in ("TLOAM_B200_DENSE", "TLOAM_B200_DENSE_CHECK", "TLOAM_B200_FINE"):
            os.environ.pop(k, None)


def run(sc, dense, **cfg):
    r = make_reg(dense, **cfg)
    r.set_input_target(sc["map"])
    r.set_input_source(sc["scan"])
    T, st = r.scan_matching(sc["predict"], want_stats=True)
    if dense:
        cnt = r.dense_check_counters()
        fn = cfg.get("factor_num", 4)
        nq = sum(len(sc["scan"][c]) for c in ((2, 3) if fn == 2 else (0, 2, 3)))
        assert cnt[1] == 0, f"dense kNN differs from the plain search for {cnt[1]} of {cnt[0]} queries: {cnt}"
        assert cnt[0] == st.n_outer * nq, (cnt, st.n_outer, nq)        # every query searched exactly once per outer iteration
    r.close()
    return T, st


def same_trace(a, b):
    assert a.n_outer == b.n_outer
    for o in range(a.n_outer):
        assert list(a.outer[o].n_factors) == list(b.outer[o].n_factors)
        assert np.array_equal(np.array(a.outer[o].H0), np.array(b.outer[o].H0))
        assert np.array_equal(np.array(a.outer[o].x_end), np.array(b.outer[o].x_end))


def very_dense_scene(seed=5, n_floor=300_000, n_wall=120_000, n_scan=3000):
    """A 4 x 4 m floor patch at ~18700 points / m^2 (4700 per 0.5 m cell, 42k in a 9-cell neighbourhood: several staging
    passes of kDenseCap = 10240) and a wall."""
    rng = np.random.default_rng(seed)
    floor = np.stack([rng.uniform(0, 4, n_floor), rng.uniform(0, 4, n_floor), rng.normal(0, 0.002, n_floor)], 1)
    wall = np.stack([rng.normal(0, 0.002, n_wall), rng.uniform(0, 4, n_wall), rng.uniform(0, 2, n_wall)], 1)
    wall2 = np.stack([rng.uniform(0, 4, n_wall // 2), rng.normal(0, 0.002, n_wall // 2), rng.uniform(0, 2, n_wall // 2)], 1)
    dummy = np.array([[2.0, 2.0, 1.0]]) + rng.normal(0, 0.05, (16, 3))
    T_gt = synth.se3_exp([1.5, 2.2, 0.8, 0.02, -0.01, 0.4])
    Ti = np.linalg.inv(T_gt)

    def to_scan(p, n):
        q = p[rng.choice(len(p), n, replace=False)] + rng.normal(0, 0.004, (n, 3))
        q[rng.random(n) < 0.1] += rng.uniform(-0.3, 0.3, 3)
        return np.ascontiguousarray(q @ Ti[:3, :3].T + Ti[:3, 3])

    return dict(map=[dummy, dummy.copy(), np.concatenate([wall, wall2]), floor],
                scan=[to_scan(dummy, 16), to_scan(dummy, 16), to_scan(np.concatenate([wall, wall2]), n_scan), to_scan(floor, n_scan)],
                predict=T_gt @ synth.se3_exp([0.02, -0.015, 0.01, 0.002, -0.0015, 0.003]), T_gt=T_gt)


def test_dense_path_is_bit_identical_on_config3_and_within_parity_of_the_oracle(oracle):
    f = quantized(synth.config3(20_000, 80_000))
    cfg = dict(factor_num=2, **CAPS)
    Td, sd = run(f, True, **cfg)
    Ts, ss = run(f, False, **cfg)
    assert sd.gpu_launches > ss.gpu_launches          # the dense kernels did run
    assert np.array_equal(Td, Ts)
    same_trace(sd, ss)
    o = oracle.Oracle(threads_mode=1, **cfg)
    o.set_input_target(f["map"])
    o.set_input_source(f["scan"])
    rc, To, so = o.scan_matching(f["predict"])
    dt, dr = pose_err(Td, To)
    assert rc == 0 and dt < 1e-4 and dr < 1e-5, (dt, dr)
    assert [list(sd.outer[i].n_factors) for i in range(sd.n_outer)] == [list(so.outer[i].n_factors) for i in range(so.n_outer)]


def test_dense_path_with_many_staging_passes():
    sc = very_dense_scene()
    cfg = dict(factor_num=2, **CAPS)
    Td, sd = run(sc, True, **cfg)
    Ts, ss = run(sc, False, **cfg)
    assert np.array_equal(Td, Ts)
    same_trace(sd, ss)
    assert pose_err(Td, sc["T_gt"])[0] < 0.05


def test_dense_path_is_picked_automatically_for_a_dense_map():
    """points per occupied brick >= 256 and >= 2048 queries per cloud: the handle waits ONCE for the statistics of its
    first map (later maps are judged by the statistics that come home with every frame's result) and takes the dense path
    under TLOAM_B200_DENSE=auto; the pose equals the lane-pair search's."""
    import tloam_b200
    sc = very_dense_scene()
    cfg = dict(factor_num=2, **CAPS)
    os.environ["TLOAM_B200_DENSE"] = "auto"
    try:
        r = tloam_b200.LocalRegistration(**cfg)
    finally:
        os.environ.pop("TLOAM_B200_DENSE", None)
    r.set_input_target(sc["map"])
    r.set_input_source(sc["scan"])
    T1, s1 = r.scan_matching(sc["predict"], want_stats=True)
    r.set_input_target(sc["map"])
    T2, s2 = r.scan_matching(sc["predict"], want_stats=True)
    r.close()
    for st in (s1, s2):
        assert st.gpu_launches == 1 + 4 * (2 + 4 + 4)      # un-fused + 3 binning kernels + the dense search per outer
    Ts, _ = run(sc, False, **cfg)                          # un-fused lane-pair search: same reduction tree
    assert np.array_equal(T1, Ts) and np.array_equal(T2, Ts)


def test_dense_path_on_a_sparse_outdoor_scene_and_with_binding_caps(oracle):
    """Forced onto a config-2-shaped (sparse) scene with the reference's default caps: same factors, same pose."""
    cfg = synth.scaled(0.05, seed=321)
    T_gt = synth.se3_exp([3.0, -1.0, 0.0, 0.01, 0.0, 0.2])
    sc = dict(map=synth.make_map(cfg, T_gt), scan=synth.make_scan(cfg, T_gt, 3), predict=T_gt @ synth.se3_exp(synth.CONFIG1_PERTURB))
    for caps in ({}, CAPS):
        Td, sd = run(sc, True, **caps)
        Ts, ss = run(sc, False, **caps)
        assert np.array_equal(Td, Ts)
        same_trace(sd, ss)


def test_build_factors_through_the_dense_path_matches_the_oracle(oracle):
    f = quantized(synth.config3(6_000, 60_000))
    r = make_reg(True, factor_num=2, **CAPS)
    r.set_input_target(f["map"])
    r.set_input_source(f["scan"])
    from oracle import pyoracle
    x = pyoracle.se3_log(f["predict"])
    for cloud in (2, 3):
        valid, prim = r.build_factors(cloud, x)
        o = pyoracle.Oracle(factor_num=2, **CAPS)
        o.set_input_target(f["map"])
        o.set_input_source(f["scan"])
        vo, po = o.build_factors(cloud, x)
        assert np.array_equal(valid, vo)
        assert np.allclose(prim, po, rtol=0, atol=1e-9)
    r.close()
