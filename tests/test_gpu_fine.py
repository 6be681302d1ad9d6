"""Two-level grid (fine_search.cuh + k_map_fine): cells with >= 64 points get their points ordered by a 4 x 4 x 4 grid of
fine bins at map-build time, and the correspondence search of dense clouds becomes an expanding-box search over those
bins, one thread per feature.  It must return the SAME exact radius-truncated kNN as the lane-pair search -- so poses,
factor counts and normal equations are bit-identical whichever path runs -- and the re-ordered map must stay valid for
every other consumer (kNN queries, fitness, batch).  BASELINE config 3 at a reduced size is the parity case."""
import os

import numpy as np
import pytest

from tloam_b200 import synth
from test_gpu_dense import CAPS, pose_err, quantized, same_trace, very_dense_scene

pytestmark = pytest.mark.gpu


def make_reg(fine, check=True, **cfg):
    import tloam_b200
    os.environ["TLOAM_B200_FINE"] = fine
    os.environ["TLOAM_B200_DENSE"] = "0"
    if check:
        os.environ["TLOAM_B200_DENSE_CHECK"] = "1"   # every query of the two-level search is re-searched by knn_search on the device
    try:
        return tloam_b200.LocalRegistration(**cfg)
    finally:
        for k in ("TLOAM_B200_FINE", "TLOAM_B200_DENSE", "TLOAM_B200_DENSE_CHECK"):
            os.environ.pop(k, None)


def enabled_clouds(cfg):
    fn = cfg.get("factor_num", 4)
    return (2, 3) if fn == 2 else ((0, 2, 3) if fn == 3 else (0, 1, 2, 3))


def run(sc, fine, **cfg):
    r = make_reg("1" if fine else "0", **cfg)
    r.set_input_target(sc["map"])
    r.set_input_source(sc["scan"])
    T, st = r.scan_matching(sc["predict"], want_stats=True)
    if fine:
        cnt = r.dense_check_counters()
        nq = sum(len(sc["scan"][c]) for c in enabled_clouds(cfg) if len(sc["map"][c]) >= 64)
        assert cnt[1] == 0, f"two-level kNN differs from the plain search for {cnt[1]} of {cnt[0]} queries: {cnt}"
        assert cnt[0] == st.n_outer * nq, (cnt, st.n_outer, nq)
    r.close()
    return T, st


def test_two_level_search_is_bit_identical_on_config3_and_within_parity_of_the_oracle(oracle):
    f = quantized(synth.config3(20_000, 80_000))
    cfg = dict(factor_num=2, **CAPS)
    Tf, sf = run(f, True, **cfg)
    Ts, ss = run(f, False, **cfg)
    assert sf.gpu_launches == ss.gpu_launches + 4      # one k_correspond_fine per outer iteration
    assert np.array_equal(Tf, Ts)
    same_trace(sf, ss)
    o = oracle.Oracle(threads_mode=1, **cfg)
    o.set_input_target(f["map"])
    o.set_input_source(f["scan"])
    rc, To, so = o.scan_matching(f["predict"])
    dt, dr = pose_err(Tf, To)
    assert rc == 0 and dt < 1e-4 and dr < 1e-5, (dt, dr)
    assert [list(sf.outer[i].n_factors) for i in range(sf.n_outer)] == [list(so.outer[i].n_factors) for i in range(so.n_outer)]


def test_two_level_search_with_thousands_of_points_per_cell():
    sc = very_dense_scene()                            # 4700 points per cell; outliers up to 0.3 m off the surfaces
    cfg = dict(factor_num=2, **CAPS)
    Tf, sf = run(sc, True, **cfg)
    Ts, ss = run(sc, False, **cfg)
    assert np.array_equal(Tf, Ts)
    same_trace(sf, ss)
    assert pose_err(Tf, sc["T_gt"])[0] < 0.05


def test_two_level_search_on_a_sparse_outdoor_scene_all_four_clouds_and_binding_caps():
    """Forced onto a config-2-shaped (sparse) scene: almost no cell reaches 64 points, every cell is a single segment,
    K = 1 (sphere) and K = 5 clouds, with the reference's default caps and without."""
    cfg = synth.scaled(0.05, seed=321)
    T_gt = synth.se3_exp([3.0, -1.0, 0.0, 0.01, 0.0, 0.2])
    sc = dict(map=synth.make_map(cfg, T_gt), scan=synth.make_scan(cfg, T_gt, 3), predict=T_gt @ synth.se3_exp(synth.CONFIG1_PERTURB))
    for caps in ({}, CAPS):
        Tf, sf = run(sc, True, **caps)
        Ts, ss = run(sc, False, **caps)
        assert np.array_equal(Tf, Ts)
        same_trace(sf, ss)


def test_mixed_density_map_and_queries_far_from_any_point():
    """One cloud with a dense patch next to a sparse one, queries in free space (fewer than K neighbours inside the radius,
    the box grows to the full radius), on the patch boundary and outside the map."""
    rng = np.random.default_rng(17)
    dense_patch = np.stack([rng.uniform(0, 1.5, 60_000), rng.uniform(0, 1.5, 60_000), rng.normal(0, 0.003, 60_000)], 1)
    sparse_patch = np.stack([rng.uniform(1.5, 12, 3_000), rng.uniform(-5, 5, 3_000), rng.normal(0, 0.01, 3_000)], 1)
    ground = np.concatenate([dense_patch, sparse_patch])
    wall = np.stack([rng.normal(6, 0.003, 30_000), rng.uniform(-3, 3, 30_000), rng.uniform(0, 2.5, 30_000)], 1)
    dummy = np.array([[2.0, 2.0, 1.0]]) + rng.normal(0, 0.05, (16, 3))
    T_gt = synth.se3_exp([0.3, 0.2, 0.0, 0.0, 0.0, 0.1])
    Ti = np.linalg.inv(T_gt)

    def to_scan(p, n, lift):
        q = p[rng.choice(len(p), n, replace=False)] + rng.normal(0, 0.004, (n, 3))
        q[: n // 4] += rng.uniform(-lift, lift, (n // 4, 3))           # a quarter of the queries hang in free space
        return np.ascontiguousarray(q @ Ti[:3, :3].T + Ti[:3, 3])

    sc = dict(map=[dummy, dummy.copy(), wall, ground],
              scan=[to_scan(dummy, 16, 0.0), to_scan(dummy, 16, 0.0), to_scan(wall, 4000, 0.9), to_scan(ground, 6000, 0.9)],
              predict=T_gt @ synth.se3_exp([0.02, -0.015, 0.01, 0.002, -0.0015, 0.003]))
    cfg = dict(factor_num=2, **CAPS)
    Tf, sf = run(sc, True, **cfg)
    Ts, ss = run(sc, False, **cfg)
    assert np.array_equal(Tf, Ts)
    same_trace(sf, ss)


def test_reordered_map_serves_knn_queries_and_fitness_unchanged():
    f = quantized(synth.config3(6_000, 60_000))
    rng = np.random.default_rng(3)
    out = {}
    for fine in ("1", "0"):
        r = make_reg(fine, check=False, factor_num=2, **CAPS)
        r.set_input_target(f["map"])
        r.set_input_source(f["scan"])
        q = f["map"][2][rng.choice(len(f["map"][2]), 500)] + rng.normal(0, 0.05, (500, 3))
        rng = np.random.default_rng(3)                 # same queries for both handles
        idx, d2, cnt = r.knn(2, q, 0.5, 5)
        out[fine] = (idx, d2, cnt, r.get_fitness_score())
        r.close()
    assert np.array_equal(out["1"][0], out["0"][0]) and np.array_equal(out["1"][1], out["0"][1]) and np.array_equal(out["1"][2], out["0"][2])
    assert out["1"][3] == out["0"][3]


def test_second_level_is_built_and_used_automatically_for_a_big_first_map_then_by_density():
    """No environment variable: the first map of a handle gets the second level for clouds >= 65536 points (no statistics
    yet) and the frame takes the two-level search; the statistics that come home with the result keep it on for a dense map."""
    import tloam_b200
    sc = very_dense_scene()
    cfg = dict(factor_num=2, **CAPS)
    r = tloam_b200.LocalRegistration(**cfg)
    r.set_input_target(sc["map"])
    r.set_input_source(sc["scan"])
    T1, s1 = r.scan_matching(sc["predict"], want_stats=True)
    r.set_input_target(sc["map"])
    T2, s2 = r.scan_matching(sc["predict"], want_stats=True)
    r.close()
    for st in (s1, s2):
        assert st.gpu_launches == 1 + 4 * (2 + 1 + 4)      # un-fused + k_correspond_fine per outer
    Ts, _ = run(sc, False, **cfg)
    assert np.array_equal(T1, Ts) and np.array_equal(T2, Ts)


def test_sparse_maps_drop_the_second_level_once_their_statistics_are_known():
    import tloam_b200
    cfgs = synth.scaled(1.0, seed=5)
    T_gt = synth.se3_exp([3.0, -1.0, 0.0, 0.01, 0.0, 0.2])
    mp, scan = synth.make_map(cfgs, T_gt), synth.make_scan(cfgs, T_gt, 3)
    predict = T_gt @ synth.se3_exp(synth.CONFIG1_PERTURB)
    r = tloam_b200.LocalRegistration(**CAPS)
    launches = []
    for k in range(3):
        r.set_input_target(mp)
        r.set_input_source(scan)
        _, st = r.scan_matching(predict, want_stats=True)
        launches.append(st.gpu_launches)
    r.close()
    assert launches[-1] == 1 + 4 * (2 + 4), launches       # plain path from the second or third frame on


def test_build_factors_through_the_two_level_search_matches_the_oracle(oracle):
    f = quantized(synth.config3(6_000, 60_000))
    r = make_reg("1", factor_num=2, **CAPS)
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


def test_batch_of_dense_maps_uses_the_two_level_search_and_matches_single():
    import tloam_b200
    scs = [quantized(synth.config3(8_000 + 500 * s, 70_000, seed=77 + s)) for s in range(2)]
    cfg = dict(factor_num=2, **CAPS)
    singles = [run(sc, False, **cfg)[0] for sc in scs]
    os.environ["TLOAM_B200_FINE"] = "1"
    try:
        b = tloam_b200.BatchRegistration(2, **cfg)
    finally:
        os.environ.pop("TLOAM_B200_FINE", None)
    b.set_input_target(b.pack_host([sc["map"] for sc in scs]))
    b.set_input_source(b.pack_host([sc["scan"] for sc in scs]))
    T, st = b.scan_matching(np.stack([sc["predict"] for sc in scs]))
    b.close()
    for s in range(2):
        assert st[s] == 0 and np.array_equal(T[s], singles[s])
