"""Pins the oracle's building blocks against INDEPENDENT implementations available in this image
(scipy expm/logm, numpy eigh, scipy cKDTree, central differences).  The reference itself has no tests or
golden vectors for this path => "parity unpinned"; these cross-checks are what pins it."""
import numpy as np
import pytest
from scipy.linalg import expm, logm
from scipy.spatial import cKDTree


def hat6(a):
    u, w = a[:3], a[3:]
    M = np.zeros((4, 4))
    M[:3, :3] = [[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]
    M[:3, 3] = u
    return M


def vee6(M):
    return np.array([M[0, 3], M[1, 3], M[2, 3], M[2, 1], M[0, 2], M[1, 0]])


@pytest.mark.parametrize("seed", range(8))
def test_se3_exp_log_vs_scipy(oracle, seed):
    rng = np.random.default_rng(seed)
    om = rng.normal(0, 1, 3)
    om *= rng.choice([1e-3, 0.1, 3.0]) / np.linalg.norm(om)      # |omega| < pi so that log(exp(a)) == a
    a = np.concatenate([rng.normal(0, 5, 3), om])
    T = oracle.se3_exp(a)
    assert np.allclose(T, expm(hat6(a)), atol=1e-12)
    assert np.allclose(oracle.se3_log(T), vee6(np.real(logm(T))), atol=1e-10)
    assert np.allclose(oracle.se3_log(oracle.se3_exp(a)), a, atol=1e-12)


def test_se3_small_angle_branches(oracle):
    # Sophus Taylor branch below epsilon = 1e-10 (sophus common.hpp:94, so3.hpp:594-606)
    for th in (0.0, 1e-12, 9e-11, 1.1e-10, 1e-8):
        a = np.array([0.3, -0.2, 0.1, th, 0, 0])
        T = oracle.se3_exp(a)
        assert np.allclose(T, expm(hat6(a)), atol=1e-14)
        assert np.allclose(oracle.se3_log(T), a, atol=1e-14)


def test_se3_log_large_angle_negative_w(oracle):
    # rotation by > pi: quaternion w < 0, Sophus' atan(n/w) returns the equivalent short rotation
    a = np.array([1.0, 2.0, 3.0, 0.0, 0.0, 4.0])
    T = oracle.se3_exp(a)
    b = oracle.se3_log(T)
    assert np.allclose(oracle.se3_exp(b), T, atol=1e-12)
    assert abs(np.linalg.norm(b[3:]) - (2 * np.pi - 4.0)) < 1e-12


def test_se3_plus_is_left_perturbation(oracle):
    rng = np.random.default_rng(3)
    x, d = rng.normal(0, 0.5, 6), rng.normal(0, 0.1, 6)
    lhs = oracle.se3_exp(oracle.se3_plus(x, d))
    assert np.allclose(lhs, expm(hat6(d)) @ expm(hat6(x)), atol=1e-12)  # ref: registration.cpp:167-170


@pytest.mark.parametrize("seed", range(5))
def test_sym_eig3_vs_numpy(oracle, seed):
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(3, 3))
    cov = A @ A.T * rng.choice([1e-4, 1.0, 1e3])
    eig, vec = oracle.sym_eig3(cov)
    w, v = np.linalg.eigh(cov)
    assert np.allclose(eig, w, rtol=1e-12, atol=1e-15 * w.max())
    for k in range(3):
        assert abs(abs(vec[:, k] @ v[:, k]) - 1) < 1e-10


@pytest.mark.parametrize("k,radius", [(1, 0.5), (5, 0.5), (5, 1.0), (3, 0.05)])
def test_knn_vs_ckdtree_and_bruteforce(oracle, k, radius):
    rng = np.random.default_rng(k)
    pts = rng.uniform(-10, 10, size=(20000, 3)) * [1, 1, 0.2]
    q = rng.uniform(-10.5, 10.5, size=(3000, 3)) * [1, 1, 0.2]
    idx, d2, cnt = oracle.knn(pts, q, radius, k)
    idx_b, d2_b, cnt_b = oracle.knn(pts, q, radius, k, brute_force=True)
    assert np.array_equal(idx, idx_b) and np.array_equal(cnt, cnt_b) and np.array_equal(d2, d2_b)
    dd, ii = cKDTree(pts).query(q, k=k, distance_upper_bound=radius)
    dd = dd.reshape(len(q), k)
    ii = ii.reshape(len(q), k)
    assert np.array_equal(cnt, np.isfinite(dd).sum(1))
    m = np.isfinite(dd)
    assert np.array_equal(idx[m], ii[m])
    assert np.allclose(d2[m], dd[m] ** 2, rtol=1e-12)      # squared distances (Open3D convention, Q14)
    if k > 1:
        both = m[:, 1:] & m[:, :-1]
        assert np.all((d2[:, 1:] - d2[:, :-1])[both] >= 0)          # ascending


def test_knn_edge_cases(oracle):
    pts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0.0]])
    idx, d2, cnt = oracle.knn(pts, [[0.1, 0, 0]], 0.5, 5)       # fewer points than k
    assert cnt[0] == 1 and idx[0, 0] == 0 and idx[0, 1] == -1
    idx, d2, cnt = oracle.knn(pts, [[0.5, 0, 0]], 0.5, 5)       # d2 == r^2 exactly is excluded (strict <)
    assert cnt[0] == 0
    idx, d2, cnt = oracle.knn(np.array([[1, 1, 1.0]] * 4), [[1, 1, 1.0]], 0.5, 2)  # duplicates: lowest index first
    assert list(idx[0]) == [0, 1] and cnt[0] == 2


def test_fit_plane_matches_svd_normal(oracle):
    rng = np.random.default_rng(0)
    n = np.array([0.3, -0.5, 0.8])
    n /= np.linalg.norm(n)
    B = np.linalg.svd(n[None])[2][1:]
    pts = (rng.normal(size=(5, 2)) @ B) * 0.3 + 7.0 * n + rng.normal(0, 1e-3, size=(5, 3))
    nd = oracle.fit_plane(pts)
    assert abs(np.linalg.norm(nd[:3]) - 1) < 1e-12
    c = pts.mean(0)
    assert abs(nd[:3] @ c + nd[3]) < 1e-12               # d = -n.c, ref: registration.cpp:366
    _, _, vt = np.linalg.svd(pts - c)
    assert abs(abs(nd[:3] @ vt[2]) - 1) < 1e-4           # close to the least-squares normal
    assert np.allclose(oracle.fit_plane(np.zeros((5, 3))), 0)   # degenerate -> zero vector (:362-364)


def test_fit_line_thresholds(oracle):
    z = np.linspace(0, 1, 5)
    vertical = np.stack([0.01 * np.sin(z * 9), 0.01 * np.cos(z * 7), z], axis=1) + [5, 5, 0]
    ok, a, b, mean, d, eig = oracle.fit_line(vertical, 0.85)
    assert ok == 1 and abs(d[2]) > 0.99 and np.allclose(a - b, 0.2 * d, atol=1e-12)
    assert np.allclose(mean, vertical.mean(0), atol=1e-12) and abs(np.linalg.norm(a - b) - 0.2) < 1e-12
    horizontal = vertical[:, [2, 1, 0]]
    assert oracle.fit_line(horizontal, 0.85)[0] == 0            # |dir.z| <= 0.85 (:481)
    blob = np.random.default_rng(1).normal(size=(5, 3))
    assert oracle.fit_line(blob, 0.0)[0] == 0 or oracle.fit_line(blob, 0.0)[5][2] > 3 * oracle.fit_line(blob, 0.0)[5][1]


def numeric_jac(f, x, h=1e-6):
    """d r / d delta for the LEFT perturbation x+ = log(exp(delta) exp(x)) at delta = 0."""
    cols = []
    for j in range(6):
        d = np.zeros(6)
        d[j] = h
        cols.append((f(+d) - f(-d)) / (2 * h))
    return np.stack(cols, axis=1)


def test_functor_jacobians_vs_central_differences(oracle):
    rng = np.random.default_rng(5)
    x = np.array([0.8, 0.02, 0.01, 0.01, -0.02, 0.15])
    p, q = rng.normal(0, 10, 3), rng.normal(0, 10, 3)
    a = rng.normal(0, 10, 3)
    dvec = rng.normal(size=3)
    dvec /= np.linalg.norm(dvec)
    b = a - 0.2 * dvec
    n = rng.normal(size=3)
    n /= np.linalg.norm(n)
    w = 0.7
    # point-to-point: analytic J equals the true derivative (ref: registration.cpp:36-42)
    r, J, c = oracle.eval_point_to_point(x, p, q, w)
    Jn = numeric_jac(lambda d: oracle.eval_point_to_point(oracle.se3_plus(x, d), p, q, w)[0], x)
    assert np.allclose(J, Jn, atol=1e-6)
    assert abs(c - r.sum() ** 2) < 1e-15                  # Q3: (r0+r1+r2)^2
    # point-to-plane: residual unweighted, Jacobian weighted (Q4) => J = w * true derivative
    r, J, c = oracle.eval_point_to_plane(x, p, n, 0.3, w)
    Jn = numeric_jac(lambda d: oracle.eval_point_to_plane(oracle.se3_plus(x, d), p, n, 0.3, w)[0], x)
    assert np.allclose(J, w * Jn, atol=1e-6) and abs(c - r[0] ** 2) < 1e-15
    # point-to-line: the reference Jacobian is hat(b-a)[I|-hat(c)] w/|a-b| (:75-83) = MINUS the true derivative
    # of r = w (c-a)x(c-b)/|a-b|; the oracle restates the reference, and this test documents the sign.
    r, J, c = oracle.eval_point_to_line(x, p, a, b, w)
    Jn = numeric_jac(lambda d: oracle.eval_point_to_line(oracle.se3_plus(x, d), p, a, b, w)[0], x)
    assert np.allclose(np.abs(J), np.abs(Jn), atol=1e-5)
    sign = np.sign((J * Jn).sum())
    assert np.allclose(J, sign * Jn, atol=1e-5)


def test_update_weight_rule(oracle):
    c2, mu = 1e-4, 0.5
    th1, th2 = (mu + 1) / mu * c2, mu / (mu + 1) * c2
    slots = np.array([0.0, th1 * 2, th2 / 2, (th1 + th2) / 2])
    w = oracle.update_weight(np.full(4, 0.37), slots, c2, th1, th2, mu)
    assert w[0] == 0.37 and w[1] == 0.0 and w[2] == 1.0   # Q13 untouched, truncated, inlier
    assert abs(w[3] - (np.sqrt(c2 * mu * (mu + 1) / slots[3]) - mu)) < 1e-15 and 0 <= w[3] <= 1
