"""Reference solution of the 2-D trust-region boundary problem by the quartic Ceres solves (numpy.roots)."""
import numpy as np


def boundary_min_quartic(B, g, r):
    """y = r (2t, 1 - t^2) / (1 + t^2); stationarity of f = 0.5 y^T B y + g^T y on the circle is a quartic in t
    (dogleg_strategy.cc MakePolynomialForBoundaryConstrainedProblem); t = inf is y = (0, -r).  Evaluate f at every
    real root and keep the minimum (FindMinimumOnTrustRegionBoundary)."""
    B = np.asarray(B, float).reshape(2, 2)
    B = 0.5 * (B + B.T)
    a, b, c = float(B[0, 0]), float(B[0, 1]), float(B[1, 1])
    g0, g1, r = float(g[0]), float(g[1]), float(r)
    s_num = np.poly1d([2.0, 0.0])              # 2t
    c_num = np.poly1d([-1.0, 0.0, 1.0])        # 1 - t^2
    den = np.poly1d([1.0, 0.0, 1.0])           # 1 + t^2
    poly = (s_num * c_num * (a - c) + (c_num * c_num - s_num * s_num) * b) * (r * r) + (c_num * g0 - s_num * g1) * den * r
    cands = [np.array([0.0, -r])]
    for t in np.roots(poly.coeffs):
        if abs(t.imag) < 1e-9 * max(1.0, abs(t.real)):
            t = t.real
            cands.append(r * np.array([2 * t, 1 - t * t]) / (1 + t * t))
    f = [0.5 * y @ B @ y + np.dot(np.array([g0, g1]), y) for y in cands]
    k = int(np.argmin(f))
    return cands[k], f[k]
