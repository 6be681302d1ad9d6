"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT -- see oracle/tloam_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may import this module.
PARITY UNPINNED: the reference ships no tests or golden vectors for this path and cannot be built here.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libtloam_oracle.so")

MAX_OUTER = 16
MAX_INNER = 8
CLOUDS = ("edge", "sphere", "planar", "ground")


class Config(C.Structure):
    _fields_ = [
        ("k_corr", C.c_int), ("factor_num", C.c_int),
        ("edge_dist_thres", C.c_double), ("sphere_dist_thres", C.c_double),
        ("planar_dist_thres", C.c_double), ("ground_dist_thres", C.c_double),
        ("edge_dir_thres", C.c_double),
        ("edge_maxnum", C.c_int), ("sphere_maxnum", C.c_int), ("planar_maxnum", C.c_int), ("ground_maxnum", C.c_int),
        ("max_iterations", C.c_int),
        ("cost_threshold", C.c_double), ("gnc_factor", C.c_double), ("noise_bound", C.c_double),
        ("fitness_thres", C.c_double),
        ("ceres_max_num_iterations", C.c_int),
        ("reinit_dir", C.c_double * 3),
        ("initial_trust_region_radius", C.c_double),
        ("threads_mode", C.c_int), ("num_threads", C.c_int),
    ]


class FeatureConfig(C.Structure):
    """oracle_feature_config (ref: config/mapping/feature.yaml)."""
    _fields_ = [("radius", C.c_double), ("K", C.c_int), ("min_neigh", C.c_int), ("planar_num", C.c_int),
                ("sphere_num", C.c_int), ("cvr_scan", C.c_double), ("cvr_submap", C.c_double),
                ("planar_scan_thres", C.c_double), ("planar_submap_thres", C.c_double),
                ("planar_vertic_thres", C.c_double)]


class SubmapConfig(C.Structure):
    _fields_ = [("ground_down_sample", C.c_double), ("ground_down_sample_submap", C.c_double),
                ("edge_down_sample_submap", C.c_double), ("planar_frame_size", C.c_int), ("sphere_frame_size", C.c_int),
                ("edge_crop_box_length", C.c_double), ("ground_crop_box_length", C.c_double)]


class DcvcConfig(C.Structure):
    """oracle_dcvc_config (ref: config/mapping/segmentation.yaml DCVC + velodyne ranges)."""
    _fields_ = [("start_r", C.c_double), ("delta_r", C.c_double), ("delta_p", C.c_double), ("delta_a", C.c_double),
                ("min_seg", C.c_int), ("sensor_min_range", C.c_double), ("sensor_max_range", C.c_double),
                ("min_pitch_init", C.c_double), ("max_pitch_init", C.c_double), ("min_polar_init", C.c_double),
                ("max_polar_init", C.c_double)]


class GroundConfig(C.Structure):
    """oracle_ground_config (ref: config/mapping/segmentation.yaml)."""
    _fields_ = [("sensor_model", C.c_int), ("sensor_height", C.c_double), ("vertical_res", C.c_double), ("init_angle", C.c_double),
                ("sensor_min_range", C.c_double), ("sensor_max_range", C.c_double), ("quadrant", C.c_int), ("num_sec", C.c_int),
                ("plane_dis", C.c_double), ("max_iter", C.c_int), ("ground_seed_num", C.c_int)]


class InnerTrace(C.Structure):
    _fields_ = [
        ("x_candidate", C.c_double * 6), ("candidate_cost", C.c_double), ("model_cost_change", C.c_double),
        ("relative_decrease", C.c_double), ("step_norm_scaled", C.c_double), ("radius", C.c_double),
        ("accepted", C.c_int), ("used_gauss_newton", C.c_int),
    ]


class OuterTrace(C.Structure):
    _fields_ = [
        ("x_start", C.c_double * 6), ("x_end", C.c_double * 6),
        ("initial_cost", C.c_double), ("final_cost", C.c_double),
        ("H0", C.c_double * 36), ("g0", C.c_double * 6),
        ("mu", C.c_double), ("th1", C.c_double), ("th2", C.c_double),
        ("slot_sum", C.c_double * 4), ("n_factors", C.c_int * 4),
        ("n_inner", C.c_int), ("termination", C.c_int),
        ("inner", InnerTrace * MAX_INNER),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("n_outer", C.c_int), ("converged_early", C.c_int),
        ("x_init", C.c_double * 6), ("x_final", C.c_double * 6),
        ("t_kdtree", C.c_double), ("t_factors", C.c_double), ("t_solve", C.c_double),
        ("t_weights", C.c_double), ("t_total", C.c_double),
        ("outer", OuterTrace * MAX_OUTER),
    ]


def build(force=False):
    """Compile oracle/_build/libtloam_oracle.so with the system g++ (the image's $CXX has no OpenMP)."""
    src = [os.path.join(_HERE, f) for f in ("tloam_oracle.cpp", "tloam_oracle.h", "feature_oracle.cpp", "segmentation_oracle.cpp", "Makefile")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp = C.POINTER(C.c_double)
        L.oracle_default_config.argtypes = [C.POINTER(Config)]
        L.oracle_create.argtypes = [C.POINTER(Config)]
        L.oracle_create.restype = C.c_void_p
        L.oracle_destroy.argtypes = [C.c_void_p]
        for f in (L.oracle_set_source, L.oracle_set_target):
            f.argtypes = [C.c_void_p, C.POINTER(dp), C.POINTER(C.c_size_t)]
            f.restype = C.c_int
        L.oracle_scan_match.argtypes = [C.c_void_p, dp, dp, C.POINTER(Stats)]
        L.oracle_scan_match.restype = C.c_int
        L.oracle_fitness.argtypes = [C.c_void_p, dp, dp]
        L.oracle_get_pose_increment.argtypes = [C.c_void_p, dp]
        L.oracle_get_weights.argtypes = [C.c_void_p, C.c_int, dp, C.c_size_t]
        L.oracle_get_weights.restype = C.c_int
        L.oracle_se3_exp.argtypes = [dp, dp]
        L.oracle_se3_log.argtypes = [dp, dp]
        L.oracle_se3_plus.argtypes = [dp, dp, dp]
        L.oracle_se3_exp_quat.argtypes = [dp, dp]
        L.oracle_knn.argtypes = [dp, C.c_size_t, dp, C.c_size_t, C.c_double, C.c_int,
                                 C.POINTER(C.c_int), dp, C.POINTER(C.c_int), C.c_int]
        L.oracle_fit_plane.argtypes = [dp, C.c_int, dp]
        L.oracle_fit_line.argtypes = [dp, C.c_int, C.c_double, dp, dp, dp, dp, dp]
        L.oracle_fit_line.restype = C.c_int
        L.oracle_sym_eig3.argtypes = [dp, dp, dp]
        L.oracle_eval_point_to_point.argtypes = [dp, dp, dp, C.c_double, dp, dp, dp]
        L.oracle_eval_point_to_line.argtypes = [dp, dp, dp, dp, C.c_double, dp, dp, dp]
        L.oracle_eval_point_to_plane.argtypes = [dp, dp, dp, C.c_double, C.c_double, dp, dp, dp]
        L.oracle_build_factors.argtypes = [C.c_void_p, C.c_int, dp, C.POINTER(C.c_int), dp, C.c_size_t]
        L.oracle_build_factors.restype = C.c_int
        L.oracle_update_weight.argtypes = [dp, dp, C.c_size_t, C.c_double, C.c_double, C.c_double, C.c_double]
        L.oracle_min_on_boundary_2d.argtypes = [dp, dp, C.c_double, dp]
        L.oracle_min_on_boundary_2d.restype = None
        L.oracle_submap_default_config.argtypes = [C.POINTER(SubmapConfig)]
        L.oracle_voxel_down_sample.argtypes = [dp, C.c_size_t, C.c_double, dp]
        L.oracle_voxel_down_sample.restype = C.c_size_t
        L.oracle_crop.argtypes = [dp, C.c_size_t, dp, dp, dp]
        L.oracle_crop.restype = C.c_size_t
        L.oracle_submap_create.argtypes = [C.POINTER(SubmapConfig)]
        L.oracle_submap_create.restype = C.c_void_p
        L.oracle_submap_destroy.argtypes = [C.c_void_p]
        L.oracle_submap_init.argtypes = [C.c_void_p, dp, C.c_size_t, dp, C.c_size_t, dp, C.c_size_t, dp, C.c_size_t]
        L.oracle_submap_update.argtypes = [C.c_void_p, dp, dp, C.c_size_t, dp, C.c_size_t, dp, C.c_size_t, dp, C.c_size_t]
        L.oracle_submap_size.argtypes = [C.c_void_p, C.c_int]
        L.oracle_submap_size.restype = C.c_size_t
        L.oracle_submap_data.argtypes = [C.c_void_p, C.c_int]
        L.oracle_submap_data.restype = dp
        szp, ip = C.POINTER(C.c_size_t), C.POINTER(C.c_int)
        L.oracle_feature_default_config.argtypes = [C.POINTER(FeatureConfig)]
        L.oracle_pca_info.argtypes = [dp, C.c_size_t, C.POINTER(FeatureConfig), dp, dp, dp, dp, ip, ip]
        L.oracle_pca_info.restype = C.c_int
        L.oracle_extract_planar_sphere.argtypes = [dp, C.c_size_t, C.POINTER(FeatureConfig), szp, szp, szp, szp, szp, szp,
                                                   szp, szp, szp]
        L.oracle_extract_planar_sphere.restype = C.c_int
        L.oracle_ground_default_config.argtypes = [C.POINTER(GroundConfig)]
        L.oracle_fast_atan2.argtypes = [C.c_float, C.c_float]
        L.oracle_fast_atan2.restype = C.c_float
        L.oracle_ground_section_bounds.argtypes = [C.POINTER(GroundConfig), C.POINTER(C.c_float), C.c_int]
        L.oracle_ground_extract.argtypes = [dp, C.c_size_t, C.POINTER(GroundConfig), szp, szp, szp, szp, ip, ip, dp, dp]
        L.oracle_ground_extract.restype = C.c_int
        L.oracle_extract_edge.argtypes = [dp, dp, C.c_size_t, C.c_int, C.c_int, C.c_int, szp, szp, szp, szp]
        L.oracle_extract_edge.restype = C.c_int
        L.oracle_dcvc_default_config.argtypes = [C.POINTER(DcvcConfig)]
        L.oracle_dcvc_polar.argtypes = [dp, C.c_size_t, C.POINTER(DcvcConfig), dp, dp]
        L.oracle_dcvc_from_polar.argtypes = [dp, dp, dp, C.c_size_t, C.POINTER(DcvcConfig), C.c_int, ip, ip, ip, szp, szp, ip, ip, dp]
        L.oracle_dcvc_from_polar.restype = C.c_int
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def default_config(**overrides):
    cfg = Config()
    lib().oracle_default_config(C.byref(cfg))
    for k, v in overrides.items():
        if k == "reinit_dir":
            for i in range(3):
                cfg.reinit_dir[i] = float(v[i])
        else:
            setattr(cfg, k, v)
    return cfg


def _cloud_args(clouds):
    arrs = [_f64(clouds[i] if not isinstance(clouds, dict) else clouds[CLOUDS[i]]).reshape(-1, 3) for i in range(4)]
    ptrs = (C.POINTER(C.c_double) * 4)(*[_dp(a) for a in arrs])
    ns = (C.c_size_t * 4)(*[a.shape[0] for a in arrs])
    return arrs, ptrs, ns


class Oracle:
    """Mirror of tloam::LocalRegistration (ref: include/tloam/models/registration/registration.hpp:142-165)."""

    def __init__(self, cfg=None, **overrides):
        self.cfg = cfg if cfg is not None else default_config(**overrides)
        self._h = lib().oracle_create(C.byref(self.cfg))
        self._n_src = [0, 0, 0, 0]

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_destroy(self._h)
            self._h = None

    def set_input_source(self, clouds):
        arrs, ptrs, ns = _cloud_args(clouds)
        self._n_src = [a.shape[0] for a in arrs]
        return lib().oracle_set_source(self._h, ptrs, ns)

    def set_input_target(self, clouds):
        arrs, ptrs, ns = _cloud_args(clouds)
        return lib().oracle_set_target(self._h, ptrs, ns)

    def scan_matching(self, predict):
        """predict: 4x4 (row-major numpy). Returns (status, result 4x4, Stats)."""
        p = _f64(np.asarray(predict).T).reshape(16)  # column-major
        out = np.zeros(16)
        st = Stats()
        rc = lib().oracle_scan_match(self._h, _dp(p), _dp(out), C.byref(st))
        return rc, out.reshape(4, 4).T.copy(), st

    def fitness(self):
        a = C.c_double(0)
        b = C.c_double(0)
        lib().oracle_fitness(self._h, C.byref(a), C.byref(b))
        return a.value, b.value

    def pose_increment(self):
        out = np.zeros(16)
        lib().oracle_get_pose_increment(self._h, _dp(out))
        return out.reshape(4, 4).T.copy()

    def weights(self, cloud):
        w = np.zeros(self._n_src[cloud])
        rc = lib().oracle_get_weights(self._h, cloud, _dp(w), w.size)
        assert rc == 0
        return w

    def build_factors(self, cloud, x):
        n = self._n_src[cloud]
        valid = np.zeros(n, dtype=np.int32)
        prim = np.zeros((n, 6))
        x = _f64(x)
        rc = lib().oracle_build_factors(self._h, cloud, _dp(x), valid.ctypes.data_as(C.POINTER(C.c_int)), _dp(prim), n)
        assert rc == 0
        return valid, prim


def se3_exp(a):
    a = _f64(a)
    T = np.zeros(16)
    lib().oracle_se3_exp(_dp(a), _dp(T))
    return T.reshape(4, 4).T.copy()


def se3_log(T):
    t = _f64(np.asarray(T).T).reshape(16)
    a = np.zeros(6)
    lib().oracle_se3_log(_dp(t), _dp(a))
    return a


def se3_plus(x, d):
    x = _f64(x)
    d = _f64(d)
    out = np.zeros(6)
    lib().oracle_se3_plus(_dp(x), _dp(d), _dp(out))
    return out


def knn(pts, queries, radius, k, brute_force=False):
    pts = _f64(pts).reshape(-1, 3)
    q = _f64(queries).reshape(-1, 3)
    nq = q.shape[0]
    idx = np.full((nq, k), -1, dtype=np.int32)
    d2 = np.full((nq, k), np.inf)
    cnt = np.zeros(nq, dtype=np.int32)
    lib().oracle_knn(_dp(pts), pts.shape[0], _dp(q), nq, float(radius), int(k),
                     idx.ctypes.data_as(C.POINTER(C.c_int)), _dp(d2), cnt.ctypes.data_as(C.POINTER(C.c_int)),
                     1 if brute_force else 0)
    return idx, d2, cnt


def fit_plane(pts):
    pts = _f64(pts).reshape(-1, 3)
    out = np.zeros(4)
    lib().oracle_fit_plane(_dp(pts), pts.shape[0], _dp(out))
    return out


def fit_line(pts, dir_thres=0.85):
    pts = _f64(pts).reshape(-1, 3)
    a, b, mean, d, eig = (np.zeros(3) for _ in range(5))
    ok = lib().oracle_fit_line(_dp(pts), pts.shape[0], float(dir_thres), _dp(a), _dp(b), _dp(mean), _dp(d), _dp(eig))
    return ok, a, b, mean, d, eig


def sym_eig3(cov):
    cov = _f64(cov).reshape(9)
    eig = np.zeros(3)
    vec = np.zeros(9)
    lib().oracle_sym_eig3(_dp(cov), _dp(eig), _dp(vec))
    return eig, vec.reshape(3, 3).T.copy()  # columns = eigenvectors


def eval_point_to_point(x, p, q, w):
    x, p, q = _f64(x), _f64(p), _f64(q)
    r, J, c = np.zeros(3), np.zeros(18), C.c_double(0)
    lib().oracle_eval_point_to_point(_dp(x), _dp(p), _dp(q), float(w), _dp(r), _dp(J), C.byref(c))
    return r, J.reshape(3, 6), c.value


def eval_point_to_line(x, p, a, b, w):
    x, p, a, b = _f64(x), _f64(p), _f64(a), _f64(b)
    r, J, c = np.zeros(3), np.zeros(18), C.c_double(0)
    lib().oracle_eval_point_to_line(_dp(x), _dp(p), _dp(a), _dp(b), float(w), _dp(r), _dp(J), C.byref(c))
    return r, J.reshape(3, 6), c.value


def eval_point_to_plane(x, p, n, d, w):
    x, p, n = _f64(x), _f64(p), _f64(n)
    r, J, c = np.zeros(1), np.zeros(6), C.c_double(0)
    lib().oracle_eval_point_to_plane(_dp(x), _dp(p), _dp(n), float(d), float(w), _dp(r), _dp(J), C.byref(c))
    return r, J.reshape(1, 6), c.value


def update_weight(weights, slots, noise_bound_sq, th1, th2, mu):
    w = _f64(weights).copy()
    s = _f64(slots)
    lib().oracle_update_weight(_dp(w), _dp(s), w.size, noise_bound_sq, th1, th2, mu)
    return w


def voxel_down_sample(pts, voxel):
    a = _f64(pts).reshape(-1, 3)
    out = np.zeros_like(a)
    n = lib().oracle_voxel_down_sample(_dp(a), a.shape[0], float(voxel), _dp(out))
    return out[:n].copy()


def crop(pts, lo, hi):
    a = _f64(pts).reshape(-1, 3)
    out = np.zeros_like(a)
    lo, hi = _f64(lo), _f64(hi)
    n = lib().oracle_crop(_dp(a), a.shape[0], _dp(lo), _dp(hi), _dp(out))
    return out[:n].copy()


class Submap:
    """Mirror of FrontEnd's submap handling (ref: src/front_end/front_end.cpp:201-267, 285-305)."""

    def __init__(self, **overrides):
        self.cfg = SubmapConfig()
        lib().oracle_submap_default_config(C.byref(self.cfg))
        for k, v in overrides.items():
            setattr(self.cfg, k, v)
        self._s = lib().oracle_submap_create(C.byref(self.cfg))

    def __del__(self):
        if getattr(self, "_s", None):
            lib().oracle_submap_destroy(self._s)
            self._s = None

    def init(self, edge, ground_raw, planar_sub, sphere_sub):
        a = [_f64(x).reshape(-1, 3) for x in (edge, ground_raw, planar_sub, sphere_sub)]
        lib().oracle_submap_init(self._s, _dp(a[0]), a[0].shape[0], _dp(a[1]), a[1].shape[0], _dp(a[2]), a[2].shape[0],
                                 _dp(a[3]), a[3].shape[0])

    def update(self, pose, edge_scan, ground_scan, planar_sub, sphere_sub):
        p = _f64(np.asarray(pose).T).reshape(16)
        a = [_f64(x).reshape(-1, 3) for x in (edge_scan, ground_scan, planar_sub, sphere_sub)]
        lib().oracle_submap_update(self._s, _dp(p), _dp(a[0]), a[0].shape[0], _dp(a[1]), a[1].shape[0], _dp(a[2]),
                                   a[2].shape[0], _dp(a[3]), a[3].shape[0])

    def cloud(self, c):
        n = lib().oracle_submap_size(self._s, c)
        if n == 0:
            return np.zeros((0, 3))
        ptr = lib().oracle_submap_data(self._s, c)
        return np.ctypeslib.as_array(ptr, shape=(n * 3,)).reshape(n, 3).copy()

    def clouds(self):
        return [self.cloud(c) for c in range(4)]


# ---- "next" row (f)-2: PCA feature extraction (ref: src/models/feature_extraction/feature_extract.cpp) ----
def feature_config(**overrides):
    c = FeatureConfig()
    lib().oracle_feature_default_config(C.byref(c))
    for k, v in overrides.items():
        setattr(c, k, v)
    return c


def pca_info(pts, **overrides):
    """calculatePCAInfo: dict of per-point cvr / flatness / sphericity / normal (n,3) / num_sum / neigh (n,K)."""
    a = _f64(pts).reshape(-1, 3)
    c = feature_config(**overrides)
    n = a.shape[0]
    out = {"cvr": np.zeros(n), "flatness": np.zeros(n), "sphericity": np.zeros(n), "normal": np.zeros((n, 3)),
           "num_sum": np.zeros(n, dtype=np.int32), "neigh": np.full((n, c.K), -1, dtype=np.int32)}
    ip = C.POINTER(C.c_int)
    rc = lib().oracle_pca_info(_dp(a), n, C.byref(c), _dp(out["cvr"]), _dp(out["flatness"]), _dp(out["sphericity"]),
                               _dp(out["normal"]), out["num_sum"].ctypes.data_as(ip), out["neigh"].ctypes.data_as(ip))
    if rc != 0:
        raise ValueError("oracle_pca_info: empty cloud or bad configuration")
    return out


def extract_planar_sphere(pts, **overrides):
    """extractPlanarSphere: (planar_scan, planar_submap, sphere_scan, sphere_submap, sphere_candidates)."""
    a = _f64(pts).reshape(-1, 3)
    c = feature_config(**overrides)
    n = a.shape[0]
    bufs = [np.zeros(max(n, 1), dtype=np.uintp) for _ in range(5)]
    cnt = [C.c_size_t(0) for _ in range(4)]
    szp = C.POINTER(C.c_size_t)
    lib().oracle_extract_planar_sphere(_dp(a), n, C.byref(c), bufs[0].ctypes.data_as(szp), C.byref(cnt[0]),
                                       bufs[1].ctypes.data_as(szp), C.byref(cnt[1]), bufs[2].ctypes.data_as(szp),
                                       C.byref(cnt[2]), bufs[3].ctypes.data_as(szp), C.byref(cnt[3]),
                                       bufs[4].ctypes.data_as(szp))
    return (bufs[0][:cnt[0].value].copy(), bufs[1][:cnt[1].value].copy(), bufs[2][:cnt[2].value].copy(),
            bufs[3][:cnt[3].value].copy(), bufs[4][:cnt[3].value].copy())


def min_on_boundary_2d(B, g, radius):
    B = _f64(B).reshape(4)
    g = _f64(g).reshape(2)
    y = np.zeros(2)
    lib().oracle_min_on_boundary_2d(_dp(B), _dp(g), float(radius), _dp(y))
    return y


def ground_config(**overrides):
    c = GroundConfig()
    lib().oracle_ground_default_config(C.byref(c))
    for k, v in overrides.items():
        setattr(c, k, v)
    return c


def fast_atan2(y, x):
    return float(lib().oracle_fast_atan2(float(y), float(x)))


def ground_section_bounds(**overrides):
    c = ground_config(**overrides)
    out = (C.c_float * 8)()
    n = lib().oracle_ground_section_bounds(C.byref(c), out, 8)
    return [float(out[i]) for i in range(n)]


def ground_extract(pts, **overrides):
    """Segmentation::groundRemove restated.  Returns dict(ground, object, beam, region, height_threshold, planes)."""
    a = _f64(pts).reshape(-1, 3)
    n = a.shape[0]
    c = ground_config(**overrides)
    g = np.zeros(max(n, 1), dtype=np.uintp)
    o = np.zeros(max(n, 1), dtype=np.uintp)
    ng, no = C.c_size_t(0), C.c_size_t(0)
    beam = np.zeros(max(n, 1), dtype=np.int32)
    region = np.zeros(max(n, 1), dtype=np.int32)
    thr = C.c_double(0)
    planes = np.zeros((12, 8, 4))
    szp, ip = C.POINTER(C.c_size_t), C.POINTER(C.c_int)
    rc = lib().oracle_ground_extract(_dp(a), n, C.byref(c), g.ctypes.data_as(szp), C.byref(ng), o.ctypes.data_as(szp), C.byref(no),
                                     beam.ctypes.data_as(ip), region.ctypes.data_as(ip), C.byref(thr), _dp(planes))
    assert rc == 0
    return dict(ground=g[:ng.value].copy(), object=o[:no.value].copy(), beam=beam[:n].copy(), region=region[:n].copy(),
                height_threshold=thr.value, planes=planes)


def extract_edge(pts, intensity, sensor_model=64, ring_min_num=16, max_section=4096):
    """Segmentation::extractEdgePoint restated (ref: segmentation.cpp:1144-1304).  Returns dict(edge, non_edge) of index
    lists into the input, in the reference's append order; None if a sector exceeds max_section (the device limit)."""
    a = _f64(pts).reshape(-1, 3)
    it = _f64(intensity).reshape(-1)
    n = a.shape[0]
    assert it.shape[0] == n
    e = np.zeros(max(n, 1), dtype=np.uintp)
    o = np.zeros(max(n, 1), dtype=np.uintp)
    ne, no = C.c_size_t(0), C.c_size_t(0)
    szp = C.POINTER(C.c_size_t)
    rc = lib().oracle_extract_edge(_dp(a), _dp(it), n, sensor_model, ring_min_num, max_section, e.ctypes.data_as(szp), C.byref(ne),
                                   o.ctypes.data_as(szp), C.byref(no))
    if rc != 0:
        return None
    return dict(edge=e[:ne.value].copy(), non_edge=o[:no.value].copy())


def dcvc_config(**overrides):
    c = DcvcConfig()
    lib().oracle_dcvc_default_config(C.byref(c))
    for k, v in overrides.items():
        setattr(c, k, v)
    return c


def dcvc_polar(pts, **overrides):
    """convertToPolar's triples (range, pitch deg, azimuth deg; zeros for out-of-range points) and the 4 extrema."""
    a = _f64(pts).reshape(-1, 3)
    n = a.shape[0]
    c = dcvc_config(**overrides)
    polar = np.zeros((max(n, 1), 3))
    ext = np.zeros(4)
    lib().oracle_dcvc_polar(_dp(a), n, C.byref(c), _dp(polar), _dp(ext))
    return polar[:n], ext


def dcvc_from_polar(pts, polar, ext, max_bounds=0, **overrides):
    """createHashTable + DCVC + labelAnalysis + colorSegmentation on given polar triples (literal restatement).  Returns
    dict(voxel, root, cluster, segmented, sizes, boxes) or None if polarNum exceeds max_bounds."""
    a = _f64(pts).reshape(-1, 3)
    n = a.shape[0]
    pol = _f64(polar).reshape(-1, 3)
    if n == 0:
        pol = np.zeros((1, 3))
    ex = _f64(ext).reshape(4)
    c = dcvc_config(**overrides)
    szp, ip = C.POINTER(C.c_size_t), C.POINTER(C.c_int)
    m = max(n, 1)
    voxel, root, cluster, sizes = (np.zeros(m, dtype=np.int32) for _ in range(4))
    seg = np.zeros(m, dtype=np.uintp)
    boxes = np.zeros((m, 6))
    nseg, ncl = C.c_size_t(0), C.c_int(0)
    rc = lib().oracle_dcvc_from_polar(_dp(a), _dp(pol), _dp(ex), n, C.byref(c), max_bounds, voxel.ctypes.data_as(ip),
                                      root.ctypes.data_as(ip), cluster.ctypes.data_as(ip), seg.ctypes.data_as(szp), C.byref(nseg),
                                      C.byref(ncl), sizes.ctypes.data_as(ip), _dp(boxes))
    if rc != 0:
        return None
    k = ncl.value
    return dict(voxel=voxel[:n].copy(), root=root[:n].copy(), cluster=cluster[:n].copy(), segmented=seg[:nseg.value].copy(),
                sizes=sizes[:k].copy(), boxes=boxes[:k].copy())


def dcvc(pts, max_bounds=0, **overrides):
    """Segmentation::objectSegmentation restated.  dict(polar, ext, voxel, root, cluster, segmented, sizes, boxes)."""
    polar, ext = dcvc_polar(pts, **overrides)
    r = dcvc_from_polar(pts, polar, ext, max_bounds=max_bounds, **overrides)
    if r is not None:
        r["polar"], r["ext"] = polar, ext
    return r
