"""(f)-2: PCA feature extraction.  CPU tests pin the restatement (oracle/feature_oracle.cpp) against independent
numpy / scipy computations; GPU tests hold the CUDA path (tloam_b200/csrc/feature_extract.cuh, through the C ABI) to
BIT-EXACT equality with the restatement -- the outputs are index lists.
ref: src/models/feature_extraction/feature_extract.cpp:47-122 (calculatePCAInfo), 133-197 (extractPlanarSphere)."""
import numpy as np
import pytest

from tloam_b200 import synth

CFG = dict(radius=0.2, K=20, min_neigh=10, planar_num=500, sphere_num=300, cvr_scan=0.25, cvr_submap=0.15,
           planar_scan_thres=0.75, planar_submap_thres=0.65, planar_vertic_thres=0.25)


def numpy_pca(pts, radius, K, min_neigh):
    """Independent restatement: scipy cKDTree neighbours, numpy raw-moment covariance, numpy eigh."""
    from scipy.spatial import cKDTree
    tree = cKDTree(pts)
    d, idx = tree.query(pts, k=K, distance_upper_bound=radius)
    n = pts.shape[0]
    cvr, flat, sph, nrm, num = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros((n, 3)), np.zeros(n, dtype=int)
    neigh = []
    for i in range(n):
        ok = np.isfinite(d[i]) & (d[i] ** 2 < radius * radius)
        nb = idx[i][ok]
        neigh.append(nb)
        if nb.size <= min_neigh:
            continue
        q = pts[nb]
        mean = q.mean(0)
        cov = (q[:, :, None] * q[:, None, :]).mean(0) - np.outer(mean, mean)
        w, v = np.linalg.eigh(cov)
        cvr[i] = 0.0 if w.sum() == 0 else w[0] / w.sum()
        flat[i] = (w[1] - w[0]) / w[2]
        sph[i] = w[0] / w[2]
        nrm[i] = v[:, 0]
        num[i] = nb.size
    return cvr, flat, sph, nrm, num, neigh


def numpy_select(info, cfg):
    """Independent restatement of the selection, quirks included (sphere lists = ranks, thresholded by flatness)."""
    cvr, flat, nrm, neigh = info["cvr"], info["flatness"], info["normal"], info["neigh"]
    planar, sphere = [], []
    for i in range(cvr.size):
        if flat[i] > cfg["planar_submap_thres"] and abs(nrm[i, 2]) < cfg["planar_vertic_thres"]:
            planar.append(i)
        elif cvr[i] > cfg["cvr_submap"]:
            nb = neigh[i][neigh[i] >= 0]
            if not np.any(cvr[i] < cvr[nb]):
                sphere.append(i)
    planar = sorted(planar, key=lambda i: (-flat[i], i))
    sphere = sorted(sphere, key=lambda i: (-flat[i], i))
    p_scan = [i for r, i in enumerate(planar) if r < cfg["planar_num"] or flat[i] > cfg["planar_scan_thres"]]
    s_scan = [r for r, i in enumerate(sphere) if r < cfg["sphere_num"] or flat[i] > cfg["cvr_scan"]]
    return p_scan, planar, s_scan, list(range(len(sphere))), sphere


# ------------------------------------------------------------------------------------------------
# oracle pins (CPU)
# ------------------------------------------------------------------------------------------------
def test_oracle_pca_vs_numpy_scipy(oracle):
    pts = synth.general_cloud(4000, seed=11)
    info = oracle.pca_info(pts, **CFG)
    cvr, flat, sph, nrm, num, neigh = numpy_pca(pts, CFG["radius"], CFG["K"], CFG["min_neigh"])
    assert np.array_equal(info["num_sum"], num)
    kept = num > 0
    assert 0.3 < kept.mean() < 1.0 and (num == CFG["K"]).any() and ((num > 0) & (num < CFG["K"])).any()
    for i in np.flatnonzero(kept)[:500]:                           # neighbour SETS (and order: ascending distance)
        assert np.array_equal(info["neigh"][i][: num[i]], neigh[i])
    # raw second moments of coordinates up to 45 m cancel down to ~1e-3 m^2: ~1e-9 relative noise is inherent
    assert np.allclose(info["cvr"], cvr, rtol=1e-5, atol=1e-8)
    assert np.allclose(info["flatness"], flat, rtol=1e-5, atol=1e-7)
    assert np.allclose(info["sphericity"], sph, rtol=1e-5, atol=1e-8)
    good = kept & (flat > 0.3)                                     # normal well defined when lambda1 - lambda0 is not tiny
    dots = np.abs(np.sum(info["normal"][good] * nrm[good], axis=1))
    assert dots.min() > 1 - 1e-6
    # skipped points keep the value-initialised PCAInfo
    assert not info["cvr"][~kept].any() and not info["normal"][~kept].any() and (info["neigh"][~kept] == -1).all()


def test_oracle_selection_vs_numpy(oracle):
    pts = synth.general_cloud(6000, seed=12)
    info = oracle.pca_info(pts, **CFG)
    for over in ({}, {"planar_num": 50, "sphere_num": 3}, {"planar_scan_thres": 0.9, "cvr_scan": 0.6}):
        cfg = dict(CFG, **over)
        got = oracle.extract_planar_sphere(pts, **cfg)
        ref = numpy_select(info, cfg)
        for g, r in zip(got, ref):
            assert np.array_equal(g, np.asarray(r, dtype=np.uintp))
    p_scan, p_sub, s_scan, s_sub, s_cand = oracle.extract_planar_sphere(pts, **CFG)
    assert len(p_sub) > CFG["planar_num"] and len(p_scan) >= CFG["planar_num"] and len(s_sub) > 5
    # quirk FE-1: the sphere lists are ranks
    assert np.array_equal(s_sub, np.arange(len(s_sub))) and np.array_equal(s_scan, np.arange(len(s_scan)))
    # the planar lists are sorted by descending flatness, scan is a prefix of submap
    assert np.all(np.diff(info["flatness"][p_sub]) <= 0) and np.array_equal(p_scan, p_sub[: len(p_scan)])


def test_oracle_feature_edge_cases(oracle):
    assert [len(x) for x in oracle.extract_planar_sphere(np.zeros((0, 3)), **CFG)] == [0, 0, 0, 0, 0]
    with pytest.raises(ValueError):
        oracle.pca_info(np.zeros((0, 3)), **CFG)
    few = np.random.default_rng(0).normal(0, 0.05, (8, 3))          # fewer points than min_neigh: nothing kept
    info = oracle.pca_info(few, **CFG)
    assert not info["num_sum"].any()
    assert [len(x) for x in oracle.extract_planar_sphere(few, **CFG)] == [0, 0, 0, 0, 0]
    dup = np.tile(np.array([[1.0, 2.0, 3.0]]), (30, 1))              # zero covariance: cvr = 0 (sum == 0), flatness NaN
    info = oracle.pca_info(dup, **CFG)
    assert (info["num_sum"] == 20).all() and not info["cvr"].any() and np.isnan(info["flatness"]).all()
    assert [len(x) for x in oracle.extract_planar_sphere(dup, **CFG)] == [0, 0, 0, 0, 0]


# ------------------------------------------------------------------------------------------------
# CUDA path vs the oracle (GPU): bit-exact
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def reg():
    import tloam_b200
    r = tloam_b200.LocalRegistration()
    yield r
    r.close()


def assert_info_equal(a, b):
    for k in ("num_sum", "neigh", "cvr", "flatness", "sphericity"):
        assert np.array_equal(a[k], b[k], equal_nan=True), k
    # eigenvector sign is a free choice of the solver; both sides run the same iteration, so it is equal as well
    assert np.array_equal(a["normal"], b["normal"])


@pytest.mark.gpu
@pytest.mark.parametrize("n,seed", [(3000, 1), (20000, 2), (50000, 3)])
def test_gpu_pca_info_bit_exact(reg, oracle, n, seed):
    pts = synth.general_cloud(n, seed=seed)
    assert_info_equal(reg.pca_info(pts, **CFG), oracle.pca_info(pts, **CFG))


@pytest.mark.gpu
def test_gpu_extract_planar_sphere_bit_exact(reg, oracle):
    pts = synth.general_cloud(50000, seed=4)
    for over in ({}, {"K": 12, "min_neigh": 6}, {"planar_num": 20, "sphere_num": 2, "radius": 0.3},
                 {"planar_scan_thres": 0.9, "cvr_scan": 0.6, "cvr_submap": 0.05}):
        cfg = dict(CFG, **over)
        got, ref = reg.extract_planar_sphere(pts, **cfg), oracle.extract_planar_sphere(pts, **cfg)
        assert len(ref[1]) > 100
        for g, r in zip(got, ref):
            assert np.array_equal(g, r)


@pytest.mark.gpu
def test_gpu_extract_planar_sphere_long_lists_bit_exact(reg, oracle):
    """More than 16 384 candidates in a list: the sort leaves its single shared-memory tile (distances >= the tile size in
    global memory, the rest tile by tile)."""
    pts = synth.general_cloud(160000, seed=8)
    cfg = dict(CFG, planar_submap_thres=0.3, planar_vertic_thres=0.9, cvr_submap=0.01)
    got, ref = reg.extract_planar_sphere(pts, **cfg), oracle.extract_planar_sphere(pts, **cfg)
    assert len(ref[1]) > 16384
    for g, r in zip(got, ref):
        assert np.array_equal(g, r)


@pytest.mark.gpu
def test_gpu_feature_far_from_origin_and_lattice_ties(reg, oracle):
    """Coordinates around 1 km (raw-moment cancellation is then ~1e-6 relative: still identical on both sides) and a
    regular lattice, where many neighbours are exactly equidistant (ordering by (d2, index))."""
    pts = synth.general_cloud(8000, seed=5) + np.array([1000.0, -750.0, 20.0])
    assert_info_equal(reg.pca_info(pts, **CFG), oracle.pca_info(pts, **CFG))
    g = np.arange(0, 1.5, 0.0625)
    lat = np.stack(np.meshgrid(g, g, g[:6], indexing="ij"), -1).reshape(-1, 3)
    assert_info_equal(reg.pca_info(lat, **CFG), oracle.pca_info(lat, **CFG))
    for a, b in zip(reg.extract_planar_sphere(lat, **CFG), oracle.extract_planar_sphere(lat, **CFG)):
        assert np.array_equal(a, b)


@pytest.mark.gpu
def test_gpu_feature_edge_cases(reg, oracle):
    from tloam_b200 import _lib
    from tloam_b200.registration import RegistrationError
    assert [len(x) for x in reg.extract_planar_sphere(np.zeros((0, 3)), **CFG)] == [0, 0, 0, 0, 0]
    few = np.random.default_rng(0).normal(0, 0.05, (8, 3))
    assert [len(x) for x in reg.extract_planar_sphere(few, **CFG)] == [0, 0, 0, 0, 0]
    dup = np.tile(np.array([[1.0, 2.0, 3.0]]), (30, 1))
    assert_info_equal(reg.pca_info(dup, **CFG), oracle.pca_info(dup, **CFG))
    with pytest.raises(RegistrationError) as ei:                     # K above the compiled list capacity
        reg.extract_planar_sphere(few, **dict(CFG, K=21))
    assert ei.value.status == _lib.ERR_INVALID_ARG
    with pytest.raises(RegistrationError):
        reg.extract_planar_sphere(few, **dict(CFG, radius=0.0))


@pytest.mark.gpu
def test_gpu_feature_full_size_properties(reg):
    """At the size of a real general cloud (120k points): structural properties that need no oracle."""
    pts = synth.general_cloud(120000, seed=6)
    info = reg.pca_info(pts, **CFG)
    p_scan, p_sub, s_scan, s_sub, s_cand = reg.extract_planar_sphere(pts, **CFG)
    f = info["flatness"]
    assert np.all(np.diff(f[p_sub]) <= 0) and np.array_equal(p_scan, p_sub[: len(p_scan)])
    assert np.all(f[p_sub] > CFG["planar_submap_thres"]) and np.all(np.abs(info["normal"][p_sub, 2]) < CFG["planar_vertic_thres"])
    assert len(p_scan) == max(min(CFG["planar_num"], len(p_sub)), int((f[p_sub] > CFG["planar_scan_thres"]).sum()))
    assert np.array_equal(s_sub, np.arange(len(s_sub))) and np.all(np.diff(f[s_cand]) <= 0)
    assert np.all(info["cvr"][s_cand] > CFG["cvr_submap"]) and len(set(p_sub) & set(s_cand)) == 0
    nb = info["neigh"][s_cand]
    nbc = np.where(nb >= 0, info["cvr"][np.maximum(nb, 0)], -1.0)
    assert np.all(nbc <= info["cvr"][s_cand][:, None])               # local maxima of the curvature
    # every point is its own nearest neighbour
    kept = info["num_sum"] > 0
    assert np.array_equal(info["neigh"][kept, 0], np.flatnonzero(kept))
