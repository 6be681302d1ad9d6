"""(f)-4, third part: object segmentation = Dynamic Curved-Voxel Clustering (Segmentation::objectSegmentation, ref:
src/models/segmentation/segmentation.cpp:772-1112).

* the LITERAL oracle (label vector relabelled by full sweeps, unordered_map of point lists) against `structural_dcvc`, an
  independent voxel-level restatement (labelled-prefix states + events + union-find) -- the model the device executes;
* the device path against the oracle: polar triples within 4 ulp of libm's (libdevice asin / atan2, then the conversion to degrees), every integer
  output (voxel index, classes, cluster numbers, segmented scan, sizes) bit-exact, boxes bit-exact."""
import numpy as np
import pytest

from tloam_b200 import synth


def structural_dcvc(polar, ext, start_r=0.35, delta_r=0.0004, delta_p=1.2, delta_a=1.2):
    """Voxel-level restatement of DCVC (:915-990).  State of a voxel = length of its labelled prefix (0, 1, all).  A visit
    of a still-unlabelled point (an EVENT) finds the first labelled entry p of its searchKNN list, labels every listed
    voxel at or after p (all of them if there is none) and joins them with the visitor.  Voxels that are not in their own
    list are in nobody's: every one of their points is an event and a union-find node of its own.
    Returns (root, key): root[i] = smallest point index of i's class, key = the reference's voxelIndex."""
    n = len(polar)
    min_pitch, max_pitch, min_polar, max_polar = ext
    width = int(round(360.0 / delta_a) + 1)
    height = int((max_pitch - min_pitch) / delta_p)
    bounds = []
    rng, step = min_polar, 1
    while rng <= max_polar:
        rng += (start_r - step * delta_r)
        bounds.append(rng)
        step += 1
    polar_num = len(bounds)
    pi = np.minimum(np.searchsorted(np.array(bounds), polar[:, 0], side="right"), polar_num - 1)
    ti = np.floor((polar[:, 1] - min_pitch) / delta_p + 0.5).astype(np.int64)          # arguments are >= 0: round half up
    ai = np.floor(polar[:, 2] / delta_a + 0.5).astype(np.int64)
    key = (ai * (polar_num + 1) + pi) + ti * (polar_num + 1) * (width + 1)
    members, coords = {}, {}
    for i, k in enumerate(key):
        members.setdefault(int(k), []).append(i)
        coords.setdefault(int(k), (int(pi[i]), int(ti[i]), int(ai[i])))
    parent = list(range(n))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    def union(a, c):
        a, c = find(a), find(c)
        if a != c:
            parent[max(a, c)] = min(a, c)

    def knn(V):
        p_, t_, a_ = coords[V]
        out = []
        for z in (t_ - 1, t_, t_ + 1):
            if z < 0 or z > height:
                continue
            for y in (p_ - 1, p_, p_ + 1):
                if y < 0 or y > polar_num:
                    continue
                for x in (a_ - 1, a_, a_ + 1):
                    ax = width - 1 if x < 0 else x
                    ax = 300 if ax > 300 else ax
                    out.append((ax * (polar_num + 1) + y) + z * (polar_num + 1) * (width + 1))
        return out

    lists = {V: knn(V) for V in members}
    own_listed = {V: V in lists[V] for V in members}
    state = {V: 0 for V in members}
    for i in range(n):
        V = int(key[i])
        m = members[V]
        if own_listed[V]:
            if not ((state[V] == 0 and i == m[0]) or (state[V] == 1 and len(m) > 1 and i == m[1])):
                continue
        lst = [k for k in lists[V] if k in members]
        first = next((q for q, k in enumerate(lst) if state[k] > 0), 0)
        seed = m[0] if own_listed[V] else i
        for k in lst[first:]:
            state[k] = len(members[k])
            union(seed, members[k][0])
        if own_listed[V] and state[V] == 0:
            state[V] = 1
    root = np.array([find(members[int(key[i])][0] if own_listed[int(key[i])] else i) for i in range(n)])
    return root, key


def python_dcvc_labels(polar, ext, start_r=0.35, delta_r=0.0004, delta_p=1.2, delta_a=1.2):
    """Literal transcription of createHashTable + searchKNN + DCVC (:843-990) in plain Python: dict of point lists, the
    label list relabelled by full sweeps.  Returns the raw label list (labels are arbitrary: compare partitions)."""
    n = len(polar)
    min_pitch, max_pitch, min_polar, max_polar = (float(v) for v in ext)
    width = int(round(360.0 / delta_a) + 1)
    height = int((max_pitch - min_pitch) / delta_p)
    bounds, rng, step = [], min_polar, 1
    while rng <= max_polar:
        rng += (start_r - step * delta_r)
        bounds.append(rng)
        step += 1
    polar_num = len(bounds)

    def polar_index(radius):
        for r in range(polar_num):
            if radius < bounds[r]:
                return r
        return polar_num - 1

    def c_round(x):                                    # std::round: half away from zero
        return int(np.floor(abs(x) + 0.5)) * (1 if x >= 0 else -1)

    idx3, voxel_map = [], {}
    for i in range(n):
        pi_, ti_, ai_ = polar_index(polar[i][0]), c_round((polar[i][1] - min_pitch) / delta_p), c_round(polar[i][2] / delta_a)
        idx3.append((pi_, ti_, ai_))
        voxel_map.setdefault((ai_ * (polar_num + 1) + pi_) + ti_ * (polar_num + 1) * (width + 1), []).append(i)
    label, count = [-1] * n, 0
    for i in range(n):
        if label[i] != -1:
            continue
        pi_, ti_, ai_ = idx3[i]
        neighbors = []
        for z in range(ti_ - 1, ti_ + 2):
            if z < 0 or z > height:
                continue
            for y in range(pi_ - 1, pi_ + 2):
                if y < 0 or y > polar_num:
                    continue
                for x in range(ai_ - 1, ai_ + 2):
                    ax = x
                    if ax < 0:
                        ax = width - 1
                    if ax > 300:
                        ax = 300
                    neighbors += voxel_map.get((ax * (polar_num + 1) + y) + z * (polar_num + 1) * (width + 1), [])
        for j in neighbors:
            curr, neigh = label[i], label[j]
            if curr != -1 and neigh != -1 and curr != neigh:
                label = [neigh if s == curr else s for s in label]
            elif neigh != -1:
                label[i] = neigh
            elif curr != -1:
                label[j] = curr
        if label[i] == -1:
            count += 1
            label[i] = count
            for j in neighbors:
                label[j] = count
    return label


def canonical(label):
    first = {}
    for i, l in enumerate(label):
        first.setdefault(l, i)
    return np.array([first[l] for l in label])


def test_literal_oracle_against_a_literal_python_transcription(oracle):
    """Second, independent pin of the C++ restatement: the reference's loops transcribed statement by statement in Python
    (small clouds: the label sweeps are quadratic)."""
    for p in random_clouds(16, seed=21, nmax=700):
        for kw in ({}, dict(delta_a=0.6)):
            r = oracle.dcvc(p, **kw)
            lab = python_dcvc_labels(r["polar"], r["ext"], **kw)
            assert np.array_equal(canonical(lab), r["root"]), kw


def object_scan(oracle, seed=20260924 + 5151, **kw):
    scan = synth.raw_scan(seed=seed, **kw)
    g = oracle.ground_extract(scan)
    return np.ascontiguousarray(scan[g["object"]])                # the reference's object_scan: region by region, then the tall points


@pytest.fixture(scope="module")
def obj(oracle):
    return object_scan(oracle)


def random_clouds(count, seed=0, nmax=3000):
    rng = np.random.default_rng(seed)
    for trial in range(count):
        n = int(rng.integers(1, nmax))
        kind = trial % 4
        if kind == 0:                                          # a shell of isolated points, out-of-range ones included
            p = rng.normal(0, 1, (n, 3))
            p = p / np.linalg.norm(p, axis=1)[:, None] * rng.uniform(0.5, 130, (n, 1))
        elif kind == 1:                                        # a dense patch across azimuth 0 / 360 (the 300 clamp, the -1 wrap)
            az, el, r = rng.normal(0, 0.05, n), rng.uniform(-0.05, 0.05, n), rng.uniform(5, 8, n)
            p = np.stack([r * np.cos(el) * np.cos(az), r * np.cos(el) * np.sin(az), r * np.sin(el)], 1)
        elif kind == 2:                                        # a blob in a handful of voxels
            p = rng.normal(0, 0.3, (n, 3)) + np.array([10, 3, 1.0])
        else:                                                  # a lidar-like sweep
            az = np.sort(rng.uniform(0, 2 * np.pi, n))
            el = rng.choice(np.radians(np.arange(-24, 3, 0.4)), n)
            r = 10 + 3 * np.sin(3 * az) + rng.normal(0, 0.02, n)
            p = np.stack([r * np.cos(el) * np.cos(az), r * np.cos(el) * np.sin(az), r * np.sin(el)], 1)
        yield np.ascontiguousarray(p)


def test_literal_oracle_against_the_structural_model(oracle, obj):
    for p in (obj, obj[::-1].copy(), object_scan(oracle, seed=9, n_az=900)):
        r = oracle.dcvc(p)
        root, key = structural_dcvc(r["polar"], r["ext"])
        assert np.array_equal(key, r["voxel"])
        assert np.array_equal(root, r["root"])
    top = 0
    for p in random_clouds(120):
        r = oracle.dcvc(p)
        root, key = structural_dcvc(r["polar"], r["ext"])
        assert np.array_equal(key, r["voxel"]) and np.array_equal(root, r["root"])
        height = int((r["ext"][1] - r["ext"][0]) / 1.2)
        top += int((np.round((r["polar"][:, 1] - r["ext"][0]) / 1.2) > height).any())
    assert top > 10                                            # voxels in pitch layer height + 1 (not their own neighbours) occur


def test_literal_oracle_against_the_structural_model_other_configs(oracle):
    """delta_a = 0.6: azimuth indices up to 600 but searchKNN clamps at the literal 300 -> half of the voxels are in nobody's
    list; the first-frame members (min / max polar = 5.0)."""
    for p in random_clouds(24, seed=5):
        for kw in (dict(delta_a=0.6), dict(delta_p=0.4, start_r=0.5), dict(min_polar_init=5.0, max_polar_init=5.0)):
            r = oracle.dcvc(p, **kw)
            mk = {k: v for k, v in kw.items() if k in ("start_r", "delta_r", "delta_p", "delta_a")}
            root, key = structural_dcvc(r["polar"], r["ext"], **mk)
            assert np.array_equal(key, r["voxel"]) and np.array_equal(root, r["root"]), kw


def test_oracle_segmentation_outputs(oracle, obj):
    r = oracle.dcvc(obj)
    n = len(obj)
    # polar triples: independent numpy
    rad = np.linalg.norm(obj, axis=1)
    ok = (rad < 120.0) & (rad > 1.0)
    assert np.allclose(r["polar"][ok, 0], rad[ok], rtol=1e-15)
    assert np.allclose(r["polar"][ok, 1], np.degrees(np.arcsin(obj[ok, 2] / rad[ok])), rtol=1e-13)
    assert np.allclose(r["polar"][ok, 2], np.degrees(np.arctan2(obj[ok, 1], obj[ok, 0])) % 360.0, rtol=1e-13)
    assert r["ext"][2] == 0.0 and r["ext"][3] == rad[ok].max()           # minPolar starts from the member value 0
    # clusters: sizes descending, > min_seg, ties by smallest index; segmented = members in index order
    sizes = r["sizes"]
    assert len(sizes) > 3 and (sizes > 80).all() and (np.diff(sizes) <= 0).all()
    off = 0
    for c, sz in enumerate(sizes):
        mem = r["segmented"][off:off + sz]
        assert (np.diff(mem.astype(np.int64)) > 0).all()
        assert (r["cluster"][mem] == c + 1).all() and len(set(r["root"][mem])) == 1
        lo, hi = obj[mem].min(0), obj[mem].max(0)
        assert np.array_equal(r["boxes"][c, :3], lo + (hi - lo) / 2.0) and np.array_equal(r["boxes"][c, 3:], hi - lo)
        off += sz
    assert off == len(r["segmented"]) == (r["cluster"] > 0).sum()
    # filtered classes have <= min_seg points
    roots, counts = np.unique(r["root"][r["cluster"] == 0], return_counts=True)
    assert (counts <= 80).all()
    # the big structures of the scene (two walls) come out as big clusters
    assert sizes[0] > 0.1 * n
    # the device limit on the polar table is mirrored
    assert oracle.dcvc(obj, max_bounds=100) is None
    r0 = oracle.dcvc(np.zeros((0, 3)))
    assert len(r0["segmented"]) == 0 and len(r0["sizes"]) == 0


def ulp_diff(a, b):
    ia, ib = a.view(np.int64), b.view(np.int64)
    return np.abs(ia - ib)


def check_gpu_against_oracle(oracle, reg, p, **kw):
    g = reg.object_segmentation(p, **kw)
    o = oracle.dcvc(p, **kw)
    if len(p):
        # trigonometry: libdevice's asin / atan2 are within 2 ulp of libm's; the conversion to degrees (x 180, / pi) can
        # stretch that across a binade boundary: tolerance 4 ulp on the triples (range is sqrt of the same sums: exact)
        assert np.array_equal(g["polar"][:, 0], o["polar"][:, 0])
        assert ulp_diff(g["polar"][:, 1:], o["polar"][:, 1:]).max() <= 4
        assert ulp_diff(g["ext"], o["ext"]).max() <= 4
    # everything downstream of the triples: bit-exact against the literal oracle run on the DEVICE's triples ...
    o2 = oracle.dcvc_from_polar(p, g["polar"], g["ext"], **kw)
    for k in ("voxel", "root", "cluster", "segmented", "sizes", "boxes"):
        assert np.array_equal(g[k], o2[k]), k
    # ... and, unless a last-bit difference of asin / atan2 lands exactly on a rounding boundary of a voxel index (it does not
    # on these inputs), against the oracle's own triples as well
    for k in ("voxel", "root", "cluster", "segmented", "sizes", "boxes"):
        assert np.array_equal(g[k], o[k]), k
    return g


@pytest.mark.gpu
def test_gpu_object_segmentation_matches_the_oracle(oracle, obj):
    import tloam_b200
    reg = tloam_b200.LocalRegistration()
    g = check_gpu_against_oracle(oracle, reg, obj)
    assert len(g["sizes"]) > 3
    check_gpu_against_oracle(oracle, reg, obj[::-1].copy())
    check_gpu_against_oracle(oracle, reg, obj[::3].copy(), min_seg=5)
    check_gpu_against_oracle(oracle, reg, obj, min_seg=0)                              # thousands of clusters: 2 partition passes
    check_gpu_against_oracle(oracle, reg, obj, min_polar_init=5.0, max_polar_init=5.0)   # first frame
    check_gpu_against_oracle(oracle, reg, object_scan(oracle, seed=9, n_az=900))
    check_gpu_against_oracle(oracle, reg, synth.raw_scan())                             # 116k points, ground included
    check_gpu_against_oracle(oracle, reg, synth.raw_scan(n_az=4200))                    # 260k points: 2-bit voxel states (no byte per voxel)
    check_gpu_against_oracle(oracle, reg, np.zeros((0, 3)))
    check_gpu_against_oracle(oracle, reg, np.array([[10.0, 0.0, 0.0]]), min_seg=0)
    reg.close()


@pytest.mark.gpu
def test_gpu_object_segmentation_fuzz_and_configs(oracle):
    import tloam_b200
    reg = tloam_b200.LocalRegistration()
    for k, p in enumerate(random_clouds(60, seed=11)):
        check_gpu_against_oracle(oracle, reg, p, min_seg=k % 7)
    for p in random_clouds(12, seed=5):
        check_gpu_against_oracle(oracle, reg, p, delta_a=0.6, min_seg=2)
        check_gpu_against_oracle(oracle, reg, p, delta_p=0.4, start_r=0.5, min_seg=2)
    # more polar rings than the device table holds: rejected on both sides
    p = next(random_clouds(1, seed=3))
    assert oracle.dcvc(p, start_r=0.001, delta_r=0.0, max_bounds=4096) is None
    with pytest.raises(Exception):
        reg.object_segmentation(p, start_r=0.001, delta_r=0.0)
    reg.close()


@pytest.mark.gpu
def test_gpu_segmentation_front_half_chain(oracle):
    """groundRemove -> objectSegmentation -> extractEdgePoint on the device, every stage fed by the previous DEVICE stage,
    against the same chain on the oracle (ref: segmentation.cpp:47-66)."""
    import tloam_b200
    reg = tloam_b200.LocalRegistration()
    scan = synth.raw_scan()
    out = {}
    for name, impl in (("gpu", reg), ("oracle", None)):
        ge = reg.ground_extract(scan) if impl else oracle.ground_extract(scan)
        object_pts = np.ascontiguousarray(scan[ge["object"]])
        object_beam = ge["beam"][ge["object"]].astype(np.float64)
        os_ = reg.object_segmentation(object_pts) if impl else oracle.dcvc(object_pts)
        seg_pts = np.ascontiguousarray(object_pts[os_["segmented"]])
        seg_beam = object_beam[os_["segmented"]]
        ee = reg.extract_edge(seg_pts, seg_beam, ring_min_num=131) if impl else oracle.extract_edge(seg_pts, seg_beam, ring_min_num=131)
        out[name] = (ge["ground"], ge["object"], os_["segmented"], os_["sizes"], ee["edge"], ee["non_edge"])
    for a, b in zip(out["gpu"], out["oracle"]):
        assert np.array_equal(a, b)
    assert len(out["gpu"][4]) > 100 and len(out["gpu"][5]) > 5000
    reg.close()


@pytest.mark.gpu
def test_gpu_segment_scan_single_call_matches_the_chain(oracle):
    """tloam_b200_segment_scan (one upload, stages fed on the device, indices into the ORIGINAL scan) against the three
    separate calls with host gathers in between, on the oracle."""
    import tloam_b200
    reg = tloam_b200.LocalRegistration()
    for scan, rmin, dc in ((synth.raw_scan(), 131, {}), (synth.raw_scan(seed=9, n_az=900), 60, dict(min_seg=30)),
                           (synth.raw_scan(seed=4, n_az=300)[:4000], 16, dict(min_seg=5))):
        got = reg.segment_scan(scan, ring_min_num=rmin, dcvc=dc)
        ge = oracle.ground_extract(scan)
        obj = np.ascontiguousarray(scan[ge["object"]])
        beam = ge["beam"][ge["object"]].astype(np.float64)
        os_ = oracle.dcvc(obj, **dc)
        seg_orig = ge["object"][os_["segmented"]]
        ee = oracle.extract_edge(np.ascontiguousarray(obj[os_["segmented"]]), beam[os_["segmented"]], ring_min_num=rmin)
        assert np.array_equal(got["ground"], ge["ground"])
        assert np.array_equal(got["edge"], seg_orig[ee["edge"]])
        assert np.array_equal(got["general"], seg_orig[ee["non_edge"]])
        assert np.array_equal(got["sizes"], os_["sizes"]) and np.array_equal(got["boxes"], os_["boxes"])
        assert np.array_equal(got["beam"], ge["beam"])
    e = reg.segment_scan(np.zeros((0, 3)))
    assert len(e["ground"]) == 0 and len(e["edge"]) == 0 and len(e["general"]) == 0
    # the stage entry points are unaffected by a chained call before them
    scan = synth.raw_scan(seed=4, n_az=300)
    a, b = reg.ground_extract(scan), oracle.ground_extract(scan)
    assert np.array_equal(a["ground"], b["ground"]) and np.array_equal(a["object"], b["object"])
    reg.close()
