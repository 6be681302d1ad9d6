"""(f)-4, second part: LOAM-style edge extraction (Segmentation::extractEdgePoint + extractFromSection, ref:
src/models/segmentation/segmentation.cpp:1144-1304) -- the CPU restatement against a literal pure-Python transcription of
the reference's control flow (picked_points as a LIST searched with `in`, like std::find), and the device path bit-exact
against the restatement (both index lists, in the reference's append order)."""
import numpy as np
import pytest

from tloam_b200 import synth


def object_cloud(oracle, scan):
    """What the reference hands to extractEdgePoint: non-ground points with the beam id in the intensity channel."""
    g = oracle.ground_extract(scan)
    idx = np.sort(g["object"])                      # beam-major order, like a segmented scan assembled cluster by cluster
    return np.ascontiguousarray(scan[idx]), g["beam"][idx].astype(np.float64)


@pytest.fixture(scope="module")
def cloud(oracle):
    return object_cloud(oracle, synth.raw_scan())


def python_edges(pts, intensity, sensor_model=64, ring_min_num=16):
    """Literal transcription (Python floats are IEEE doubles, no contraction: bit-exact against the C++ oracle)."""
    rings = [[] for _ in range(sensor_model)]
    for i, b in enumerate(intensity):
        b = int(b)
        if 0 <= b < sensor_model:
            rings[b].append(i)
    edge, non_edge = [], []
    for ring in rings:
        total = len(ring)
        if total < ring_min_num or total - 10 <= 0:
            continue
        P = [tuple(float(v) for v in pts[i]) for i in ring]
        curv = []
        for j in range(5, total - 5):
            d = []
            for a in range(3):
                s = P[j - 5][a] + P[j - 4][a] + P[j - 3][a] + P[j - 2][a] + P[j - 1][a] - 10 * P[j][a]
                s = s + P[j + 1][a] + P[j + 2][a] + P[j + 3][a] + P[j + 4][a] + P[j + 5][a]
                d.append(s)
            curv.append((j, d[0] * d[0] + d[1] * d[1] + d[2] * d[2]))
        total_points = total - 10
        for sct in range(6):
            sl = total_points // 6
            start = sl * sct
            end = sl * (sct + 1) - 1 if sct != 5 else total_points - 1
            sub = sorted(curv[start:end], key=lambda t: (t[1], t[0]))
            picked, largest = [], 0
            for i in range(len(sub) - 1, -1, -1):
                pid = sub[i][0]
                if pid in picked:
                    continue
                if sub[i][1] <= 0.1:
                    break
                largest += 1
                picked.append(pid)
                if largest <= 20:
                    edge.append(ring[pid])
                else:
                    break
                for k in range(1, 6):
                    dx, dy, dz = (P[pid + k][a] - P[pid + k - 1][a] for a in range(3))
                    if dx * dx + dy * dy + dz * dz > 0.05:
                        break
                    picked.append(pid + k)
                for k in range(-1, -6, -1):
                    dx, dy, dz = (P[pid + k][a] - P[pid + k + 1][a] for a in range(3))
                    if dx * dx + dy * dy + dz * dz > 0.05:
                        break
                    picked.append(pid + k)
            for pid, _ in sub:
                if pid not in picked:
                    non_edge.append(ring[pid])
    return np.array(edge, dtype=np.uintp), np.array(non_edge, dtype=np.uintp)


def test_oracle_edges_against_literal_python(oracle, cloud):
    pts, beam = cloud
    sel = np.flatnonzero((beam >= 30) & (beam < 38))           # 8 beams: the pure-Python loops stay within seconds
    p, b = pts[sel], beam[sel]
    r = oracle.extract_edge(p, b)
    e, ne = python_edges(p, b)
    assert len(e) > 50 and len(ne) > 1000
    assert np.array_equal(r["edge"], e)
    assert np.array_equal(r["non_edge"], ne)


def test_oracle_edges_properties(oracle, cloud):
    pts, beam = cloud
    r = oracle.extract_edge(pts, beam)
    e, ne = r["edge"], r["non_edge"]
    assert len(np.intersect1d(e, ne)) == 0 and len(np.unique(e)) == len(e) and len(np.unique(ne)) == len(ne)
    # <= 20 edge points per (beam, sector); 6 sectors per beam with >= ring_min_num points
    nb = len([b for b in range(64) if (beam == b).sum() >= 16])
    assert len(e) <= 20 * 6 * nb
    # the first and last 5 points of every ring and the last curvature of every sector are in neither list
    for b in range(64):
        ring = np.flatnonzero(beam == b)
        if len(ring) >= 16:
            both = np.concatenate([e, ne])
            assert not (set(ring[:5]) | set(ring[-5:])) & set(both)
    # rings with fewer than ring_min_num points contribute nothing
    few = np.flatnonzero(beam == 3)[:15]
    r2 = oracle.extract_edge(pts[few], beam[few])
    assert len(r2["edge"]) == 0 and len(r2["non_edge"]) == 0


def test_oracle_edges_edge_cases(oracle):
    r = oracle.extract_edge(np.zeros((0, 3)), np.zeros(0))
    assert len(r["edge"]) == 0 and len(r["non_edge"]) == 0
    # beam ids outside [0, sensor_model) are dropped (the reference's `<=` would write past ringScans: not reproduced)
    rng = np.random.default_rng(4)
    p = rng.normal(0, 5, (400, 3))
    b = np.repeat([64.0, -1.0, 7.9, 200.0], 100)             # (int)7.9 = 7
    r = oracle.extract_edge(p, b)
    assert set(np.concatenate([r["edge"], r["non_edge"]])) <= set(range(200, 300))
    # exact curvature ties (a lattice): ties are ordered by ring position on both sides
    q = np.zeros((120, 3))
    q[:, 0] = np.arange(120) % 3
    r = oracle.extract_edge(q, np.zeros(120))
    e, ne = python_edges(q, np.zeros(120))
    assert np.array_equal(r["edge"], e) and np.array_equal(r["non_edge"], ne)
    # the device limit is mirrored
    assert oracle.extract_edge(np.zeros((30000, 3)), np.zeros(30000), max_section=4096) is None


@pytest.mark.gpu
def test_gpu_edges_are_bit_exact(oracle, cloud):
    import tloam_b200
    reg = tloam_b200.LocalRegistration()
    pts, beam = cloud
    rng = np.random.default_rng(2)
    shuffled = rng.permutation(len(pts))                        # beams interleaved: the stable partition by beam is exercised
    lattice = np.zeros((5000, 3))
    lattice[:, 0] = np.arange(5000) % 7
    lattice[:, 1] = (np.arange(5000) // 7) % 3
    cases = [(pts, beam, 64, 16), (pts[::2], beam[::2], 64, 16), (pts[shuffled], beam[shuffled], 64, 16), (pts, beam, 32, 16),
             (pts, beam, 64, 2000), (lattice, np.arange(5000) % 5, 64, 16), (np.zeros((0, 3)), np.zeros(0), 64, 16),
             (pts[:40], np.zeros(40), 64, 16), (rng.normal(0, 30, (24000, 3)), np.zeros(24000), 1, 16)]
    for p, b, sm, rmin in cases:
        g = reg.extract_edge(p, b, sensor_model=sm, ring_min_num=rmin)
        o = oracle.extract_edge(p, b, sensor_model=sm, ring_min_num=rmin, max_section=0)
        assert np.array_equal(g["edge"], o["edge"]), (len(p), sm, rmin)
        assert np.array_equal(g["non_edge"], o["non_edge"]), (len(p), sm, rmin)
    # a sector beyond the shared-memory sort capacity is rejected, not truncated
    with pytest.raises(Exception):
        reg.extract_edge(np.zeros((30000, 3)), np.zeros(30000))
    reg.close()
