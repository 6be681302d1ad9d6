"""The C++ shim (include/tloam_b200/local_registration_b200.hpp) compiles against stand-in host types and, on
a GPU, produces the same pose as the Python mirror when driven through the RegistrationInterface base class."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "mock", "_build", "shim_driver")


def build_driver(name="shim_driver", header="local_registration_b200.hpp"):
    from tloam_b200 import build
    lib = build.build()
    exe = os.path.join(os.path.dirname(EXE), name)
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    src = os.path.join(ROOT, "tests", "mock", name + ".cpp")
    hdr = os.path.join(ROOT, "include", "tloam_b200", header)
    mock = os.path.join(ROOT, "tests", "mock", "mock_tloam.hpp")
    if os.path.exists(exe) and os.path.getmtime(exe) > max(os.path.getmtime(p) for p in (src, hdr, mock, lib)):
        return exe
    cmd = ["/usr/bin/g++", "-std=c++14", "-O2", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "tests", "mock"), src,
           "-o", exe, lib, "-Wl,-rpath," + os.path.dirname(lib), "-ldl", "-lpthread", "-lrt"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def test_shim_compiles_as_cpp14_against_host_types():
    """C++14 like the reference (CMakeLists.txt:4); -Wall -Wextra clean."""
    assert os.path.exists(build_driver())
    assert os.path.exists(build_driver("feature_driver", "feature_extract_b200.hpp"))
    assert os.path.exists(build_driver("ground_driver", "ground_extract_b200.hpp"))
    assert os.path.exists(build_driver("segmentation_driver", "segmentation_b200.hpp"))


def test_shim_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    exe = build_driver()
    path = os.path.join(os.path.dirname(EXE), "empty.bin")
    with open(path, "wb") as f:
        f.write(struct.pack("8Q", *([0] * 8)))
        f.write(np.eye(4).T.tobytes())
    res = subprocess.run([exe, path], capture_output=True, text=True)
    assert res.returncode == 3 and "no CUDA device" in res.stderr


@pytest.mark.gpu
def test_shim_matches_python_mirror():
    import tloam_b200
    from tloam_b200 import synth
    exe = build_driver()
    cfg = synth.scaled(0.03, seed=31)
    T_gt = synth.se3_exp([5.0, -1.0, 0.0, 0.0, 0.01, 0.7])
    predict = T_gt @ synth.se3_exp(synth.CONFIG1_PERTURB)
    mp, scan = synth.make_map(cfg, T_gt), synth.make_scan(cfg, T_gt, 0)
    path = os.path.join(os.path.dirname(EXE), "frame.bin")
    with open(path, "wb") as f:
        f.write(struct.pack("8Q", *[c.shape[0] for c in mp + scan]))
        for c in mp + scan:
            f.write(np.ascontiguousarray(c, dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(predict.T, dtype=np.float64).tobytes())     # column-major
    res = subprocess.run([exe, path], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    vals = np.array([float(x) for x in res.stdout.split()])
    T_cpp = vals[:16].reshape(4, 4).T
    reg = tloam_b200.LocalRegistration()
    reg.set_input_target(mp)
    reg.set_input_source(scan)
    T_py = reg.scan_matching(predict)
    assert np.array_equal(T_cpp, T_py)
    f_py = reg.get_fitness_score()
    assert np.allclose(vals[16:18], f_py, rtol=1e-12, atol=0)
    reg.close()


@pytest.mark.gpu
def test_feature_shim_matches_oracle(oracle):
    """featureExtractB200::extractPlanarSphere appends exactly the lists of the CPU restatement."""
    from tloam_b200 import synth
    exe = build_driver("feature_driver", "feature_extract_b200.hpp")
    pts = synth.general_cloud(20000, seed=9)
    path = os.path.join(os.path.dirname(EXE), "cloud.bin")
    with open(path, "wb") as f:
        f.write(struct.pack("Q", pts.shape[0]))
        f.write(np.ascontiguousarray(pts, dtype=np.float64).tobytes())
    res = subprocess.run([exe, path], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    vals = [int(x) for x in res.stdout.split()]
    ref = oracle.extract_planar_sphere(pts)
    k = 0
    for r in ref[:4]:
        n = vals[k]; lst = vals[k + 1:k + 1 + n]; k += 1 + n
        assert lst[0] == 987654321 and np.array_equal(np.asarray(lst[1:], dtype=np.uintp), r)
    assert k == len(vals)


@pytest.mark.gpu
def test_ground_shim_matches_oracle(oracle):
    """tloam::GroundExtractB200::groundRemove driven like Segmentation::spinOnce: ground_scan / object_scan receive (+=)
    exactly the points the CPU restatement selects, in its order, with the intensities of the reference."""
    from tloam_b200 import synth
    exe = build_driver("ground_driver", "ground_extract_b200.hpp")
    pts = synth.raw_scan(seed=4, n_az=600)
    path = os.path.join(os.path.dirname(EXE), "scan.bin")
    with open(path, "wb") as f:
        f.write(struct.pack("Q", pts.shape[0]))
        f.write(np.ascontiguousarray(pts, dtype=np.float64).tobytes())
    res = subprocess.run([exe, path], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    lines = res.stdout.strip().split("\n")
    ng, no, ncur, thr = lines[0].split()
    ref = oracle.ground_extract(pts)
    assert int(ng) == 1 + len(ref["ground"]) and int(no) == 1 + len(ref["object"]) and float(thr) == ref["height_threshold"]
    assert int(ncur) == int(np.sum(ref["region"] != 12))
    rows = [l.split() for l in lines[1:]]
    xbits = pts[:, 0].copy().view(np.uint64)
    g = rows[:int(ng)]
    o = rows[int(ng):]
    assert float(g[0][1]) == -1.0 and float(o[0][1]) == -1.0                     # the sentinels stay in front (+=)
    assert [int(r[0]) for r in g[1:]] == [int(xbits[i]) for i in ref["ground"]]
    assert all(float(r[1]) == 0.0 for r in g[1:])
    assert [int(r[0]) for r in o[1:]] == [int(xbits[i]) for i in ref["object"]]
    assert [float(r[1]) for r in o[1:]] == [float(ref["beam"][i]) for i in ref["object"]]


@pytest.mark.gpu
def test_segmentation_shim_chain_matches_oracle(oracle):
    """tloam::SegmentationB200 driven like Segmentation::spinOnce for two frames (groundRemove -> objectSegmentation ->
    extractEdgePoint): clouds, boxes and the final edge / general clouds equal the oracle chain's; frame 0 runs with the
    members' initial polar extrema (5.0), frame 1 with what resetParams() leaves (0.0)."""
    from tloam_b200 import synth
    exe = build_driver("segmentation_driver", "segmentation_b200.hpp")
    pts = synth.raw_scan(seed=4, n_az=1200)
    path = os.path.join(os.path.dirname(EXE), "scan_seg.bin")
    with open(path, "wb") as f:
        f.write(struct.pack("Q", pts.shape[0]))
        f.write(np.ascontiguousarray(pts, dtype=np.float64).tobytes())
    res = subprocess.run([exe, path], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    lines = res.stdout.strip().split("\n")
    ge = oracle.ground_extract(pts)
    obj = np.ascontiguousarray(pts[ge["object"]])
    beam = ge["beam"][ge["object"]].astype(np.float64)
    pos = 0
    for frame, init in enumerate((5.0, 0.0)):
        seg = oracle.dcvc(obj, min_polar_init=init, max_polar_init=init)
        sp, sb = np.ascontiguousarray(obj[seg["segmented"]]), beam[seg["segmented"]]
        ee = oracle.extract_edge(sp, sb, ring_min_num=131)
        ng, no, ns, nb, ne, nn = (int(v) for v in lines[pos].split())
        pos += 1
        assert (ng, no, ns, nb, ne, nn) == (len(ge["ground"]), len(ge["object"]), len(seg["segmented"]), len(seg["sizes"]),
                                            len(ee["edge"]), len(ee["non_edge"]))
        for c in range(nb):
            v = lines[pos + c].split()
            assert int(v[0]) == c + 1 and int(v[1]) == seg["sizes"][c]
            assert [float(x) for x in v[2:]] == list(seg["boxes"][c])
        pos += nb
        rows = [l.split() for l in lines[pos:pos + ne + nn]]
        pos += ne + nn
        xbits = sp[:, 0].copy().view(np.uint64)
        want = list(ee["edge"]) + list(ee["non_edge"])
        assert [int(r[0]) for r in rows] == [int(xbits[i]) for i in want]
        assert [float(r[1]) for r in rows] == [float(sb[i]) for i in want]
        last_rows, last_counts = rows, (ng, nb, ne, nn)
    # third block: SegmentationB200::segmentScan (one device pass) = the clouds of frame 1
    ng, nb, ne, nn = (int(v) for v in lines[pos].split())
    assert (ng, nb, ne, nn) == last_counts
    rows = [l.split() for l in lines[pos + 1:pos + 1 + ne + nn]]
    assert rows == last_rows
    pos += 1 + ne + nn
    assert pos == len(lines)
