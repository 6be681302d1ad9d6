"""The "next" rows chained like the reference's odometry (segmentation nodelet -> FrontEnd::processCloud ->
scanMatching; ref: src/models/segmentation/segmentation.cpp:39-92, src/front_end/front_end.cpp:181-199, 278-337): a raw
scan goes through groundRemove / objectSegmentation / extractEdgePoint ((f)-4), the general cloud through the PCA
feature extraction ((f)-2), and the resulting four feature clouds of frame 1 are registered against those of frame 0
(the hot path) -- every stage on the device, fed by the previous device stage, against the same chain on the oracle:
identical feature clouds, pose within 1e-4 m / 1e-5 rad."""
import numpy as np
import pytest

from tloam_b200 import synth

FE = dict(cvr_submap=0.005, cvr_scan=0.01)       # the synthetic street scene has few curvature maxima: lower the sphere thresholds


def frame_features(scan, segment, planar_sphere):
    """segment(scan) -> dict(ground, edge, general) index lists; planar_sphere(cloud) -> the 5 lists of (f)-2."""
    seg = segment(scan)
    edge = np.ascontiguousarray(scan[seg["edge"]])
    general = np.ascontiguousarray(scan[seg["general"]])
    ground = np.ascontiguousarray(scan[seg["ground"]][::8])              # (the front end voxel-down-samples the ground scan)
    p_scan, p_sub, s_scan, s_sub, s_cand = planar_sphere(general)
    return dict(edge=edge, ground=ground, p_scan=general[p_scan], p_sub=general[p_sub], s_scan=general[s_cand[s_scan]],
                s_sub=general[s_cand[s_sub]])


def oracle_segment(oracle, scan):
    ge = oracle.ground_extract(scan)
    obj = np.ascontiguousarray(scan[ge["object"]])
    beam = ge["beam"][ge["object"]].astype(np.float64)
    seg = oracle.dcvc(obj)
    orig = ge["object"][seg["segmented"]]
    ee = oracle.extract_edge(np.ascontiguousarray(obj[seg["segmented"]]), beam[seg["segmented"]], ring_min_num=131)
    return dict(ground=ge["ground"], edge=orig[ee["edge"]], general=orig[ee["non_edge"]])


@pytest.mark.gpu
def test_raw_scan_to_pose_on_the_device_matches_the_oracle_chain(oracle):
    import tloam_b200
    scan0 = synth.raw_scan()
    T = synth.se3_exp([0.4, 0.05, 0.0, 0.0, 0.0, 0.01])                  # frame 1 seen from a sensor moved by T
    Ti = np.linalg.inv(T)
    scan1 = np.ascontiguousarray((scan0 @ Ti[:3, :3].T + Ti[:3, 3]) + np.random.default_rng(3).normal(0, 0.005, scan0.shape))
    predict = T @ synth.se3_exp(synth.CONFIG1_PERTURB)
    reg = tloam_b200.LocalRegistration()
    g = [frame_features(s, lambda x: reg.segment_scan(x), lambda c: reg.extract_planar_sphere(c, **FE)) for s in (scan0, scan1)]
    o = [frame_features(s, lambda x: oracle_segment(oracle, x), lambda c: oracle.extract_planar_sphere(c, **FE)) for s in (scan0, scan1)]
    for a, b in zip(g, o):
        for k in a:
            assert np.array_equal(a[k], b[k]), k
        assert len(a["edge"]) > 500 and len(a["p_scan"]) > 500 and len(a["s_scan"]) > 50 and len(a["ground"]) > 1000
    # frame 0 is the local map, frame 1 the scan (cloud order: edge, sphere, planar, ground)
    reg.set_input_target([g[0]["edge"], g[0]["s_sub"], g[0]["p_sub"], g[0]["ground"]])
    reg.set_input_source([g[1]["edge"], g[1]["s_scan"], g[1]["p_scan"], g[1]["ground"]])
    Tg = reg.scan_matching(predict)
    orc = oracle.Oracle(threads_mode=1)
    orc.set_input_target([o[0]["edge"], o[0]["s_sub"], o[0]["p_sub"], o[0]["ground"]])
    orc.set_input_source([o[1]["edge"], o[1]["s_scan"], o[1]["p_scan"], o[1]["ground"]])
    rc, To, _ = orc.scan_matching(predict)
    assert rc == 0
    d = np.linalg.inv(To) @ Tg
    dt, dr = np.linalg.norm(d[:3, 3]), np.arccos(np.clip((np.trace(d[:3, :3]) - 1) / 2, -1, 1))
    assert dt < 1e-4 and dr < 1e-5, (dt, dr)
    # against the motion that generated frame 1: the street canyon constrains the along-street translation weakly (5 cm of
    # prediction error shrink to ~3 cm), everything else is recovered to millimetres / 1e-3 rad
    e = np.linalg.inv(T) @ Tg
    assert abs(e[1, 3]) < 5e-3 and abs(e[2, 3]) < 1e-2 and np.linalg.norm(e[:3, 3]) < 5e-2
    assert np.arccos(np.clip((np.trace(e[:3, :3]) - 1) / 2, -1, 1)) < 2e-3
    reg.close()
