"""Experiment: does spatially sorting the scan features (so that the lanes of a warp query neighbouring cells)
speed up k_correspond?  Compares per-kernel times with the features in generator order vs sorted by a coarse
sensor-frame cell key."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench, tloam_b200

frames, prev_gt = bench.gen_frames("00", 4)


def sorted_scan(scan, cell):
    out = []
    for c in scan:
        k = np.floor(c / cell).astype(np.int64)
        key = (k[:, 0] + 4096) * (1 << 26) + (k[:, 1] + 4096) * (1 << 13) + (k[:, 2] + 4096)
        out.append(np.ascontiguousarray(c[np.argsort(key, kind="stable")]))
    return out


for label, cell in (("generator order", None), ("sorted, 1.0 m cells", 1.0), ("sorted, 0.5 m cells", 0.5), ("sorted, 4 m cells", 4.0)):
    reg = tloam_b200.LocalRegistration(**bench.CAPS)
    last, cur = prev_gt.copy(), None
    poses = []
    for k, fr in enumerate(frames):
        if k == 1:
            reg.set_profiling(True)
        predict = bench.first_predict(fr) if cur is None else bench.predict_next(last, cur)
        reg.set_input_target(fr["map"])
        reg.set_input_source(fr["scan"] if cell is None else sorted_scan(fr["scan"], cell))
        T = reg.scan_matching(predict)
        poses.append(T)
        last, cur = (cur if cur is not None else prev_gt), T
    prof = reg.get_profile()
    print(label, {k: round(1e3 * v[1] / max(v[0], 1), 1) for k, v in prof.items() if k in ("correspond", "eval", "eval_first")},
          "pose[-1] t:", np.round(poses[-1][:3, 3], 6))
    reg.close()
