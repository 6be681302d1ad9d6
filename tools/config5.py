"""BASELINE config 5: IRLS / GNC threshold ablation -- noise_bound in {0.002, 0.005, 0.01, 0.02, 0.05} on the
config-2 stream; reports pose error vs ground truth (drift proxy: the map is ground truth, so errors do not
accumulate) and throughput.

    python tools/config5.py [frames]
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
import tloam_b200  # noqa: E402


def main():
    nframes = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    frames, prev_gt = bench.gen_frames("00", nframes)
    rows = []
    for nb in (0.002, 0.005, 0.01, 0.02, 0.05):
        reg = tloam_b200.LocalRegistration(noise_bound=nb, **bench.CAPS)
        last, cur = prev_gt.copy(), None
        errs, ms, inner = [], [], []
        for k, fr in enumerate(frames):
            predict = bench.first_predict(fr) if cur is None else bench.predict_next(last, cur)
            reg.set_input_target(fr["map"])
            reg.set_input_source(fr["scan"])
            t0 = time.perf_counter()
            T, st = reg.scan_matching(predict, want_stats=True)
            ms.append((time.perf_counter() - t0) * 1e3)
            errs.append(bench.pose_err(T, fr["T_gt"]))
            inner.append(sum(st.outer[i].n_inner for i in range(st.n_outer)))
            last, cur = (cur if cur is not None else prev_gt), T
        e = np.array(errs[2:])
        rows.append({"noise_bound": nb, "mean_err_m": float(e[:, 0].mean()), "max_err_m": float(e[:, 0].max()),
                     "mean_rot_err_rad": float(e[:, 1].mean()), "scan_match_ms": float(np.median(ms[2:])),
                     "frames_per_s_scan_match_only": float(1e3 / np.median(ms[2:])), "tr_iterations_per_frame": float(np.mean(inner[2:]))})
        reg.close()
    print(json.dumps({"config": 5, "frames": nframes, "rows": rows}))


if __name__ == "__main__":
    main()
