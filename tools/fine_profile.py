"""Two-level grid on BASELINE config 3 at full size: exactness self-check (every query re-searched by knn_search on the
device) and per-kernel times of the three search paths.

    python tools/fine_profile.py [scale]
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import tloam_b200  # noqa: E402
from tloam_b200 import synth  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
f = synth.config3(int(500_000 * scale), int(2_000_000 * scale))
BIG = 10 ** 9
CFG = dict(edge_maxnum=BIG, sphere_maxnum=BIG, planar_maxnum=BIG, ground_maxnum=BIG, factor_num=2)


def run(env, check=False):
    for k in ("TLOAM_B200_FINE", "TLOAM_B200_DENSE", "TLOAM_B200_DENSE_CHECK"):
        os.environ.pop(k, None)
    os.environ.update(env)
    if check:
        os.environ["TLOAM_B200_DENSE_CHECK"] = "1"
    reg = tloam_b200.LocalRegistration(**CFG)
    out = {}
    for _ in range(3):
        reg.set_input_target(f["map"])
        reg.set_input_source(f["scan"])
        T, st = reg.scan_matching(f["predict"], want_stats=True)
    out["gpu_ms_scan_match"] = st.gpu_ms
    if check:
        c = reg.dense_check_counters()
        out["check"] = {"queries": int(c[0]), "mismatches": int(c[1])}
    else:
        reg.set_profiling(True)
        for _ in range(3):
            reg.set_input_target(f["map"])
            reg.scan_matching(f["predict"])
        out["kernels_us"] = {k: round(1e3 * ms / n, 2) for k, (n, ms) in reg.get_profile().items() if n > 0}
    reg.close()
    return T, out


res = {}
T_check, res["two_level_self_check"] = run({"TLOAM_B200_FINE": "1"}, check=True)
T_fine, res["two_level"] = run({})
T_pair, res["lane_pair"] = run({"TLOAM_B200_FINE": "0"})
res["bit_identical"] = bool(np.array_equal(T_fine, T_pair) and np.array_equal(T_check, T_pair))
res["err_vs_gt_m"] = float(np.linalg.norm(T_fine[:3, 3] - f["T_gt"][:3, 3]))
print(json.dumps(res))
