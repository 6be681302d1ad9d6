"""Phase breakdown of the dense-map correspondence kernel on BASELINE config 3 (in-kernel SM cycle counters)."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ["TLOAM_B200_DENSE"] = "1"
os.environ["TLOAM_B200_DENSE_CHECK"] = "1"
import tloam_b200  # noqa: E402
from tloam_b200 import synth  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
f = synth.config3(int(500_000 * scale), int(2_000_000 * scale))
BIG = 10 ** 9
reg = tloam_b200.LocalRegistration(edge_maxnum=BIG, sphere_maxnum=BIG, planar_maxnum=BIG, ground_maxnum=BIG, factor_num=2)
reg.set_input_target(f["map"])
reg.set_input_source(f["scan"])
T, st = reg.scan_matching(f["predict"], want_stats=True)
c = reg.dense_check_counters()
items = max(c[2], 1)
print(json.dumps({"queries": c[0], "mismatches": c[1], "items": c[2], "passes": c[3], "items_le_8_queries": c[12],
                  "queries_per_item": c[0] / items, "kcycles_per_item": {"total": c[8] * 64 / items / 1e3, "tma_wait": c[9] * 64 / items / 1e3,
                                                                           "fine_sort": c[10] * 64 / items / 1e3, "search": c[11] * 64 / items / 1e3,
                                                                           "phase1_27cells": c[14] * 64 / items / 1e3, "phase2_hard": c[15] * 64 / items / 1e3},
                  "hard_queries": c[13],
                  "gpu_ms": st.gpu_ms}))
