"""Where does the end-to-end frame (pinned host buffers through the ABI) spend its time?  Host-side wall clock per call.
    python tools/e2e_probe.py [frames]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
import bench  # noqa: E402
import tloam_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 23
frames, prev_gt = bench.gen_frames("00", n)
reg = tloam_b200.LocalRegistration(stream=torch.cuda.current_stream().cuda_stream, **bench.CAPS)
pageable = len(sys.argv) > 2 and sys.argv[2] == "pageable"
pin = (lambda a: np.array(a, copy=True)) if pageable else (lambda a: torch.from_numpy(a).pin_memory().numpy())
data = [([pin(c) for c in fr["map"]], [pin(c) for c in fr["scan"]]) for fr in frames]
for mp, sc in data:
    for a in mp + sc:
        torch.from_numpy(a).cuda(non_blocking=True)
torch.cuda.synchronize()
last, cur = prev_gt.copy(), None
rows = []
for k, fr in enumerate(frames):
    predict = bench.first_predict(fr) if cur is None else bench.predict_next(last, cur)
    mp, sc = data[k]
    t0 = time.perf_counter()
    reg.set_input_target(mp)
    t1 = time.perf_counter()
    reg.set_input_source(sc)
    t2 = time.perf_counter()
    T = reg.scan_matching(predict)
    t3 = time.perf_counter()
    rows.append((t1 - t0, t2 - t1, t3 - t2, t3 - t0))
    last, cur = (cur if cur is not None else prev_gt), T
r = np.array(rows[3:]) * 1e3
print("pageable" if pageable else "pinned", end=" ")
print("median ms: set_target %.3f set_source %.3f scan_match %.3f total %.3f" % tuple(np.median(r, axis=0)))
print("map points", [len(c) for c in data[0][0]], "scan points", [len(c) for c in data[0][1]])
reg.close()
