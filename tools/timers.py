"""In-kernel timer breakdown (profiling mode) for a few frames of the bench workload."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench, tloam_b200

frames, prev_gt = bench.gen_frames("00", 5)
reg = tloam_b200.LocalRegistration(**bench.CAPS)
last, cur = prev_gt.copy(), None
for k, fr in enumerate(frames):
    if k == 2:
        reg.set_profiling(True)
    predict = bench.first_predict(fr) if cur is None else bench.predict_next(last, cur)
    reg.set_input_target(fr["map"]); reg.set_input_source(fr["scan"])
    T, st = reg.scan_matching(predict, want_stats=True)
    last, cur = (cur if cur is not None else prev_gt), T
    print("frame", k, "gpu_ms", round(st.gpu_ms, 3))
prof = reg.get_profile()
d = reg.last_dbg
n = max(d[4], 1)
print({k: (v[0], round(1e3 * v[1] / max(v[0], 1), 2)) for k, v in prof.items()})
print(f"k_eval active launches={d[4]}: parallel {d[1]/n/1e3:.2f} us, partial-sum {d[2]/n/1e3:.2f} us, solver {d[3]/n/1e3:.2f} us")
print(f"k_correspond: blocks={d[9]} avg kNN-phase cycles/block={d[8]/max(d[9],1):.0f} fit-phase cycles/block={d[10]/max(d[9],1):.0f}")
print("avg SM cycles per active eval: warp 2 gn_model %.0f, gn_model + speculative step %.0f | warp 0 wait-for-cand + decision %.0f, wait-for-model + step %.0f | warp 1 proj %.0f | warp 3 log(cand) %.0f" % tuple(d[i] / n for i in (6, 13, 11, 12, 5, 7)))
nb = max(d[9], 1)
print("k_correspond thread-0 stages (avg SM cycles; NOTE slots 11-14 are shared with the solver stamps, run with few frames): brick probes + cell list %.0f, (unused) %.0f, candidates %.0f, merge %.0f" % tuple(d[i] / nb for i in (11, 12, 13, 14)))
