"""BASELINE config 3 (dense indoor, planar-only, HBM-bound stress): parity vs the oracle at a reduced size and the
per-kernel timing / roofline of the full size (F = 500k features vs M = 2M map points).

    python tools/config3.py [scale]      # scale 1.0 = full size (GPU only; the oracle check runs at 0.04)
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
import tloam_b200  # noqa: E402
from tloam_b200 import synth  # noqa: E402

BIG = 10 ** 9


def run(scale, check):
    f = synth.config3(int(500_000 * scale), int(2_000_000 * scale))
    caps = dict(edge_maxnum=BIG, sphere_maxnum=BIG, planar_maxnum=BIG, ground_maxnum=BIG, factor_num=2)
    reg = tloam_b200.LocalRegistration(**caps)
    reg.set_input_target(f["map"])
    reg.set_input_source(f["scan"])
    for _ in range(3):
        T, st = reg.scan_matching(f["predict"], want_stats=True)
    out = {"scale": scale, "features": int(sum(c.shape[0] for c in f["scan"])), "map_points": int(sum(c.shape[0] for c in f["map"])),
           "gpu_ms_per_frame": st.gpu_ms, "err_vs_gt_m": float(np.linalg.norm(T[:3, 3] - f["T_gt"][:3, 3]))}
    reg.set_profiling(True)
    for _ in range(3):
        reg.set_input_target(f["map"])
        reg.scan_matching(f["predict"])
    prof = reg.get_profile()
    nfeat = [c.shape[0] for c in f["scan"]]
    alg = bench.algorithmic_bytes(nfeat, [c.shape[0] for c in f["map"]])
    peak, src = bench.measured_peak_hbm()
    kern = {}
    for k, (n, ms) in prof.items():
        if n:
            kern[k] = {"launches": n, "avg_us": 1e3 * ms / n}
    for k in ("correspond", "eval", "eval_first"):
        if k in kern:
            gbs = alg[k] / (kern[k]["avg_us"] * 1e-6) / 1e9
            kern[k].update(algorithmic_MB=alg[k] / 1e6, achieved_GBps=gbs, frac_of_peak=gbs / peak)
    mb = sum(v["avg_us"] for k, v in kern.items() if k.startswith("map_"))
    out["map_build"] = {"us": mb, "achieved_GBps": alg["map_build"] / (mb * 1e-6) / 1e9}
    out["kernels"] = kern
    out["peak_GBps"] = peak
    out["peak_source"] = src
    if check:
        from oracle import pyoracle
        o = pyoracle.Oracle(threads_mode=1, **caps)
        o.set_input_target(f["map"])
        o.set_input_source(f["scan"])
        t0 = time.perf_counter()
        rc, To, so = o.scan_matching(f["predict"])
        out["oracle_s"] = time.perf_counter() - t0
        d = np.linalg.inv(To) @ T
        out["dt_vs_oracle_m"] = float(np.linalg.norm(d[:3, 3]))
        out["dr_vs_oracle_rad"] = float(np.arccos(np.clip((np.trace(d[:3, :3]) - 1) / 2, -1, 1)))
        out["factors"] = [list(st.outer[i].n_factors) for i in range(st.n_outer)]
        out["factors_oracle"] = [list(so.outer[i].n_factors) for i in range(so.n_outer)]
    reg.close()
    return out


if __name__ == "__main__":
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    print(json.dumps(run(0.04, True)))
    if scale > 0.04:
        print(json.dumps(run(scale, False)))
