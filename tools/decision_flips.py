"""How many discrete decisions of one registration flip under +-1 ulp input perturbations, and how far the pose moves.

The path is full of thresholds (5 nearest within r, lambda2 > 3 lambda1, |dir.z| > 0.85, plane_dis > 0.2, TLS weight
cut-offs, trust-region accept / reject): an implementation that differs from the reference's Ceres / Eigen arithmetic
by one rounding can flip some of them.  This tool puts an error bar on the 1e-4 m / 1e-5 rad parity claim: BASELINE
config 1 (F = 40k vs M = 500k) is registered once as is and R times with every scan AND map coordinate moved by one ulp
in a random direction.

    python tools/decision_flips.py [R]        # JSON on stdout (GPU)
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import tloam_b200  # noqa: E402
from tloam_b200 import synth  # noqa: E402

BIG = 10 ** 9
CAPS = dict(edge_maxnum=BIG, sphere_maxnum=BIG, planar_maxnum=BIG, ground_maxnum=BIG)


def ulp_jitter(a, rng, level="f64"):
    """+-1 ulp in a random direction: of the FP64 value, or of its FP32 rounding (the device stores map coordinates in
    FP32 relative to the map origin, so only an FP32-sized step reaches the search)."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    up = rng.random(a.shape) < 0.5
    if level == "f32":
        f = a.astype(np.float32)
        step = np.where(up, np.nextafter(f, np.float32(np.inf)), np.nextafter(f, np.float32(-np.inf))).astype(np.float64) - f.astype(np.float64)
        return a + step
    return np.where(up, np.nextafter(a, np.inf), np.nextafter(a, -np.inf))


def register(mp, scan, predict, x0):
    r = tloam_b200.LocalRegistration(**CAPS)
    r.set_input_target(mp)
    r.set_input_source(scan)
    valid = [r.build_factors(c, x0)[0] for c in range(4)]
    T, st = r.scan_matching(predict, want_stats=True)
    trace = {"n_factors": [list(st.outer[i].n_factors) for i in range(st.n_outer)],
             "accepted": [[st.outer[i].inner[k].accepted for k in range(min(st.outer[i].n_inner, 8))] for i in range(st.n_outer)],
             "n_inner": [st.outer[i].n_inner for i in range(st.n_outer)]}
    r.close()
    return T, valid, trace


def run(R=5, scale=1.0, seed=1, level="f64"):
    f = synth.config1() if scale >= 1.0 else None
    if f is None:
        cfg = synth.scaled(scale, seed=20260924 + 1000)
        T_gt = synth.se3_exp(synth.CONFIG1_GT)
        f = dict(map=synth.make_map(cfg, T_gt), scan=synth.make_scan(cfg, T_gt, 0), predict=T_gt @ synth.se3_exp(synth.CONFIG1_PERTURB), T_gt=T_gt)
    reg = tloam_b200.LocalRegistration()
    x0 = reg.se3_log(f["predict"])
    reg.close()
    T0, v0, tr0 = register(f["map"], f["scan"], f["predict"], x0)
    rng = np.random.default_rng(seed)
    rows = []
    for _ in range(R):
        mp = [ulp_jitter(c, rng, level) for c in f["map"]]
        sc = [ulp_jitter(c, rng, level) for c in f["scan"]]
        T, v, tr = register(mp, sc, f["predict"], x0)
        d = np.linalg.inv(T0) @ T
        rows.append({"validity_flips_outer0": [int(np.sum(a != b)) for a, b in zip(v0, v)],
                     "factor_count_delta_per_outer": [[int(a - b) for a, b in zip(x, y)] for x, y in zip(tr0["n_factors"], tr["n_factors"])],
                     "same_accept_reject_sequence": tr0["accepted"] == tr["accepted"] and tr0["n_inner"] == tr["n_inner"],
                     "dt_m": float(np.linalg.norm(d[:3, 3])),
                     "dr_rad": float(np.arccos(np.clip((np.trace(d[:3, :3]) - 1) / 2, -1, 1)))})
    nfeat = [int(len(c)) for c in f["scan"]]
    return {"features": nfeat, "perturbation": f"+-1 {level} ulp on every scan and map coordinate, random sign", "runs": rows,
            "max_validity_flips_outer0": int(max(sum(r["validity_flips_outer0"]) for r in rows)),
            "max_dt_m": max(r["dt_m"] for r in rows), "max_dr_rad": max(r["dr_rad"] for r in rows),
            "note": "map coordinates are stored in FP32 on the device, so a 1-ulp FP64 change of a map point only matters when it crosses "
                    "an FP32 rounding boundary; scan coordinates stay FP64 end to end"}


if __name__ == "__main__":
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    print(json.dumps({"f64_ulp": run(R, level="f64"), "f32_ulp": run(R, level="f32")}))
