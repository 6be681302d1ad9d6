"""(f)-2 measurement: PCA feature extraction of a synthetic general cloud, GPU (through the C ABI, host buffers in and
out) vs the CPU restatement on all host cores.

    python tools/feature_bench.py [n_points ...]
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import tloam_b200  # noqa: E402
from tloam_b200 import synth  # noqa: E402


def measure(reg, n, reps=20, cpu=True):
    pts = synth.general_cloud(n, seed=77)
    for _ in range(3):
        out = reg.extract_planar_sphere(pts)
    t0 = time.perf_counter()
    for _ in range(reps):
        out = reg.extract_planar_sphere(pts)
    e2e_ms = 1e3 * (time.perf_counter() - t0) / reps
    reg.set_profiling(True)
    for _ in range(5):
        reg.extract_planar_sphere(pts)
    prof = reg.get_profile()
    reg.set_profiling(False)
    launches, ms = prof["feature"]
    dbg = reg.last_dbg
    phases = None
    if dbg[11]:
        phases = {"blocks_sampled": dbg[11], "search_cycles": dbg[8] / dbg[11], "cumulant_cycles": dbg[9] / dbg[11],
                  "eigen_cycles": dbg[10] / dbg[11]}
    res = {"points": int(pts.shape[0]), "gpu_e2e_ms": e2e_ms, "gpu_kernels_ms": ms / 5, "kernel_launches_per_call": launches / 5 + 2,
           "k_fe_pca_thread0_phases": phases, "h2d_bytes": int(pts.nbytes), "lists": [int(len(x)) for x in out[:4]],
           "what": "tloam_b200_extract_planar_sphere: host cloud in, four host index lists out (grid build + kNN/PCA + "
                   "classification + 2 CUB radix sorts)"}
    if cpu:
        from oracle import pyoracle
        pyoracle.build()
        ref = pyoracle.extract_planar_sphere(pts)
        t0 = time.perf_counter()
        for _ in range(3):
            ref = pyoracle.extract_planar_sphere(pts)
        res["cpu_port_ms"] = 1e3 * (time.perf_counter() - t0) / 3
        res["cpu_cores"] = os.cpu_count()
        res["identical_to_cpu_port"] = bool(all(np.array_equal(a, b) for a, b in zip(out, ref)))
        res["speedup"] = res["cpu_port_ms"] / e2e_ms
    return res


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [50_000, 120_000]
    reg = tloam_b200.LocalRegistration()
    for n in sizes:
        print(json.dumps(measure(reg, n)))
    reg.close()


if __name__ == "__main__":
    main()
