"""(f)-4 measurement: the segmentation front half (groundRemove -> objectSegmentation = DCVC -> extractEdgePoint) of a raw
64-beam scan.  Per stage: wall time through the C ABI (host buffers in and out), device time of its kernels (CUDA events
around every launch, tloam_b200_set_profiling), launches per call, and the CPU restatement on one host thread.

    python tools/segmentation_bench.py [n_az ...]        # points ~ 62 * n_az
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import tloam_b200  # noqa: E402
from tloam_b200 import synth  # noqa: E402


def timed(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    return out, 1e3 * (time.perf_counter() - t0) / reps


def kernels(reg, key, fn, reps=5):
    reg.set_profiling(True)
    for _ in range(reps):
        fn()
    launches, ms = reg.get_profile()[key]
    reg.set_profiling(False)
    return launches / reps, ms / reps


def measure(reg, n_az, cpu=True):
    scan = synth.raw_scan(n_az=n_az)
    res = {"points": int(len(scan))}
    ge, ms = timed(lambda: reg.ground_extract(scan), 10)
    res["ground_extract"] = {"gpu_e2e_ms": ms}
    res["ground_extract"]["launches"], res["ground_extract"]["gpu_kernels_ms"] = kernels(reg, "ground", lambda: reg.ground_extract(scan))
    opts = np.ascontiguousarray(scan[ge["object"]])
    obeam = ge["beam"][ge["object"]].astype(np.float64)
    os_, ms = timed(lambda: reg.object_segmentation(opts, details=False), 10)     # what the C++ shim asks for
    det = reg.object_segmentation(opts)                                             # + per-point classes, voxel indices, polar triples
    res["object_segmentation"] = {"points": int(len(opts)), "clusters": int(len(os_["sizes"])), "voxels": int(len(np.unique(det["voxel"]))),
                                  "gpu_e2e_ms": ms}
    res["object_segmentation"]["launches"], res["object_segmentation"]["gpu_kernels_ms"] = kernels(
        reg, "object", lambda: reg.object_segmentation(opts, details=False))
    spts = np.ascontiguousarray(opts[os_["segmented"]])
    sbeam = obeam[os_["segmented"]]
    ee, ms = timed(lambda: reg.extract_edge(spts, sbeam, ring_min_num=131), 10)
    res["extract_edge"] = {"points": int(len(spts)), "edge": int(len(ee["edge"])), "non_edge": int(len(ee["non_edge"])), "gpu_e2e_ms": ms}
    res["extract_edge"]["launches"], res["extract_edge"]["gpu_kernels_ms"] = kernels(
        reg, "edge", lambda: reg.extract_edge(spts, sbeam, ring_min_num=131))
    if cpu:
        from oracle import pyoracle
        pyoracle.build()
        o, ms = timed(lambda: pyoracle.ground_extract(scan), 3)
        res["ground_extract"]["cpu_port_ms"] = ms
        res["ground_extract"]["identical"] = bool(np.array_equal(o["ground"], ge["ground"]) and np.array_equal(o["object"], ge["object"]))
        o, ms = timed(lambda: pyoracle.dcvc(opts), 3)
        res["object_segmentation"]["cpu_port_ms"] = ms
        res["object_segmentation"]["identical"] = bool(np.array_equal(o["segmented"], os_["segmented"]) and np.array_equal(o["root"], det["root"]))
        o, ms = timed(lambda: pyoracle.extract_edge(spts, sbeam, ring_min_num=131), 3)
        res["extract_edge"]["cpu_port_ms"] = ms
        res["extract_edge"]["identical"] = bool(np.array_equal(o["edge"], ee["edge"]) and np.array_equal(o["non_edge"], ee["non_edge"]))
        res["cpu_threads"] = 1
    return res


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [1875]
    reg = tloam_b200.LocalRegistration()
    for n_az in sizes:
        print(json.dumps(measure(reg, n_az)))
    reg.close()


if __name__ == "__main__":
    main()
