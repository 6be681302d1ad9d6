"""One pass of the segmentation front half (after one warm-up pass) for ncu:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_seg_launches.csv python tools/segmentation_profile.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import tloam_b200  # noqa: E402
from tloam_b200 import synth  # noqa: E402

reg = tloam_b200.LocalRegistration()
scan = synth.raw_scan()
for _ in range(2):
    ge = reg.ground_extract(scan)
    opts = np.ascontiguousarray(scan[ge["object"]])
    obeam = ge["beam"][ge["object"]].astype(np.float64)
    os_ = reg.object_segmentation(opts, details=False)
    spts = np.ascontiguousarray(opts[os_["segmented"]])
    ee = reg.extract_edge(spts, obeam[os_["segmented"]], ring_min_num=131)
print(len(scan), len(opts), len(spts), len(ee["edge"]), len(ee["non_edge"]))
reg.close()
