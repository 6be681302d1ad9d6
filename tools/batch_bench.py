"""Batched-registration throughput only (A/B tool): S sequences per launch, inputs resident in HBM.

    [TLOAM_B200_LIB=build/variants/x.so] python tools/batch_bench.py [S ...]
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
import bench  # noqa: E402
import tloam_b200  # noqa: E402
from tloam_b200 import multi  # noqa: E402

for S in [int(a) for a in sys.argv[1:]] or [8]:
    seqs = [multi.sequence_for_rank(i) for i in range(S)]
    data = bench.gen_many(seqs, 3 + 12)
    b = tloam_b200.BatchRegistration(S, **bench.CAPS)
    ms = [bench.run_batch(b, data, torch, 3, 12, "device")[0] for _ in range(3)]
    b.set_profiling(True)
    bench.run_batch(b, [(d[0][:4], d[1]) for d in data], torch, 0, 4, "device")
    prof = {k: round(1e3 * v[1] / v[0], 2) for k, v in b.get_profile().items() if v[0]}
    print(json.dumps({"S": S, "frames_per_s": S * 12 / (np.median(ms) * 1e-3), "ms_per_batch_frame": float(np.median(ms)) / 12,
                      "avg_us": prof, "lib": os.environ.get("TLOAM_B200_LIB", "default")}), flush=True)
    b.close()
    for G in (2, 4):
        if S % G:
            continue
        bs = [tloam_b200.BatchRegistration(S // G, **bench.CAPS) for _ in range(G)]
        msg = [bench.run_batch_groups(bs, data, torch, 3, 12)[0] for _ in range(3)]
        print(json.dumps({"S": S, "groups": G, "frames_per_s": S * 12 / (np.median(msg) * 1e-3)}), flush=True)
        for b in bs:
            b.close()
