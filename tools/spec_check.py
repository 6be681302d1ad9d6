"""Speculative solver step vs the serial advance(): poses and traces must be bit-identical (run with
TLOAM_B200_LIB=build/variants/no_spec.so for the serial side)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import tloam_b200
from tloam_b200 import synth
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from test_gpu_batch import scenes, CAPS

out = []
for i, sc in enumerate(scenes(4)):
    for fuse in ("0", "1"):
        os.environ["TLOAM_B200_FUSE"] = fuse
        r = tloam_b200.LocalRegistration(**CAPS)
        r.set_input_target(sc["map"]); r.set_input_source(sc["scan"])
        T, st = r.scan_matching(sc["predict"], want_stats=True)
        r.close()
        out.append({"scene": i, "fuse": fuse, "T": T.reshape(-1).tolist(), "n_inner": [st.outer[o].n_inner for o in range(st.n_outer)],
                    "term": [st.outer[o].termination for o in range(st.n_outer)],
                    "acc": [[st.outer[o].inner[k].accepted for k in range(st.outer[o].n_inner)] for o in range(st.n_outer)],
                    "mcc": [[st.outer[o].inner[k].model_cost_change for k in range(st.outer[o].n_inner)] for o in range(st.n_outer)],
                    "x_end": [list(st.outer[o].x_end) for o in range(st.n_outer)],
                    "H0": [list(st.outer[o].H0) for o in range(st.n_outer)], "nf": [list(st.outer[o].n_factors) for o in range(st.n_outer)]})
json.dump(out, open(sys.argv[1], "w"))
print("wrote", sys.argv[1])
