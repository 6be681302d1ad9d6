"""Run-to-run determinism of the fused / un-fused frame on one scene: poses, H0 traces, accept sequences."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import tloam_b200
from test_gpu_batch import scenes, CAPS

sc = scenes(1)[0]
res = {"0": [], "1": []}
for rep in range(8):
    for fuse in ("0", "1"):
        os.environ["TLOAM_B200_FUSE"] = fuse
        r = tloam_b200.LocalRegistration(**CAPS)
        r.set_input_target(sc["map"]); r.set_input_source(sc["scan"])
        T, st = r.scan_matching(sc["predict"], want_stats=True)
        r.close()
        res[fuse].append((T.copy(), [np.array(st.outer[o].H0) for o in range(st.n_outer)], [list(st.outer[o].n_factors) for o in range(st.n_outer)],
                          [st.outer[o].n_inner for o in range(st.n_outer)]))
for fuse in ("0", "1"):
    T0, H0, nf0, ni0 = res[fuse][0]
    for k, (T, H, nf, ni) in enumerate(res[fuse][1:], 1):
        dH = [float(np.abs(a - b).max() / np.abs(a).max()) for a, b in zip(H0, H)]
        print("fuse", fuse, "run", k, "T equal", bool(np.array_equal(T, T0)), "H0 rel diff per outer", dH, "nf equal", nf == nf0, "n_inner", ni)
Tu, Hu, _, _ = res["0"][0]
Tf, Hf, _, _ = res["1"][0]
for o, (a, b) in enumerate(zip(Hu, Hf)):
    d = np.abs(a - b); i = int(d.argmax())
    print("fused vs unfused outer", o, "max abs diff", float(d.max()), "at", i, "values", a[i], b[i], "scale", float(np.abs(a).max()))
