"""Minimal driver for ncu captures: a few frames of the bench workload (config 2), no torch.

    ncu ... python tools/profile_frame.py [frames] [scale]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
import tloam_b200  # noqa: E402
from tloam_b200 import synth  # noqa: E402


def main():
    nframes = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    if scale == 1.0:
        frames, prev_gt = bench.gen_frames("00", nframes)
    else:
        st = synth.Stream(cfg=synth.scaled(scale, seed=5), seq="00", start=100)
        frames = [st.frame() for _ in range(nframes)]
        prev_gt = frames[0]["T_gt"]
    reg = tloam_b200.LocalRegistration(**bench.CAPS)
    last, cur = prev_gt.copy(), None
    for fr in frames:
        predict = bench.first_predict(fr) if cur is None else bench.predict_next(last, cur)
        reg.set_input_target(fr["map"])
        reg.set_input_source(fr["scan"])
        T, st = reg.scan_matching(predict, want_stats=True)
        dt = np.linalg.norm(T[:3, 3] - fr["T_gt"][:3, 3])
        print(f"frame {fr['frame_id']}: gpu_ms={st.gpu_ms:.3f} launches={st.gpu_launches} |t - t_gt|={dt:.4f} "
              f"outer={st.n_outer} inner={[st.outer[i].n_inner for i in range(st.n_outer)]}")
        last, cur = (cur if cur is not None else prev_gt), T
    reg.close()


if __name__ == "__main__":
    main()
