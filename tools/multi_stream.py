"""Throughput of S independent sequences sharing ONE B200 (S handles, S CUDA streams, async ABI).

One frame is only 8.5-17 warps per SM, so a single stream is latency-bound; independent sequences (BASELINE
config 4 puts one per GPU) can also be co-scheduled on one GPU.  Each sequence keeps its own map / scan / pose.

    python tools/multi_stream.py [S ...]          inputs resident in HBM
    python tools/multi_stream.py --e2e [S ...]    inputs in pinned host memory: every frame uploads its 13 MB; the
                                                  upload of one sequence overlaps the solve of the others
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
import bench  # noqa: E402
import tloam_b200  # noqa: E402
from tloam_b200 import multi  # noqa: E402


def run(S, nframes=16, warm=3, e2e=False):
    seqs = [multi.sequence_for_rank(i) for i in range(S)]
    data = []
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    for s in seqs:
        frames, prev_gt = bench.gen_frames(s, nframes + warm, start=100)
        if e2e:
            dev = [([pin(c) for c in fr["map"]], [pin(c) for c in fr["scan"]]) for fr in frames]
            for mp, sc in dev:                             # first-touch DMA of freshly pinned pages is slow: warm it
                for a in mp + sc:
                    torch.from_numpy(a).cuda(non_blocking=True)
        else:
            dev = [([torch.from_numpy(c).cuda() for c in fr["map"]], [torch.from_numpy(c).cuda() for c in fr["scan"]]) for fr in frames]
        data.append((frames, prev_gt, dev))
    regs = [tloam_b200.LocalRegistration(**bench.CAPS) for _ in range(S)]      # each with its own stream
    state = [[d[1].copy(), None] for d in data]
    torch.cuda.synchronize()
    t0 = None
    for k in range(nframes + warm):
        if k == warm:
            for r in regs:
                r.synchronize()
            t0 = time.perf_counter()
        for i, r in enumerate(regs):                       # enqueue frame k of every sequence
            frames, prev_gt, dev = data[i]
            last, cur = state[i]
            predict = bench.first_predict(frames[k]) if cur is None else bench.predict_next(last, cur)
            if e2e:
                r.set_input_target(dev[k][0])
                r.set_input_source(dev[k][1])
            else:
                r.set_input_target_device(dev[k][0])
                r.set_input_source_device(dev[k][1])
            r.scan_matching_async(predict)
        for i, r in enumerate(regs):                       # collect
            T = r.get_result()
            last, cur = state[i]
            state[i] = [(cur if cur is not None else data[i][1]), T]
    for r in regs:
        r.synchronize()
    dt = time.perf_counter() - t0
    err = max(bench.pose_err(state[i][1], data[i][0][-1]["T_gt"])[0] for i in range(S))
    for r in regs:
        r.close()
    return {"streams": S, "inputs": "pinned host (13 MB uploaded per frame)" if e2e else "HBM-resident",
            "frames_per_s": S * nframes / dt, "ms_per_round": 1e3 * dt / nframes, "max_err_vs_gt_m": err}


if __name__ == "__main__":
    args = sys.argv[1:]
    e2e = "--e2e" in args
    for S in [int(a) for a in args if a != "--e2e"] or [1, 2, 4, 8]:
        print(json.dumps(run(S, e2e=e2e)), flush=True)
