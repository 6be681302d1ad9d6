"""The chained device flow (registration -> submap update -> next frame, (f)-1 + (f)-3) for an ncu launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_chain_launches.csv python tools/profile_chain.py [frames]
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
import bench  # noqa: E402
import tloam_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
frames, prev_gt = bench.gen_frames("00", n)
reg = tloam_b200.LocalRegistration(stream=torch.cuda.current_stream().cuda_stream, **bench.CAPS)
ms, h2d, err, poses, fit = bench.run_stream_device_submap(reg, frames, prev_gt, torch, 2, n - 2)
print("ms_per_frame", ms / (n - 2), "err", err)
reg.close()
