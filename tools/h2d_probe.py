"""Times the host->device leg of set_target / set_source for pageable vs pinned host buffers."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import tloam_b200
from tloam_b200 import synth

f = synth.config1()
reg = tloam_b200.LocalRegistration()
pinned = [torch.from_numpy(c).pin_memory().numpy() for c in f["map"]]
dev = [torch.from_numpy(c).cuda() for c in f["map"]]
for name, arrs, fn in (("pageable", f["map"], reg.set_input_target), ("pinned", pinned, reg.set_input_target),
                       ("device", dev, reg.set_input_target_device)):
    for _ in range(3):
        fn(arrs); reg.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        fn(arrs)
    reg.synchronize()
    print(f"set_target {name}: {(time.perf_counter() - t0) / 10 * 1e3:.3f} ms")
t = torch.from_numpy(np.concatenate(f["map"])).pin_memory()
d = torch.empty_like(t, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    d.copy_(t, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
print(f"torch pinned H2D 12 MB: {dt * 1e3:.3f} ms = {t.numel() * 8 / dt / 1e9:.1f} GB/s")
