"""Which kernels make up one FP32 regression step of HNeRV-Bunny-3M (methods/regress.DecoderTrainer)?  torch.profiler,
CUDA activities only, 5 steps after warm-up; prints the top kernels by device time.  GPU only."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from neuroquant_b200.methods.regress import DecoderTrainer  # noqa: E402
from neuroquant_b200.models import HNeRV  # noqa: E402
from neuroquant_b200.workloads import WORKLOADS  # noqa: E402

if __name__ == "__main__":
    arch, cfg = WORKLOADS["hnerv-bunny-3m"]
    torch.manual_seed(903)
    model = HNeRV(dict(cfg)).cuda().train()
    frames = torch.rand(2, 3, cfg["crop_h"], cfg["crop_w"]).cuda()
    tr = DecoderTrainer(model, arch, 1e-3)
    for _ in range(5):
        tr.step(frames, frames)
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            tr.step(frames, frames)
        torch.cuda.synchronize()
    rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:14]
    total = sum(e.device_time_total for e in prof.key_averages())
    print(f"device time per step: {total / 5 / 1e3:.3f} ms")
    for e in rows:
        print(f"{e.device_time_total / 5 / 1e3:8.3f} ms  x{e.count // 5:<4d} {e.key[:110]}")
