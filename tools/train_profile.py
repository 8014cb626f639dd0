"""Where an RC-Net training step (bench.py --workload train, one GPU) spends its device time: torch.profiler kernel table."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from riders_b200 import synth  # noqa: E402
from riders_b200.rcnet.rcnet_model import RCNetModel  # noqa: E402
from riders_b200.train import TrainStep  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = RCNetModel(device=torch.device("cpu"), **bench.MODEL_ARGS)
model.device = dev
model.to(dev)
model.channels_last()
torch.backends.cudnn.benchmark = True
ts = TrainStep(model, bf16="--f32" not in sys.argv)
FR, KP, PATCH, H, W = 4, 30, bench.PATCH, bench.H, bench.W
frames = [synth.chain_frame(i, KP, H, W, PATCH) for i in range(FR)]
img = torch.from_numpy(np.stack([np.pad(f["image"], ((0, 0), (PATCH[0] // 2,) * 2, (PATCH[1] // 2,) * 2), mode="edge") for f in frames])).to(dev)
img = img.contiguous(memory_format=torch.channels_last)
pts = torch.from_numpy(np.concatenate([f["points"] for f in frames]).astype(np.float32)).to(dev)
boxes = [torch.from_numpy(f["boxes"]).to(dev) for f in frames]
gt = (torch.rand((FR * KP, 1) + PATCH, device=dev) < 0.1).float() * (pts[:, 2].view(-1, 1, 1, 1) + torch.randn((FR * KP, 1) + PATCH, device=dev))
for _ in range(4):
    ts.step(img, pts, boxes, gt)
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        ts.step(img, pts, boxes, gt)
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, "device_time_total", None)
    if t is None:
        t = getattr(e, "cuda_time_total", 0)
    if e.device_type == torch.autograd.DeviceType.CUDA and t > 0:
        rows.append((t / 3e3, e.count // 3, e.key[:110]))
rows.sort(reverse=True)
total = sum(r[0] for r in rows)
print(json.dumps({"kernel_ms_per_step": round(total, 2)}))
for ms, n, name in rows[:28]:
    print("%8.3f ms  %5d x  %s" % (ms, n, name))
