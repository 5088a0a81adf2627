"""Decode a few frames of a preset (for ncu captures).  Usage: python tools/frame_once.py [config] [frames]
The last frame is bracketed by cudaProfilerStart/Stop: `ncu --profile-from-start off` captures exactly one warm frame."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "boosting-nerv_b200"))
import bench  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "hnerv_l"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 2
model, args = bench.build_model(cfg)
model = model.cuda()
fh, fw = [int(v) for v in args.fc_hw.split("_")]
emb = torch.rand(1, 16, fh, fw, device="cuda")
t = torch.tensor([0.5], dtype=torch.float64, device="cuda")
with torch.no_grad():
    for i in range(frames):
        if i == frames - 1:
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStart()
        img = model.decode(emb, t) if args.model == "HNeRV_Boost" else model.decode(t)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print(tuple(img.shape))
