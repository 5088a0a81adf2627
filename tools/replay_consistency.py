"""Decode the same frame N times through the captured graph of a preset and check that every replay is bit-identical to the first
(a race in one of the multi-stage streaming kernels would show up as rare differing frames).
Usage: python tools/replay_consistency.py [config] [replays]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "boosting-nerv_b200"))
import bench  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "nerv_s"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 300
model, args = bench.build_model(cfg)
model = model.cuda().eval()
fh, fw = [int(v) for v in args.fc_hw.split("_")]
emb = torch.rand(1, 16, fh, fw, device="cuda")
t = torch.tensor([0.37], dtype=torch.float64, device="cuda")
is_h = args.model == "HNeRV_Boost"
with torch.no_grad():
    first = (model.decode(emb, t) if is_h else model.decode(t)).clone()
    bad = 0
    for i in range(n):
        img = model.decode(emb, t) if is_h else model.decode(t)
        if not torch.equal(img, first):
            bad += 1
torch.cuda.synchronize()
print(f"{cfg}: {n} replays, {bad} differ from the first; finite {bool(torch.isfinite(first).all())}, range [{first.min().item():.3f}, {first.max().item():.3f}]")
sys.exit(1 if bad else 0)
