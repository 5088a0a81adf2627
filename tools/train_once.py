"""N native training steps (forward + MSE + backward) of a preset, for profiling.  Usage: python tools/train_once.py [config] [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "boosting-nerv_b200"))
import bench  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "hnerv_l"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
model, args = bench.build_model(cfg)
model = model.cuda().train()
is_h = args.model == "HNeRV_Boost"
fh, fw = [int(v) for v in args.fc_hw.split("_")]
emb = torch.rand(1, 16, fh, fw, device="cuda") if is_h else None
t = torch.tensor([0.5], dtype=torch.float64, device="cuda")
target = None
for _ in range(steps):
    model.zero_grad(set_to_none=True)
    img = (model.forward_decoder(emb, t) if is_h else model(t))[0]
    if target is None:
        target = torch.rand_like(img)
    ((img - target) ** 2).mean().backward()
torch.cuda.synchronize()
print("ok", tuple(img.shape))
