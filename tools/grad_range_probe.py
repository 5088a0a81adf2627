"""How far the f16 gradient maps of the native backward are from saturation: trains a preset for a few Adam steps and prints,
per step, the loss, the loss scale S and max |scaled gradient| of every block's dy / dc0 maps (65504 = saturated).
Usage: python tools/grad_range_probe.py [config] [steps] [lr]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "boosting-nerv_b200"))
import bench  # noqa: E402
from bnerv_b200 import ops, train  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "hnerv_l"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
lr = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-4
model, a = bench.build_model(cfg)
model = model.cuda().train()
model.engine().train_graph = False          # eager: the wrappers below see every launch
os.environ["BNERV_TRAIN_NO_FALLBACK"] = "1"
rec = []
_front, _mid, _hb = ops.block_front_bwd, ops.resblock_mid_bwd, ops.head_bwd


def front(du, dout, x0, dact, g0p, C, want_dy_sums=False):
    r = _front(du, dout, x0, dact, g0p, C, want_dy_sums)
    rec.append(("front", tuple(du.shape[2:4]), C, float(du.abs().max()), float(dout.abs().max()), float(r[0].abs().max())))
    return r


def mid(dw, v, dact, g1p, C):
    r = _mid(dw, v, dact, g1p, C)
    rec.append(("mid", tuple(dw.shape[2:4]), C, float(dw.abs().max()), 0.0, float(r[0].abs().max())))
    return r


scales = []


def hb(dimg, img, scale):
    r = _hb(dimg, img, scale)
    scales.append(scale)
    return r


ops.block_front_bwd, ops.resblock_mid_bwd, ops.head_bwd = front, mid, hb
train.ops = ops
fh, fw = [int(v) for v in a.fc_hw.split("_")]
up = 1
for s_ in a.dec_strds:
    up *= s_
H, W = fh * up, fw * up
is_h = a.model == "HNeRV_Boost"
t = torch.tensor([0.5], dtype=torch.float64, device="cuda")
emb = torch.rand(1, 16, fh, fw, device="cuda") if is_h else None
yy, xx = torch.meshgrid(torch.linspace(0, 1, H, device="cuda"), torch.linspace(0, 1, W, device="cuda"), indexing="ij")
frame = torch.stack([0.5 + 0.45 * torch.sin(6.2832 * ((1 + c) * xx + (2 - 0.5 * c) * yy)) for c in range(3)])[None]
opt = torch.optim.Adam(model.parameters(), lr=lr)
for it in range(steps):
    rec.clear()
    scales.clear()
    opt.zero_grad(set_to_none=True)
    out = (model.forward_decoder(emb, t) if is_h else model(t))[0]
    loss = ((out - frame) ** 2).mean()
    loss.backward()
    opt.step()
    if it < 4 or it % 8 == 0 or it == steps - 1:
        S = float(scales[0][0]) if scales else float("nan")
        worst = max(rec, key=lambda r: max(r[3], r[4], r[5])) if rec else None
        print(f"step {it:3d} loss {loss.item():.5f} S {S:.3g} status {train.gradient_range_status()} ctrl {train.loss_scale_state()} worst {worst}")
        if it in (0, steps - 1):
            for r in rec:
                print("    ", r[0], r[1], "C", r[2], f"in {r[3]:.1f} {r[4]:.1f} out {r[5]:.1f}")
