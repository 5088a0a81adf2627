"""Prints, for the three reference-trained goldens, how far the device decode is from the f32 reference and from the oracle's
f16-operand emulation (image and every intermediate map) and the PSNR against the ground-truth frames.  GPU; reads tests/golden."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "boosting-nerv_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from conftest import elementwise_rel, load_golden, max_rel  # noqa: E402
from oracle import nerv_oracle as orc  # noqa: E402  (checker only)
from bnerv_b200 import ENeRV_Boost, HNeRV_Boost, NeRV_Boost, tiny_args  # noqa: E402

for model, gold in (("HNeRV_Boost", "hnerv_tiny_trained.npz"), ("NeRV_Boost", "nerv_tiny_trained.npz"), ("ENeRV_Boost", "enerv_tiny_trained.npz")):
    sd, g = load_golden(gold)
    a = tiny_args(model)
    m = (NeRV_Boost(1, a) if model == "NeRV_Boost" else ENeRV_Boost(3, a) if model == "ENeRV_Boost" else HNeRV_Boost(a)).eval()
    m.load_state_dict(sd)
    m = m.cuda()
    m.keep_intermediates = True
    inputs = (g["emb"], g["t"]) if model == "HNeRV_Boost" else (g["t"],)
    with torch.no_grad():
        img, outs, _ = m.forward_decoder(*[v.cuda() for v in inputs]) if model == "HNeRV_Boost" else m(g["t"].cuda())
    orc.EMULATE = torch.float16
    emu_img, emu_outs = orc.forward(model, sd, orc.cfg_from_args(a), *inputs)
    orc.EMULATE = None
    f = lambda xs: "[" + ", ".join(f"{x:.1e}" for x in xs) + "]"
    print(f"{gold}: image vs ref {max_rel(img.cpu(), g['img']):.2e} (element-wise |a-b|/(|b|+1e-3): {elementwise_rel(img.cpu(), g['img']):.2e}), "
          f"vs emulation {max_rel(img.cpu(), emu_img):.2e}; "
          f"maps vs ref {f([max_rel(o.cpu(), g[f'out{i}']) for i, o in enumerate(outs)])}, vs emulation "
          f"{f([max_rel(o.cpu(), emu_outs[i]) for i, o in enumerate(outs)])}; PSNR vs frames: ours {orc.psnr(img.cpu(), g['frame']):.4f} dB, "
          f"reference {orc.psnr(g['img'], g['frame']):.4f} dB")
