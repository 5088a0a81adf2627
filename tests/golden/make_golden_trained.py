"""Mint a golden from a BRIEFLY TRAINED reference model (trained pre-sin magnitudes differ from the
initialisation's, so parity has to be shown on trained weights too).  Build container only:
    python tests/golden/make_golden_trained.py
The UNMODIFIED reference HNeRV_Boost / NeRV_Boost / ENeRV_Boost (tiny configs of bnerv_b200.config.tiny_args, imported from /root/reference with the
sys.modules stubs of make_golden.py) are trained on CPU for 1200 Adam steps, the way train_nerv_all.py:342-348 steps it
(model(frame, norm_idx) -> L2 loss -> backward -> optimiser step; batch 1), on 8 synthetic 40x80 frames
(0.5 + 0.5 sin(2 pi (fx x + fy y + ft t)) per channel + 0.05 uniform noise).  Written to
tests/golden/{hnerv,nerv,enerv}_tiny_trained.npz: the trained state_dict, the frames, and for two frames the encoder output, every decoder
block output and the image - all produced by the reference's own forward in eval mode.
"""
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
from make_golden import import_reference, pack  # noqa: E402


def synthetic_frames(n, h, w, seed=1):
    g = torch.Generator().manual_seed(seed)
    y, x = torch.meshgrid(torch.arange(h) / h, torch.arange(w) / w, indexing="ij")
    frames = []
    for i in range(n):
        t = i / n
        chans = [0.5 + 0.5 * torch.sin(2 * math.pi * (fx * x + fy * y + ft * t)) for fx, fy, ft in ((1.0, 2.0, 1.0), (3.0, 1.0, 2.0), (2.0, 3.0, 1.0))]
        f = torch.stack(chans) + 0.05 * (torch.rand(3, h, w, generator=g) - 0.5)
        frames.append(f.clamp(0, 1))
    return torch.stack(frames)


def train(m, frames, idx, is_h, steps, lr):
    opt = torch.optim.Adam(m.parameters(), lr=lr, betas=(0.9, 0.999))
    n = frames.shape[0]
    m.train()
    for step in range(steps):
        i = step % n
        img = (m(frames[i:i + 1], norm_idx=idx[i:i + 1]) if is_h else m(idx[i:i + 1]))[0]
        loss = torch.nn.functional.mse_loss(img, frames[i:i + 1])
        opt.zero_grad()
        loss.backward()
        opt.step()
        if step % 300 == 0 or step == steps - 1:
            print(f"  step {step:4d} loss {loss.item():.5f}  psnr {-10 * math.log10(loss.item()):.2f} dB")
    return m.eval()


def main():
    mb, mn, me, mh = import_reference()
    sys.path.insert(0, os.path.join(ROOT, "boosting-nerv_b200"))
    from bnerv_b200.config import tiny_args
    n = 8
    frames = synthetic_frames(n, 40, 80)
    idx = torch.tensor([(i + 1) / n for i in range(n)], dtype=torch.float64)      # hnerv_utils.py:47
    sel = [1, 6]
    t = idx[sel]
    only = sys.argv[1:]            # e.g. `nerv enerv` re-mints only those (hnerv_tiny_trained.npz is pinned by committed gates)

    if not only or "hnerv" in only:
        torch.manual_seed(1)
        m = train(mh.HNeRV_Boost(tiny_args("HNeRV_Boost")), frames, idx, True, 1200, 3e-3)
        with torch.no_grad():
            enc = m.forward_encoder(frames[sel])
            img, outs, _ = m.forward_decoder(enc, t)
            img_full, _, _ = m(frames[sel], norm_idx=t)
        sd = m.state_dict()
        gam = max(float(v.abs().max()) for k, v in sd.items() if k.endswith("gamma"))
        print("max |gamma| after training", gam, " max |pre-sin weight|", float(sd["decoder.1.conv.upconv.0.weight"].abs().max()))
        np.savez(os.path.join(HERE, "hnerv_tiny_trained.npz"),
                 **pack(sd, t=t, emb=enc, img=img, frame=frames[sel], enc=enc, img_full=img_full, **{f"out{i}": o for i, o in enumerate(outs)}))
    for key, ctor, name in (("nerv", lambda a: mn.NeRV_Boost(1, a), "NeRV_Boost"), ("enerv", lambda a: me.ENeRV_Boost(3, a), "ENeRV_Boost")):
        if only and key not in only:
            continue
        torch.manual_seed(1)
        print(name)
        m = train(ctor(tiny_args(name)), frames, idx, False, 1200, 3e-3)
        with torch.no_grad():
            img, outs, _ = m(t)
        np.savez(os.path.join(HERE, f"{key}_tiny_trained.npz"),
                 **pack(m.state_dict(), t=t, img=img, frame=frames[sel], **{f"out{i}": o for i, o in enumerate(outs)}))
    for f in sorted(os.listdir(HERE)):
        if f.endswith("_trained.npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
