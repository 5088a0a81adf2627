"""Model-level check of the native training path: gradients of every parameter, native (sm_100a fwd+bwd) vs torch
autograd through the plain fp32 forward, on the same weights/inputs/target; then a short optimisation run with both.
Usage: python tools/train_probe.py [family ...]"""
import copy
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "boosting-nerv_b200"))
from bnerv_b200 import ENeRV_Boost, HNeRV_Boost, NeRV_Boost, tiny_args, preset  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda")


def build(family, args=None):
    a = args or tiny_args(family)
    torch.manual_seed(3)
    m = NeRV_Boost(1, a) if family == "NeRV_Boost" else ENeRV_Boost(3, a) if family == "ENeRV_Boost" else HNeRV_Boost(a)
    return m.to(dev).train(), a


def run(m, family, t, emb, target):
    if family == "HNeRV_Boost":
        img, lst, _ = m.forward_decoder(emb, t)
    else:
        img, lst, _ = m(t)
    loss = ((img - target) ** 2).mean() + 0.3 * (img - target).abs().mean()
    return img, loss


def grads(family, args=None, B=2):
    m, a = build(family, args)
    fh, fw = [int(v) for v in a.fc_hw.split("_")]
    t = torch.tensor([(i + 1) / 8 for i in range(B)], dtype=torch.float64, device=dev)
    emb = torch.rand(B, 16, fh, fw, device=dev, requires_grad=True) if family == "HNeRV_Boost" else None
    with torch.no_grad():
        m.train_backend = "torch"
        img0, _ = run(m, family, t, emb, 0.0)
    target = torch.rand_like(img0)
    res = {}
    for mode in ("torch", "b200"):
        m.train_backend = mode
        m.zero_grad(set_to_none=True)
        if emb is not None:
            emb.grad = None
        img, loss = run(m, family, t, emb, target)
        loss.backward()
        res[mode] = ({n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None},
                     None if emb is None else emb.grad.detach().clone(), img.detach(), loss.item())
    gt, et, it, lt = res["torch"]
    gn, en, inn, ln = res["b200"]
    print(f"{family}: loss torch {lt:.6f} native {ln:.6f}; img max-rel {((it - inn).abs().max() / it.abs().max()).item():.2e}; "
          f"params with grad torch {len(gt)} native {len(gn)}")
    worst = []
    for n in gt:
        if n not in gn:
            print("   MISSING native grad:", n)
            continue
        a_, b_ = gn[n].double().flatten(), gt[n].double().flatten()
        rel = ((a_ - b_).abs().max() / b_.abs().max().clamp_min(1e-30)).item()
        cos = torch.nn.functional.cosine_similarity(a_, b_, dim=0).item()
        worst.append((rel, cos, n, b_.abs().max().item()))
    worst.sort(reverse=True)
    for rel, cos, n, mx in worst[:8]:
        print(f"   {n:45s} max-rel {rel:.2e} cos {cos:.6f} |g|max {mx:.2e}")
    print(f"   median max-rel {sorted(w[0] for w in worst)[len(worst) // 2]:.2e}; min cos {min(w[1] for w in worst):.6f}")
    if et is not None:
        print(f"   d/d(embed) max-rel {((et - en).abs().max() / et.abs().max()).item():.2e}")


def fit(family, steps=60):
    out = {}
    for mode in ("torch", "b200"):
        m, a = build(family)
        m.train_backend = mode
        fh, fw = [int(v) for v in a.fc_hw.split("_")]
        B = 4
        t = torch.tensor([(i + 1) / B for i in range(B)], dtype=torch.float64, device=dev)
        g = torch.Generator(device="cpu").manual_seed(5)
        emb = torch.rand(B, 16, fh, fw, generator=g).to(dev) if family == "HNeRV_Boost" else None
        with torch.no_grad():
            m.train_backend = "torch"
            shape = run(m, family, t, emb, 0.0)[0].shape
            m.train_backend = mode
        yy, xx = torch.meshgrid(torch.linspace(0, 1, shape[2]), torch.linspace(0, 1, shape[3]), indexing="ij")
        target = torch.stack([0.5 + 0.5 * torch.sin(6.28 * (2 * xx + 3 * yy + 0.25 * i + 0.1 * c)) for i in range(B) for c in range(3)])
        target = target.view(B, 3, *shape[2:]).to(dev)
        opt = torch.optim.Adam(m.parameters(), lr=2e-3)
        losses = []
        torch.cuda.synchronize()
        t0 = time.time()
        for s in range(steps):
            opt.zero_grad(set_to_none=True)
            _, loss = run(m, family, t, emb, target)
            loss.backward()
            opt.step()
            losses.append(loss.item())
        torch.cuda.synchronize()
        out[mode] = losses
        print(f"{family} fit [{mode}]: loss {losses[0]:.4f} -> {losses[steps // 2]:.4f} -> {losses[-1]:.4f}  ({(time.time() - t0) / steps * 1e3:.1f} ms/step)")
        if mode == "b200":       # trained-weights parity of the decode path
            m.eval()
            with torch.no_grad():
                nat = (m.forward_decoder(emb, t) if family == "HNeRV_Boost" else m(t))[0]
                m.backend = "torch"
                ref = (m.forward_decoder(emb, t) if family == "HNeRV_Boost" else m(t))[0]
            print(f"   decode parity on the trained weights: max-rel {((nat - ref).abs().max() / ref.abs().max()).item():.2e}")


if __name__ == "__main__":
    fams = sys.argv[1:] or ["HNeRV_Boost", "NeRV_Boost", "ENeRV_Boost"]
    for f in fams:
        try:
            grads(f)
            fit(f)
        except Exception as ex:
            import traceback
            traceback.print_exc()
            try:
                torch.cuda.synchronize()
            except Exception:
                print("context dead")
                sys.exit(3)
