"""Loss forward+backward time at 1080p (batch 1): native SSIM / MS-SSIM kernels (bnerv_b200.losses) vs the torch formulation
pytorch_msssim executes on the GPU (grouped separable conv2d per moment map, autograd backward).
Usage: python tools/loss_bench.py"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "boosting-nerv_b200"))
from bnerv_b200 import losses  # noqa: E402


class mo:
    """The torch formulation pytorch_msssim executes (grouped separable conv2d per moment map), inlined here so that the
    tool does not import oracle/ (reserved for tests, smoke and bench.py's CPU leg)."""
    W = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)

    @staticmethod
    def _blur(x, win):
        c = x.shape[1]
        w = win.view(1, 1, 1, -1).repeat(c, 1, 1, 1)
        return F.conv2d(F.conv2d(x, w.transpose(2, -1), groups=c), w, groups=c)

    @staticmethod
    def _stats(x, y):
        co = torch.arange(11, dtype=x.dtype, device=x.device) - 5
        win = torch.exp(-(co ** 2) / (2 * 1.5 ** 2))
        win = win / win.sum()
        c1, c2 = 0.01 ** 2, 0.03 ** 2
        mu1, mu2 = mo._blur(x, win), mo._blur(y, win)
        s11, s22, s12 = mo._blur(x * x, win) - mu1 * mu1, mo._blur(y * y, win) - mu2 * mu2, mo._blur(x * y, win) - mu1 * mu2
        cs = (2 * s12 + c2) / (s11 + s22 + c2)
        return (((2 * mu1 * mu2 + c1) / (mu1 * mu1 + mu2 * mu2 + c1)) * cs).flatten(2).mean(-1), cs.flatten(2).mean(-1)

    @staticmethod
    def ssim(x, y, data_range=1.0, size_average=False):
        return torch.relu(mo._stats(x, y)[0]).mean(1)

    @staticmethod
    def ms_ssim(x, y, data_range=1.0, size_average=False):
        mcs = []
        for i in range(5):
            s, cs = mo._stats(x, y)
            if i < 4:
                mcs.append(torch.relu(cs))
                pad = [sz % 2 for sz in x.shape[2:]]
                x, y = F.avg_pool2d(x, 2, padding=pad), F.avg_pool2d(y, 2, padding=pad)
        vals = torch.stack(mcs + [torch.relu(s)], dim=0)
        return torch.prod(vals ** torch.tensor(mo.W, device=x.device).view(-1, 1, 1), dim=0).mean(1)

torch.manual_seed(0)
target = torch.rand(1, 3, 1080, 1920, device="cuda")
pred0 = (target + 0.05 * torch.randn_like(target)).clamp(0, 1)


def ref_loss(p, kind):
    l1 = F.l1_loss(p, target, reduction="none").flatten(1).mean(1)
    if kind == "Fusion10_freq":
        ms = 1 - mo.ms_ssim(p, target, 1.0, False)
        return (60 * (0.7 * l1 + 0.3 * ms) + losses._freq_l1(p, target)).mean()
    if kind == "Fusion6":
        return (0.7 * l1 + 0.3 * (1 - mo.ssim(p, target, 1.0, False))).mean()
    return (1 - mo.ms_ssim(p, target, 1.0, False)).mean()


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print(torch.cuda.get_device_name(0), "loss forward+backward @1080x1920, batch 1, ms")
for kind in ("Fusion10_freq", "Fusion6", "ms_ssim only"):
    def native():
        p = pred0.clone().requires_grad_(True)
        (losses.loss_fn(p, target, kind) if kind != "ms_ssim only" else (1 - losses.ms_ssim(p, target)).mean()).backward()

    def torch_ref():
        p = pred0.clone().requires_grad_(True)
        ref_loss(p, kind).backward()

    print(f"{kind:16s} native {timed(native):7.3f}   torch restatement of pytorch_msssim {timed(torch_ref):7.3f}", flush=True)
