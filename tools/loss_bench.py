"""Loss forward+backward time at 1080p (batch 1): native SSIM / MS-SSIM kernels (bnerv_b200.losses) vs the torch restatement of
pytorch_msssim (oracle/msssim_oracle.py run on the GPU - what the reference's loss_fn executes through the package).
Usage: python tools/loss_bench.py"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "boosting-nerv_b200"))
from bnerv_b200 import losses  # noqa: E402
from oracle import msssim_oracle as mo  # noqa: E402  (tool = measurement harness, like bench.py's cpu leg)

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
