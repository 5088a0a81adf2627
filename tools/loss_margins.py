import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/boosting-nerv_b200"); sys.path.insert(0, "/root/repo/tests")
from test_gpu_losses import _pair
from bnerv_b200 import losses
from oracle import msssim_oracle as mo
for shape in [(1, 3, 180, 320), (2, 3, 181, 203), (1, 3, 360, 640), (1, 3, 1080, 1920)]:
    pred, target = _pair(*shape, seed=1)
    p1 = pred.clone().requires_grad_(True); p2 = pred.clone().requires_grad_(True)
    v1 = losses.ms_ssim(p1, target); v2 = mo.ms_ssim(p2, target, 1.0, False)
    v1.sum().backward(); v2.sum().backward()
    print(shape, "value rel", ((v1 - v2).abs().max() / v2.abs().max()).item(), "grad rel", ((p1.grad - p2.grad).abs().max() / p2.grad.abs().max()).item())
for shape in [(2, 3, 40, 64), (1, 3, 37, 53), (2, 3, 180, 320)]:
    pred, target = _pair(*shape)
    p1 = pred.clone().requires_grad_(True); p2 = pred.clone().requires_grad_(True)
    v1 = losses.ssim(p1, target); v2 = mo.ssim(p2, target, 1.0, False)
    v1.sum().backward(); v2.sum().backward()
    print("ssim", shape, "value rel", ((v1 - v2).abs().max() / v2.abs().max()).item(), "grad rel", ((p1.grad - p2.grad).abs().max() / p2.grad.abs().max()).item())
