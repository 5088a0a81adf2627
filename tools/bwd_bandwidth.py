import sys, torch
sys.path.insert(0, "/root/repo/boosting-nerv_b200")
from bnerv_b200 import ops
C, H, W = 112, 1080, 1920
mk = lambda: torch.randn(ops.c8_shape(1, C, H, W), dtype=torch.float16, device="cuda")
a, b, c, d = mk(), mk(), mk(), mk()
g = torch.ones(1, 112, device="cuda")
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ms = timed(lambda: ops.block_front_bwd(a, b, c, d, g, C, want_dy_sums=True))
print(f"front_bwd distinct maps: {ms:.3f} ms  {5 * a.numel() * 2 / ms / 1e6:.0f} GB/s")
ms = timed(lambda: ops.resblock_mid_bwd(a, b, c, g, C))
print(f"mid_bwd distinct maps: {ms:.3f} ms  {4 * a.numel() * 2 / ms / 1e6:.0f} GB/s")
ms = timed(lambda: ops.unshuffle_c8(a, C, 2, want_sums=True))
print(f"unshuffle_sum s=2: {ms:.3f} ms  {2 * a.numel() * 2 / ms / 1e6:.0f} GB/s")
