"""Launch the backward kernels at HNeRV-L's largest layer shape a few times (for ncu).  Usage: python tools/prof_wgrad.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "boosting-nerv_b200"))
from bnerv_b200 import ops  # noqa: E402

cin = cout = 112
H, W = 1080, 1920
torch.manual_seed(0)
x = ops.nchw_to_c8(torch.randn(1, cin, H, W, device="cuda"))
dy = ops.nchw_to_c8(torch.randn(1, cout, H, W, device="cuda"))
w = torch.randn(cout, cin, 3, 3, device="cuda") / (cin * 9) ** 0.5
pd = ops.PackedDgrad(w, 1)
dx = torch.empty_like(x)
for _ in range(3):
    ops.conv_wgrad(x, dy, cin, 3)
    ops.conv_fused(dy, pd, pd.cin, H, W, act="none", out_pre=dx)
torch.cuda.synchronize()
print("ok")
