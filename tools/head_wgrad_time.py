import sys, torch
sys.path.insert(0, "/root/repo/boosting-nerv_b200")
from bnerv_b200 import ops
x = ops.nchw_to_c8(torch.randn(1, 112, 1080, 1920, device="cuda"))
dy = ops.nchw_to_c8(torch.randn(1, 3, 1080, 1920, device="cuda"))
for _ in range(3): ops.conv_wgrad(x, dy, 112, 3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.conv_wgrad(x, dy, 112, 3)
e1.record(); torch.cuda.synchronize()
print("head wgrad 112->3 @1080p: %.3f ms" % (e0.elapsed_time(e1) / 10))
