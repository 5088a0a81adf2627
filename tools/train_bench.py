"""Training-step time (forward + loss + backward, batch 1 like the reference's -b 1) of a preset at full size:
native sm_100a path vs torch autograd (cuDNN, TF32 allowed = the reference's default GPU arithmetic, and strict fp32).
Also times the individual backward kernels at the largest layer shape.
Usage: python tools/train_bench.py [config] [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "boosting-nerv_b200"))
import bench  # noqa: E402
from bnerv_b200 import ops  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "enerv_m"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda")


def timed(fn, n, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def model_step(mode, tf32):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    model, args = bench.build_model(cfg)
    model = model.to(dev).train()
    model.train_backend = mode
    is_h = args.model == "HNeRV_Boost"
    fh, fw = [int(v) for v in args.fc_hw.split("_")]
    emb = torch.rand(1, 16, fh, fw, device=dev) if is_h else None
    t = torch.tensor([0.5], dtype=torch.float64, device=dev)
    with torch.no_grad():
        model.eval()
        shape = (model.forward_decoder(emb, t) if is_h else model(t))[0].shape
        model.train()
    target = torch.rand(shape, device=dev)

    def fwd():
        img = (model.forward_decoder(emb, t) if is_h else model(t))[0]
        return ((img - target) ** 2).mean()

    def step():
        model.zero_grad(set_to_none=True)
        fwd().backward()

    torch.cuda.reset_peak_memory_stats()
    ms = timed(step, steps)
    with torch.enable_grad():
        ms_f = timed(lambda: fwd(), steps)
    mem = torch.cuda.max_memory_allocated() / 2 ** 30
    del model
    torch.cuda.empty_cache()
    return ms, ms_f, mem


print(torch.cuda.get_device_name(0), cfg, flush=True)
gflop = bench.ALG_GFLOP.get(cfg)
for mode, tf32, label in (("b200", False, "native sm_100a fwd+bwd"), ("torch", True, "torch autograd, cuDNN TF32 allowed (reference default)"),
                          ("torch", False, "torch autograd, strict fp32")):
    try:
        ms, ms_f, mem = model_step(mode, tf32)
        extra = f", {3 * gflop / ms:.0f} TFLOP/s algorithmic (3x forward FLOPs)" if gflop else ""
        print(f"{label:58s}: {ms:8.2f} ms/step (forward+loss {ms_f:7.2f} ms), peak {mem:5.1f} GiB{extra}", flush=True)
    except Exception as ex:
        print(f"{label:58s}: FAILED {type(ex).__name__}: {str(ex)[:200]}", flush=True)
        torch.cuda.empty_cache()

# HNeRV: the reference trains THROUGH the ConvNeXt encoder (model(frame) -> encoder -> decoder, train_nerv_all.py:342);
# the encoder stays a torch module, so this is what a full step costs with it in the loop
def _torch_fusion10_freq(p, target):
    """The shipped loss (scripts/regression/UVG/hnerv_boost.sh: --loss Fusion10_freq) as the reference computes it:
    pytorch_msssim's torch formulation (grouped separable conv2d per moment map) + cuFFT."""
    import torch.nn.functional as F
    from bnerv_b200 import losses

    def blur(x, win):
        c = x.shape[1]
        w = win.view(1, 1, 1, -1).repeat(c, 1, 1, 1)
        return F.conv2d(F.conv2d(x, w.transpose(2, -1), groups=c), w, groups=c)

    co = torch.arange(11, dtype=p.dtype, device=p.device) - 5
    win = torch.exp(-(co ** 2) / 4.5)
    win = win / win.sum()
    x, y, mcs = p, target, []
    for i in range(5):
        mu1, mu2 = blur(x, win), blur(y, win)
        s11, s22, s12 = blur(x * x, win) - mu1 * mu1, blur(y * y, win) - mu2 * mu2, blur(x * y, win) - mu1 * mu2
        cs = (2 * s12 + 9e-4) / (s11 + s22 + 9e-4)
        ss = ((2 * mu1 * mu2 + 1e-4) / (mu1 * mu1 + mu2 * mu2 + 1e-4)) * cs
        if i < 4:
            mcs.append(torch.relu(cs.flatten(2).mean(-1)))
            pad = [sz % 2 for sz in x.shape[2:]]
            x, y = F.avg_pool2d(x, 2, padding=pad), F.avg_pool2d(y, 2, padding=pad)
    vals = torch.stack(mcs + [torch.relu(ss.flatten(2).mean(-1))], dim=0)
    ms = torch.prod(vals ** torch.tensor(losses.WEIGHTS, device=p.device).view(-1, 1, 1), dim=0).mean(1)
    l1 = F.l1_loss(p, target, reduction="none").flatten(1).mean(1)
    return (60 * (0.7 * l1 + 0.3 * (1 - ms)) + losses._freq_l1(p, target)).mean()


if cfg.startswith("hnerv"):
    from bnerv_b200 import losses as _losses
    for mode, tf32, label in (("b200", True, "full step incl. torch encoder: native decoder"), ("torch", True, "full step incl. torch encoder: torch decoder (TF32)"),
                              ("b200+loss", True, "encoder + native decoder + native Fusion10_freq loss"),
                              ("torch+loss", True, "encoder + torch decoder + torch Fusion10_freq loss")):
        torch.backends.cudnn.allow_tf32 = tf32
        model, args = bench.build_model(cfg)
        model = model.to(dev).train()
        model.train_backend = mode.split("+")[0]
        frame = torch.rand(1, 3, 1080, 1920, device=dev)
        t = torch.tensor([0.5], dtype=torch.float64, device=dev)

        def step():
            model.zero_grad(set_to_none=True)
            img = model(frame, norm_idx=t)[0]
            if mode == "b200+loss":
                _losses.loss_fn(img, frame, "Fusion10_freq").backward()
            elif mode == "torch+loss":
                _torch_fusion10_freq(img, frame).backward()
            else:
                ((img - frame) ** 2).mean().backward()

        try:
            print(f"{label:58s}: {timed(step, steps):8.2f} ms/step", flush=True)
        except Exception as ex:
            print(f"{label:58s}: FAILED {type(ex).__name__}: {str(ex)[:200]}", flush=True)
        del model
        torch.cuda.empty_cache()
    torch.backends.cudnn.allow_tf32 = False

# individual backward kernels at the biggest layer of the preset
shapes = {"hnerv_l": (112, 112, 1080, 1920), "hnerv_m": (89, 89, 1080, 1920), "enerv_m": (21, 21, 1080, 1920)}.get(cfg, (12, 12, 720, 1280))
cin, cout, H, W = shapes
torch.manual_seed(0)
x = ops.nchw_to_c8(torch.randn(1, cin, H, W, device=dev))
dy = ops.nchw_to_c8(torch.randn(1, cout, H, W, device=dev))
w = torch.randn(cout, cin, 3, 3, device=dev) / (cin * 9) ** 0.5
pd = ops.PackedDgrad(w, 1)
dx = torch.empty_like(x)
fl = 2.0 * cout * cin * 9 * H * W
ms = timed(lambda: ops.conv_wgrad(x, dy, cin, 3), 10)
print(f"wgrad  {cin}->{cout} 3x3 @{H}x{W}: {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s (incl. zeroing the f32 accumulator)")
ms = timed(lambda: ops.conv_fused(dy, pd, pd.cin, H, W, act="none", out_pre=dx), 10)
print(f"dgrad  {cout}->{cin} 3x3 @{H}x{W}: {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s")
g = torch.ones(1, ops.round_up(cout, 16), device=dev)
ms = timed(lambda: ops.resblock_mid_bwd(dy, dy, dy, g, cout), 10)
nbytes = 4 * dy.numel() * 2
print(f"mid_bwd   @{H}x{W}x{cout}: {ms:.3f} ms  {nbytes / ms / 1e6:.0f} GB/s (3 maps read, 1 written)")
ms = timed(lambda: ops.block_front_bwd(dy, dy, dy, dy, g, cout, want_dy_sums=True), 10)
nbytes = 5 * dy.numel() * 2
print(f"front_bwd @{H}x{W}x{cout}: {ms:.3f} ms  {nbytes / ms / 1e6:.0f} GB/s (4 maps read, 1 written)")
ms = timed(lambda: ops.channel_sum(dy), 10)
print(f"channel_sum @{H}x{W}x{cout}: {ms:.3f} ms  {dy.numel() * 2 / ms / 1e6:.0f} GB/s")
