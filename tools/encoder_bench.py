"""ConvNeXt encoder forward at 1080p (HNeRV-Boost L, batch 1): native f32 kernels (bnerv_b200.encoder) vs the torch module on the
same GPU (cuDNN/cuBLAS, TF32 allowed = PyTorch's default for convs, and strict f32), CUDA-event timed, plus per-kernel times.
Usage: python tools/encoder_bench.py [--ncu-once]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "boosting-nerv_b200"))
import bench  # noqa: E402
from bnerv_b200.encoder import convnext_forward  # noqa: E402


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def main():
    model, args = bench.build_model("hnerv_l")
    enc = model.encoder.cuda().eval()
    x = torch.rand(1, 3, 1080, 1920, device="cuda")
    if "--ncu-once" in sys.argv:
        with torch.no_grad():
            convnext_forward(enc, x)
        torch.cuda.synchronize()
        return
    with torch.no_grad():
        t_nat, got = timed(lambda: convnext_forward(enc, x))
        torch.backends.cudnn.allow_tf32 = True
        t_tf32, _ = timed(lambda: enc(x))
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        t_f32, ref = timed(lambda: enc(x))
        xc = x.contiguous(memory_format=torch.channels_last)
        encl = enc.to(memory_format=torch.channels_last)
        t_cl, _ = timed(lambda: encl(xc))
    rel = float((got - ref).abs().max() / ref.abs().max())
    macs = 0
    H, W, cin = 1080, 1920, 3
    for down, stage in zip(enc.downsample_layers, enc.stages):
        conv = down[0] if isinstance(down[0], torch.nn.Conv2d) else down[1]
        s, c = conv.kernel_size[0], conv.out_channels
        H, W = H // s, W // s
        macs += H * W * c * cin * s * s + len(stage) * H * W * (49 * c + 8 * c * c)
        cin = c
    print(f"encoder forward @1080x1920, batch 1, {2 * macs / 1e9:.2f} GFLOP: native {t_nat:.3f} ms ({2 * macs / t_nat / 1e9:.1f} TFLOP/s f32), "
          f"torch cuDNN TF32-allowed {t_tf32:.3f} ms, torch strict f32 {t_f32:.3f} ms (channels_last {t_cl:.3f} ms); max-rel vs torch f32 {rel:.2e}")


if __name__ == "__main__":
    main()
