"""GPU bring-up probe: runs conv_fused cases of increasing complexity against a torch reference with
fp16-rounded operands and prints error statistics + timings.  Usage: python tools/gpu_probe.py [case ...]"""
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "boosting-nerv_b200"))
from bnerv_b200 import ops  # noqa: E402

dev = torch.device("cuda")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
h16 = lambda t: t.half().float()


def ref_conv(x, w, b, s, act, resid, g1p, beta):
    k = w.shape[-1]
    y = F.conv2d(h16(x), h16(w), b, 1, (k - 1) // 2)
    if s > 1:
        y = F.pixel_shuffle(y, s)
    y = {"none": lambda v: v, "sin": torch.sin, "gelu": F.gelu, "tanh01": lambda v: torch.tanh(v) * 0.5 + 0.5}[act](y)
    if resid is not None:
        y = y + resid
    aff = None if g1p is None else y * g1p[:, :y.shape[1], None, None] + beta[:, :y.shape[1], None, None]
    return y, aff


def run_case(name, B, cin, cout, H, W, k, s, act="none", resid=False, affine=False, nchw=False, time_it=False, scale=1.0, pre=True,
             head=False):
    torch.manual_seed(0)
    x = torch.randn(B, cin, H, W, device=dev)
    w = torch.randn(cout * s * s, cin, k, k, device=dev) * (scale / (cin * k * k) ** 0.5)
    b = torch.randn(cout * s * s, device=dev) * 0.1
    cp = ops.round_up(cout, 16)
    r = h16(torch.randn(B, cout, H * s, W * s, device=dev)) if resid else None
    g1p = beta = None
    if affine:
        g1p = torch.zeros(B, cp, device=dev); beta = torch.zeros(B, cp, device=dev)
        g1p[:, :cout] = 1 + 0.3 * torch.randn(B, cout, device=dev); beta[:, :cout] = 0.3 * torch.randn(B, cout, device=dev)
    pc = ops.PackedHead(w, b) if head else ops.PackedConv(w, b, s)
    xc = ops.nchw_to_c8(x)
    rc = ops.nchw_to_c8(r) if resid else None
    out_pre = torch.full(ops.c8_shape(B, cout, H * s, W * s), float("nan"), dtype=torch.float16, device=dev) if pre else None
    out_aff = torch.full(ops.c8_shape(B, cout, H * s, W * s), float("nan"), dtype=torch.float16, device=dev) if affine else None
    out_n = torch.full((B, cout, H * s, W * s), float("nan"), device=dev) if nchw else None
    ops.conv_fused(xc, pc, cin, H, W, act=act, resid=rc, g1p=g1p, beta=beta, out_pre=out_pre, out_aff=out_aff, out_nchw=out_n)
    torch.cuda.synchronize()
    y_ref, a_ref = ref_conv(x, w, b, s, act, r, g1p, beta)
    err = 0.0
    msg = f"{name:28s}"
    if pre:
        got = ops.c8_to_nchw(out_pre, cout)
        err = (got - y_ref).abs().max().item() / y_ref.abs().max().item()
        msg += f" pre rel-err {err:.2e}"
    pad_ok = True
    if pre and cp != cout:
        pad_ok = bool((out_pre.view(B, cp // 8, H * s, W * s, 8).permute(0, 1, 4, 2, 3).reshape(B, cp, H * s, W * s)[:, cout:] == 0).all())
        msg += f" pad0={pad_ok}"
    if affine:
        ga = ops.c8_to_nchw(out_aff, cout)
        msg += f" aff rel-err {(ga - a_ref).abs().max().item() / a_ref.abs().max().item():.2e}"
    if nchw:
        msg += f" nchw rel-err {(out_n - y_ref).abs().max().item() / y_ref.abs().max().item():.2e}"
    if time_it:
        for _ in range(3):
            ops.conv_fused(xc, pc, cin, H, W, act=act, resid=rc, g1p=g1p, beta=beta, out_pre=out_pre, out_aff=out_aff, out_nchw=out_n)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        n = 10
        e0.record()
        for _ in range(n):
            ops.conv_fused(xc, pc, cin, H, W, act=act, resid=rc, g1p=g1p, beta=beta, out_pre=out_pre, out_aff=out_aff, out_nchw=out_n)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        fl = 2.0 * B * cout * s * s * cin * k * k * H * W
        msg += f" | {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s (algorithmic)"
    print(msg, flush=True)
    return err


CASES = {
    "k1_min":      dict(B=1, cin=16, cout=16, H=16, W=16, k=1, s=1),
    "k3_min":      dict(B=1, cin=16, cout=16, H=16, W=16, k=3, s=1),
    "k3_2ksteps":  dict(B=1, cin=32, cout=32, H=16, W=16, k=3, s=1),
    "k3_multi":    dict(B=2, cin=48, cout=64, H=40, W=50, k=3, s=1),
    "k3_odd":      dict(B=1, cin=13, cout=27, H=17, W=33, k=3, s=1, act="sin", affine=True),
    "k3_s2":       dict(B=1, cin=27, cout=13, H=9, W=7, k=3, s=2, act="sin", affine=True),
    "k3_s5":       dict(B=1, cin=15, cout=15, H=9, W=16, k=3, s=5, act="sin", affine=True),
    "k1_s5":       dict(B=2, cin=24, cout=21, H=4, W=3, k=1, s=5, act="sin", affine=True),
    "k3_resid":    dict(B=1, cin=43, cout=43, H=30, W=50, k=3, s=1, act="none", resid=True),
    "k3_gelu":     dict(B=1, cin=43, cout=43, H=30, W=50, k=3, s=1, act="gelu", affine=True),
    "head_k3":     dict(B=1, cin=21, cout=3, H=36, W=64, k=3, s=1, act="tanh01", nchw=True),
    "head_k1":     dict(B=1, cin=12, cout=3, H=36, W=64, k=1, s=1, act="tanh01", nchw=True),
    "headk_min":   dict(B=1, cin=16, cout=3, H=16, W=32, k=3, s=1, act="tanh01", nchw=True, pre=False, head=True),
    "headk_odd":   dict(B=2, cin=21, cout=3, H=37, W=53, k=3, s=1, act="tanh01", nchw=True, pre=False, head=True),
    "headk_c2":    dict(B=1, cin=40, cout=2, H=20, W=70, k=3, s=1, act="none", nchw=True, pre=False, head=True),
    "L_headk":     dict(B=1, cin=112, cout=3, H=1080, W=1920, k=3, s=1, act="tanh01", nchw=True, pre=False, head=True, time_it=True),
    "n144":        dict(B=1, cin=135, cout=135, H=64, W=64, k=3, s=1, act="gelu", affine=True),
    "big_in":      dict(B=1, cin=16, cout=16, H=16, W=16, k=3, s=1, scale=300.0, act="sin"),
    # HNeRV-L shapes, timed
    "L_dec8_c0":   dict(B=1, cin=112, cout=112, H=1080, W=1920, k=3, s=1, act="gelu", affine=True, pre=False, time_it=True),
    "L_dec8_up":   dict(B=1, cin=112, cout=112, H=1080, W=1920, k=3, s=1, act="sin", affine=True, time_it=True),
    "L_dec8_c1":   dict(B=1, cin=112, cout=112, H=1080, W=1920, k=3, s=1, act="none", resid=True, time_it=True),
    "L_dec7_up":   dict(B=1, cin=135, cout=112, H=540, W=960, k=3, s=2, act="sin", affine=True, time_it=True),
    "L_dec6_c0":   dict(B=1, cin=135, cout=135, H=540, W=960, k=3, s=1, act="gelu", affine=True, pre=False, time_it=True),
    "L_dec5_up":   dict(B=1, cin=162, cout=135, H=270, W=480, k=3, s=2, act="sin", affine=True, time_it=True),
    "L_dec4_c1":   dict(B=1, cin=162, cout=162, H=270, W=480, k=3, s=1, act="none", resid=True, time_it=True),
    "L_dec3_up":   dict(B=1, cin=194, cout=162, H=135, W=240, k=3, s=2, act="sin", affine=True, time_it=True),
    "L_dec1_up":   dict(B=1, cin=280, cout=233, H=9, W=16, k=1, s=5, act="sin", affine=True, time_it=True),
    "L_head":      dict(B=1, cin=112, cout=3, H=1080, W=1920, k=3, s=1, act="tanh01", nchw=True, time_it=True),
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    print(torch.cuda.get_device_name(0), flush=True)
    for n in names:
        try:
            run_case(n, **CASES[n])
        except Exception as ex:  # keep going only if the context survived
            print(f"{n:28s} FAILED: {type(ex).__name__}: {str(ex)[:300]}", flush=True)
            try:
                torch.cuda.synchronize()
            except Exception as ex2:
                print("context dead:", str(ex2)[:200], flush=True)
                sys.exit(3)
