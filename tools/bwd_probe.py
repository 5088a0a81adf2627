"""Bring-up probe for the backward kernels: every case compares one native kernel with torch autograd on f16-rounded
operands and prints max-relative errors.  Usage: python tools/bwd_probe.py [case ...]"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "boosting-nerv_b200"))
from bnerv_b200 import ops  # noqa: E402

dev = torch.device("cuda")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
h16 = lambda t: t.half().float()


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def unshuffle_ref(dy, s):
    """NCHW [B, C, H*s, W*s] -> [B, s*s*Cp, H, W] in the un-shuffled order m = (i*s+j)*Cp + c."""
    B, C, Hs, Ws = dy.shape
    cp = ops.round_up(C, 16)
    H, W = Hs // s, Ws // s
    out = torch.zeros(B, s * s * cp, H, W, device=dy.device)
    for i in range(s):
        for j in range(s):
            out[:, (i * s + j) * cp:(i * s + j) * cp + C] = dy[:, :, i::s, j::s]
    return out


def to_c8_padded(x, cp_total):
    """NCHW f32 whose channel count is already a multiple of 16 -> C8."""
    assert x.shape[1] == cp_total
    return ops.nchw_to_c8(x.contiguous())


def case_wgrad(B, cin, cout, H, W, k, s):
    torch.manual_seed(0)
    x = h16(torch.randn(B, cin, H, W, device=dev))
    w = torch.randn(cout * s * s, cin, k, k, device=dev, requires_grad=True)
    y = F.conv2d(x, w, None, 1, (k - 1) // 2)
    if s > 1:
        y = F.pixel_shuffle(y, s)
    dy = h16(torch.randn_like(y))
    y.backward(dy)
    cp = ops.round_up(cout, 16)
    dyu = to_c8_padded(unshuffle_ref(dy, s), s * s * cp)
    acc = ops.conv_wgrad(ops.nchw_to_c8(x), dyu, cin, k)
    one = torch.ones(1, device=dev)
    g = ops.wgrad_finalize(acc, cout, cin, k, s, one)
    torch.cuda.synchronize()
    return rel(g, w.grad)


def case_dgrad(B, cin, cout, H, W, k, s):
    torch.manual_seed(0)
    x = torch.randn(B, cin, H, W, device=dev, requires_grad=True)
    w = torch.randn(cout * s * s, cin, k, k, device=dev) / (cin * k * k) ** 0.5
    y = F.conv2d(x, h16(w), None, 1, (k - 1) // 2)
    if s > 1:
        y = F.pixel_shuffle(y, s)
    dy = h16(torch.randn_like(y))
    y.backward(dy)
    cp = ops.round_up(cout, 16)
    dyu_nchw = unshuffle_ref(dy, s)
    # the unshuffle kernel against the reference permutation (bit-exact)
    dy_c8 = ops.nchw_to_c8(dy)
    un = ops.unshuffle_c8(dy_c8, cout, s)
    dyu = to_c8_padded(dyu_nchw, s * s * cp)
    exact = bool((un == dyu).all())
    pd = ops.PackedDgrad(w, s)
    dx = torch.empty(ops.c8_shape(B, cin, H, W), dtype=torch.float16, device=dev)
    ops.conv_fused(un, pd, pd.cin, H, W, act="none", out_pre=dx)
    torch.cuda.synchronize()
    return rel(ops.c8_to_nchw(dx, cin), x.grad), exact


def case_deriv(act, B=1, cin=24, cout=20, H=20, W=36, k=3, s=1):
    torch.manual_seed(0)
    x = torch.randn(B, cin, H, W, device=dev)
    w = torch.randn(cout * s * s, cin, k, k, device=dev) * (2.0 / (cin * k * k) ** 0.5)
    b = torch.randn(cout * s * s, device=dev) * 0.1
    cp = ops.round_up(cout, 16)
    g1p = torch.zeros(B, cp, device=dev); beta = torch.zeros(B, cp, device=dev)
    g1p[:, :cout] = 1 + 0.3 * torch.randn(B, cout, device=dev); beta[:, :cout] = 0.3 * torch.randn(B, cout, device=dev)
    pc = ops.PackedConv(w, b, s)
    shp = ops.c8_shape(B, cout, H * s, W * s)
    pre, aff, der = [torch.full(shp, float("nan"), dtype=torch.float16, device=dev) for _ in range(3)]
    ops.conv_fused(ops.nchw_to_c8(x), pc, cin, H, W, act=act, g1p=g1p, beta=beta, out_pre=pre, out_aff=aff, out_deriv=der)
    torch.cuda.synchronize()
    z = F.conv2d(h16(x), h16(w), b, 1, (k - 1) // 2)
    if s > 1:
        z = F.pixel_shuffle(z, s)
    z = z.double().requires_grad_(True)
    a = torch.sin(z) if act == "sin" else F.gelu(z)
    a.sum().backward()
    return rel(ops.c8_to_nchw(pre, cout), a.detach().float()), rel(ops.c8_to_nchw(der, cout), z.grad.float())


def case_elementwise(B=2, C=21, H=18, W=30):
    torch.manual_seed(0)
    cp = ops.round_up(C, 16)
    mk = lambda: h16(torch.randn(B, C, H, W, device=dev))
    du, dout, x0, dact, dw, v = mk(), mk(), mk(), mk(), mk(), mk()
    g = torch.zeros(B, cp, device=dev); g[:, :C] = 1 + 0.3 * torch.randn(B, C, device=dev)
    c8 = ops.nchw_to_c8
    dy, dG, dB, db1, _ = ops.block_front_bwd(c8(du), c8(dout), c8(x0), c8(dact), g, C)
    gb = g[:, :C, None, None]
    e = [rel(ops.c8_to_nchw(dy, C), (dout + du * gb) * dact), rel(dG[:, :C], (du * x0).sum((2, 3))), rel(dB[:, :C], du.sum((2, 3))),
         rel(db1[:C], dout.sum((0, 2, 3)))]
    dc0, dG1, dB1, db0 = ops.resblock_mid_bwd(c8(dw), c8(v), c8(dact), g, C)
    ref = dw * gb * dact
    e += [rel(ops.c8_to_nchw(dc0, C), ref), rel(dG1[:, :C], (dw * v).sum((2, 3))), rel(dB1[:, :C], dw.sum((2, 3))),
          rel(db0[:C], h16(ref).sum((0, 2, 3)))]
    e += [rel(ops.channel_sum(c8(du))[:C], du.sum((0, 2, 3))), rel(ops.channel_sum(c8(du), True)[:, :C], du.sum((2, 3)))]
    # head
    img = torch.rand(B, 3, H, W, device=dev)
    dimg = torch.randn(B, 3, H, W, device=dev) * 1e-7
    scale = torch.zeros(2, device=dev)
    dz = ops.head_bwd(dimg, img, scale)
    torch.cuda.synchronize()
    S = scale[0].item()
    e += [rel(ops.c8_to_nchw(dz, 3) / S, dimg * 2 * img * (1 - img)), S]
    return e


CASES = {
    "elementwise": lambda: case_elementwise(),
    "deriv_sin": lambda: case_deriv("sin"),
    "deriv_gelu": lambda: case_deriv("gelu"),
    "deriv_sin_s2": lambda: case_deriv("sin", s=2),
    "wgrad_k1_min": lambda: case_wgrad(1, 16, 16, 8, 16, 1, 1),
    "wgrad_k3_min": lambda: case_wgrad(1, 16, 16, 8, 16, 3, 1),
    "wgrad_k3_odd": lambda: case_wgrad(2, 21, 43, 19, 37, 3, 1),
    "wgrad_k3_wide": lambda: case_wgrad(1, 176, 162, 24, 40, 3, 1),
    "wgrad_k3_s2": lambda: case_wgrad(1, 43, 21, 10, 18, 3, 2),
    "wgrad_k1_s5": lambda: case_wgrad(1, 40, 33, 9, 16, 1, 5),
    "wgrad_many_jobs": lambda: case_wgrad(1, 200, 60, 6, 8, 3, 5),
    "dgrad_k3": lambda: case_dgrad(2, 21, 43, 19, 37, 3, 1),
    "dgrad_k1": lambda: case_dgrad(1, 16, 30, 9, 16, 1, 1),
    "dgrad_k3_s2": lambda: case_dgrad(1, 43, 21, 10, 18, 3, 2),
    "dgrad_k3_s3_wide": lambda: case_dgrad(1, 345, 172, 12, 20, 3, 3),
    "dgrad_k1_s5": lambda: case_dgrad(1, 40, 33, 9, 16, 1, 5),
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    print(torch.cuda.get_device_name(0), "SWAP" if os.environ.get("BNERV_WGRAD_SWAP") else "", flush=True)
    for n in names:
        try:
            print(f"{n:20s}", CASES[n](), flush=True)
        except Exception as ex:
            print(f"{n:20s} FAILED: {type(ex).__name__}: {str(ex)[:300]}", flush=True)
            try:
                torch.cuda.synchronize()
            except Exception as ex2:
                print("context dead:", str(ex2)[:200], flush=True)
                sys.exit(3)
