"""Pixel / structural losses of the reference's `loss_fn` (hnerv_utils.py:335-395) with the SSIM / MS-SSIM terms on the
device kernels of csrc/loss_ops.cu.

`ssim` / `ms_ssim` follow pytorch_msssim==0.2.1 (the reference's pinned dependency; algorithm restated in
oracle/msssim_oracle.py - parity unpinned, see there): per pyramid level one launch computes the per-(batch, channel)
means of the ssim and cs maps, the combination across levels is a handful of scalar torch ops under autograd, and the
backward pass is two launches per level (partial-derivative maps, then their Gaussian transpose-filter) chained through
the 2x2 average pools.  Only `pred` receives a gradient (the reference detaches the target, hnerv_utils.py:336).
The FFT terms of the '*_freq' losses stay on cuFFT through torch.
"""
import torch
import torch.nn.functional as F

from . import ops
from ._capi import check, lib, ptr

WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)
K1, K2 = 0.01, 0.03


def _stats_level(x, y, c1, c2):
    B, C, H, W = x.shape
    stats = torch.zeros((B * C, 2), dtype=torch.float64, device=x.device)
    check("bnerv_ssim_stats", lib.bnerv_ssim_stats(ptr(x), ptr(y), B * C, H, W, c1, c2, ptr(stats), ops._stream()))
    n = float((H - 10) * (W - 10))
    return (stats / n).float().view(B, C, 2)


def _pool(x):
    return F.avg_pool2d(x, kernel_size=2, padding=[s % 2 for s in x.shape[2:]])


def _unpool(g, shape):
    """Transpose of _pool: every input pixel receives a quarter of its output pixel's gradient."""
    H, W = shape[-2:]
    ih = (torch.arange(H, device=g.device) + H % 2) // 2
    iw = (torch.arange(W, device=g.device) + W % 2) // 2
    return g[..., ih, :][..., iw] * 0.25


class _SsimStats(torch.autograd.Function):
    """(pred, target) -> [levels, B, C, 2] per-channel means of the ssim and cs maps of every pyramid level."""

    @staticmethod
    def forward(ctx, pred, target, levels, data_range):
        if not pred.is_cuda:
            raise RuntimeError("bnerv_b200 losses need CUDA tensors (no CPU path)")
        c1, c2 = (K1 * data_range) ** 2, (K2 * data_range) ** 2
        x, y = pred.detach().float().contiguous(), target.detach().float().contiguous()
        pyr, out = [], []
        ops.require_current_device(pred.device)
        for lv in range(levels):
            pyr.append((x, y))
            out.append(_stats_level(x, y, c1, c2))
            if lv < levels - 1:
                x, y = _pool(x).contiguous(), _pool(y).contiguous()
        ctx.pyr, ctx.c = pyr, (c1, c2)
        return torch.stack(out)

    @staticmethod
    def backward(ctx, g):
        c1, c2 = ctx.c
        g = g.contiguous().float()
        dx = None
        for lv in range(len(ctx.pyr) - 1, -1, -1):
            x, y = ctx.pyr[lv]
            B, C, H, W = x.shape
            gw = (g[lv] / float((H - 10) * (W - 10))).reshape(B * C, 2).contiguous()
            acc = dx is not None
            cur = _unpool(dx, x.shape).contiguous() if acc else torch.empty_like(x)
            scratch = torch.empty(lib.bnerv_ssim_scratch_floats(B * C, H, W), dtype=torch.float32, device=x.device)
            check("bnerv_ssim_grad", lib.bnerv_ssim_grad(ptr(x), ptr(y), B * C, H, W, c1, c2, ptr(gw), ptr(scratch), int(acc), ptr(cur),
                                                         ops._stream()))
            dx = cur
        ctx.pyr = None
        return dx, None, None, None


def ssim(pred, target, data_range=1.0, size_average=False):
    """pytorch_msssim.ssim as hnerv_utils.loss_fn calls it (:343-366): `nonnegative_ssim` keeps its default False in the
    pinned 0.2.1, so a channel with a negative mean SSIM is NOT clamped (only ms_ssim clamps, unconditionally) and still
    receives a gradient."""
    s = _SsimStats.apply(pred, target, 1, float(data_range))[0, :, :, 0]
    return s.mean() if size_average else s.mean(1)


def ms_ssim(pred, target, data_range=1.0, size_average=False):
    if min(pred.shape[-2:]) <= 10 * 2 ** 4:
        raise ValueError("image too small for a 5-level MS-SSIM (smaller side must exceed 160)")
    st = _SsimStats.apply(pred, target, len(WEIGHTS), float(data_range))          # [L, B, C, 2]
    w = torch.tensor(WEIGHTS, dtype=st.dtype, device=st.device)
    vals = torch.cat([torch.relu(st[:-1, :, :, 1]), torch.relu(st[-1:, :, :, 0])], dim=0)
    out = torch.prod(vals ** w.view(-1, 1, 1), dim=0)
    return out.mean() if size_average else out.mean(1)


def _freq_l1(pred, target):
    pf, tf = torch.fft.fft2(pred, dim=(-2, -1)), torch.fft.fft2(target, dim=(-2, -1))
    pf, tf = torch.stack([pf.real, pf.imag], -1), torch.stack([tf.real, tf.imag], -1)
    return F.l1_loss(pf, tf, reduction="none").flatten(1).mean(1)


def loss_fn(pred, target, loss_type="L2", batch_average=True):
    """hnerv_utils.loss_fn (:335-395) - same loss_type names and weights."""
    target = target.detach()
    l2 = lambda: F.mse_loss(pred, target, reduction="none").flatten(1).mean(1)
    l1 = lambda: F.l1_loss(pred, target, reduction="none").flatten(1).mean(1)
    s1 = lambda: 1 - ssim(pred, target, data_range=1, size_average=False)
    ms = lambda: 1 - ms_ssim(pred, target, data_range=1, size_average=False)
    table = {
        "L2": lambda: l2(), "L1": lambda: l1(), "SSIM": lambda: s1(),
        "Fusion1": lambda: 0.3 * l2() + 0.7 * s1(), "Fusion2": lambda: 0.3 * l1() + 0.7 * s1(),
        "Fusion3": lambda: 0.5 * l2() + 0.5 * s1(), "Fusion4": lambda: 0.5 * l1() + 0.5 * s1(),
        "Fusion5": lambda: 0.7 * l2() + 0.3 * s1(), "Fusion6": lambda: 0.7 * l1() + 0.3 * s1(),
        "Fusion7": lambda: 0.7 * l2() + 0.3 * l1(), "Fusion8": lambda: 0.5 * l2() + 0.5 * l1(),
        "Fusion9": lambda: 0.9 * l1() + 0.1 * s1(), "Fusion10": lambda: 0.7 * l1() + 0.3 * ms(),
        "Fusion11": lambda: 0.9 * l1() + 0.1 * ms(), "Fusion12": lambda: 0.8 * l1() + 0.2 * ms(),
        "Fusion10_freq": lambda: 60 * (0.7 * l1() + 0.3 * ms()) + _freq_l1(pred, target),
        "L1_freq": lambda: 60 * l1() + _freq_l1(pred, target),
        "L1_ssim_freq": lambda: 60 * (0.7 * l1() + 0.3 * s1()) + _freq_l1(pred, target),
    }
    if loss_type not in table:
        raise KeyError(f"unknown loss_type {loss_type!r}")
    loss = table[loss_type]()
    return loss.mean() if batch_average else loss
