"""Native forward of HNeRV_Boost's ConvNeXt encoder (inference): one bnerv_convnext_stage_fwd call per stage on the
module's own f32 parameter storage (no packed copies, so optimiser steps / load_state_dict need no invalidation).

Replaces ConvNeXt.forward (model_blocks.py:314-320) as HNeRV_Boost.forward_encoder runs it under torch.no_grad()
(model_hnerv.py:230-234; evaluate() calls it per frame, train_nerv_all.py:482-486).  With autograd enabled the torch module
is used (models.HNeRV_Boost.forward_encoder decides); CPU tensors raise - there is no fallback here.
"""
import ctypes

import torch

from ._capi import ConvNextBlock, ConvNextStage, check, lib, ptr
from .ops import _need_cuda, _stream, require_current_device


def _p(t):
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise TypeError("encoder parameters must be contiguous float32")
    return ctypes.c_void_p(t.data_ptr())


def _stage_desc(down, blocks, first):
    """-> (ConvNextStage, keep-alive list).  Stage 0 is Conv2d -> LayerNorm, the others LayerNorm -> Conv2d."""
    conv, ln = (down[0], down[1]) if first else (down[1], down[0])
    s = conv.kernel_size[0]
    if conv.kernel_size != (s, s) or conv.stride != (s, s) or conv.padding != (0, 0) or conv.groups != 1 or conv.dilation != (1, 1):
        raise ValueError("encoder down-sampling conv must be kernel = stride, no padding")
    arr = (ConvNextBlock * max(1, len(blocks)))()
    for i, b in enumerate(blocks):
        if b.gamma is None or tuple(b.dwconv.kernel_size) != (7, 7) or b.dwconv.groups != b.dwconv.in_channels:
            raise ValueError("encoder block must be the ConvNeXt block with layer scale (model_blocks.py:234-244)")
        arr[i] = ConvNextBlock(_p(b.dwconv.weight), _p(b.dwconv.bias), _p(b.norm.weight), _p(b.norm.bias), _p(b.pwconv1.weight),
                               _p(b.pwconv1.bias), _p(b.pwconv2.weight), _p(b.pwconv2.bias), _p(b.gamma))
    st = ConvNextStage()
    if first:
        st.ln_in_w = st.ln_in_b = None
        st.ln_out_w, st.ln_out_b = _p(ln.weight), _p(ln.bias)
    else:
        st.ln_in_w, st.ln_in_b = _p(ln.weight), _p(ln.bias)
        st.ln_out_w = st.ln_out_b = None
    st.down_w, st.down_b = _p(conv.weight), _p(conv.bias)
    st.blocks = arr
    st.n_blocks, st.Cin, st.Cout, st.s = len(blocks), conv.in_channels, conv.out_channels, s
    return st, arr


def convnext_forward(encoder, x):
    """encoder: layers.ConvNeXt (or the reference's, same attribute names); x: [B, 3, H, W] f32 CUDA -> [B, C_last, h, w]."""
    _need_cuda(x)
    require_current_device(x.device)
    if x.dtype != torch.float32 or x.dim() != 4:
        raise TypeError("convnext_forward expects a [B, C, H, W] float32 frame batch")
    cur, nchw = x.contiguous(), 1
    B, _, H, W = x.shape
    for i, (down, stage) in enumerate(zip(encoder.downsample_layers, encoder.stages)):
        st, keep = _stage_desc(down, list(stage), i == 0)
        if (cur.shape[1] if nchw else cur.shape[3]) != st.Cin:
            raise ValueError(f"encoder stage {i}: expected {st.Cin} input channels")
        Ho, Wo = H // st.s, W // st.s
        y = torch.empty((B, Ho, Wo, st.Cout), dtype=torch.float32, device=x.device)
        work = torch.empty(lib.bnerv_convnext_stage_work_floats(B, H, W, st.s, st.Cout), dtype=torch.float32, device=x.device)
        check("bnerv_convnext_stage_fwd",
              lib.bnerv_convnext_stage_fwd(ctypes.byref(st), ptr(cur), nchw, B, H, W, ptr(y), ptr(work), _stream()))
        del keep
        cur, nchw, H, W = y, 0, Ho, Wo
    out = torch.empty((B, cur.shape[3], H, W), dtype=torch.float32, device=x.device)
    check("bnerv_nhwc_to_nchw", lib.bnerv_nhwc_to_nchw(ptr(cur), B, H, W, cur.shape[3], ptr(out), _stream()))
    return out
