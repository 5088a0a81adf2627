"""Thin torch-tensor wrappers over the C-ABI (one function per entry point of include/bnerv_b200.h).

torch is used for device memory and the current stream only; every computation below happens inside
libbnerv_b200.so.  All functions raise on CPU tensors — there is no fallback.
"""
import ctypes
import os

import torch

from . import _capi
from ._capi import ACT_CODES, lib, check, ptr


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("bnerv_b200 ops need CUDA tensors (no CPU path)")


def require_current_device(dev):
    """The C-ABI launches on the CURRENT device's stream (one process per GPU, as torchrun / DistributedDataParallel run
    it); a tensor on another device would be dereferenced on the wrong GPU, so refuse it loudly."""
    if dev.type != "cuda":
        raise RuntimeError("bnerv_b200 needs CUDA tensors (no CPU path)")
    cur = torch.cuda.current_device()
    idx = cur if dev.index is None else dev.index
    if idx != cur:
        raise RuntimeError(f"bnerv_b200: tensors live on cuda:{idx} but the current device is cuda:{cur}; call "
                           f"torch.cuda.set_device({idx}) first (one process per GPU)")


def round_up(x, m):
    return (x + m - 1) // m * m


def c8_shape(B, C, H, W):
    return (B, round_up(C, 16) // 8, H, W, 8)


def nchw_to_c8(x):
    """[B,C,H,W] f32 -> C8 f16 tensor of shape [B, Cp/8, H, W, 8]."""
    _need_cuda(x)
    assert x.dtype == torch.float32 and x.dim() == 4
    x = x.contiguous()
    B, C, H, W = x.shape
    y = torch.empty(c8_shape(B, C, H, W), dtype=torch.float16, device=x.device)
    check("bnerv_nchw_to_c8", lib.bnerv_nchw_to_c8(ptr(x), B, C, H, W, ptr(y), _stream()))
    return y


def c8_to_nchw(y, C):
    _need_cuda(y)
    assert y.dtype == torch.float16 and y.dim() == 5 and y.is_contiguous()
    B, _, H, W, _ = y.shape
    x = torch.empty((B, C, H, W), dtype=torch.float32, device=y.device)
    check("bnerv_c8_to_nchw", lib.bnerv_c8_to_nchw(ptr(y), B, C, H, W, ptr(x), _stream()))
    return x


def pixel_shuffle(x, s):
    _need_cuda(x)
    assert x.dtype == torch.float32 and x.dim() == 4
    x = x.contiguous()
    B, Cs, H, W = x.shape
    assert Cs % (s * s) == 0
    C = Cs // (s * s)
    y = torch.empty((B, C, H * s, W * s), dtype=torch.float32, device=x.device)
    check("bnerv_pixel_shuffle", lib.bnerv_pixel_shuffle(ptr(x), B, C, H, W, s, ptr(y), _stream()))
    return y


def split_weight(weight):
    """OIHW f32 weight -> the weight of the same conv over a split input map [hi | lo | hi] (bnerv_conv_fused_split): input
    channel blocks [W_hi ; W_hi ; W_lo], each zero-padded to a multiple of 16 channels; both halves are exact in f16."""
    w = weight.detach().float()
    pad = round_up(w.shape[1], 16) - w.shape[1]
    if pad:
        w = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, pad))
    hi = w.half().float()
    lo = (w - hi).half().float()
    return torch.cat([hi, hi, lo], dim=1).contiguous()


class PackedConv:
    """Kernel-layout copy of one conv's effective weight: f16 [taps][Kp/8][Np][8] + f32 bias [Np].  split_in: the conv reads a
    split map (3 * round_up(Cin, 16) input channels, see split_weight)."""

    def __init__(self, weight, bias, s=1, split_in=False):
        _need_cuda(weight, bias)
        self.split_in = split_in
        if split_in:
            weight = split_weight(weight)
        co_s2, self.cin, k, k2 = weight.shape
        assert k == k2 and co_s2 % (s * s) == 0
        self.k, self.s, self.cout = k, s, co_s2 // (s * s)
        dev = weight.device
        self.w = torch.empty(lib.bnerv_packed_weight_numel(self.cout, self.cin, k, s), dtype=torch.float16, device=dev)
        self.b = torch.empty(lib.bnerv_packed_bias_numel(self.cout, s), dtype=torch.float32, device=dev)
        self.repack(weight, bias)

    def repack(self, weight, bias):
        w = split_weight(weight) if self.split_in and weight.shape[1] != self.cin else weight.detach().contiguous().float()
        b = None if bias is None else bias.detach().contiguous().float()
        check("bnerv_pack_conv_weight",
              lib.bnerv_pack_conv_weight(ptr(w), ptr(b), self.cout, self.cin, self.k, self.s, ptr(self.w), ptr(self.b), _stream()))


    _CODE_DTYPES = {torch.int8: 1, torch.int16: 2, torch.int32: 4}

    def repack_codes(self, w_codes, w_scale, b_codes=None, b_scale=None):
        """Ingest integer codes + scale(s) directly (bnerv_pack_conv_weight_q): bit-identical to
        repack(w_codes.float() * w_scale, b_codes.float() * b_scale), without materialising the dequantised tensors."""
        _need_cuda(w_codes, w_scale, b_codes, b_scale)
        nb = self._CODE_DTYPES.get(w_codes.dtype)
        if nb is None or (b_codes is not None and b_codes.dtype != w_codes.dtype):
            raise TypeError("codes must be int8 / int16 / int32 tensors of one dtype")
        assert tuple(w_codes.shape) == (self.cout * self.s * self.s, self.cin, self.k, self.k)
        w_codes, w_scale = w_codes.contiguous(), w_scale.detach().contiguous().float()
        if b_codes is not None:
            b_codes, b_scale = b_codes.contiguous(), b_scale.detach().contiguous().float()
        check("bnerv_pack_conv_weight_q",
              lib.bnerv_pack_conv_weight_q(ptr(w_codes), ptr(w_scale), int(w_scale.numel() > 1), ptr(b_codes), ptr(b_scale),
                                           int(b_scale is not None and b_scale.numel() > 1), nb, self.cout, self.cin, self.k,
                                           self.s, ptr(self.w), ptr(self.b), _stream()))


class PackedHead:
    """Weights of a 3x3 conv to <= 3 channels in the head kernel's layout (bnerv_pack_head_weight): one [Kp][32] slab
    whose column tap*Cout + c holds W[c][:, tap]; the bias stays a raw f32 vector."""

    def __init__(self, weight, bias, split_in=False):
        _need_cuda(weight, bias)
        self.split_in = split_in
        if split_in:
            weight = split_weight(weight)
        self.cout, self.cin, k, k2 = weight.shape
        assert k == 3 and k2 == 3 and self.cout <= 3
        self.k, self.s = 3, 1
        dev = weight.device
        self.w = torch.empty(32 * round_up(self.cin, 16), dtype=torch.float16, device=dev)
        self.b = torch.zeros(self.cout, dtype=torch.float32, device=dev)
        self.repack(weight, bias)

    def repack(self, weight, bias):
        w = split_weight(weight) if self.split_in and weight.shape[1] != self.cin else weight.detach().contiguous().float()
        check("bnerv_pack_head_weight", lib.bnerv_pack_head_weight(ptr(w), self.cout, self.cin, ptr(self.w), _stream()))
        if bias is None:
            self.b.zero_()
        else:
            self.b.copy_(bias.detach())


class PackedHead1:
    """A 1x1 head conv to <= 4 channels for bnerv_head_conv1: the raw f32 weights [Cout, Cin] and bias (copies, so that
    a captured graph reads stable storage)."""

    def __init__(self, weight, bias):
        _need_cuda(weight, bias)
        self.cout, self.cin, k, k2 = weight.shape
        assert k == 1 and k2 == 1 and self.cout <= 4
        self.k, self.s = 1, 1
        self.w = torch.empty((self.cout, self.cin), dtype=torch.float32, device=weight.device)
        self.b = torch.zeros(self.cout, dtype=torch.float32, device=weight.device)
        self.repack(weight, bias)

    def repack(self, weight, bias):
        self.w.copy_(weight.detach().reshape(self.cout, self.cin))
        if bias is None:
            self.b.zero_()
        else:
            self.b.copy_(bias.detach())


# When set to a list, every conv_fused launch appends (algorithmic_flops, start_event, end_event) — used by
# bench.py to time the dominant kernel inside the timed region on the launching stream.
TIMING = None


def conv_flops(B, cin, cout, k, s, H, W):
    """Algorithmic FLOPs of one conv launch: 2*Cout*Cin*k*k*Hout*Wout with unpadded channels."""
    return 2.0 * B * (cout * s * s) * cin * k * k * H * W


def conv_fused(x_c8, pc, cin, H, W, act="none", resid=None, g1p=None, beta=None, out_pre=None, out_aff=None,
               out_nchw=None, out_deriv=None, split=0):
    """Launch the tcgen05 fused conv.  Output tensors are caller-provided (see bnerv_conv_fused[_ex]).
    split: bit 0 = C8 outputs are split maps, bit 1 = resid is a split map (bnerv_conv_fused_split)."""
    _need_cuda(x_c8)
    B = x_c8.shape[0]
    assert cin == pc.cin
    if TIMING is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    if isinstance(pc, PackedHead1):
        if out_nchw is None or resid is not None or g1p is not None or out_pre is not None or out_deriv is not None:
            raise ValueError("the head kernel writes the NCHW f32 image only")
        check("bnerv_head_conv1", lib.bnerv_head_conv1(ptr(x_c8), B, cin, H, W, ptr(pc.w), ptr(pc.b), pc.cout, ACT_CODES[act],
                                                       ptr(out_nchw), _stream()))
    elif isinstance(pc, PackedHead):
        if out_nchw is None or resid is not None or g1p is not None or out_pre is not None or out_deriv is not None:
            raise ValueError("the head kernel writes the NCHW f32 image only")
        check("bnerv_head_conv3", lib.bnerv_head_conv3(ptr(x_c8), B, cin, H, W, ptr(pc.w), ptr(pc.b), pc.cout, ACT_CODES[act],
                                                       ptr(out_nchw), _stream()))
    elif split:
        if out_deriv is not None:
            raise ValueError("the split form is a decode-path form (no derivative output)")
        check("bnerv_conv_fused_split",
              lib.bnerv_conv_fused_split(ptr(x_c8), B, cin, H, W, ptr(pc.w), ptr(pc.b), pc.cout, pc.k, pc.s, ACT_CODES[act],
                                         ptr(resid), ptr(g1p), ptr(beta), ptr(out_pre), ptr(out_aff), ptr(out_nchw), split,
                                         _stream()))
    elif out_deriv is None:
        check("bnerv_conv_fused",
              lib.bnerv_conv_fused(ptr(x_c8), B, cin, H, W, ptr(pc.w), ptr(pc.b), pc.cout, pc.k, pc.s, ACT_CODES[act],
                                   ptr(resid), ptr(g1p), ptr(beta), ptr(out_pre), ptr(out_aff), ptr(out_nchw), _stream()))
    else:
        check("bnerv_conv_fused_ex",
              lib.bnerv_conv_fused_ex(ptr(x_c8), B, cin, H, W, ptr(pc.w), ptr(pc.b), pc.cout, pc.k, pc.s, ACT_CODES[act],
                                      ptr(resid), ptr(g1p), ptr(beta), ptr(out_pre), ptr(out_aff), ptr(out_nchw),
                                      ptr(out_deriv), _stream()))
    if TIMING is not None:
        e1.record()
        TIMING.append((conv_flops(B, cin, pc.cout, pc.k, pc.s, H, W), e0, e1, (cin, pc.cout, pc.k, pc.s, H, W, act)))


def conv_fused_f32(x, weight, bias, s=1, act="none", resid=None, g1p=None, beta=None, want_pre=True):
    """f32 CUDA-core fused conv on NCHW / OIHW.  Returns (out_pre or None, out_aff or None)."""
    _need_cuda(x, weight)
    x, weight = x.contiguous(), weight.contiguous()
    B, cin, H, W = x.shape
    co_s2, cin_w, k, _ = weight.shape
    assert cin_w == cin
    cout = co_s2 // (s * s)
    out_pre = torch.empty((B, cout, H * s, W * s), dtype=torch.float32, device=x.device) if want_pre else None
    out_aff = torch.empty((B, cout, H * s, W * s), dtype=torch.float32, device=x.device) if g1p is not None else None
    ldg = 0 if g1p is None else g1p.shape[-1]
    check("bnerv_conv_fused_f32",
          lib.bnerv_conv_fused_f32(ptr(x), B, cin, H, W, ptr(weight), ptr(bias), cout, k, s, ACT_CODES[act], ptr(resid),
                                   ptr(g1p), ptr(beta), ldg, ptr(out_pre), ptr(out_aff), _stream()))
    return out_pre, out_aff


def linear_act(x, weight, bias, act="none"):
    """y = act(W x + b); x [B,Cin] f32, weight [Cout,Cin(,1,1)] f32."""
    _need_cuda(x, weight)
    x = x.contiguous()
    B, cin = x.shape
    cout = weight.shape[0]
    assert weight.numel() == cout * cin and weight.is_contiguous()
    y = torch.empty((B, cout), dtype=torch.float32, device=x.device)
    check("bnerv_linear_act", lib.bnerv_linear_act(ptr(x), B, cin, ptr(weight), ptr(bias), cout, ACT_CODES[act], ptr(y), _stream()))
    return y


def _linear_problems(specs, B, need_x):
    """specs: two dicts(x=, w=, b=, act=, y=, y_c8=, hw=) -> (ctypes array, tensors kept alive)."""
    arr = (_capi.LinearProblem * 2)()
    keep = []
    for k, sp in enumerate(specs):
        w = sp["w"]
        cout = w.shape[0]
        cin = w.numel() // cout
        assert w.is_cuda and w.dtype == torch.float32 and w.is_contiguous()
        p = arr[k]
        if need_x:
            x = sp["x"].contiguous()
            assert x.dtype == torch.float32 and tuple(x.shape) == (B, cin)
            p.x = x.data_ptr()
            keep.append(x)
        b = sp.get("b")
        p.w, p.bias = w.data_ptr(), (b.data_ptr() if b is not None else None)
        y, y_c8 = sp.get("y"), sp.get("y_c8")
        p.y = y.data_ptr() if y is not None else None
        p.y_c8 = y_c8.data_ptr() if y_c8 is not None else None
        p.Cin, p.Cout, p.act, p.hw = cin, cout, ACT_CODES[sp["act"]], int(sp.get("hw") or 0)
        keep += [w, b, y, y_c8]
    return arr, keep


def pe_linear_pair(t, bases, specs):
    """First layers of two MLPs on the position encoding of t ([B] f32) in one launch (bnerv_pe_linear_pair)."""
    _need_cuda(t, bases)
    assert t.dtype == torch.float32 and bases.dtype == torch.float32 and t.dim() == 1
    t, bases = t.contiguous(), bases.contiguous()
    arr, keep = _linear_problems(specs, t.shape[0], False)
    check("bnerv_pe_linear_pair", lib.bnerv_pe_linear_pair(ptr(t), t.shape[0], ptr(bases), bases.numel(), arr, _stream()))


def linear_pair(specs, B):
    """Two independent y = act(W x + b) layers in one launch (bnerv_linear_pair); outputs are caller-provided."""
    arr, keep = _linear_problems(specs, B, True)
    check("bnerv_linear_pair", lib.bnerv_linear_pair(arr, B, _stream()))


def conv_stream(x_c8, pc, cin, H, W, act="none", resid=None, g1p=None, beta=None, out_pre=None, out_aff=None):
    """bnerv_conv_stream: conv_fused for 3x3 / stride 1 / equal padded widths of 32 or 48 channels in the row-streaming form.
    Returns True when launched, None for shapes outside that class (nothing launched)."""
    _need_cuda(x_c8)
    if not isinstance(pc, PackedConv) or pc.k != 3 or pc.s != 1 or pc.cin != cin:
        return None
    rc = lib.bnerv_conv_stream(ptr(x_c8), x_c8.shape[0], cin, H, W, ptr(pc.w), ptr(pc.b), pc.cout, ACT_CODES[act], ptr(resid), ptr(g1p),
                               ptr(beta), ptr(out_pre), ptr(out_aff), _stream())
    if rc == _capi.E_UNSUPPORTED:
        return None
    check("bnerv_conv_stream", rc)
    return True


def upconv_stream(x_c8, up, cin, H, W, act, g0p, beta0, x0, u):
    """bnerv_upconv_stream: a 17..32-channel block's 3x3 up-conv (no PixelShuffle) + activation + TAT affine in the row-streaming
    form; writes x0 and u.  None = unsupported shape (nothing launched)."""
    _need_cuda(x_c8, x0, u)
    if up.k != 3 or up.s != 1 or not isinstance(up, PackedConv) or up.cin != cin:
        return None
    rc = lib.bnerv_upconv_stream(ptr(x_c8), x_c8.shape[0], cin, H, W, ptr(up.w), ptr(up.b), up.cout, ACT_CODES[act], ptr(g0p), ptr(beta0),
                                 ptr(x0), ptr(u), _stream())
    if rc == _capi.E_UNSUPPORTED:
        return None
    check("bnerv_upconv_stream", rc)
    return x0


def resblock_head_fused(u_c8, x0_c8, c0, c1, C, H, W, act_inner, g1p, beta1, head, img, head_act="tanh01"):
    """bnerv_resblock_stream_head: the last block's ResBlock_SFT half + the 1x1 head conv + OutImg in one kernel; `head` is a
    PackedHead1, `img` the [B, Cout, H, W] f32 output.  None = unsupported shape (nothing launched)."""
    _need_cuda(u_c8, x0_c8, img)
    assert isinstance(head, PackedHead1) and head.cin == C and img.dtype == torch.float32 and img.is_contiguous()
    B = u_c8.shape[0]
    rc = lib.bnerv_resblock_stream_head(ptr(u_c8), ptr(x0_c8), B, C, H, W, ptr(c0.w), ptr(c0.b), ptr(c1.w), ptr(c1.b),
                                        ACT_CODES[act_inner], ptr(g1p), ptr(beta1), ptr(head.w), ptr(head.b), head.cout,
                                        ACT_CODES[head_act], ptr(img), _stream())
    if rc == _capi.E_UNSUPPORTED:
        return None
    check("bnerv_resblock_stream_head", rc)
    return img


def nerv_block_head_fused(x_c8, up, c0, c1, cin, H, W, g0p, beta0, g1p, beta1, head, img, head_act="tanh01"):
    """bnerv_nerv_block_stream_head: a whole sin / GELU NeRVBlock (3x3 up-conv, no PixelShuffle, <= 16 channels) + the 1x1 head
    conv + OutImg in one kernel.  None = unsupported shape (nothing launched)."""
    _need_cuda(x_c8, img)
    assert isinstance(head, PackedHead1) and head.cin == up.cout and img.dtype == torch.float32 and img.is_contiguous()
    if up.k != 3 or up.s != 1:
        return None
    B = x_c8.shape[0]
    rc = lib.bnerv_nerv_block_stream_head(ptr(x_c8), B, cin, H, W, ptr(up.w), ptr(up.b), ptr(c0.w), ptr(c0.b), ptr(c1.w), ptr(c1.b),
                                          up.cout, ptr(g0p), ptr(beta0), ptr(g1p), ptr(beta1), ptr(head.w), ptr(head.b), head.cout,
                                          ACT_CODES[head_act], ptr(img), _stream())
    if rc == _capi.E_UNSUPPORTED:
        return None
    check("bnerv_nerv_block_stream_head", rc)
    return img


class SftTable:
    """Device-side array of bnerv_sft_layer descriptors + the g1p/beta output tables for a batch size."""

    def __init__(self, layers, B, device):
        """layers: list of (ws0, bs0, ws1, bs1, wh0, bh0, wh1, bh1) f32 CUDA tensors."""
        self.B = B
        self.keep = layers
        self.C = [l[2].shape[0] for l in layers]
        self.Cp = [round_up(c, 16) for c in self.C]
        self.ch_t = layers[0][0].shape[1]
        offs, tot = [], 0
        for cp in self.Cp:
            offs.append(tot)
            tot += B * cp
        self.g1p_all = torch.empty(max(tot, 1), dtype=torch.float32, device=device)
        self.beta_all = torch.empty(max(tot, 1), dtype=torch.float32, device=device)
        self.g1p = [self.g1p_all[o:o + B * cp].view(B, cp) for o, cp in zip(offs, self.Cp)]
        self.beta = [self.beta_all[o:o + B * cp].view(B, cp) for o, cp in zip(offs, self.Cp)]
        arr = (_capi.SftLayer * len(layers))()
        for i, l in enumerate(layers):
            for name, t in zip(("ws0", "bs0", "ws1", "bs1", "wh0", "bh0", "wh1", "bh1"), l):
                assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
                setattr(arr[i], name, t.data_ptr())
            arr[i].g1p, arr[i].beta = self.g1p[i].data_ptr(), self.beta[i].data_ptr()
            arr[i].C, arr[i].Cp = self.C[i], self.Cp[i]
        host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        self.desc = host.to(device)

    def run(self, e):
        """e: [B, ch_t] f32.  Fills every layer's g1p/beta with one launch."""
        _need_cuda(e)
        e = e.contiguous()
        assert e.shape == (self.B, self.ch_t)
        check("bnerv_sft_affine", lib.bnerv_sft_affine(ptr(self.desc), len(self.C), ptr(e), self.B, self.ch_t, _stream()))


# ------------------------------------------------------------------------------------------------
# backward of the cascade (include/bnerv_b200.h, "Backward of the cascade")
# ------------------------------------------------------------------------------------------------
class PackedDgrad:
    """Packed weights of the conv that computes dL/dx from the un-shuffled dL/dy: a conv with Cin' = s*s*Cout_p,
    Cout' = Cin, the same k, no shuffle, zero bias (weights transposed and tap-flipped)."""

    def __init__(self, weight, s=1):
        _need_cuda(weight)
        co_s2, cin_fwd, k, _ = weight.shape
        self.k, self.s = k, 1
        self.s_fwd = s
        self.cout_fwd = co_s2 // (s * s)
        self.cin = s * s * round_up(self.cout_fwd, 16)      # K of the dgrad conv (already padded)
        self.cout = cin_fwd                                 # N of the dgrad conv
        dev = weight.device
        self.w = torch.empty(lib.bnerv_packed_weight_numel(self.cout, self.cin, k, 1), dtype=torch.float16, device=dev)
        self.b = torch.zeros(lib.bnerv_packed_bias_numel(self.cout, 1), dtype=torch.float32, device=dev)
        self.repack(weight)

    def repack(self, weight):
        w = weight.detach().contiguous().float()
        check("bnerv_pack_conv_weight_dgrad",
              lib.bnerv_pack_conv_weight_dgrad(ptr(w), self.cout_fwd, self.cout, self.k, self.s_fwd, ptr(self.w), _stream()))


def head_bwd(dimg, img, scale):
    """dz (C8 f16) = S * dimg * d(tanh01)/dz; fills scale = [S, 1/S] (2-float CUDA tensor) on the device."""
    _need_cuda(dimg, img, scale)
    dimg, img = dimg.contiguous().float(), img.contiguous().float()
    B, C, H, W = img.shape
    dz = torch.empty(c8_shape(B, C, H, W), dtype=torch.float16, device=img.device)
    scratch = torch.empty(1, dtype=torch.float32, device=img.device)
    check("bnerv_head_bwd", lib.bnerv_head_bwd(ptr(dimg), ptr(img), B, C, H, W, ptr(scratch), ptr(scale), ptr(dz), _stream()))
    return dz


def conv_wgrad(x_c8, dy_c8, cin, k):
    """acc [k*k][M_p][Cin_p] f32 = sum_pixels dy (x) shifted x (scaled by the loss scale)."""
    _need_cuda(x_c8, dy_c8)
    B, _, H, W, _ = x_c8.shape
    m_p = dy_c8.shape[1] * 8
    assert dy_c8.shape[0] == B and dy_c8.shape[2] == H and dy_c8.shape[3] == W
    acc = torch.zeros(lib.bnerv_wgrad_acc_numel(m_p, cin, k), dtype=torch.float32, device=x_c8.device)
    check("bnerv_conv_wgrad", lib.bnerv_conv_wgrad(ptr(x_c8), ptr(dy_c8), B, cin, H, W, m_p, k, ptr(acc), _stream()))
    return acc


def wgrad_finalize(acc, cout, cin, k, s, inv_scale, grad=None, out=None):
    """-> grad OIHW f32 [cout*s*s, cin, k, k]: a new tensor, written into `out`, or accumulated into `grad`."""
    accumulate = grad is not None
    if grad is None:
        grad = out if out is not None else torch.empty((cout * s * s, cin, k, k), dtype=torch.float32, device=acc.device)
    assert grad.is_contiguous() and grad.dtype == torch.float32 and grad.numel() == cout * s * s * cin * k * k
    check("bnerv_wgrad_finalize", lib.bnerv_wgrad_finalize(ptr(acc), cout, cin, k, s, ptr(inv_scale), int(accumulate), ptr(grad), _stream()))
    return grad


def bias_finalize(acc, cout, s, inv_scale, out=None):
    grad = out if out is not None else torch.empty(cout * s * s, dtype=torch.float32, device=acc.device)
    assert grad.is_contiguous() and grad.dtype == torch.float32 and grad.numel() == cout * s * s
    check("bnerv_bias_finalize", lib.bnerv_bias_finalize(ptr(acc), cout, s, ptr(inv_scale), 0, ptr(grad), _stream()))
    return grad


def channel_sum(x_c8, per_b=False):
    _need_cuda(x_c8)
    B, G, H, W, _ = x_c8.shape
    out = torch.zeros((B, G * 8) if per_b else (G * 8,), dtype=torch.float32, device=x_c8.device)
    check("bnerv_channel_sum", lib.bnerv_channel_sum(ptr(x_c8), B, G * 8, H, W, int(per_b), ptr(out), _stream()))
    return out


def resblock_mid_bwd(dw, v, dact, g1p, C):
    """-> (dc0 C8, dG [B,Cp], dB [B,Cp], dbias0 [Cp]) all reductions scaled by the loss scale."""
    B, G, H, W, _ = dw.shape
    dev = dw.device
    dc0 = torch.empty_like(dw)
    red = torch.zeros((2 * B + 1, G * 8), dtype=torch.float32, device=dev)
    dG, dB, db = red[:B], red[B:2 * B], red[2 * B]
    check("bnerv_resblock_mid_bwd", lib.bnerv_resblock_mid_bwd(ptr(dw), ptr(v), ptr(dact), ptr(g1p), B, C, H, W, ptr(dc0),
                                                               ptr(dG), ptr(dB), ptr(db), _stream()))
    return dc0, dG, dB, db


def block_front_bwd(du, dout, x0, dact, g0p, C, want_dy_sums=False):
    """-> (dy C8 at the block's output resolution, dG, dB, dbias1, channel sums of dy or None)."""
    B, G, H, W, _ = du.shape
    dev = du.device
    dy = torch.empty_like(du)
    red = torch.zeros((2 * B + 2, G * 8), dtype=torch.float32, device=dev)       # one memset: dG | dB | dbias1 | dy sums
    dG, dB, db, dsum = red[:B], red[B:2 * B], red[2 * B], red[2 * B + 1]
    check("bnerv_block_front_bwd", lib.bnerv_block_front_bwd(ptr(du), ptr(dout), ptr(x0), ptr(dact), ptr(g0p), B, C, H, W,
                                                             ptr(dy), ptr(dG), ptr(dB), ptr(db),
                                                             ptr(dsum) if want_dy_sums else None, _stream()))
    return dy, dG, dB, db, (dsum if want_dy_sums else None)


def unshuffle_c8(src, C, s, want_sums=False):
    """PixelShuffle(s) transposed: [B][Cp/8][H*s][W*s][8] -> [B][s*s*Cp/8][H][W][8].
    want_sums: also return the channel sums of the result (f32 [s*s*Cp]) from the same pass (s = 2, 3) or a second one."""
    if s == 1:
        return (src, channel_sum(src)) if want_sums else src
    B, G, Hs, Ws, _ = src.shape
    H, W = Hs // s, Ws // s
    dst = torch.empty((B, s * s * G, H, W, 8), dtype=torch.float16, device=src.device)
    fused = want_sums and s in (2, 3) and not os.environ.get("BNERV_NO_FUSED_SUMS")
    sums = torch.zeros(s * s * G * 8, dtype=torch.float32, device=src.device) if fused else None
    check("bnerv_unshuffle_c8", lib.bnerv_unshuffle_c8(ptr(src), B, C, H, W, s, ptr(dst), ptr(sums), _stream()))
    if want_sums:
        return dst, (sums if fused else channel_sum(dst))
    return dst


def frame_metrics(img, gt):
    """-> f32 [B, 3] on the device: (mse, mae, psnr) per frame (hnerv_utils.py:338-341, 400-403), no host sync."""
    _need_cuda(img, gt)
    assert img.shape == gt.shape and img.dtype == torch.float32 and gt.dtype == torch.float32
    img, gt = img.contiguous(), gt.contiguous()
    B = img.shape[0]
    n = img.numel() // B
    scratch = torch.empty(lib.bnerv_frame_metrics_scratch_doubles(B), dtype=torch.float64, device=img.device)
    out = torch.empty((B, 3), dtype=torch.float32, device=img.device)
    check("bnerv_frame_metrics", lib.bnerv_frame_metrics(ptr(img), ptr(gt), B, n, ptr(scratch), ptr(out), _stream()))
    return out


def nerv_block_fwd(x_c8, up, c0, c1, cin, H, W, act_up, act_inner, g0p, beta0, g1p, beta1):
    """One NeRVBlock through bnerv_nerv_block_fwd.  up / c0 / c1: PackedConv.  Returns (out, x0) C8 f16."""
    _need_cuda(x_c8)
    B = x_c8.shape[0]
    C, s = up.cout, up.s
    mk = lambda: torch.empty(c8_shape(B, C, H * s, W * s), dtype=torch.float16, device=x_c8.device)
    x0, u, wmap, out = mk(), mk(), mk(), mk()
    check("bnerv_nerv_block_fwd",
          lib.bnerv_nerv_block_fwd(ptr(x_c8), B, cin, H, W, ptr(up.w), ptr(up.b), up.k, s, ACT_CODES[act_up], ptr(c0.w), ptr(c0.b),
                                   ptr(c1.w), ptr(c1.b), C, ACT_CODES[act_inner], ptr(g0p), ptr(beta0), ptr(g1p), ptr(beta1),
                                   ptr(x0), ptr(u), ptr(wmap), ptr(out), _stream()))
    return out, x0


def nerv_block_fused(x_c8, up, c0, c1, cin, H, W, act_up, act_inner, g0p, beta0, g1p, beta1, out=None, form="tile"):
    """One NeRVBlock as ONE kernel: form "tile" = bnerv_nerv_block_fused (region-tiled, <= 48 channels), "stream" =
    bnerv_nerv_block_stream (row-streaming with the A operand in tensor memory, <= 16 channels).  Returns out C8 f16, or
    None when the shape is outside that kernel's range (BNERV_E_UNSUPPORTED: nothing was launched)."""
    _need_cuda(x_c8)
    B = x_c8.shape[0]
    C, s = up.cout, up.s
    if out is None:
        out = torch.empty(c8_shape(B, C, H * s, W * s), dtype=torch.float16, device=x_c8.device)
    if TIMING is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    fn = lib.bnerv_nerv_block_stream if form == "stream" else lib.bnerv_nerv_block_fused
    rc = fn(ptr(x_c8), B, cin, H, W, ptr(up.w), ptr(up.b), up.k, s, ACT_CODES[act_up], ptr(c0.w), ptr(c0.b), ptr(c1.w), ptr(c1.b),
            C, ACT_CODES[act_inner], ptr(g0p), ptr(beta0), ptr(g1p), ptr(beta1), ptr(out), _stream())
    if rc == _capi.E_UNSUPPORTED:
        return None
    check("bnerv_nerv_block_" + ("stream" if form == "stream" else "fused"), rc)
    if TIMING is not None:
        e1.record()
        fl = conv_flops(B, cin, C, up.k, s, H, W) + 2 * conv_flops(B, C, C, 3, 1, H * s, W * s)
        TIMING.append((fl, e0, e1, (cin, C, up.k, s, H, W, "block")))
    return out


def resblock_fused(u_c8, x0_c8, c0, c1, C, H, W, act_inner, g1p, beta1, out=None, form="tile"):
    """ResBlock_SFT as ONE kernel (bnerv_resblock_fused / bnerv_resblock_stream): out = x0 + conv3(act(conv3(u))*g1p + beta1).
    None = unsupported."""
    _need_cuda(u_c8, x0_c8)
    B = u_c8.shape[0]
    if out is None:
        out = torch.empty(c8_shape(B, C, H, W), dtype=torch.float16, device=u_c8.device)
    if TIMING is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    fn = lib.bnerv_resblock_stream if form == "stream" else lib.bnerv_resblock_fused
    rc = fn(ptr(u_c8), ptr(x0_c8), B, C, H, W, ptr(c0.w), ptr(c0.b), ptr(c1.w), ptr(c1.b), ACT_CODES[act_inner], ptr(g1p),
            ptr(beta1), ptr(out), _stream())
    if rc == _capi.E_UNSUPPORTED:
        return None
    check("bnerv_resblock_" + ("stream" if form == "stream" else "fused"), rc)
    if TIMING is not None:
        e1.record()
        TIMING.append((2 * conv_flops(B, C, C, 3, 1, H, W), e0, e1, (C, C, 3, 1, H, W, "resblock")))
    return out
