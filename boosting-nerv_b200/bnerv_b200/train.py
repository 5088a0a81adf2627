"""Native training path of the block cascade: forward AND backward on the sm_100a kernels.

``cascade_train(engine, x, cond)`` is a ``torch.autograd.Function`` over the whole conv cascade
(model_blocks.py:34-46, 74-105; head model_blocks.py:57-63).  What stays in torch autograd is everything that is a
few kFLOP per frame: position encoding, the stem MLPs / transformer and the SFT (TAT) MLPs that turn the time
embedding into per-channel (scale, shift) — their outputs enter the Function as ordinary differentiable tensors, so
``loss.backward()`` (train_nerv_all.py:346) fills ``.grad`` of every parameter exactly as with the reference modules.

Forward (3 launches per block, as in decode) additionally writes act'(pre-activation) maps; every intermediate map
is kept for the backward pass.  Backward per block, from dL/dout:

    conv1 : wgrad(w, dout)                      dw  = dgrad(dout, W1)
    mid   : dc0 = dw * g1p * gelu'(c0)          dG1, dB1 (TAT grads), db0           [one element-wise + reduction pass]
    conv0 : wgrad(u, dc0)                       du  = dgrad(dc0, W0)
    front : dy = (dout + du*g0p) * sin'(y)      dG0, dB0, db1                        [one pass]
    up    : un-shuffle(dy); wgrad(x, dy_u); db_up; dx = dgrad(dy_u, Wup)  ->  dL/dout of the previous block

Gradient maps are C8 f16 scaled by one power-of-two loss scale chosen on the device (bnerv_head_bwd); all
reductions are f32.  Precision: f16 operands (11 significant bits) / f32 accumulation in both passes; parameter gradients
are checked against torch fp32 autograd in tests/test_gpu_train.py.

Both passes are pure stream work on static shapes, so they are captured ONCE per input shape into two CUDA graphs
(weight packing included: the pack kernels read the live parameter storage, so an optimiser step needs no re-capture)
and replayed: a training step is two graph launches plus the torch stems instead of ~400 host-driven launches.
``engine.train_graph = False`` runs the same bodies eagerly.  A second forward issued while the previous one still
awaits its backward (gradient accumulation over several frames) runs eagerly, because the graph owns one set of maps.
"""
import os
import weakref

import torch
import torch.nn.functional as F

from . import ops
from ._capi import check, lib, ptr
from .layers import effective_weight


# ------------------------------------------------------------------------------------------------
# argument layout: flat f32 buffers for the SFT tables and for all gradients
# ------------------------------------------------------------------------------------------------
class _Layout:
    """Order and offsets of everything the Function exchanges with autograd for one engine and batch size."""

    def __init__(self, eng, B):
        self.B = B
        self.convs = []
        for blk in eng.blocks:
            if blk.pre is not None:
                self.convs.append(blk.pre)
            self.convs += [blk.up, blk.c0, blk.c1]
        self.convs.append(eng.head)
        self.slot_pos = {id(s): i for i, s in enumerate(self.convs)}
        self.sft_layers = [l for blk in eng.blocks for l in blk.sfts]       # 2 per block: sft0, sft1
        self.sft_cols, off = [], 0                # column range of each SFT layer in the [B, Ctot] scale / shift tables
        for blk in eng.blocks:
            for _ in range(2):
                self.sft_cols.append((off, blk.cout))
                off += blk.cout
        self.sft_total = off
        self.row_layer = None                     # [Ctot] index of the SFT layer each column belongs to (built lazily)
        self.conv_off, off = [], 0                # (w offset, w numel, b offset, b numel)
        for s in self.convs:
            wn = s.m.weight.numel()
            bn = 0 if s.m.bias is None else s.m.bias.numel()
            self.conv_off.append((off, wn, off + wn, bn))
            off += wn + bn
        self.conv_numel = off

    def conv_tensors(self):
        out = []
        for s in self.convs:
            w, b = effective_weight(s.m)
            out += [w, b]
        return out

    def sft_view(self, table, i):
        off, c = self.sft_cols[i]
        return table[:, off:off + c]

    def sft_tables(self, cond):
        """All SFT (TAT) MLPs of the cascade at once (model_blocks.py:101-104): (scale, shift), each [B, Ctot], through
        a handful of batched torch ops instead of 4 tiny convs per layer - autograd still reaches every SFT parameter."""
        e = cond.flatten(1)
        dev = e.device
        if self.row_layer is None or self.row_layer.device != dev:
            self.row_layer = torch.cat([torch.full((c,), i, dtype=torch.long) for i, (_, c) in enumerate(self.sft_cols)]).to(dev)

        def eff(conv):
            w, b = effective_weight(conv)
            w = w.reshape(w.shape[0], -1)
            return w, (b if b is not None else torch.zeros(w.shape[0], dtype=w.dtype, device=w.device))

        def branch(first, second):
            p0 = [eff(getattr(l, first)) for l in self.sft_layers]
            p1 = [eff(getattr(l, second)) for l in self.sft_layers]
            w0, b0 = torch.stack([p[0] for p in p0]), torch.stack([p[1] for p in p0])      # [L, hid, ch_t], [L, hid]
            h = torch.relu(torch.einsum("lhc,bc->blh", w0, e) + b0)                       # [B, L, hid]
            w1, b1 = torch.cat([p[0] for p in p1]), torch.cat([p[1] for p in p1])          # [Ctot, hid], [Ctot]
            return (h[:, self.row_layer] * w1).sum(-1) + b1                                # [B, Ctot]

        return branch("SFT_scale_conv0", "SFT_scale_conv1"), branch("SFT_shift_conv0", "SFT_shift_conv1")

    def grad_views(self, flat, i):
        s = self.convs[i]
        wo, wn, bo, bn = self.conv_off[i]
        gw = flat[wo:wo + wn].view(s.m.weight.shape)
        gb = flat[bo:bo + bn] if bn else None
        return gw, gb


def _table(t, cp, plus_one):
    """[B, C] (scale or shift) -> contiguous f32 [B, Cp] table (scale + 1 like model_blocks.py:105)."""
    if plus_one:
        t = t + 1.0
    return F.pad(t, (0, cp - t.shape[1])).contiguous()


def _dgrad_pack(slot, force):
    w, _ = effective_weight(slot.m)
    key = (id(w), w._version, w.data_ptr())
    if slot.pd is None:
        slot.pd = ops.PackedDgrad(w, slot.s)
    elif force or slot.dkey != key:
        slot.pd.repack(w)
    slot.dkey = key
    return slot.pd


def _dgrad(slot, dy_u, H, W, force):
    """dL/dx (C8 f16 [B][Cin_p/8][H][W][8]) of a conv from its un-shuffled output gradient."""
    pd = _dgrad_pack(slot, force)
    dx = torch.empty(ops.c8_shape(dy_u.shape[0], slot.cin, H, W), dtype=torch.float16, device=dy_u.device)
    ops.conv_fused(dy_u, pd, pd.cin, H, W, act="none", out_pre=dx)
    return dx


# ------------------------------------------------------------------------------------------------
# the two bodies: pure stream work (capturable)
# ------------------------------------------------------------------------------------------------
def _forward_body(eng, lay, x, scale_all, shift_all, force):
    """x: [B,C,h,w] f32 contiguous; scale_all / shift_all: [B, Ctot] SFT tables.  Returns (img, first, state)."""
    B, C, h, w = x.shape
    dev = x.device
    c8 = lambda c, H, W: torch.empty(ops.c8_shape(B, c, H, W), dtype=torch.float16, device=dev)
    cur = ops.nchw_to_c8(x)
    cin, H, W = C, h, w
    saved = []
    for bi, blk in enumerate(eng.blocks):
        cp = ops.round_up(blk.cout, 16)
        g0p, b0 = _table(lay.sft_view(scale_all, 2 * bi), cp, True), _table(lay.sft_view(shift_all, 2 * bi), cp, False)
        g1p, b1 = _table(lay.sft_view(scale_all, 2 * bi + 1), cp, True), _table(lay.sft_view(shift_all, 2 * bi + 1), cp, False)
        rec = {"in": cur, "in_hw": (H, W), "g0p": g0p, "g1p": g1p}
        if blk.pre is not None:                 # E-NeRV stage 0: up-conv without activation feeds a 3x3 conv
            mid = c8(blk.pre.cout, H * blk.pre.s, W * blk.pre.s)
            ops.conv_fused(cur, blk.pre.packed(force), cin, H, W, act="none", out_pre=mid)
            cur, cin, H, W = mid, blk.pre.cout, H * blk.pre.s, W * blk.pre.s
            rec["mid"], rec["mid_hw"] = mid, (H, W)
        Ho, Wo = H * blk.up.s, W * blk.up.s
        x0, u, d0 = c8(blk.cout, Ho, Wo), c8(blk.cout, Ho, Wo), c8(blk.cout, Ho, Wo)
        ops.conv_fused(cur, blk.up.packed(force), cin, H, W, act=blk.act, g1p=g0p, beta=b0, out_pre=x0, out_aff=u, out_deriv=d0)
        v, wmap, d1 = c8(blk.cout, Ho, Wo), c8(blk.cout, Ho, Wo), c8(blk.cout, Ho, Wo)
        ops.conv_fused(u, blk.c0.packed(force), blk.cout, Ho, Wo, act=blk.inner_act, g1p=g1p, beta=b1, out_pre=v, out_aff=wmap,
                       out_deriv=d1)
        out = c8(blk.cout, Ho, Wo)
        ops.conv_fused(wmap, blk.c1.packed(force), blk.cout, Ho, Wo, act="none", resid=x0, out_pre=out)
        rec.update(x0=x0, u=u, d0=d0, v=v, w=wmap, d1=d1, hw=(Ho, Wo))
        saved.append(rec)
        cur, cin, H, W = out, blk.cout, Ho, Wo
    img = torch.empty((B, 3, H, W), dtype=torch.float32, device=dev)
    ops.conv_fused(cur, eng.head.packed(force), cin, H, W, act="tanh01", out_nchw=img)
    first = ops.c8_to_nchw(saved[1]["in"] if len(saved) > 1 else cur, eng.blocks[0].cout)   # block 0's output (callers keep [0])
    state = {"saved": saved, "last": cur, "last_hw": (H, W), "img": img, "x_shape": (B, C, h, w)}
    return img, first, state


def _backward_body(eng, lay, st, dimg, need_x, force):
    """Returns (gx NCHW f32 or None, gscale [B,Ctot], gshift [B,Ctot], gconv_flat): true (un-scaled) gradients."""
    saved = st["saved"]
    B, C, h, w = st["x_shape"]
    dev = dimg.device
    gscale = torch.empty((B, lay.sft_total), dtype=torch.float32, device=dev)
    gshift = torch.empty((B, lay.sft_total), dtype=torch.float32, device=dev)
    gconv = torch.empty(max(lay.conv_numel, 1), dtype=torch.float32, device=dev)

    def param_grads(slot, x_in, dy_u, inv, dbias_acc=None):
        gw, gb = lay.grad_views(gconv, lay.slot_pos[id(slot)])
        acc = ops.conv_wgrad(x_in, dy_u, slot.cin, slot.k)
        ops.wgrad_finalize(acc, slot.cout, slot.cin, slot.k, slot.s, inv, out=gw)
        if gb is not None:
            if dbias_acc is None:
                dbias_acc = ops.channel_sum(dy_u)
            ops.bias_finalize(dbias_acc, slot.cout, slot.s, inv, out=gb)

    scale = torch.zeros(2, dtype=torch.float32, device=dev)
    dz = ops.head_bwd(dimg, st["img"], scale)
    inv = scale[1:2]
    H, W = st["last_hw"]
    param_grads(eng.head, st["last"], dz, inv)
    dout = _dgrad(eng.head, dz, H, W, force)
    for bi in range(len(eng.blocks) - 1, -1, -1):
        blk, rec = eng.blocks[bi], saved[bi]
        Ho, Wo = rec["hw"]
        Cb = blk.cout
        # conv1 (input w, output gradient dout); its bias gradient = channel sums of dout, produced by the front pass
        gw1, gb1 = lay.grad_views(gconv, lay.slot_pos[id(blk.c1)])
        acc = ops.conv_wgrad(rec["w"], dout, blk.c1.cin, blk.c1.k)
        ops.wgrad_finalize(acc, Cb, blk.c1.cin, blk.c1.k, 1, inv, out=gw1)
        dw = _dgrad(blk.c1, dout, Ho, Wo, force)
        dc0, dG1, dB1, db0 = ops.resblock_mid_bwd(dw, rec["v"], rec["d1"], rec["g1p"], Cb)
        del dw
        # conv0 (input u, output gradient dc0)
        param_grads(blk.c0, rec["u"], dc0, inv, dbias_acc=db0)
        du = _dgrad(blk.c0, dc0, Ho, Wo, force)
        del dc0
        dy, dG0, dB0, db1, dy_sums = ops.block_front_bwd(du, dout, rec["x0"], rec["d0"], rec["g0p"], Cb,
                                                         want_dy_sums=(blk.up.s == 1))
        del du, dout
        if gb1 is not None:
            ops.bias_finalize(db1, Cb, 1, inv, out=gb1)
        for table, l, t in ((gscale, 2 * bi, dG0), (gshift, 2 * bi, dB0), (gscale, 2 * bi + 1, dG1), (gshift, 2 * bi + 1, dB1)):
            lay.sft_view(table, l).copy_(t[:, :Cb] * inv)
        # up-conv (input: block input or the E-NeRV pre-conv's output)
        up_in = rec.get("mid", rec["in"])
        Hi, Wi = rec.get("mid_hw", rec["in_hw"])
        if blk.up.s == 1:
            dyu = dy
        else:                                    # un-shuffle and the up-conv's bias sums in one pass
            dyu, dy_sums = ops.unshuffle_c8(dy, Cb, blk.up.s, want_sums=True)
        del dy
        param_grads(blk.up, up_in, dyu, inv, dbias_acc=dy_sums)
        first_conv = bi == 0 and blk.pre is None
        dprev = _dgrad(blk.up, dyu, Hi, Wi, force) if (not first_conv or need_x) else None
        del dyu
        if blk.pre is not None:
            Hi0, Wi0 = rec["in_hw"]
            dmu, dm_sums = ops.unshuffle_c8(dprev, blk.pre.cout, blk.pre.s, want_sums=True)
            param_grads(blk.pre, rec["in"], dmu, inv, dbias_acc=dm_sums)
            dprev = _dgrad(blk.pre, dmu, Hi0, Wi0, force) if (bi > 0 or need_x) else None
        dout = dprev
    gx = ops.c8_to_nchw(dout, C) * inv if need_x else None
    return gx, gscale, gshift, gconv


def _unpack_grads(lay, need, gx, gscale, gshift, gconv):
    """Gradient tuple in the Function's argument order (owner, x, scale_all, shift_all, *conv tensors)."""
    need_t = need[4:]
    g_conv = []
    for i in range(len(lay.convs)):
        gw, gb = lay.grad_views(gconv, i)
        g_conv += [gw if need_t[2 * i] else None, gb if (gb is not None and need_t[2 * i + 1]) else None]
    return (None, gx if need[1] else None, gscale if need[2] else None, gshift if need[3] else None) + tuple(g_conv)


# ------------------------------------------------------------------------------------------------
# eager Function
# ------------------------------------------------------------------------------------------------
class _CascadeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, owner, x, scale_all, shift_all, *tensors):
        eng, lay = owner
        img, first, st = _forward_body(eng, lay, x.detach().float().contiguous(), scale_all.detach().float(),
                                       shift_all.detach().float(), False)
        ctx.owner, ctx.st = owner, st
        ctx.mark_non_differentiable(first)
        return img, first

    @staticmethod
    def backward(ctx, dimg, _dfirst):
        eng, lay = ctx.owner
        if ctx.st is None:
            raise RuntimeError("bnerv_b200 native training: backward through the cascade a second time is not supported")
        grads = _backward_body(eng, lay, ctx.st, dimg.contiguous().float(), ctx.needs_input_grad[1], False)
        ctx.st = None
        return _unpack_grads(lay, ctx.needs_input_grad, *grads)


# ------------------------------------------------------------------------------------------------
# graph-captured Function
# ------------------------------------------------------------------------------------------------
class _Token:
    done = False


class _TrainGraph:
    """Forward and backward of the cascade captured for one (input shape, parameter storage) signature."""

    def __init__(self, eng, lay, x, scale_all, shift_all, need_x):
        self.need_x = need_x
        self.x = x.detach().float().contiguous().clone()
        self.scale, self.shift = scale_all.detach().float().clone(), shift_all.detach().float().clone()
        self.pending = None                     # weakref to the token of the forward that awaits its backward
        # eager warm-up of both passes: creates the packed-weight holders, sets kernel attributes, primes the allocator
        img, _, st = _forward_body(eng, lay, self.x, self.scale, self.shift, False)
        _backward_body(eng, lay, st, torch.zeros_like(img), need_x, False)
        del st, img
        torch.cuda.synchronize()
        torch.cuda.empty_cache()                # the warm-up's maps would otherwise stay cached next to the graph's private pool
        self.gf = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.gf, capture_error_mode="thread_local"):
            self.img, self.first, self.st = _forward_body(eng, lay, self.x, self.scale, self.shift, True)
        self.dimg = torch.zeros_like(self.img)
        self.gb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.gb, pool=self.gf.pool(), capture_error_mode="thread_local"):
            self.gx, self.gscale, self.gshift, self.gconv = _backward_body(eng, lay, self.st, self.dimg, need_x, True)

    def busy(self):
        tok = self.pending() if self.pending is not None else None
        return tok is not None and not tok.done


class _CascadeGraphFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, owner, x, scale_all, shift_all, *tensors):
        eng, lay, tg = owner
        tg.x.copy_(x.detach().reshape(tg.x.shape))
        tg.scale.copy_(scale_all.detach())
        tg.shift.copy_(shift_all.detach())
        tg.gf.replay()
        ctx.owner = owner
        ctx.token = _Token()
        tg.pending = weakref.ref(ctx.token)
        img, first = tg.img.clone(), tg.first.clone()
        ctx.mark_non_differentiable(first)
        return img, first

    @staticmethod
    def backward(ctx, dimg, _dfirst):
        eng, lay, tg = ctx.owner
        tok = tg.pending() if tg.pending is not None else None
        if tok is not ctx.token or tok.done:
            raise RuntimeError("bnerv_b200 native training: this forward's activation maps were overwritten by a later "
                               "forward of the same captured graph (or backward ran twice); set engine.train_graph = False "
                               "for such schedules")
        tg.dimg.copy_(dimg)
        tg.gb.replay()
        tok.done = True
        gx = tg.gx.clone() if tg.gx is not None else None
        return _unpack_grads(lay, ctx.needs_input_grad, gx, tg.gscale.clone(), tg.gshift.clone(), tg.gconv.clone())


def _graph_key(lay, x):
    """Shapes + parameter storage: a captured graph bakes device pointers of the weights it packs."""
    key = [tuple(x.shape), bool(x.requires_grad)]
    for t in lay.conv_tensors():
        if t is None:
            key.append(None)
            continue
        if not t.is_leaf or t.dtype != torch.float32 or not t.is_contiguous():
            return None                         # e.g. dequant_w re-created by cal_params every step: no stable storage
        key.append(t.data_ptr())
    return tuple(key)


# Gradient-range monitor (bnerv_bwd_set_status): one int per device that the backward kernels OR bits into when an f16
# gradient map saturates (bit 0) or turns non-finite (bit 1).  Read every CHECK_EVERY native steps (and after steps 1, 2, 4:
# a bad loss scale shows at once); on a hit the model falls back to torch autograd for the following steps and a
# RuntimeWarning says so - a diverging or clipped run must not look like a healthy one.
_STATUS = {}
CHECK_EVERY = 32


def _status_tensor(device):
    st = _STATUS.get(device)
    if st is None:
        st = _STATUS[device] = [torch.zeros(4, dtype=torch.int32, device=device), 0]
        check("bnerv_bwd_set_status", lib.bnerv_bwd_set_status(ptr(st[0])))
    return st


def gradient_range_status(device=None, reset=True):
    """Bits seen by the native backward since the last reset: 1 = an f16 gradient map saturated (clipped at 65504 behind the
    loss scale), 2 = a non-finite gradient.  Synchronises the device."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    st = _STATUS.get(device)
    if st is None:
        return 0
    bits = int(st[0][0].item())
    if reset and bits:
        st[0][0].zero_()
    return bits


def loss_scale_state(device=None):
    """(target for S * max|dL/dz_head| chosen by the device-side controller, largest scaled gradient of the last backward)."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    st = _STATUS.get(device)
    if st is None:
        return None
    v = st[0][1:3].clone().view(torch.float32).tolist()
    return {"target": v[1] if v[1] > 0 else 8.0, "last_max_scaled_gradient": v[0]}


def _check_gradient_range(eng, device):
    st = _status_tensor(device)
    st[1] += 1
    n = st[1]
    if n in (2, 3, 5) or n % CHECK_EVERY == 0:
        bits = gradient_range_status(device)
        if bits:
            import warnings
            what = " and ".join(w for b, w in ((1, "saturated f16 gradient maps (clipped at 65504 behind the loss scale)"),
                                                 (2, "non-finite gradients")) if bits & b)
            fallback = not os.environ.get("BNERV_TRAIN_NO_FALLBACK")
            warnings.warn(f"bnerv_b200 native backward: {what} in the last {min(n, CHECK_EVERY)} step(s)"
                          + ("; switching this model to train_backend = 'torch' (fp32 autograd)" if fallback else ""),
                          RuntimeWarning, stacklevel=3)
            if fallback:
                eng.model.train_backend = "torch"


def cascade_train(eng, x, cond):
    """x: [B, C, h, w] stem output (differentiable); cond: time embedding fed to every SFT layer.
    Returns (img [B,3,H,W] f32, first block output NCHW f32 (non-differentiable))."""
    if not x.is_cuda:
        raise RuntimeError("bnerv_b200 native training needs CUDA tensors (no CPU path)")
    ops.require_current_device(x.device)
    _check_gradient_range(eng, x.device)
    B = x.shape[0]
    lays = eng.__dict__.setdefault("_train_layouts", {})
    lay = lays.get(B)
    if lay is None:
        lay = lays[B] = _Layout(eng, B)
    scale_all, shift_all = lay.sft_tables(cond)
    tensors = lay.conv_tensors()
    if getattr(eng, "train_graph", True):
        key = _graph_key(lay, x)
        if key is not None:
            graphs = eng.__dict__.setdefault("_train_graphs", {})
            tg = graphs.get(key)
            if tg is None:
                if len(graphs) >= 4:            # parameter storage moved (.to(), new optimiser state ...): drop stale captures
                    graphs.clear()
                tg = graphs[key] = _TrainGraph(eng, lay, x, scale_all, shift_all, bool(x.requires_grad))
            if not tg.busy():
                return _CascadeGraphFn.apply((eng, lay, tg), x, scale_all, shift_all, *tensors)
    return _CascadeFn.apply((eng, lay), x, scale_all, shift_all, *tensors)
