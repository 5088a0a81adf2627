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
reductions are f32.  Precision: f16 operands / f32 accumulation in both passes — the reference's own GPU training
runs its convs in TF32 (torch.backends.cudnn.allow_tf32 defaults to True), which has the same 11-bit significand.
"""
import torch
import torch.nn.functional as F

from . import ops
from .layers import effective_weight


def _pad_table(t, cp, plus_one):
    """[B, C, 1, 1] (scale or shift) -> contiguous f32 [B, Cp] table (scale + 1 like model_blocks.py:105)."""
    t = t.detach().reshape(t.shape[0], -1).float()
    if plus_one:
        t = t + 1.0
    return F.pad(t, (0, cp - t.shape[1])).contiguous()


class _Slots:
    """The conv slots of an engine in autograd-argument order, with their dgrad packings."""

    def __init__(self, eng):
        self.convs = []
        for blk in eng.blocks:
            if blk.pre is not None:
                self.convs.append(blk.pre)
            self.convs += [blk.up, blk.c0, blk.c1]
        self.convs.append(eng.head)

    def tensors(self):
        out = []
        for slot in self.convs:
            w, b = effective_weight(slot.m)
            out += [w, b]
        return out


def _dgrad_pack(slot):
    w, _ = effective_weight(slot.m)
    key = (id(w), w._version, w.data_ptr())
    if getattr(slot, "dkey", None) != key:
        if getattr(slot, "pd", None) is None:
            slot.pd = ops.PackedDgrad(w, slot.s)
        else:
            slot.pd.repack(w)
        slot.dkey = key
    return slot.pd


def _dgrad(slot, dy_u, H, W):
    """dL/dx (C8 f16 [B][Cin_p/8][H][W][8]) of a conv from its un-shuffled output gradient."""
    pd = _dgrad_pack(slot)
    dx = torch.empty(ops.c8_shape(dy_u.shape[0], slot.cin, H, W), dtype=torch.float16, device=dy_u.device)
    ops.conv_fused(dy_u, pd, pd.cin, H, W, act="none", out_pre=dx)
    return dx


def _conv_param_grads(slot, x_in, dy_u, inv, need_w, need_b, dbias_acc=None):
    """(dW OIHW f32, db f32) of one conv.  dbias_acc: un-shuffled per-channel sums if a fused pass already made them."""
    gw = gb = None
    if need_w:
        acc = ops.conv_wgrad(x_in, dy_u, slot.cin, slot.k)
        gw = ops.wgrad_finalize(acc, slot.cout, slot.cin, slot.k, slot.s, inv)
    if need_b:
        if dbias_acc is None:
            dbias_acc = ops.channel_sum(dy_u)
        gb = ops.bias_finalize(dbias_acc, slot.cout, slot.s, inv)
    return gw, gb


class _CascadeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eng, x, *tensors):
        nb = len(eng.blocks)
        sft = tensors[:4 * nb]                      # per block: scale0, shift0, scale1, shift1  ([B, C, 1, 1])
        B, C, h, w = x.shape
        dev = x.device
        if not x.is_cuda:
            raise RuntimeError("bnerv_b200 native training needs CUDA tensors (no CPU path)")
        c8 = lambda c, H, W: torch.empty(ops.c8_shape(B, c, H, W), dtype=torch.float16, device=dev)
        cur = ops.nchw_to_c8(x.detach().float().contiguous())
        cin, H, W = C, h, w
        saved = []
        for bi, blk in enumerate(eng.blocks):
            cp = ops.round_up(blk.cout, 16)
            g0p, b0 = _pad_table(sft[4 * bi], cp, True), _pad_table(sft[4 * bi + 1], cp, False)
            g1p, b1 = _pad_table(sft[4 * bi + 2], cp, True), _pad_table(sft[4 * bi + 3], cp, False)
            rec = {"in": cur, "in_hw": (H, W), "g0p": g0p, "g1p": g1p}
            if blk.pre is not None:                 # E-NeRV stage 0: up-conv without activation feeds a 3x3 conv
                mid = c8(blk.pre.cout, H * blk.pre.s, W * blk.pre.s)
                ops.conv_fused(cur, blk.pre.packed(), cin, H, W, act="none", out_pre=mid)
                cur, cin, H, W = mid, blk.pre.cout, H * blk.pre.s, W * blk.pre.s
                rec["mid"], rec["mid_hw"] = mid, (H, W)
            Ho, Wo = H * blk.up.s, W * blk.up.s
            x0, u, d0 = c8(blk.cout, Ho, Wo), c8(blk.cout, Ho, Wo), c8(blk.cout, Ho, Wo)
            ops.conv_fused(cur, blk.up.packed(), cin, H, W, act=blk.act, g1p=g0p, beta=b0, out_pre=x0, out_aff=u, out_deriv=d0)
            v, wmap, d1 = c8(blk.cout, Ho, Wo), c8(blk.cout, Ho, Wo), c8(blk.cout, Ho, Wo)
            ops.conv_fused(u, blk.c0.packed(), blk.cout, Ho, Wo, act=blk.inner_act, g1p=g1p, beta=b1, out_pre=v, out_aff=wmap,
                           out_deriv=d1)
            out = c8(blk.cout, Ho, Wo)
            ops.conv_fused(wmap, blk.c1.packed(), blk.cout, Ho, Wo, act="none", resid=x0, out_pre=out)
            rec.update(x0=x0, u=u, d0=d0, v=v, w=wmap, d1=d1, hw=(Ho, Wo))
            saved.append(rec)
            cur, cin, H, W = out, blk.cout, Ho, Wo
        img = torch.empty((B, 3, H, W), dtype=torch.float32, device=dev)
        ops.conv_fused(cur, eng.head.packed(), cin, H, W, act="tanh01", out_nchw=img)
        ctx.eng, ctx.saved, ctx.last, ctx.last_hw, ctx.img, ctx.x_shape = eng, saved, cur, (H, W), img, (B, C, h, w)
        first = ops.c8_to_nchw(saved[1]["in"] if nb > 1 else cur, eng.blocks[0].cout)     # block 0's output (callers keep [0])
        ctx.mark_non_differentiable(first)
        return img, first

    @staticmethod
    def backward(ctx, dimg, _dfirst):
        eng, saved = ctx.eng, ctx.saved
        nb = len(eng.blocks)
        need = ctx.needs_input_grad            # (eng, x, *tensors)
        need_x, need_t = need[1], need[2:]
        B, C, h, w = ctx.x_shape
        dev = dimg.device
        slots = _Slots(eng).convs
        slot_pos = {id(s): i for i, s in enumerate(slots)}
        g_sft = [None] * (4 * nb)
        g_conv = [None] * (2 * len(slots))

        def put(slot, gw, gb):
            i = slot_pos[id(slot)]
            g_conv[2 * i], g_conv[2 * i + 1] = gw, gb

        def needs(slot):
            i = 4 * nb + 2 * slot_pos[id(slot)]
            return need_t[i], need_t[i + 1]

        scale = torch.zeros(2, dtype=torch.float32, device=dev)
        dz = ops.head_bwd(dimg, ctx.img, scale)
        inv = scale[1:2]
        H, W = ctx.last_hw
        put(eng.head, *_conv_param_grads(eng.head, ctx.last, dz, inv, *needs(eng.head)))
        dout = _dgrad(eng.head, dz, H, W)
        for bi in range(nb - 1, -1, -1):
            blk, rec = eng.blocks[bi], saved[bi]
            Ho, Wo = rec["hw"]
            Cb = blk.cout
            # conv1 (input w, output gradient dout)
            nw, nb_ = needs(blk.c1)
            gw1, _ = _conv_param_grads(blk.c1, rec["w"], dout, inv, nw, False)
            dw = _dgrad(blk.c1, dout, Ho, Wo)
            dc0, dG1, dB1, db0 = ops.resblock_mid_bwd(dw, rec["v"], rec["d1"], rec["g1p"], Cb)
            del dw
            # conv0 (input u, output gradient dc0)
            nw0, nb0 = needs(blk.c0)
            gw0, gb0 = _conv_param_grads(blk.c0, rec["u"], dc0, inv, nw0, nb0, dbias_acc=db0)
            du = _dgrad(blk.c0, dc0, Ho, Wo)
            del dc0
            dy, dG0, dB0, db1 = ops.block_front_bwd(du, dout, rec["x0"], rec["d0"], rec["g0p"], Cb)
            del du, dout
            put(blk.c0, gw0, gb0)
            put(blk.c1, gw1, ops.bias_finalize(db1, Cb, 1, inv) if nb_ else None)
            for j, t in enumerate((dG0, dB0, dG1, dB1)):
                if need_t[4 * bi + j]:
                    g_sft[4 * bi + j] = (t[:, :Cb] * inv).view(B, Cb, 1, 1)
            # up-conv (input: block input or the E-NeRV pre-conv's output)
            up_in = rec.get("mid", rec["in"])
            Hi, Wi = rec.get("mid_hw", rec["in_hw"])
            dyu = ops.unshuffle_c8(dy, Cb, blk.up.s)
            del dy
            put(blk.up, *_conv_param_grads(blk.up, up_in, dyu, inv, *needs(blk.up)))
            first_conv = bi == 0 and blk.pre is None
            dprev = None
            if not first_conv or need_x:
                dprev = _dgrad(blk.up, dyu, Hi, Wi)
            del dyu
            if blk.pre is not None:
                Hi0, Wi0 = rec["in_hw"]
                dmu = ops.unshuffle_c8(dprev, blk.pre.cout, blk.pre.s)
                put(blk.pre, *_conv_param_grads(blk.pre, rec["in"], dmu, inv, *needs(blk.pre)))
                dprev = _dgrad(blk.pre, dmu, Hi0, Wi0) if (bi > 0 or need_x) else None
            dout = dprev
        gx = None
        if need_x:
            gx = ops.c8_to_nchw(dout, C) * inv
        ctx.saved = None
        return (None, gx) + tuple(g_sft) + tuple(g_conv)


def cascade_train(eng, x, cond):
    """x: [B, C, h, w] stem output (differentiable); cond: time embedding fed to every SFT layer.
    Returns (img [B,3,H,W] f32, first block output NCHW f32 (non-differentiable))."""
    sft = []
    for blk in eng.blocks:
        for layer in blk.sfts:
            scale, shift = layer.affine(cond)
            sft += [scale, shift]
    return _CascadeFn.apply(eng, x, *sft, *_Slots(eng).tensors())
