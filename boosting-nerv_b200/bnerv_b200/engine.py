"""Decode engine: walks a *_Boost model's parameter tree once, packs weights into kernel layout, owns the
activation workspaces, and runs the block cascade with three fused-conv launches per NeRVBlock:

    up-conv (+PixelShuffle +sin)  -> x0 (kept for the residual)  and  u = x0*(g0+1)+b0
    conv0   (+GELU)               -> w = gelu(.)*(g1+1)+b1
    conv1   (+residual x0)        -> block output

plus ONE launch that evaluates every SFT layer's (scale, shift) MLP for the frame(s), and the head conv
(+tanh*0.5+0.5) that writes the NCHW f32 image.  (model_blocks.py:34-46, 74-105; model_*.py forward.)

The whole per-frame sequence (position encoding, stem MLPs, SFT table, cascade, head) is captured ONCE per input
shape into a CUDA graph and replayed: ~40 launches per frame with no host work in between, the fused convs chained
by programmatic dependent launch.  ``engine.use_graph = False`` runs the same sequence eagerly.

Weights are re-packed only when the effective tensor (``dequant_w ?? weight``) changed — detected through
tensor identity + in-place version counters, so optimiser steps and ``cal_params`` invalidate the cache.
"""
import os

import torch
import torch.nn as nn

from . import ops
from .layers import (Conv_Up_Block, CustomConv2d, DownConv, NeRVBlock, UpConv, act_name, effective_weight)


class Unsupported(RuntimeError):
    pass


def _tensor_key(t):
    return None if t is None else (id(t), t._version, t.data_ptr())


class _ConvSlot:
    """One conv of the cascade: module + geometry + lazily (re)packed weights."""

    def __init__(self, module, s=1, head=False):
        if not isinstance(module, nn.Conv2d) or module.stride != (1, 1) or module.dilation != (1, 1) or module.groups != 1:
            raise Unsupported(f"conv {module} is not a stride-1 dense conv")
        k = module.kernel_size[0]
        if module.kernel_size != (k, k) or k not in (1, 3) or module.padding != ((k - 1) // 2,) * 2:
            raise Unsupported(f"conv {module}: only 1x1 / 3x3 'same' convs are accelerated")
        self.m, self.k, self.s = module, k, s
        self.cin, self.cout = module.in_channels, module.out_channels // (s * s)
        self.key, self.pc = None, None
        self.skey, self.ps = None, None          # packing for a split input map (precise blocks, ops.split_weight)
        self.dkey, self.pd = None, None          # dgrad packing of the same weights (training path, train.py)
        # HNeRV's 3x3 head to 3 channels has its own kernel form (bnerv_head_conv3); BNERV_NO_HEAD_KERNEL=1 = generic path
        self.head3 = bool(head and k == 3 and self.cout <= 3 and not os.environ.get("BNERV_NO_HEAD_KERNEL"))
        # NeRV / E-NeRV's 1x1 head: HBM-bound CUDA-core kernel (bnerv_head_conv1)
        self.head1 = bool(head and k == 1 and self.cout <= 4 and not os.environ.get("BNERV_NO_HEAD_KERNEL"))

    def packed(self, force=False, split_in=False):
        """force: re-pack unconditionally (inside a captured training graph the pack kernels ARE the per-step ingest).
        split_in: the packing for a split input map (3 * Cin_p input channels)."""
        w, b = effective_weight(self.m)
        key = (_tensor_key(w), _tensor_key(b))
        if split_in:
            if key != self.skey or force:
                if self.ps is None:
                    self.ps = ops.PackedHead(w, b, split_in=True) if self.head3 else ops.PackedConv(w, b, self.s, split_in=True)
                else:
                    self.ps.repack(w, b)
                self.skey = key
            return self.ps
        if key != self.key or force:
            if self.pc is None:
                self.pc = (ops.PackedHead(w, b) if self.head3 else ops.PackedHead1(w, b) if self.head1
                           else ops.PackedConv(w, b, self.s))
            else:
                self.pc.repack(w, b)
            self.key = key
        return self.pc


def _upconv_slot(conv):
    """(slot) for an UpConv / DownConv wrapper in one of the accelerated forms."""
    if isinstance(conv, UpConv):
        if conv.kind not in ("pshuffel", "pshuffel_3x3"):
            raise Unsupported(f"UpConv conv_type {conv.kind!r} (no shipped script selects it)")
        shuffle = conv.upconv[1]
        s = shuffle.upscale_factor if isinstance(shuffle, nn.PixelShuffle) else 1
        return _ConvSlot(conv.upconv[0], s)
    if isinstance(conv, DownConv):
        if conv.kind != "conv":
            raise Unsupported(f"DownConv conv_type {conv.kind!r}")
        return _ConvSlot(conv.downconv, 1)      # HNeRV_Boost decoder[0]: 1x1, stride 1 (model_blocks.py:185)
    if isinstance(conv, nn.Conv2d):
        return _ConvSlot(conv, 1)
    raise Unsupported(f"unknown conv wrapper {type(conv).__name__}")


class _BlockPlan:
    def __init__(self, blk):
        if not isinstance(blk.norm, nn.Identity):
            raise Unsupported("norm layers other than 'none' are not accelerated (no shipped script selects them)")
        self.act = act_name(blk.act)
        if self.act is None:
            raise Unsupported(f"activation {blk.act} is not accelerated")
        if isinstance(blk, Conv_Up_Block):
            self.pre = _upconv_slot(blk.conv1)
            self.up = _upconv_slot(blk.conv2)
        elif isinstance(blk, NeRVBlock):
            if not blk.dec_block:
                raise Unsupported("NeRVBlock without dec_block (un-boosted HNeRV stem) is out of scope")
            self.pre = None
            self.up = _upconv_slot(blk.conv)
        else:
            raise Unsupported(f"block type {type(blk).__name__}")
        rb = getattr(blk, "sft_block", None)
        if rb is None or not all(hasattr(rb, n) for n in ("conv0", "conv1", "sft0", "sft1", "act")):
            raise Unsupported("blocks without a ResBlock_SFT (sft_block other than 'res_sft') are not accelerated")
        self.inner_act = act_name(rb.act)
        if self.inner_act is None:
            raise Unsupported(f"activation {rb.act} is not accelerated")
        self.c0, self.c1 = _ConvSlot(rb.conv0), _ConvSlot(rb.conv1)
        self.sfts = [rb.sft0, rb.sft1]
        for sft in self.sfts:
            if act_name(sft.act) != "relu":
                raise Unsupported("SFT inner activation other than relu")
        self.cout = self.up.cout
        self.scale = (self.pre.s if self.pre else 1) * self.up.s
        # One kernel per block for the narrow stages.  <= 16 channels: the row-streaming form (bnerv_nerv_block_stream, A operand
        # in tensor memory); when only the up-conv is outside it (PixelShuffle, wide input) the ResBlock_SFT half alone is one
        # kernel (bnerv_resblock_stream).  The region-tiled form (bnerv_nerv_block_fused, <= 48 channels) is correct but not
        # faster than three launches above 16 channels (profiles/r02_block_fused_ss_vs_three_launches.txt): BNERV_BLOCK_FORM=tile
        # selects it for comparison, BNERV_NO_BLOCK_FUSION=1 switches every fused form off.
        cp, cin_p = ops.round_up(self.cout, 16), ops.round_up(self.up.cin, 16)
        self.fuse = None
        form = os.environ.get("BNERV_BLOCK_FORM", "stream")
        if not os.environ.get("BNERV_NO_BLOCK_FUSION"):
            if form == "tile" and cp <= 48:
                whole = self.up.k == 3 and self.up.s in (1, 2) and cin_p <= 64 and self.up.s ** 2 * cp <= 256
                self.fuse = ("tile", "block" if whole else "res")
            elif form == "stream" and cp <= 16:
                whole = self.up.k == 3 and self.up.s in (1, 2) and cin_p <= 16
                self.fuse = ("stream", "block" if whole else "res")
            elif form == "stream" and cp == 32 and not os.environ.get("BNERV_NO_STREAM32"):
                self.fuse = ("stream", "res")       # 17..32 channels: up-conv launch + the ResBlock_SFT half as one kernel


def _sft_tensors(sft):
    """The eight effective tensors of one SFTLayer in bnerv_sft_layer order (raw, for identity keys)."""
    out = []
    for conv in (sft.SFT_scale_conv0, sft.SFT_scale_conv1, sft.SFT_shift_conv0, sft.SFT_shift_conv1):
        out += list(effective_weight(conv))
    return out


class _Captured:
    """One captured decode: static input buffers, the graph, and its (static) outputs."""
    __slots__ = ("graph", "inputs", "outputs")


class DecoderEngine:
    use_graph = True
    fuse_blocks = True          # narrow NeRVBlocks as one kernel each (False: always three fused-conv launches)
    # Blocks decoded in the split ("precise", hi + lo f16) form of bnerv_conv_fused_split: a set of block indices, plus "head"
    # for the head conv and "stem" for the cascade input.  Three times the tensor work of those blocks for ~2^-21 operands; the
    # escape hatch for checkpoints whose f16-operand decode is not close enough to f32 (DESIGN.md "precision").  Set it with
    # set_precise() or BNERV_PRECISE_BLOCKS ("all" | comma-separated indices / head / stem).
    precise = frozenset()

    def __init__(self, model):
        self.model = model
        self.kind = {"HNeRV_Boost": "hnerv", "NeRV_Boost": "nerv", "ENeRV_Boost": "enerv"}.get(type(model).__name__)
        self._graphs = {}      # (input shapes/dtypes, keep) -> _Captured
        self._wkey = None
        enc = getattr(model, "encoder", None)
        skip = set() if enc is None else {id(m) for m in enc.modules()}     # the encoder is not on the decode path
        self._wmods = [(m.__dict__, m._parameters) for m in model.modules()
                       if isinstance(m, (nn.Conv2d, nn.Linear)) and id(m) not in skip]
        name = type(model).__name__
        if name == "HNeRV_Boost":
            blocks, self.t_mlp = list(model.decoder), model.stem_t
        elif name in ("NeRV_Boost", "ENeRV_Boost"):
            blocks, self.t_mlp = list(model.layers), (model.stem_t if name == "NeRV_Boost" else None)
        else:
            raise Unsupported(f"model {name}")
        if model.out_bias != "tanh":
            raise Unsupported(f"out_bias {model.out_bias!r}: only 'tanh' is accelerated")
        self.blocks = [_BlockPlan(b) for b in blocks]
        self.head = _ConvSlot(model.head_layer, head=True)
        self._ws = {}          # workspaces per (B, h, w)
        self._sft = {}         # SftTable per B
        self._sft_key = None
        self._xy, self._xy_key = None, None     # E-NeRV: frame-independent half of the stem
        self._front = {}       # NeRV / E-NeRV: stem buffers per B
        self._pe_same = False  # E-NeRV: pe_t and pe_t_manipulate are the same encoding (decided here, outside any graph capture)
        if self.kind == "enerv":
            pe, pe2 = model.pe_t, model.pe_t_manipulate
            self._pe_same = ("pe" in getattr(pe, "pe_embed", "") and pe.pe_embed == pe2.pe_embed
                             and torch.equal(pe.pe_bases.detach().cpu(), pe2.pe_bases.detach().cpu()))
        env = os.environ.get("BNERV_PRECISE_BLOCKS")
        if env:
            self.set_precise(env)

    def set_precise(self, blocks):
        """blocks: "all", None/"" (off), or an iterable / comma-separated string of block indices, "head", "stem"."""
        if isinstance(blocks, str):
            blocks = blocks.strip()
            if blocks == "all":
                blocks = list(range(len(self.blocks))) + ["head", "stem"]
            else:
                blocks = [b if b in ("head", "stem") else int(b) for b in blocks.split(",") if b]
        sel = frozenset(blocks or ())
        bad = [b for b in sel if b not in ("head", "stem") and not (isinstance(b, int) and 0 <= b < len(self.blocks))]
        if bad:
            raise ValueError(f"precise blocks {bad}: not a block index of this model (0..{len(self.blocks) - 1}), 'head' or 'stem'")
        self.precise = sel
        self._ws.clear()
        self.invalidate()

    # -- helpers -----------------------------------------------------------------------------------
    def _mlp(self, seq, x):
        """NeRV_MLP on a [B, C] vector (1x1 convs + activation after every layer), model_blocks.py:66-71."""
        mods = list(seq)
        i = 0
        while i < len(mods):
            conv = mods[i]
            act = act_name(mods[i + 1]) if i + 1 < len(mods) else "none"
            if not isinstance(conv, nn.Conv2d) or conv.kernel_size != (1, 1) or act is None:
                raise Unsupported(f"stem layer {conv} / {mods[i + 1] if i + 1 < len(mods) else None}")
            w, b = effective_weight(conv)
            x = ops.linear_act(x, w.detach().contiguous(), None if b is None else b.detach(), act)
            i += 2
        return x

    def _sft_table(self, B, device):
        tensors = [_sft_tensors(s) for blk in self.blocks for s in blk.sfts]
        key = tuple(_tensor_key(t) for l in tensors for t in l)
        if key != self._sft_key:
            self._sft.clear()
            self._sft_key = key
        tab = self._sft.get(B)
        if tab is None:
            tab = ops.SftTable([tuple(t.detach().reshape(t.shape[0], -1).contiguous().float() if t.dim() == 4
                                      else t.detach().contiguous().float() for t in l) for l in tensors], B, device)
            self._sft[B] = tab
        return tab

    def _workspace(self, B, h, w, device):
        ws = self._ws.get((B, h, w))
        if ws is None:
            sizes = {"cur": 0, "x0": 0, "u": 0, "w": 0, "nxt": 0}
            H, W = h, w
            for blk in self.blocks:
                if blk.pre is not None:
                    H, W = H * blk.pre.s, W * blk.pre.s
                    sizes["w"] = max(sizes["w"], B * ops.round_up(blk.pre.cout, 16) * H * W)
                H, W = H * blk.up.s, W * blk.up.s
                n = B * ops.round_up(blk.cout, 16) * H * W
                for kname in ("cur", "x0", "u", "w", "nxt"):
                    sizes[kname] = max(sizes[kname], n)
            mult = 3 if self.precise else 1         # split maps: [hi | lo | hi]
            ws = {kname: torch.empty(max(n * mult, 8), dtype=torch.float16, device=device) for kname, n in sizes.items()}
            ws["out_hw"] = (H, W)
            self._ws[(B, h, w)] = ws
        return ws

    # -- entry points ------------------------------------------------------------------------------
    def weights_key(self):
        """Identity + in-place version of every effective weight/bias the decode reads (torch stems included)."""
        key = []
        for d, prm in self._wmods:              # effective_weight() without nn.Module.__getattr__ (runs on every forward)
            w = d.get("dequant_w")
            if w is None:
                w = prm["weight"]
            b = d.get("dequant_b")
            if b is None:
                b = prm.get("bias")
            key.append(id(w))
            key.append(w._version)
            key.append(w.data_ptr())        # .to(device) / .data swaps keep id and _version
            if b is not None:
                key.append(id(b))
                key.append(b._version)
                key.append(b.data_ptr())
        return tuple(key)

    def sync_weights(self):
        """Compare the effective weights with the ones the captured graphs were built from and drop the graphs on a
        mismatch.  ``decode(check_weights=False)`` (model.decode) skips this per frame; the batch entry points
        (stream.decode_to_host / evaluate_*) call it once per call, so weights changed by an optimiser step,
        load_state_dict or cal_params between two calls are never served stale."""
        wk = self.weights_key()
        if wk != self._wkey:
            self._graphs.clear()
            self._wkey = wk

    def invalidate(self):
        """Drop captured graphs (call after replacing weights when decoding through ``decode(check_weights=False)``)."""
        self._graphs.clear()
        self._wkey = None

    def _body(self, inputs, keep):
        """The full per-frame sequence for this model family; pure stream work.  Returns (img, outs, extra)."""
        m = self.model
        if self.kind == "hnerv":                # HNeRV_Boost.forward_decoder (model_hnerv.py:264-277); f64 PE -> f32 (:267)
            img_embed, norm_idx = inputs
            pe = m.pe_embed_t(norm_idx[:, None]).float()
            t_embed = self._mlp(self.t_mlp, pe.flatten(1).float())
            img, outs = self.run_cascade(img_embed.float().contiguous(), t_embed, keep)
            return img, outs, None
        if self.kind == "nerv":                 # NeRV_Boost.forward (model_nerv.py:47-57)
            (t,) = inputs
            front = self._nerv_front(t)
            if front is not None:               # PE + both MLPs in two launches, cascade input written as C8 directly
                x, x_c8, t_embed = front
                img, outs = self.run_cascade(x, t_embed, keep, x_c8=x_c8)
                return img, outs, None
            v = m.pe_t(t[:, None].float()).flatten(1).float()
            x = self._mlp(m.stem, v).view(v.size(0), m.fc_dim, m.fc_h, m.fc_w)
            t_embed = self._mlp(m.stem_t, v)
            img, outs = self.run_cascade(x, t_embed, keep)
            return img, outs, None
        (t,) = inputs                           # ENeRV_Boost.forward (model_enerv.py:281-313)
        emb, t_manip = self._enerv_stem(t)
        img, outs = self.run_cascade(emb.contiguous(), t_manip.flatten(1), keep)
        return img, outs, t_manip

    def _mlp2(self, seq):
        """[(weight [Cout, Cin] f32, bias, act)] x 2 of a two-layer NeRV_MLP, or None if it has another form."""
        mods = list(seq)
        if len(mods) != 4:
            return None
        out = []
        for conv, act in ((mods[0], mods[1]), (mods[2], mods[3])):
            a = act_name(act)
            if not isinstance(conv, nn.Conv2d) or conv.kernel_size != (1, 1) or a is None:
                return None
            w, b = effective_weight(conv)
            if w.dtype != torch.float32:
                return None
            out.append((w.detach().reshape(w.shape[0], -1).contiguous(), None if b is None else b.detach().float().contiguous(), a))
        return out

    def _nerv_front(self, t):
        """NeRV_Boost's stem (model_nerv.py:47-52) as two launches: position encoding + first layer of stem and stem_t, then
        their second layers with the stem's output stored as the cascade's C8 input (and f32 for callers that want it).
        -> (x [B, C, h, w] f32, x_c8, t_embed) or None when the stem has another shape (then the generic path runs)."""
        m = self.model
        if os.environ.get("BNERV_NO_FRONT_FUSION") or "pe" not in getattr(m.pe_t, "pe_embed", ""):
            return None
        stem, stem_t = self._mlp2(m.stem), self._mlp2(m.stem_t)
        if stem is None or stem_t is None:
            return None
        B, dev = t.shape[0], t.device
        hw = m.fc_h * m.fc_w
        (w1, b1, a1), (w2, b2, a2) = stem
        (v1, c1, e1), (v2, c2, e2) = stem_t
        if w2.shape[0] != m.fc_dim * hw or B * max(w1.shape[0], v1.shape[0], w1.shape[1]) * 4 > 48 * 1024:
            return None
        bases = m.pe_t.pe_bases
        if bases.device != dev or bases.dtype != torch.float32:
            bases = m.pe_t.pe_bases = bases.to(dev, torch.float32)
        bufs = self._front.get(B)
        if bufs is None:
            f32 = lambda n: torch.empty((B, n), dtype=torch.float32, device=dev)
            bufs = self._front[B] = (f32(w1.shape[0]), f32(v1.shape[0]), f32(w2.shape[0]), f32(v2.shape[0]),
                                     torch.zeros(ops.c8_shape(B, m.fc_dim, m.fc_h, m.fc_w), dtype=torch.float16, device=dev))
        h, ht, x, t_embed, x_c8 = bufs
        ops.pe_linear_pair(t.float(), bases, [dict(w=w1, b=b1, act=a1, y=h), dict(w=v1, b=c1, act=e1, y=ht)])
        ops.linear_pair([dict(x=h, w=w2, b=b2, act=a2, y=x, y_c8=x_c8, hw=hw), dict(x=ht, w=v2, b=c2, act=e2, y=t_embed)], B)
        return x.view(B, m.fc_dim, m.fc_h, m.fc_w), x_c8, t_embed

    def _time_mlp_pair(self, t):
        """E-NeRV's stem_t(pe_t(t)) and t_branch(pe_t_manipulate(t)) (model_enerv.py:283-286) through bnerv_pe_linear_pair +
        bnerv_linear_pair when both are two-layer NeRV_MLPs over the same position encoding; None otherwise."""
        m = self.model
        pe, pe2 = m.pe_t, m.pe_t_manipulate
        if os.environ.get("BNERV_NO_FRONT_FUSION") or not self._pe_same:
            return None
        a, c = self._mlp2(m.stem_t), self._mlp2(m.t_branch)
        if a is None or c is None:
            return None
        B, dev = t.shape[0], t.device
        (w1, b1, a1), (w2, b2, a2) = a
        (v1, c1, e1), (v2, c2, e2) = c
        if B * max(w1.shape[0], v1.shape[0], w1.shape[1]) * 4 > 48 * 1024:
            return None
        if pe.pe_bases.device != dev or pe.pe_bases.dtype != torch.float32:
            pe.pe_bases = pe.pe_bases.to(dev, torch.float32)
        bufs = self._front.get(B)
        if bufs is None:
            f32 = lambda n: torch.empty((B, n), dtype=torch.float32, device=dev)
            bufs = self._front[B] = (f32(w1.shape[0]), f32(v1.shape[0]), f32(w2.shape[0]), f32(v2.shape[0]))
        h, ht, t_emb, t_manip = bufs
        ops.pe_linear_pair(t.float(), pe.pe_bases, [dict(w=w1, b=b1, act=a1, y=h), dict(w=v1, b=c1, act=e1, y=ht)])
        ops.linear_pair([dict(x=h, w=w2, b=b2, act=a2, y=t_emb), dict(x=ht, w=v2, b=c2, act=e2, y=t_manip)], B)
        return t_emb, t_manip

    def _enerv_stem(self, t):
        """ENeRV_Boost's stem (model_enerv.py:281-303) for the decode path.  Same values as model._stem(t), arranged for a
        per-frame decode: (i) the two time MLPs run on the f32 bnerv_linear_act kernel; (ii) the coordinate branch
        trans1(stem_xy(pe_xy(grid))) does not depend on the frame, so it is computed once per weight version and captured
        graphs hold it as a constant (about half of the stem's ~65 small torch kernels per frame); (iii) the transformer and
        the 1x1 `toconv` stay torch modules but with cuDNN's TF32 convs switched off - the reference arithmetic is f32, and
        `t_manip` feeds every SFT layer of the cascade."""
        m = self.model
        b = t.size(0)
        pair = self._time_mlp_pair(t)
        if pair is not None:                    # both time MLPs read the same position encoding: two launches for the lot
            t_emb, t_manip = pair
        else:
            tt = t[:, None].float()
            t_emb = self._mlp(m.stem_t, m.pe_t(tt).flatten(1).float())
            t_manip = self._mlp(m.t_branch, m.pe_t_manipulate(tt).flatten(1).float())
        tf32 = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        try:
            key = (self.weights_key(), tuple(_tensor_key(p) for p in m.trans1.parameters()), t.device)
            if self._xy_key != key:
                with torch.no_grad():
                    xy = m._xy_grid(t.device)
                    xy_emb = torch.cat([m.pe_xy(xy[0][:, None]), m.pe_xy(xy[1][:, None])], dim=1)
                    self._xy = m.trans1(m.stem_xy(xy_emb).view(1, m.fc_h * m.fc_w, -1)).contiguous()
                self._xy_key = key
            emb = m.trans2(self._xy.expand(b, -1, -1) * t_emb[:, None, :])
            emb = m.toconv(emb.reshape(b, m.fc_h, m.fc_w, emb.shape[-1]).permute(0, 3, 1, 2))
        finally:
            torch.backends.cudnn.allow_tf32 = tf32
        return emb, t_manip.view(b, -1, 1, 1)

    def decode(self, inputs, keep=False, check_weights=True):
        """inputs: (img_embed, norm_idx) for HNeRV_Boost, (t,) otherwise.  Returns (img, outs, extra).
        With graphs enabled the returned tensors are the graph's static outputs: they are overwritten by the next
        decode of the same shape (callers that keep them must clone)."""
        inputs = tuple(inputs)
        if not inputs[0].is_cuda:
            raise RuntimeError("bnerv_b200 engine needs CUDA tensors (no CPU path)")
        ops.require_current_device(inputs[0].device)
        if not self.use_graph:
            return self._body(inputs, keep)
        if check_weights:
            self.sync_weights()
        gkey = (tuple((tuple(t.shape), t.dtype, t.device) for t in inputs), keep)
        cap = self._graphs.get(gkey)
        if cap is None:
            cap = self._capture(inputs, keep)
            self._graphs[gkey] = cap
        for dst, src in zip(cap.inputs, inputs):
            dst.copy_(src, non_blocking=True)
        cap.graph.replay()
        return cap.outputs

    def _capture(self, inputs, keep):
        cap = _Captured()
        cap.inputs = [t.detach().clone() for t in inputs]
        self._body(cap.inputs, keep)            # eager warm-up: packs weights, builds workspaces / SFT tables, sets attributes
        torch.cuda.synchronize()
        cap.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cap.graph, capture_error_mode="thread_local"):   # DataLoader pin-memory threads may call CUDA meanwhile
            cap.outputs = self._body(cap.inputs, keep)
        return cap

    def run_cascade(self, x, t_embed, keep=False, x_c8=None):
        """x: [B, C, h, w] f32 NCHW stem output (x_c8: the same map already in C8 f16); t_embed: [B, ch_t] f32.
        Returns (img, [block outputs]).  keep: True = every block output as NCHW f32, "first" = only block 0's, False = none."""
        B, C, h, w = x.shape
        dev = x.device
        ws = self._workspace(B, h, w, dev)
        tab = self._sft_table(B, dev)
        tab.run(t_embed)

        def view(buf, c, H, W):
            shp = ops.c8_shape(B, c, H, W)
            return buf[:shp[0] * shp[1] * shp[2] * shp[3] * shp[4]].view(shp)

        cur_buf, nxt_buf = ws["cur"], ws["nxt"]
        prec = self.precise
        nb = len(self.blocks)
        cur_split = bool(prec) and "stem" in prec and 0 in prec
        if cur_split:                       # the cascade input as a split map: hi / lo / hi channel blocks, each padded to 16
            xp = torch.nn.functional.pad(x, (0, 0, 0, 0, 0, ops.round_up(C, 16) - C))
            hi = xp.half().float()
            cur = ops.nchw_to_c8(torch.cat([hi, xp - hi, hi], dim=1))
        else:
            cur = x_c8 if x_c8 is not None else ops.nchw_to_c8(x)
        cin, H, W = C, h, w
        outs = []
        for bi, blk in enumerate(self.blocks):
            g0, b0 = tab.g1p[2 * bi], tab.beta[2 * bi]
            g1, b1 = tab.g1p[2 * bi + 1], tab.beta[2 * bi + 1]
            precise = bi in prec
            if blk.pre is not None:         # E-NeRV stage 0: conv1 (no activation) feeds conv2
                cmid = 3 * ops.round_up(blk.pre.cout, 16) if precise else blk.pre.cout
                mid = view(ws["w"], cmid, H * blk.pre.s, W * blk.pre.s)
                ops.conv_fused(cur, blk.pre.packed(split_in=cur_split), 3 * ops.round_up(cin, 16) if cur_split else cin, H, W,
                               act="none", out_pre=mid, split=1 if precise else 0)
                cur, cin, H, W, cur_split = mid, blk.pre.cout, H * blk.pre.s, W * blk.pre.s, precise
            Ho, Wo = H * blk.up.s, W * blk.up.s
            cp = ops.round_up(blk.cout, 16)
            out_split = ((bi + 1) in prec) if bi + 1 < nb else ("head" in prec)
            out = view(nxt_buf, 3 * cp if out_split else blk.cout, Ho, Wo)
            done = None
            fuse = blk.fuse if (self.fuse_blocks and not precise and not out_split) else None
            if precise:
                # split form: every map of the block is [hi | lo | hi], every conv reads 3*Cp channels (bnerv_conv_fused_split)
                cin_k = 3 * ops.round_up(cin, 16) if cur_split else cin
                x0 = view(ws["x0"], 3 * cp, Ho, Wo)
                u = view(ws["u"], 3 * cp, Ho, Wo)
                wbuf = view(ws["w"], 3 * cp, Ho, Wo)
                ops.conv_fused(cur, blk.up.packed(split_in=cur_split), cin_k, H, W, act=blk.act, g1p=g0, beta=b0, out_pre=x0,
                               out_aff=u, split=1)
                ops.conv_fused(u, blk.c0.packed(split_in=True), 3 * cp, Ho, Wo, act=blk.inner_act, g1p=g1, beta=b1, out_aff=wbuf,
                               split=1)
                ops.conv_fused(wbuf, blk.c1.packed(split_in=True), 3 * cp, Ho, Wo, act="none", resid=x0, out_pre=out,
                               split=2 | (1 if out_split else 0))
                done = True
            # (bnerv_nerv_block_stream_head - the last NeRV-Boost block with the 1x1 head inside - exists and is bit-identical, but its
            #  single back warpgroup becomes the critical stage: NeRV-S 5056 -> 4961 frames/s.  BNERV_HEAD_FUSION16=1 selects it.)
            if (fuse is not None and fuse == ("stream", "block") and bi + 1 == nb and self.head.head1 and keep is not True
                    and blk.up.s == 1 and blk.act == "sin" and blk.inner_act == "gelu" and os.environ.get("BNERV_HEAD_FUSION16")):
                img = torch.empty((B, self.head.cout, Ho, Wo), dtype=torch.float32, device=dev)
                if ops.nerv_block_head_fused(cur, blk.up.packed(), blk.c0.packed(), blk.c1.packed(), cin, H, W, g0, b0, g1, b1,
                                             self.head.packed(), img) is not None:
                    return img, outs
            if fuse is not None and fuse[1] == "block":
                done = ops.nerv_block_fused(cur, blk.up.packed(), blk.c0.packed(), blk.c1.packed(), cin, H, W, blk.act,
                                            blk.inner_act, g0, b0, g1, b1, out=out, form=fuse[0])
            if done is None:
                x0 = view(ws["x0"], blk.cout, Ho, Wo)
                u = view(ws["u"], blk.cout, Ho, Wo)
                streamed = None
                if (fuse is not None and fuse[0] == "stream" and cp == 32 and ops.round_up(cin, 16) == 32 and Ho * Wo >= 65536
                        and blk.up.k == 3 and blk.up.s == 1):       # 17..32 -> 17..32 channels at a large map: the streaming up-conv
                    streamed = ops.upconv_stream(cur, blk.up.packed(), cin, H, W, blk.act, g0, b0, x0, u)
                # 33..48 channels at a large map (E-NeRV-M's 540p stages): each conv in the single-conv streaming form
                wide48 = (cp == 48 and Ho * Wo >= 65536 and self.fuse_blocks and not os.environ.get("BNERV_NO_CONV_STREAM"))
                if streamed is None and wide48 and blk.up.k == 3 and blk.up.s == 1:
                    streamed = ops.conv_stream(cur, blk.up.packed(), cin, H, W, act=blk.act, g1p=g0, beta=b0, out_pre=x0, out_aff=u)
                if streamed is None:
                    ops.conv_fused(cur, blk.up.packed(), cin, H, W, act=blk.act, g1p=g0, beta=b0, out_pre=x0, out_aff=u)
                if fuse is not None and (cp <= 16 or Ho * Wo >= 65536):      # the 32-channel form pays off on large maps only
                    if (bi + 1 == nb and cp == 32 and fuse[0] == "stream" and self.head.head1 and keep is not True
                            and blk.inner_act == "gelu" and not os.environ.get("BNERV_NO_HEAD_FUSION")):
                        # last block: the 1x1 head conv + OutImg ride in the same kernel, the block output is never stored
                        img = torch.empty((B, self.head.cout, Ho, Wo), dtype=torch.float32, device=dev)
                        if ops.resblock_head_fused(u, x0, blk.c0.packed(), blk.c1.packed(), blk.cout, Ho, Wo, blk.inner_act, g1, b1,
                                                   self.head.packed(), img) is not None:
                            return img, outs
                    done = ops.resblock_fused(u, x0, blk.c0.packed(), blk.c1.packed(), blk.cout, Ho, Wo, blk.inner_act, g1, b1,
                                              out=out, form=fuse[0])
                if done is None:
                    wbuf = view(ws["w"], blk.cout, Ho, Wo)
                    if not (wide48 and ops.conv_stream(u, blk.c0.packed(), blk.cout, Ho, Wo, act=blk.inner_act, g1p=g1, beta=b1, out_aff=wbuf)):
                        ops.conv_fused(u, blk.c0.packed(), blk.cout, Ho, Wo, act=blk.inner_act, g1p=g1, beta=b1, out_aff=wbuf)
                    if not (wide48 and not out_split and ops.conv_stream(wbuf, blk.c1.packed(), blk.cout, Ho, Wo, act="none", resid=x0, out_pre=out)):
                        ops.conv_fused(wbuf, blk.c1.packed(), blk.cout, Ho, Wo, act="none", resid=x0, out_pre=out,
                                       split=1 if out_split else 0)
            if keep is True or (keep == "first" and bi == 0):
                if out_split:               # hi + lo of the split map
                    g = cp // 8
                    outs.append(ops.c8_to_nchw(out[:, :g].contiguous(), blk.cout) + ops.c8_to_nchw(out[:, g:2 * g].contiguous(), blk.cout))
                else:
                    outs.append(ops.c8_to_nchw(out, blk.cout))
            cur, cin, H, W, cur_split = out, blk.cout, Ho, Wo, out_split
            cur_buf, nxt_buf = nxt_buf, cur_buf
        img = torch.empty((B, 3, H, W), dtype=torch.float32, device=dev)
        if cur_split:
            ops.conv_fused(cur, self.head.packed(split_in=True), 3 * ops.round_up(cin, 16), H, W, act="tanh01", out_nchw=img)
        else:
            ops.conv_fused(cur, self.head.packed(), cin, H, W, act="tanh01", out_nchw=img)
        return img, outs
