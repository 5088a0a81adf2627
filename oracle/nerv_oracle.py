"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the Boosting-NeRV conditional-decoder forward.

Nothing in the product path (boosting-nerv_b200/) imports this file.  It may be imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, and only as the checker or the
timed CPU baseline.

What it is: a functional (state_dict in, tensors out) restatement of the reference's forward math using
plain torch CPU ops in f32 or f64 — the same ATen/oneDNN ops the reference itself executes on CPU, so its
wall-clock is representative of the reference's own CPU path.  Each function cites the reference
file:line it follows (tree: Xinjie-Q/Boosting-NeRV @ d59ca91).

Pinning: the reference ships no tests or golden vectors.  The oracle is pinned against
outputs of the unmodified reference executed in the build container: tests/golden/*.npz were produced by
tests/golden/make_golden.py importing /root/reference, and tests/test_oracle_golden.py checks this file
against them (state_dict -> per-block outputs -> image) on every run.
"""
import math

import torch
import torch.nn.functional as F


def _act(name):
    return {"sin": torch.sin, "gelu": F.gelu, "relu": F.relu, "none": lambda v: v}[name]


# Operand-rounding emulation.  None = the reference's arithmetic.  torch.float16 = round the input and the
# weight of every SPATIAL conv (feature maps larger than 1x1) to f16 before an f32-accumulated conv and round
# the stored x0 — the arithmetic model of the sm_100a kernels (f16 operands, f32 accumulate).  Used by tests
# to separate "kernel computes what it claims" (tight) from "f16 operands are accurate enough" (1e-3 gate).
EMULATE = None


def _q(t):
    return t if EMULATE is None else t.to(EMULATE).to(t.dtype)


def conv(sd, prefix, x, pad):
    """CustomConv2d.forward with the stored (non-quantised) weights — lib/quant_ops.py:39-41."""
    w, b = sd[prefix + ".weight"], sd.get(prefix + ".bias")
    w = w.to(x.dtype)
    if EMULATE is not None and x.shape[-1] * x.shape[-2] > 1:
        x, w = _q(x), _q(w)
    return F.conv2d(x, w, None if b is None else b.to(x.dtype), 1, pad)


def position_encoding(pos, lbase=1.25, levels=80, lfreq=math.pi):
    """PositionEncoding.forward — model_blocks.py:120-126.  `pe_bases` is an f32 tensor (model_blocks.py:115);
    the product promotes to pos.dtype-or-wider exactly as torch does in the reference."""
    bases = lbase ** torch.arange(int(levels)) * lfreq
    v = pos * bases
    return torch.cat([torch.sin(v), torch.cos(v)], dim=-1).view(pos.size(0), -1, 1, 1)


def mlp(sd, prefix, x, n_layers, act="sin"):
    """NeRV_MLP: 1x1 conv + activation after EVERY layer — model_blocks.py:66-71 (Sequential idx 0,2,..)."""
    for i in range(n_layers):
        x = _act(act)(conv(sd, f"{prefix}.{2 * i}", x, 0))
    return x


def sft_affine(sd, prefix, e):
    """SFTLayer scale/shift — model_blocks.py:103-104."""
    scale = conv(sd, prefix + ".SFT_scale_conv1", F.relu(conv(sd, prefix + ".SFT_scale_conv0", e, 0)), 0)
    shift = conv(sd, prefix + ".SFT_shift_conv1", F.relu(conv(sd, prefix + ".SFT_shift_conv0", e, 0)), 0)
    return scale, shift


def res_block_sft(sd, prefix, x0, e):
    """ResBlock_SFT.forward — model_blocks.py:83-89, with SFTLayer.forward :105 inlined."""
    s0, h0 = sft_affine(sd, prefix + ".sft0", e)
    fea = x0 * (s0 + 1) + h0
    fea = F.gelu(conv(sd, prefix + ".conv0", fea, 1))
    s1, h1 = sft_affine(sd, prefix + ".sft1", e)
    fea = fea * (s1 + 1) + h1
    return _q(x0) + conv(sd, prefix + ".conv1", fea, 1)


def up_conv(sd, prefix, x, stride):
    """UpConv 'pshuffel_3x3' — model_blocks.py:213-220: conv k (pad (k-1)//2) then nn.PixelShuffle(stride)."""
    k = sd[prefix + ".upconv.0.weight"].shape[-1]
    y = conv(sd, prefix + ".upconv.0", x, (k - 1) // 2)
    return F.pixel_shuffle(y, stride) if stride != 1 else y


def nerv_block(sd, prefix, x, e, stride, act="sin"):
    """NeRVBlock.forward with a (feature, embedding) tuple — model_blocks.py:34-46 (dec_block branch)."""
    if prefix + ".conv.downconv.weight" in sd:          # HNeRV_Boost decoder[0]: DownConv 'conv', ks=0, strd=1
        y = conv(sd, prefix + ".conv.downconv", x, 0)   # -> 1x1 conv, pad ceil(0/2)=0 — model_blocks.py:185
    else:
        y = up_conv(sd, prefix + ".conv", x, stride)
    x0 = _act(act)(y)                                    # norm is Identity for every shipped script
    return res_block_sft(sd, prefix + ".sft_block", x0, e)


def conv_up_block(sd, prefix, x, e, stride, act="sin"):
    """Conv_Up_Block.forward — model_enerv.py:95-99 (either ordering of the up-conv and the 3x3 conv)."""
    if prefix + ".conv1.upconv.0.weight" in sd:
        y = conv(sd, prefix + ".conv2", up_conv(sd, prefix + ".conv1", x, stride), 1)
    else:
        y = up_conv(sd, prefix + ".conv2", conv(sd, prefix + ".conv1", x, 1), stride)
    return res_block_sft(sd, prefix + ".sft_block", _act(act)(y), e)


def out_img(x):
    """OutImg 'tanh' — model_blocks.py:61."""
    return torch.tanh(x) * 0.5 + 0.5


def block_strides(cfg):
    """Per-block PixelShuffle factor: first block of a stage carries the stage stride — model_nerv.py:36."""
    out = []
    for strd, nblk in zip(cfg["dec_strds"], cfg["dec_blks"]):
        out += [strd] + [1] * (nblk - 1)
    return out


def hnerv_boost_decode(sd, cfg, img_embed, norm_idx, dtype=torch.float32):
    """HNeRV_Boost.forward_decoder — model_hnerv.py:264-277.  norm_idx is f64 (default collate of
    hnerv_utils.py:47), PE evaluated in f64 then cast (`.float()`, :267)."""
    sd = {k: v.to(dtype) for k, v in sd.items()}
    pe = position_encoding(norm_idx[:, None]).float().to(dtype)
    e = mlp(sd, "stem_t", pe, 2)
    x = img_embed.to(dtype)
    outs = [x]
    x = nerv_block(sd, "decoder.0", x, e, 1)
    outs.append(x)
    for i, s in enumerate(block_strides(cfg)):
        x = nerv_block(sd, f"decoder.{i + 1}", x, e, s)
        outs.append(x)
    k = sd["head_layer.weight"].shape[-1]
    return out_img(conv(sd, "head_layer", x, (k - 1) // 2)), outs


def nerv_boost_forward(sd, cfg, t_in, dtype=torch.float32):
    """NeRV_Boost.forward — model_nerv.py:45-61.  `input[:,None].float()` makes the PE f32."""
    sd = {k: v.to(dtype) for k, v in sd.items()}
    pe = position_encoding(t_in[:, None].float()).to(dtype)
    fc_h, fc_w = cfg["fc_hw"]
    x = mlp(sd, "stem", pe, 2)
    x = x.view(x.size(0), -1, fc_h, fc_w)
    e = mlp(sd, "stem_t", pe, 2)
    outs = []
    for i, s in enumerate(block_strides(cfg)):
        x = nerv_block(sd, f"layers.{i}", x, e, s)
        outs.append(x)
    return out_img(conv(sd, "head_layer", x, 0)), outs


def _linear(sd, prefix, x):
    b = sd.get(prefix + ".bias")
    return F.linear(x, sd[prefix + ".weight"], b)


def _attention(sd, prefix, x, heads):
    """Attention.forward — model_enerv.py:49-57 (scale = dim_head**-0.5, dim_head = 64)."""
    b, n, _ = x.shape
    q, k, v = [t.view(b, n, heads, -1).transpose(1, 2) for t in _linear(sd, prefix + ".to_qkv", x).chunk(3, dim=-1)]
    attn = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) * (64 ** -0.5), dim=-1)
    out = torch.matmul(attn, v).transpose(1, 2).reshape(b, n, -1)
    if prefix + ".to_out.0.weight" not in sd:           # heads == 1 and dim_head == dim -> nn.Identity (model_enerv.py:44-47)
        return out
    return _linear(sd, prefix + ".to_out.0", out)


def _transformer(sd, prefix, x, heads):
    """TransformerBlock.forward (prenorm=False) — model_enerv.py:59-71; FeedForward :19-30."""
    x = _attention(sd, prefix + ".attn", x, heads) + x
    y = _linear(sd, prefix + ".ffn.net.3", F.gelu(_linear(sd, prefix + ".ffn.net.0", x)))
    return y + x


def enerv_boost_forward(sd, cfg, t_in, dtype=torch.float32):
    """ENeRV_Boost.forward — model_enerv.py:279-317."""
    sd = {k: v.to(dtype) for k, v in sd.items()}
    fc_h, fc_w = cfg["fc_hw"]
    b = t_in.size(0)
    xy = torch.stack(torch.meshgrid(torch.arange(fc_h) / fc_h, torch.arange(fc_w) / fc_w, indexing="ij"), dim=0).flatten(1, 2)
    t = t_in[:, None].float()
    t_emb = mlp(sd, "stem_t", position_encoding(t).to(dtype), 2).view(b, -1)
    t_manip = mlp(sd, "t_branch", position_encoding(t).to(dtype), 2)
    xy_emb = torch.cat([position_encoding(xy[0][:, None]), position_encoding(xy[1][:, None])], dim=1).to(dtype)
    xy_emb = mlp(sd, "stem_xy", xy_emb, 1).view(1, fc_h * fc_w, -1).expand(b, -1, -1)
    xy_emb = _transformer(sd, "trans1", xy_emb, 1)
    emb = _transformer(sd, "trans2", xy_emb * t_emb[:, None, :], 8)
    emb = emb.reshape(b, fc_h, fc_w, emb.shape[-1]).permute(0, 3, 1, 2)
    x = mlp(sd, "toconv", emb, 1) if "toconv.0.weight" in sd else emb
    outs = [t_manip]
    for i, s in enumerate(block_strides(cfg)):
        if i == 0:
            x = conv_up_block(sd, "layers.0", x, t_manip, s)
        else:
            x = nerv_block(sd, f"layers.{i}", x, t_manip, s)
        outs.append(x)
    return out_img(conv(sd, "head_layer", x, 0)), outs


def psnr(a, b):
    """psnr_fn_single — hnerv_utils.py:400-403 (per frame, mean over batch)."""
    mse = ((a.double() - b.double()) ** 2).flatten(1).mean(1)
    return (-10 * torch.log10(mse + 1e-9)).mean().item()


def cfg_from_args(args):
    """The handful of structural fields the oracle needs, from a reference-style args namespace."""
    return {"dec_strds": list(args.dec_strds), "dec_blks": list(args.dec_blks),
            "fc_hw": tuple(int(v) for v in args.fc_hw.split("_")), "model": args.model}


def forward(model_name, sd, cfg, *inputs, dtype=torch.float32):
    if model_name == "HNeRV_Boost":
        return hnerv_boost_decode(sd, cfg, *inputs, dtype=dtype)
    if model_name == "NeRV_Boost":
        return nerv_boost_forward(sd, cfg, *inputs, dtype=dtype)
    if model_name == "ENeRV_Boost":
        return enerv_boost_forward(sd, cfg, *inputs, dtype=dtype)
    raise KeyError(model_name)
