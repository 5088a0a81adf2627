"""Parameter containers for the Boosting-NeRV decoder family.

These modules own the parameters with exactly the reference's submodule names (so ``state_dict`` keys
and shapes match a reference checkpoint one to one) and carry a plain-torch ``forward`` that is used
only for autograd (training is outside the accelerated path) and as the explicit
``backend='torch'`` debugging path.  Inference on CUDA never runs these forwards: the decode engine
(``engine.py``) reads the parameters and drives the sm_100a kernels.

Reference map (file:line in Xinjie-Q/Boosting-NeRV):
  NeRVBlock            model_blocks.py:14-46      UpConv / DownConv   model_blocks.py:196-220 / 174-193
  ResBlock_SFT         model_blocks.py:74-89      SFTLayer            model_blocks.py:92-105
  PositionEncoding     model_blocks.py:108-126    NeRV_MLP            model_blocks.py:66-71
  OutImg               model_blocks.py:57-63      ConvNeXt encoder    model_blocks.py:223-347
  Conv_Up_Block        model_enerv.py:73-102      Attention/FFN       model_enerv.py:19-71
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

try:  # inside the reference tree: reuse its classes so `type(m) in [CustomConv2d, CustomLinear]` holds
    from lib.quant_ops import CustomConv2d, CustomLinear, quant_map  # type: ignore
    HAVE_REFERENCE_QUANT = True
except Exception:  # standalone: same constructor signature, no quantisers
    HAVE_REFERENCE_QUANT = False
    quant_map = {}

    def _no_quant(args):
        if getattr(args, "quant", False):
            raise RuntimeError("args.quant=True needs the reference's lib/transform_ops.py quantisers on sys.path "
                               "(out of scope for bnerv_b200; it only consumes their dequant_w / dequant_b)")

    class CustomConv2d(nn.Conv2d):
        """nn.Conv2d that prefers ``dequant_w`` / ``dequant_b`` when a quantiser has set them
        (contract of lib/quant_ops.py:18-41)."""

        def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True, **kw):
            super().__init__(in_channels, out_channels, kernel_size, stride, padding, bias=bias)
            _no_quant(kw["args"])
            self.dequant_w, self.dequant_b, self.quant = None, None, False

        def forward(self, x):
            w = self.weight if self.dequant_w is None else self.dequant_w
            b = self.bias if self.dequant_b is None else self.dequant_b
            return F.conv2d(x, w, b, self.stride, self.padding, self.dilation, self.groups)

    class CustomLinear(nn.Linear):
        """nn.Linear with the same ``dequant_*`` override (lib/quant_ops.py:43-65)."""

        def __init__(self, in_features, out_features, bias=True, **kw):
            super().__init__(in_features, out_features, bias=bias)
            _no_quant(kw["args"])
            self.dequant_w, self.dequant_b, self.quant = None, None, False

        def forward(self, x):
            w = self.weight if self.dequant_w is None else self.dequant_w
            b = self.bias if self.dequant_b is None else self.dequant_b
            return F.linear(x, w, b)


def effective_weight(m):
    """(weight, bias) a CustomConv2d / CustomLinear actually applies — lib/quant_ops.py:40."""
    w = m.weight if getattr(m, "dequant_w", None) is None else m.dequant_w
    b = m.bias if getattr(m, "dequant_b", None) is None else m.dequant_b
    return w, b


# ------------------------------------------------------------------------------------------------
# activations / norms / output squashing
# ------------------------------------------------------------------------------------------------
class Sin(nn.Module):
    def __init__(self, inplace=False):
        super().__init__()

    def forward(self, x):
        return torch.sin(x)


_ACTS = {
    "relu": lambda: nn.ReLU(True),
    "leaky": lambda: nn.LeakyReLU(inplace=True),
    "leaky01": lambda: nn.LeakyReLU(negative_slope=0.1, inplace=True),
    "relu6": lambda: nn.ReLU6(inplace=True),
    "gelu": lambda: nn.GELU(),
    "sin": lambda: Sin(),
    "swish": lambda: nn.SiLU(inplace=True),
    "softplus": lambda: nn.Softplus(),
    "hardswish": lambda: nn.Hardswish(inplace=True),
}


def ActivationLayer(act_type):
    if act_type not in _ACTS:
        raise KeyError(f"Unknown activation function {act_type}.")
    return _ACTS[act_type]()


def NormLayer(norm_type, ch_width):
    if norm_type == "none":
        return nn.Identity()
    if norm_type == "bn":
        return nn.BatchNorm2d(num_features=ch_width)
    if norm_type == "in":
        return nn.InstanceNorm2d(num_features=ch_width)
    raise NotImplementedError


def OutImg(x, out_bias="tanh"):
    if out_bias == "sigmoid":
        return torch.sigmoid(x)
    if out_bias == "tanh":
        return torch.tanh(x) * 0.5 + 0.5
    return x + float(out_bias)


def act_name(module):
    """Name the engine understands for an activation module, or None if unsupported natively."""
    if isinstance(module, Sin):
        return "sin"
    if isinstance(module, nn.GELU) and getattr(module, "approximate", "none") == "none":
        return "gelu"
    if isinstance(module, nn.ReLU):
        return "relu"
    if isinstance(module, nn.Identity):
        return "none"
    return None


# ------------------------------------------------------------------------------------------------
# embeddings
# ------------------------------------------------------------------------------------------------
class PositionEncoding(nn.Module):
    """cat(sin(pos*b^i*f), cos(pos*b^i*f)).  Kept in torch on purpose: frequencies reach 1e8 rad, so the
    op order and dtype promotion of the reference are reproduced literally (SURVEY.md §0 parity hazards)."""

    def __init__(self, pe_embed, lfreq):
        super().__init__()
        self.pe_embed = pe_embed
        if "pe" in pe_embed:
            lbase, levels = [float(v) for v in pe_embed.split("_")[-2:]]
            freq = math.pi if lfreq == "pi" else float(lfreq)
            self.pe_bases = lbase ** torch.arange(int(levels)) * freq
            self.embed_length = int(2 * levels)

    def forward(self, pos):
        if "pe" not in self.pe_embed:
            return pos
        if self.pe_bases.device != pos.device:
            self.pe_bases = self.pe_bases.to(pos.device)
        ang = pos * self.pe_bases
        return torch.cat([torch.sin(ang), torch.cos(ang)], dim=-1).view(pos.size(0), -1, 1, 1)


def NeRV_MLP(dim_list, act="relu", bias=True, omega=1.0, args=None):
    shared_act = ActivationLayer(act)
    seq = []
    for cin, cout in zip(dim_list[:-1], dim_list[1:]):
        seq.append(CustomConv2d(cin, cout, kernel_size=1, bias=bias, args=args))
        seq.append(shared_act)
    return nn.Sequential(*seq)


# ------------------------------------------------------------------------------------------------
# conv blocks
# ------------------------------------------------------------------------------------------------
def _same_pad(ks):
    return math.ceil((ks - 1) // 2)


class UpConv(nn.Module):
    def __init__(self, **kw):
        super().__init__()
        ks, ngf, new_ngf, strd, args = kw["ks"], kw["ngf"], kw["new_ngf"], kw["strd"], kw["args"]
        kind = kw["conv_type"]
        self.kind = kind
        if kind in ("pshuffel", "pshuffel_3x3"):
            if kind == "pshuffel_3x3":
                ks = min(ks, 3)
            self.upconv = nn.Sequential(
                CustomConv2d(ngf, new_ngf * strd * strd, ks, 1, _same_pad(ks), bias=kw["bias"], args=args),
                nn.PixelShuffle(strd) if strd != 1 else nn.Identity())
        elif kind == "conv":
            self.upconv = nn.ConvTranspose2d(ngf, new_ngf, ks + strd, strd, math.ceil(ks / 2))
        elif kind == "interpolate":
            self.upconv = nn.Sequential(
                nn.Upsample(scale_factor=strd, mode="bilinear"),
                CustomConv2d(ngf, new_ngf, strd + ks, 1, math.ceil((ks + strd - 1) / 2), bias=kw["bias"], args=args))

    def forward(self, x):
        return self.upconv(x)


class DownConv(nn.Module):
    def __init__(self, **kw):
        super().__init__()
        ks, ngf, new_ngf, strd, args = kw["ks"], kw["ngf"], kw["new_ngf"], kw["strd"], kw["args"]
        kind = kw["conv_type"]
        self.kind = kind
        if kind == "pshuffel":
            self.downconv = nn.Sequential(
                nn.PixelUnshuffle(strd) if strd != 1 else nn.Identity(),
                CustomConv2d(ngf * strd ** 2, new_ngf, ks, 1, _same_pad(ks), bias=kw["bias"], args=args))
        elif kind == "conv":
            self.downconv = CustomConv2d(ngf, new_ngf, ks + strd, strd, math.ceil(ks / 2), bias=kw["bias"], args=args)
        elif kind == "interpolate":
            self.downconv = nn.Sequential(
                nn.Upsample(scale_factor=1.0 / strd, mode="bilinear"),
                CustomConv2d(ngf, new_ngf, ks + strd, 1, math.ceil((ks + strd - 1) / 2), bias=kw["bias"], args=args))

    def forward(self, x):
        return self.downconv(x)


class SFTLayer(nn.Module):
    """Temporal-aware affine transform: per-frame, per-channel scale/shift from the time embedding."""

    def __init__(self, in_ch, out_ch, factor=1, act="relu", omega=1.0, args=None):
        super().__init__()
        hid = in_ch // factor
        self.SFT_scale_conv0 = CustomConv2d(in_ch, hid, 1, args=args)
        self.SFT_scale_conv1 = CustomConv2d(hid, out_ch, 1, args=args)
        self.SFT_shift_conv0 = CustomConv2d(in_ch, hid, 1, args=args)
        self.SFT_shift_conv1 = CustomConv2d(hid, out_ch, 1, args=args)
        self.act = ActivationLayer(act_type=act)

    def affine(self, cond):
        scale = self.SFT_scale_conv1(self.act(self.SFT_scale_conv0(cond)))
        shift = self.SFT_shift_conv1(self.act(self.SFT_shift_conv0(cond)))
        return scale, shift

    def forward(self, pair):
        fea, cond = pair
        scale, shift = self.affine(cond)
        return fea * (scale + 1) + shift


class ResBlock_SFT(nn.Module):
    def __init__(self, in_ch, out_ch, cond_ch, factor=1, in_act="relu", out_act="gelu", omega=1.0, args=None):
        super().__init__()
        self.sft0 = SFTLayer(cond_ch, in_ch, factor, in_act, omega, args=args)
        self.conv0 = CustomConv2d(in_ch, out_ch, kernel_size=3, stride=1, padding=1, args=args)
        self.sft1 = SFTLayer(cond_ch, out_ch, factor, in_act, omega, args=args)
        self.conv1 = CustomConv2d(out_ch, out_ch, kernel_size=3, stride=1, padding=1, args=args)
        self.act = ActivationLayer(act_type=out_act)

    def forward(self, pair):
        x0, cond = pair
        y = self.act(self.conv0(self.sft0((x0, cond))))
        y = self.conv1(self.sft1((y, cond)))
        return x0 + y


class NeRVBlock(nn.Module):
    def __init__(self, **kw):
        super().__init__()
        args = kw["args"]
        make = UpConv if kw["dec_block"] else DownConv
        self.conv = make(ngf=kw["ngf"], new_ngf=kw["new_ngf"], strd=kw["strd"], ks=kw["ks"],
                         conv_type=kw["conv_type"], bias=kw["bias"], args=args)
        self.norm = NormLayer(kw["norm"], kw["new_ngf"])
        self.act = ActivationLayer(kw["act"])
        self.dec_block = kw["dec_block"] or len(args.enc_strds)
        if args.sft_block == "res_sft" and kw["sft_ngf"] != 0:
            if self.dec_block:
                sft_ch = kw["new_ngf"]
            else:
                self.fc_h, self.fc_w = [int(v) for v in args.fc_hw.split("_")]
                sft_ch = int(kw["new_ngf"] / (self.fc_h * self.fc_w))
            self.sft_block = ResBlock_SFT(sft_ch, sft_ch, cond_ch=kw["sft_ngf"], in_act="relu", out_act="gelu",
                                          omega=1, args=args)

    def forward(self, x):
        if not isinstance(x, tuple):
            return self.act(self.norm(self.conv(x)))
        fea, cond = x
        x0 = self.act(self.norm(self.conv(fea)))
        if not self.dec_block:
            n, _, h, w = x0.shape
            x0 = x0.view(n, -1, self.fc_h, self.fc_w, h, w).permute(0, 1, 4, 2, 5, 3)
            x0 = x0.reshape(n, -1, self.fc_h * h, self.fc_w * w)
        return self.sft_block((x0, cond))


class Conv_Up_Block(nn.Module):
    """E-NeRV stage 0: a cheap up-conv at C/4 followed by a widening 3x3 conv (or the reverse order)."""

    def __init__(self, **kw):
        super().__init__()
        ngf, new_ngf, args = kw["ngf"], kw["new_ngf"], kw["args"]
        up = dict(ks=kw["ks"], strd=kw["stride"], bias=kw["bias"], conv_type=kw["conv_type"], args=args)
        if ngf <= new_ngf:
            self.conv1 = UpConv(ngf=ngf, new_ngf=ngf // 4, **up)
            self.conv2 = CustomConv2d(ngf // 4, new_ngf, 3, 1, 1, bias=kw["bias"], args=args)
        else:
            self.conv1 = CustomConv2d(ngf, new_ngf, 3, 1, 1, bias=kw["bias"], args=args)
            self.conv2 = UpConv(ngf=new_ngf, new_ngf=new_ngf, **up)
        self.norm = NormLayer(kw["norm"], new_ngf)
        self.act = ActivationLayer(kw["act"])
        self.use_sft = "sft" in args.sft_block
        if args.sft_block == "res_sft":
            self.sft_block = ResBlock_SFT(new_ngf, new_ngf, cond_ch=kw["sft_ngf"], in_act="relu", out_act="gelu",
                                          omega=1, args=args)

    def forward(self, x):
        if not isinstance(x, tuple):
            return self.act(self.norm(self.conv2(self.conv1(x))))
        fea, cond = x
        x0 = self.act(self.norm(self.conv2(self.conv1(fea))))
        return self.sft_block((x0, cond))


# ------------------------------------------------------------------------------------------------
# E-NeRV transformer stem (stays in torch: 0.1 GFLOP of 443)
# ------------------------------------------------------------------------------------------------
class PreNorm(nn.Module):
    def __init__(self, dim, fn):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.fn = fn

    def forward(self, x, **kw):
        return self.fn(self.norm(x), **kw)


class FeedForward(nn.Module):
    def __init__(self, dim, hidden_dim, dropout=0.0, args=None):
        super().__init__()
        self.net = nn.Sequential(CustomLinear(dim, hidden_dim, args=args), nn.GELU(), nn.Dropout(dropout),
                                 CustomLinear(hidden_dim, dim, args=args), nn.Dropout(dropout))

    def forward(self, x):
        return self.net(x)


class Attention(nn.Module):
    def __init__(self, dim, heads=8, dim_head=64, dropout=0.0, args=None):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.scale = heads, dim_head ** -0.5
        self.attend = nn.Softmax(dim=-1)
        self.to_qkv = CustomLinear(dim, inner * 3, bias=False, args=args)
        if heads == 1 and dim_head == dim:
            self.to_out = nn.Identity()
        else:
            self.to_out = nn.Sequential(CustomLinear(inner, dim, args=args), nn.Dropout(dropout))

    def forward(self, x):
        b, n, _ = x.shape
        q, k, v = [t.view(b, n, self.heads, -1).transpose(1, 2) for t in self.to_qkv(x).chunk(3, dim=-1)]
        attn = self.attend(torch.matmul(q, k.transpose(-1, -2)) * self.scale)
        out = torch.matmul(attn, v).transpose(1, 2).reshape(b, n, -1)
        return self.to_out(out)


class TransformerBlock(nn.Module):
    def __init__(self, dim, heads, dim_head, mlp_dim, dropout=0.0, prenorm=False, args=None):
        super().__init__()
        attn = Attention(dim, heads=heads, dim_head=dim_head, dropout=dropout, args=args)
        ffn = FeedForward(dim, mlp_dim, dropout=dropout, args=args)
        self.attn = PreNorm(dim, attn) if prenorm else attn
        self.ffn = PreNorm(dim, ffn) if prenorm else ffn

    def forward(self, x):
        x = self.attn(x) + x
        return self.ffn(x) + x


# ------------------------------------------------------------------------------------------------
# ConvNeXt encoder of HNeRV (encode-side only; torch; kept for checkpoint / API parity)
# ------------------------------------------------------------------------------------------------
class LayerNorm(nn.Module):
    def __init__(self, normalized_shape, eps=1e-6, data_format="channels_last"):
        super().__init__()
        if data_format not in ("channels_last", "channels_first"):
            raise NotImplementedError
        self.weight = nn.Parameter(torch.ones(normalized_shape))
        self.bias = nn.Parameter(torch.zeros(normalized_shape))
        self.eps, self.data_format, self.normalized_shape = eps, data_format, (normalized_shape,)

    def forward(self, x):
        if self.data_format == "channels_last":
            return F.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps)
        mu = x.mean(1, keepdim=True)
        var = (x - mu).pow(2).mean(1, keepdim=True)
        xn = (x - mu) / torch.sqrt(var + self.eps)
        return self.weight[:, None, None] * xn + self.bias[:, None, None]


class Block(nn.Module):
    def __init__(self, dim, drop_path=0.0, layer_scale_init_value=1e-6):
        super().__init__()
        self.dwconv = nn.Conv2d(dim, dim, kernel_size=7, padding=3, groups=dim)
        self.norm = LayerNorm(dim, eps=1e-6)
        self.pwconv1 = nn.Linear(dim, 4 * dim)
        self.act = nn.GELU()
        self.pwconv2 = nn.Linear(4 * dim, dim)
        self.gamma = nn.Parameter(layer_scale_init_value * torch.ones(dim)) if layer_scale_init_value > 0 else None
        if drop_path > 0.0:
            raise NotImplementedError("stochastic depth is never enabled by the reference scripts (drop_path_rate=0)")
        self.drop_path = nn.Identity()

    def forward(self, x):
        y = self.dwconv(x).permute(0, 2, 3, 1)
        y = self.pwconv2(self.act(self.pwconv1(self.norm(y))))
        if self.gamma is not None:
            y = self.gamma * y
        return x + y.permute(0, 3, 1, 2)


class ConvNeXt(nn.Module):
    def __init__(self, stage_blocks=0, strds=(2, 2, 2, 2), dims=(96, 192, 384, 768), in_chans=3,
                 drop_path_rate=0.0, layer_scale_init_value=1e-6):
        super().__init__()
        self.downsample_layers = nn.ModuleList()
        self.stages = nn.ModuleList()
        self.stage_num = len(dims)
        for i, (dim, strd) in enumerate(zip(dims, strds)):
            if i == 0:
                down = nn.Sequential(nn.Conv2d(in_chans, dim, kernel_size=strd, stride=strd),
                                     LayerNorm(dim, eps=1e-6, data_format="channels_first"))
            else:
                down = nn.Sequential(LayerNorm(dims[i - 1], eps=1e-6, data_format="channels_first"),
                                     nn.Conv2d(dims[i - 1], dim, kernel_size=strd, stride=strd))
            self.downsample_layers.append(down)
            self.stages.append(nn.Sequential(*[Block(dim=dim, drop_path=0.0,
                                                     layer_scale_init_value=layer_scale_init_value)
                                               for _ in range(stage_blocks)]))
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                nn.init.trunc_normal_(m.weight, std=0.02)
                nn.init.constant_(m.bias, 0)

    def forward(self, x):
        for down, stage in zip(self.downsample_layers, self.stages):
            x = stage(down(x))
        return x
