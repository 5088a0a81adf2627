"""NeRV_Boost / ENeRV_Boost / HNeRV_Boost with the reference's constructor and forward() API.

  NeRV_Boost(expansion, args)   model_nerv.py:11-96     forward(input, input_embed=None, norm_idx=None)
  ENeRV_Boost(expansion, args)  model_enerv.py:253-317  forward(input, input_embed=None, norm_idx=False)
  HNeRV_Boost(args)             model_hnerv.py:178-322  forward(...), forward_encoder, forward_embed_quant,
                                                         forward_decoder(img_embed, norm_idx)
All return ``(img_out [B,3,H,W] f32 in [0,1], list, dec_time)`` where ``dec_time`` is wall-clock seconds
including a device synchronise, exactly like the reference (model_nerv.py:46,58-60).

Dispatch (``model.backend``):
  'b200'  (default) under ``torch.no_grad()`` on a CUDA device the decoder runs on the sm_100a kernels
          through the C-ABI; CPU tensors or a missing extension raise — there is no CPU fallback.
          With autograd enabled (training, outside the accelerated scope) the torch forward is used.
  'torch' always the plain torch forward (debugging / CPU wiring tests).
Training (autograd enabled, CUDA tensors), ``model.train_backend``:
  'b200'  (default) the conv cascade runs forward AND backward on the sm_100a kernels (``train.cascade_train``:
          f16 operands, f32 accumulation, loss-scaled f16 gradient maps); the stems and SFT MLPs stay in torch autograd.
  'torch' autograd through the plain torch forward (fp32 / cuDNN), the reference's own arithmetic.
``model.keep_intermediates`` (default False): fill the returned list with every block output converted
to NCHW f32 like the reference does; callers only read element 0 (train_nerv_all.py:488,495), which is
always provided.
"""
import time
import weakref

import torch
import torch.nn as nn

from .layers import (CustomConv2d, CustomLinear, Conv_Up_Block, ConvNeXt, NeRV_MLP, NeRVBlock, OutImg,
                     PositionEncoding, TransformerBlock, quant_map)

_ENGINES = weakref.WeakKeyDictionary()   # model -> DecoderEngine (never deep-copied / pickled with the model)


def _engine_for(model):
    eng = _ENGINES.get(model)
    if eng is None:
        from .engine import DecoderEngine   # imports the C-ABI; raises if the extension is not built
        eng = DecoderEngine(model)
        _ENGINES[model] = eng
    return eng


class _BoostBase(nn.Module):
    """Bookkeeping shared by the three families (quantiser plumbing of model_nerv.py:62-96)."""
    backend = "b200"
    train_backend = "b200"
    keep_intermediates = False

    def _quant_layers(self):
        return [m for m in self.modules() if type(m) in (CustomConv2d, CustomLinear)]

    def cal_params(self, entropy_model=None):
        for m in self._quant_layers():
            code_w, quant_w, m.dequant_w = m.weight_quantizer(m.weight)
            if m.bias is not None:
                code_b, quant_b, m.dequant_b = m.bias_quantizer(m.bias)
            if entropy_model is not None:
                m.bitrate_w_dict.update(entropy_model.cal_bitrate(code_w, quant_w, self.training))
                if m.bias is not None:
                    m.bitrate_b_dict.update(entropy_model.cal_bitrate(code_b, quant_b, self.training))

    def get_bitrate_sum(self, name="bitrate"):
        total = 0
        for m in self._quant_layers():
            total += m.bitrate_w_dict[name]
            if name in m.bitrate_b_dict:
                total += m.bitrate_b_dict[name]
        return total

    def init_data(self):
        for m in self._quant_layers():
            m.weight_quantizer.init_data(m.weight)
            if m.bias is not None:
                m.bias_quantizer.init_data(m.bias)

    def decoder_params(self):
        return sum(p.data.nelement() for p in self.parameters()) / 1e6

    # -- dispatch --------------------------------------------------------------------------------
    def _use_engine(self, ref_tensor):
        if self.backend == "torch":
            return False
        if self.backend != "b200":
            raise ValueError(f"unknown backend {self.backend!r}")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return False            # training: the forward() bodies go on to _use_native_train
        if not ref_tensor.is_cuda:
            raise RuntimeError("bnerv_b200: the decode path runs only on a CUDA (sm_100a) device; got a CPU tensor. "
                               "Set model.backend = 'torch' explicitly for the plain-torch debugging path.")
        return True

    def _use_native_train(self, ref_tensor):
        """Autograd is on (the caller checked _use_engine first): native fwd+bwd for CUDA tensors unless opted out."""
        if self.backend == "torch" or self.train_backend == "torch":
            return False
        if self.train_backend != "b200":
            raise ValueError(f"unknown train_backend {self.train_backend!r}")
        if not ref_tensor.is_cuda:
            raise RuntimeError("bnerv_b200: native training runs only on a CUDA (sm_100a) device; got a CPU tensor.  Set "
                               "model.train_backend = 'torch' (or model.backend = 'torch') explicitly for torch autograd.")
        return torch.is_grad_enabled()

    def _cascade_train(self, x, cond):
        from .train import cascade_train
        return cascade_train(self.engine(), x, cond)

    @staticmethod
    def _finish(t0):
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        return time.time() - t0

    def engine(self):
        return _engine_for(self)

    def _own(self, t):
        """forward() hands out tensors the caller owns (reference semantics); the engine's graph outputs are static
        buffers that the next decode overwrites."""
        return t.clone() if self.engine().use_graph else t


def _decoder_widths(args, expansion):
    """Channel schedule of NeRV/E-NeRV stages (model_nerv.py:26-38): yields (stage, j, ngf, new_ngf, stride)."""
    ngf = args.fc_dim
    for i, stride in enumerate(args.dec_strds):
        if i == 0:
            new_ngf = int(ngf * expansion)
        else:
            new_ngf = int(max(ngf // (1 if stride == 1 else args.reduce), args.lower_width))
        for j in range(args.dec_blks[i]):
            yield i, j, ngf, new_ngf, (1 if j else stride)
            ngf = new_ngf


# ================================================================================================
class NeRV_Boost(_BoostBase):
    def __init__(self, expansion=1, args=None):
        super().__init__()
        self.encoder = nn.Identity()
        self.pe_t = PositionEncoding(args.embed, args.lfreq)
        self.fc_h, self.fc_w = [int(v) for v in args.fc_hw.split("_")]
        self.fc_dim = args.fc_dim
        self.stem = NeRV_MLP(dim_list=[self.pe_t.embed_length, 256, self.fc_h * self.fc_w * self.fc_dim],
                             bias=True, act=args.act, omega=1, args=args)
        self.stem_t = NeRV_MLP(dim_list=[int(self.pe_t.embed_length), int(args.ch_t * 2), args.ch_t],
                               bias=True, act=args.act, omega=1, args=args)
        _, ks1, ks2 = [int(v) for v in args.ks.split("_")]
        self.layers = nn.ModuleList()
        ngf = self.fc_dim
        for i, j, ngf, new_ngf, strd in _decoder_widths(args, expansion):
            self.layers.append(NeRVBlock(dec_block=True, conv_type=args.conv_type[1], ngf=ngf, new_ngf=new_ngf,
                                         ks=min(ks1 + 2 * i, ks2), strd=strd, bias=True, norm=args.norm,
                                         act=args.act, sft_ngf=args.ch_t, args=args, dump_features=False))
            ngf = new_ngf
        self.head_layer = CustomConv2d(ngf, 3, 1, 1, bias=True, args=args)
        self.out_bias, self.outf = args.out_bias, args.outf

    def decode(self, input):
        """Asynchronous decode on the native path: image only, no host sync, no timing.  The per-frame weight check is
        skipped: after changing weights call ``model.engine().sync_weights()`` (stream.decode_to_host / evaluate_* do it once
        per call; forward() checks on every call).  The returned image is overwritten by the next decode."""
        self._use_engine(input)
        return self.engine().decode((input,), False, check_weights=False)[0]

    def forward(self, input, input_embed=None, norm_idx=None):
        t0 = time.time()
        if self._use_engine(input):
            img, outs, _ = self.engine().decode((input,), True if self.keep_intermediates else "first")
            return self._own(img), [self._own(o) for o in outs], self._finish(t0)
        pe = self.pe_t(input[:, None].float())
        x = self.stem(pe).view(pe.size(0), self.fc_dim, self.fc_h, self.fc_w)
        cond = self.stem_t(pe)
        if self._use_native_train(input):
            img, first = self._cascade_train(x, cond)
            return img, [first], self._finish(t0)
        outs = []
        for layer in self.layers:
            x = layer((x, cond))
            outs.append(x)
        img = OutImg(self.head_layer(x), self.out_bias)
        return img, outs, self._finish(t0)


# ================================================================================================
class ENeRV_Boost(_BoostBase):
    def __init__(self, expansion=3, args=None):
        super().__init__()
        self.encoder = nn.Identity()
        self.pe_t = PositionEncoding(args.embed, args.lfreq)
        self.fc_h, self.fc_w = [int(v) for v in args.fc_hw.split("_")]
        self.fc_dim, self.block_dim = args.fc_dim, args.block_dim
        mlp_dim = args.block_dim // 2
        self.stem_t = NeRV_MLP(dim_list=[self.pe_t.embed_length, self.block_dim * 2, self.block_dim], act=args.act, args=args)
        self.pe_t_manipulate = PositionEncoding(args.embed, args.lfreq)
        self.t_branch = NeRV_MLP(dim_list=[self.pe_t_manipulate.embed_length, args.ch_t * 2, args.ch_t],
                                 act=args.act, args=args)
        self.pe_xy = PositionEncoding(args.embed, args.lfreq)
        self.stem_xy = NeRV_MLP(dim_list=[2 * self.pe_xy.embed_length, self.block_dim], act=args.act, args=args)
        self.trans1 = TransformerBlock(dim=self.block_dim, heads=1, dim_head=64, mlp_dim=mlp_dim, dropout=0.0,
                                       prenorm=False, args=args)
        self.trans2 = TransformerBlock(dim=self.block_dim, heads=8, dim_head=64, mlp_dim=mlp_dim, dropout=0.0,
                                       prenorm=False, args=args)
        if self.block_dim == self.fc_dim:
            self.toconv = nn.Identity()
        else:
            self.toconv = NeRV_MLP(dim_list=[self.block_dim, self.fc_dim], act=args.act, args=args)
        _, ks1, ks2 = [int(v) for v in args.ks.split("_")]
        self.layers = nn.ModuleList()
        ngf = self.fc_dim
        for i, j, ngf, new_ngf, strd in _decoder_widths(args, expansion):
            common = dict(ngf=ngf, new_ngf=new_ngf, ks=min(ks1 + 2 * i, ks2), bias=True, norm=args.norm, act=args.act,
                          conv_type=args.conv_type[1], sft_ngf=args.ch_t, args=args)
            if i == 0:
                self.layers.append(Conv_Up_Block(stride=strd, **common))
            else:
                self.layers.append(NeRVBlock(dec_block=True, strd=strd, **common))
            ngf = new_ngf
        self.head_layer = CustomConv2d(ngf, 3, 1, 1, bias=True, args=args)
        self.out_bias = args.out_bias
        # the reference's un-boosted base class carries these; the Boost subclass nulls them (model_enerv.py:257)
        self.t_layers, self.norm_layers = None, None

    def _xy_grid(self, device):
        """The normalised (y, x) grid of model_enerv.py:281-282.  It is frame-invariant, so it is built once per
        device (same values, same dtype) instead of on the host for every call - which also keeps the decode
        capturable into a CUDA graph (no pageable H2D copy on the stream)."""
        cache = self.__dict__.setdefault("_xy_cache", {})
        if device not in cache:
            ys = torch.arange(self.fc_h) / self.fc_h
            xs = torch.arange(self.fc_w) / self.fc_w
            cache[device] = torch.stack(torch.meshgrid(ys, xs, indexing="ij"), dim=0).flatten(1, 2).to(device)
        return cache[device]

    def _stem(self, input):
        """Everything ahead of the conv cascade (model_enerv.py:281-303): returns (emb NCHW, t_manipulate)."""
        b = input.size(0)
        xy = self._xy_grid(input.device)
        t = input[:, None].float()
        t_emb = self.stem_t(self.pe_t(t)).view(b, -1)
        t_manip = self.t_branch(self.pe_t_manipulate(t))
        xy_emb = torch.cat([self.pe_xy(xy[0][:, None]), self.pe_xy(xy[1][:, None])], dim=1)
        xy_emb = self.stem_xy(xy_emb).view(1, self.fc_h * self.fc_w, -1).expand(b, -1, -1)
        xy_emb = self.trans1(xy_emb)
        emb = self.trans2(xy_emb * t_emb[:, None, :])
        emb = emb.reshape(b, self.fc_h, self.fc_w, emb.shape[-1]).permute(0, 3, 1, 2)
        return self.toconv(emb), t_manip

    def decode(self, input):
        """Asynchronous decode on the native path (see NeRV_Boost.decode)."""
        self._use_engine(input)
        return self.engine().decode((input,), False, check_weights=False)[0]

    def forward(self, input, input_embed=None, norm_idx=False):
        use_engine = self._use_engine(input)
        t0 = time.time()
        if use_engine:
            img, outs, t_manip = self.engine().decode((input,), self.keep_intermediates)
            return self._own(img), [self._own(t_manip)] + [self._own(o) for o in outs], self._finish(t0)
        emb, t_manip = self._stem(input)
        if self._use_native_train(input):
            img, first = self._cascade_train(emb, t_manip)
            return img, [t_manip, first], self._finish(t0)
        x, outs = emb, [t_manip]
        for layer in self.layers:
            x = layer((x, t_manip))
            outs.append(x)
        img = OutImg(self.head_layer(x), self.out_bias)
        return img, outs, self._finish(t0)


# ================================================================================================
class HNeRV_Boost(_BoostBase):
    def __init__(self, args):
        super().__init__()
        self.embed = args.embed
        _, ks1, ks2 = [int(v) for v in args.ks.split("_")]
        enc_dim1, enc_dim2 = [int(v) for v in args.enc_dim.split("_")]
        dims = [enc_dim1] * len(args.enc_strds)
        dims[-1] = enc_dim2
        self.encoder = ConvNeXt(stage_blocks=args.enc_blks, strds=args.enc_strds, dims=dims, drop_path_rate=0)
        self.pe_embed_t = PositionEncoding(args.embed, args.lfreq)
        self.stem_t = NeRV_MLP(dim_list=[int(self.pe_embed_t.embed_length), int(args.ch_t * 2), args.ch_t],
                               bias=True, act=args.act, omega=1, args=args)
        ngf = args.fc_dim
        blocks = [NeRVBlock(dec_block=False, conv_type="conv", ngf=enc_dim2, new_ngf=ngf, ks=0, strd=1, bias=True,
                            norm=args.norm, act=args.act, sft_ngf=args.ch_t, args=args)]
        for i, strd in enumerate(args.dec_strds):
            reduction = strd ** 0.5 if args.reduce == -1 else args.reduce
            new_ngf = int(max(round(ngf / reduction), args.lower_width))
            for j in range(args.dec_blks[i]):
                blocks.append(NeRVBlock(dec_block=True, conv_type=args.conv_type[1], ngf=ngf, new_ngf=new_ngf,
                                        ks=min(ks1 + 2 * i, ks2), strd=1 if j else strd, bias=True, norm=args.norm,
                                        act=args.act, sft_ngf=args.ch_t, args=args))
                ngf = new_ngf
        self.decoder = nn.ModuleList(blocks)
        self.head_layer = CustomConv2d(ngf, 3, 3, 1, 1, args=args)
        self.out_bias, self.outf = args.out_bias, args.outf
        if args.quant:
            self.embed_quantizer = quant_map[args.quantizer_e](args.quant_embed_bit, signed=False,
                                                               per_channel=args.per_channel_e)
            self.bitrate_e_dict = {}
        else:
            self.embed_quantizer = None

    def decoder_params(self):
        n_all = sum(p.data.nelement() for p in self.parameters())
        n_enc = sum(p.data.nelement() for p in self.encoder.parameters())
        return (n_all - n_enc) / 1e6

    def _encode(self, frame):
        """ConvNeXt encoder (model_hnerv.py:189): the native f32 forward under the decode path's dispatch rule (no autograd,
        CUDA tensors, backend 'b200'; CPU tensors raise), the torch module when training or with backend 'torch'."""
        if self._use_engine(frame):
            from .encoder import convnext_forward
            return convnext_forward(self.encoder, frame)
        return self.encoder(frame)

    def forward_encoder(self, input):
        return self._encode(input)

    def forward_embed_quant(self, img_embed, entropy_model=None):
        code, quant, img_embed = self.embed_quantizer(img_embed)
        if entropy_model is not None:
            self.bitrate_e_dict.update(entropy_model.cal_bitrate(code, quant, self.training))
        return code, quant, img_embed

    def decode(self, img_embed, norm_idx):
        """Asynchronous decode on the native path (see NeRV_Boost.decode)."""
        self._use_engine(img_embed)
        return self.engine().decode((img_embed, norm_idx), False, check_weights=False)[0]

    def forward_decoder(self, img_embed, norm_idx):
        use_engine = self._use_engine(img_embed)
        t0 = time.time()
        if use_engine:
            img, outs, _ = self.engine().decode((img_embed, norm_idx), self.keep_intermediates)
            return self._own(img), [img_embed] + [self._own(o) for o in outs], self._finish(t0)
        pe = self.pe_embed_t(norm_idx[:, None]).float()          # f64 PE -> f32, model_hnerv.py:267
        cond = self.stem_t(pe)
        if self._use_native_train(img_embed):
            img, first = self._cascade_train(img_embed, cond)
            return img, [img_embed, first], self._finish(t0)
        x, outs = img_embed, [img_embed]
        for blk in self.decoder:
            x = blk((x, cond))
            outs.append(x)
        img = OutImg(self.head_layer(x), self.out_bias)
        return img, outs, self._finish(t0)

    def forward(self, input, input_embed=None, entropy_model=None, pre_img=None, post_img=None, norm_idx=None):
        img_embed = input_embed if input_embed is not None else self._encode(input)
        if self.embed_quantizer is not None:
            self.embed_quantizer.init_data(img_embed)
            code_e, quant_e, img_embed = self.embed_quantizer(img_embed)
            if entropy_model is not None:
                self.bitrate_e_dict.update(entropy_model.cal_bitrate(code_e, quant_e, self.training))
        if pre_img is not None and post_img is not None:
            img_embed = 0.5 * (self._encode(pre_img) + self._encode(post_img))
        return self.forward_decoder(img_embed, norm_idx)


class _OutOfScope(nn.Module):
    """HNeRV / HNeRVDecoder (model_hnerv.py:11-175) are the un-boosted baseline comparators; they are
    importable so `from model_hnerv import HNeRV, HNeRVDecoder, HNeRV_Boost` keeps working, but are not
    part of the accelerated path."""

    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError(f"{type(self).__name__} is outside the bnerv_b200 scope (un-boosted baseline); "
                                  "use the reference implementation for it")


class HNeRV(_OutOfScope):
    pass


class HNeRVDecoder(_OutOfScope):
    pass
