"""Mint the compression-path golden (BASELINE.json config 5) by executing the UNMODIFIED reference.

    python tests/golden/make_golden_quant.py          (build container only: needs /root/reference)

A tiny HNeRV_Boost is built with the flags of scripts/compression/hnerv_boost.sh:14-16 (--quant --quantizer_w scale
--quantizer_b scale --quantizer_e scalebeta, 8 bits each), then driven the way train_nerv_compression.py drives it:
init_data() (:333) -> cal_params(DiffEntropyModel) (:354) -> forward_encoder / forward_embed_quant / forward_decoder
(:505-517) and once through forward(frame, entropy_model=...) (:356).  Saved to tests/golden/hnerv_tiny_quant.npz:
the state_dict (with *.weight_quantizer.scale, *.bias_quantizer.scale, embed_quantizer.{scale,beta}), inputs, the encoder
output, (code, quant, dequant) of the embedding, the image, the summed weight/bias bit estimate, the embedding bit
estimate, and every layer's dequant_w / dequant_b (what the decode path consumes, lib/quant_ops.py:40).

constriction / compressai are not installed; they only serve `real_bitrate` (ANS coder, lib/entropy_model.py:46-62), which
is stubbed and NOT part of the golden (parity unpinned for real bits).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
from make_golden import import_reference, pack  # noqa: E402

QUANT_FLAGS = dict(quant=True, quant_model_bit=8, quant_bias_bit=8, quant_embed_bit=8, quantizer_w="scale", quantizer_b="scale",
                   quantizer_e="scalebeta", per_channel_w=False, per_channel_b=False, per_channel_e=False)


def stub_entropy_deps():
    """constriction / compressai stand-ins: enough for lib/entropy_model.py to import and for the eval-mode real_bitrate call
    to return a number (its value is not compared anywhere)."""
    cons = types.ModuleType("constriction")
    stream = types.ModuleType("constriction.stream")
    model = types.ModuleType("constriction.stream.model")
    stack = types.ModuleType("constriction.stream.stack")
    model.QuantizedGaussian = lambda *a, **k: None

    class AnsCoder:
        def encode_reverse(self, message, entropy_model):
            self.n = len(message)

        def get_compressed(self):
            return np.zeros(self.n, dtype=np.uint32)
    stack.AnsCoder = AnsCoder
    cons.stream, stream.model, stream.stack = stream, model, stack
    cai, ans = types.ModuleType("compressai"), types.ModuleType("compressai.ans")
    ans.BufferedRansEncoder = ans.RansDecoder = object
    cai.ans = ans
    sys.modules.update({"constriction": cons, "constriction.stream": stream, "constriction.stream.model": model,
                        "constriction.stream.stack": stack, "compressai": cai, "compressai.ans": ans})


def drive(model, entropy_model, frame, t):
    """The call sequence of train_nerv_compression.py on one frame (eval mode)."""
    model.init_data()
    model.cal_params(entropy_model)
    enc = model.forward_encoder(frame)
    model.embed_quantizer.init_data(enc)                       # forward() does this on first use (model_hnerv.py:231)
    code_e, quant_e, deq_e = model.forward_embed_quant(enc, entropy_model)
    img, lst, _ = model.forward_decoder(deq_e, t)
    img_fwd, _, _ = model(frame, entropy_model=entropy_model, norm_idx=t)
    return dict(enc=enc, code_e=code_e, quant_e=quant_e, deq_e=deq_e, img=img, img_fwd=img_fwd,
                bits_wb=model.get_bitrate_sum("bitrate"), bits_e=model.bitrate_e_dict["bitrate"])


def main():
    mb, mn, me, mh = import_reference()
    stub_entropy_deps()
    from lib.entropy_model import DiffEntropyModel
    sys.path.insert(0, os.path.join(ROOT, "boosting-nerv_b200"))
    from bnerv_b200.config import tiny_args
    torch.manual_seed(1)
    a = tiny_args("HNeRV_Boost", **QUANT_FLAGS)
    m = mh.HNeRV_Boost(a).eval()
    strd = int(np.prod(a.dec_strds))
    fh, fw = [int(v) for v in a.fc_hw.split("_")]
    frame = torch.rand(2, 3, fh * strd, fw * strd)
    t = torch.tensor([1 / 8, 5 / 8], dtype=torch.float64)
    with torch.no_grad():
        out = drive(m, DiffEntropyModel(), frame, t)
    deq = {}
    for name, mod in m.named_modules():
        if getattr(mod, "dequant_w", None) is not None:
            deq["dq/" + name + ".weight"] = mod.dequant_w.detach().numpy()
        if getattr(mod, "dequant_b", None) is not None:
            deq["dq/" + name + ".bias"] = mod.dequant_b.detach().numpy()
    arrays = pack(m.state_dict(), frame=frame, t=t, **out)
    arrays.update(deq)
    path = os.path.join(HERE, "hnerv_tiny_quant.npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, {k: tuple(v.shape) for k, v in arrays.items() if not k.startswith(("sd/", "dq/"))},
          "bits w+b %.2f, bits e %.2f" % (float(out["bits_wb"]), float(out["bits_e"])))


if __name__ == "__main__":
    main()
