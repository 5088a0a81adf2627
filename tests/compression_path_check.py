"""Drop-in proof for the compression path (BASELINE.json config 5) - run as a subprocess by
tests/test_compression_path_cpu.py:

    python tests/compression_path_check.py [cpu|cuda]

sys.path = [<repo>/boosting-nerv_b200 (the drop-in model_*.py), /root/reference (its lib/ quantisers + entropy model)], i.e.
what a user of train_nerv_compression.py gets by putting the drop-in first on PYTHONPATH.  The drop-in HNeRV_Boost is built
with the flags of scripts/compression/hnerv_boost.sh:14-16 (--quant ... scale / scale / scalebeta), must expose the
reference's state_dict keys (*.weight_quantizer.scale, *.bias_quantizer.scale, embed_quantizer.{scale,beta}), load the
reference's state_dict strictly, and reproduce the reference's own results on the call sequence of
train_nerv_compression.py:333,354-361,505-517 - init_data -> cal_params(DiffEntropyModel) -> forward_encoder /
forward_embed_quant / forward_decoder and forward(frame, entropy_model=...) - against tests/golden/hnerv_tiny_quant.npz
(minted by tests/golden/make_golden_quant.py from the unmodified reference).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden_quant import QUANT_FLAGS, drive, stub_entropy_deps  # noqa: E402
import make_golden  # noqa: E402  (timm / decord stubs)


def main(device):
    # stubs for the imports the reference's lib/ and model files make at module scope
    import types
    from types import SimpleNamespace
    timm, tm, tl = types.ModuleType("timm"), types.ModuleType("timm.models"), types.ModuleType("timm.models.layers")
    tl.trunc_normal_ = torch.nn.init.trunc_normal_
    tl.DropPath = torch.nn.Identity
    timm.models, tm.layers = tm, tl
    sys.modules.update({"timm": timm, "timm.models": tm, "timm.models.layers": tl})
    dec = types.ModuleType("decord")
    dec.bridge = SimpleNamespace(set_bridge=lambda *a, **k: None)
    sys.modules["decord"] = dec
    stub_entropy_deps()
    sys.path.insert(0, REF)                                        # behind ...
    sys.path.insert(0, os.path.join(ROOT, "boosting-nerv_b200"))   # ... the drop-in
    import model_hnerv                                             # resolves to the drop-in
    assert os.path.dirname(os.path.abspath(model_hnerv.__file__)) == os.path.join(ROOT, "boosting-nerv_b200"), model_hnerv.__file__
    from lib.entropy_model import DiffEntropyModel                 # the reference's own
    import lib.quant_ops as rq
    from bnerv_b200.config import tiny_args
    from bnerv_b200.layers import CustomConv2d
    assert CustomConv2d is rq.CustomConv2d, "the drop-in must build on the reference's CustomConv2d when lib/ is importable"

    z = np.load(os.path.join(HERE, "golden", "hnerv_tiny_quant.npz"))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    a = tiny_args("HNeRV_Boost", **QUANT_FLAGS)
    m = model_hnerv.HNeRV_Boost(a).eval()
    own = m.state_dict()
    assert list(own.keys()) == list(sd.keys()), sorted(set(own) ^ set(sd))[:10]
    assert any(k.endswith("weight_quantizer.scale") for k in own) and any(k.endswith("bias_quantizer.scale") for k in own)
    assert "embed_quantizer.scale" in own and "embed_quantizer.beta" in own
    m.load_state_dict(sd, strict=True)
    frame, t = torch.from_numpy(z["frame"]), torch.from_numpy(z["t"])
    if device == "cpu":
        m.backend = "torch"
    else:
        m, frame, t = m.to(device), frame.to(device), t.to(device)
    # the quantisers of a loaded checkpoint are initialised: the golden's state_dict holds the scales init_data() produced
    for mod in m.modules():
        for qn in ("weight_quantizer", "bias_quantizer"):
            q = getattr(mod, qn, None)
            if q is not None:
                q.init = True
    with torch.no_grad():
        out = drive(m, DiffEntropyModel(), frame, t)
    rel = lambda x, y: ((x.double().cpu() - y.double()).abs().max() / y.double().abs().max().clamp_min(1e-12)).item()
    tol_img = 5e-6 if device == "cpu" else 1e-3
    res = {}
    for k in ("enc", "code_e", "quant_e", "deq_e", "img", "img_fwd"):
        res[k] = rel(out[k], torch.from_numpy(z[k]))
    res["bits_wb"] = abs(float(out["bits_wb"]) - float(z["bits_wb"])) / float(z["bits_wb"])
    res["bits_e"] = abs(float(out["bits_e"]) - float(z["bits_e"])) / float(z["bits_e"])
    print("compression path vs reference golden:", {k: f"{v:.2e}" for k, v in res.items()})
    assert res["enc"] < 1e-5 and res["deq_e"] < 1e-5 and res["code_e"] < 1e-5
    assert torch.equal(out["quant_e"].cpu(), torch.from_numpy(z["quant_e"])) or res["quant_e"] < 1e-6
    assert res["img"] < tol_img and res["img_fwd"] < tol_img, res
    assert res["bits_wb"] < 1e-6 and res["bits_e"] < 1e-5, res
    # every conv consumes dequant_w / dequant_b (lib/quant_ops.py:40): they equal the reference's
    n = 0
    for name, mod in m.named_modules():
        if getattr(mod, "dequant_w", None) is not None:
            assert rel(mod.dequant_w, torch.from_numpy(z["dq/" + name + ".weight"])) < 1e-6, name
            n += 1
    assert n > 20
    print("ok", n, "quantised layers")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "cpu")
