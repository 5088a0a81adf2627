"""Mint golden vectors by executing the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the reference tree does not exist on the GPU box):
    python tests/golden/make_golden.py
Writes tests/golden/{nerv,enerv,hnerv}_tiny.npz and blocks.npz.  Each file holds the reference
state_dict (keys prefixed 'sd/'), the inputs, every block output and the final image, all float32
(float64 for norm_idx), produced under torch.manual_seed(1) (--manualSeed default, train_nerv_all.py:100).

The reference imports timm and decord at module scope (model_blocks.py:8-10); neither is installed, so
sys.modules stubs are registered first (trunc_normal_ = torch's, DropPath = identity, decord.bridge no-op).
"""
import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


def import_reference():
    timm, tm, tl = types.ModuleType("timm"), types.ModuleType("timm.models"), types.ModuleType("timm.models.layers")
    tl.trunc_normal_ = torch.nn.init.trunc_normal_

    class DropPath(torch.nn.Module):
        def __init__(self, p=0.0):
            super().__init__()

        def forward(self, x):
            return x
    tl.DropPath = DropPath
    timm.models, tm.layers = tm, tl
    sys.modules.update({"timm": timm, "timm.models": tm, "timm.models.layers": tl})
    dec = types.ModuleType("decord")
    dec.bridge = SimpleNamespace(set_bridge=lambda *a, **k: None)
    sys.modules["decord"] = dec
    sys.path.insert(0, REF)
    import model_blocks, model_enerv, model_hnerv, model_nerv  # noqa: E401
    return model_blocks, model_nerv, model_enerv, model_hnerv


def pack(sd, **arrays):
    out = {"sd/" + k: v.detach().numpy() for k, v in sd.items()}
    out.update({k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrays.items()})
    return out


def main():
    mb, mn, me, mh = import_reference()
    sys.path.insert(0, os.path.join(ROOT, "boosting-nerv_b200"))
    from bnerv_b200.config import tiny_args   # args namespaces only (no model code from the new repo is used)

    torch.manual_seed(1)
    # ---- NeRV_Boost ---------------------------------------------------------------------------
    a = tiny_args("NeRV_Boost")
    m = mn.NeRV_Boost(1, a).eval()
    t = torch.tensor([1 / 8, 5 / 8], dtype=torch.float64)
    with torch.no_grad():
        img, outs, _ = m(t)
    np.savez(os.path.join(HERE, "nerv_tiny.npz"), **pack(m.state_dict(), t=t, img=img, **{f"out{i}": o for i, o in enumerate(outs)}))
    # ---- ENeRV_Boost --------------------------------------------------------------------------
    a = tiny_args("ENeRV_Boost")
    m = me.ENeRV_Boost(3, a).eval()
    with torch.no_grad():
        img, outs, _ = m(t)
    np.savez(os.path.join(HERE, "enerv_tiny.npz"), **pack(m.state_dict(), t=t, img=img, **{f"out{i}": o for i, o in enumerate(outs)}))
    # ---- HNeRV_Boost (decoder path + full forward through the ConvNeXt encoder) ---------------
    a = tiny_args("HNeRV_Boost")
    m = mh.HNeRV_Boost(a).eval()
    emb = torch.rand(2, 16, 2, 4)
    frame = torch.rand(2, 3, 40, 80)
    with torch.no_grad():
        img, outs, _ = m.forward_decoder(emb, t)
        enc = m.forward_encoder(frame)
        img_full, _, _ = m(frame, norm_idx=t)
    np.savez(os.path.join(HERE, "hnerv_tiny.npz"), **pack(m.state_dict(), t=t, emb=emb, img=img, frame=frame, enc=enc,
                                                           img_full=img_full, **{f"out{i}": o for i, o in enumerate(outs)}))
    # ---- single blocks with the awkward channel counts of the real configs -------------------
    blocks = {}
    ns = SimpleNamespace(enc_strds=[], sft_block="res_sft", fc_hw="9_16", quant=False)
    cases = [  # (name, ngf, new_ngf, ks, stride, H, W, batch)
        ("s5_k3", 15, 15, 3, 5, 3, 4, 1), ("s2_k3", 27, 13, 3, 2, 9, 7, 2), ("s1_k3", 12, 12, 3, 1, 17, 33, 1),
        ("s3_k3", 21, 43, 3, 3, 5, 6, 1), ("s5_k1", 24, 21, 1, 5, 4, 3, 2), ("s2_wide", 85, 43, 3, 2, 6, 10, 1),
    ]
    for name, ngf, new_ngf, ks, s, H, W, B in cases:
        blk = mb.NeRVBlock(dec_block=True, conv_type="pshuffel_3x3", ngf=ngf, new_ngf=new_ngf, ks=ks, strd=s, bias=True,
                           norm="none", act="sin", sft_ngf=32, args=ns).eval()
        x = torch.randn(B, ngf, H, W)
        e = torch.randn(B, 32, 1, 1)
        with torch.no_grad():
            y = blk((x, e))
            x0 = blk.act(blk.norm(blk.conv(x)))
        for k, v in blk.state_dict().items():
            blocks[f"{name}/sd/{k}"] = v.numpy()
        blocks[f"{name}/x"], blocks[f"{name}/e"], blocks[f"{name}/y"], blocks[f"{name}/x0"] = x.numpy(), e.numpy(), y.numpy(), x0.numpy()
        blocks[f"{name}/meta"] = np.array([ngf, new_ngf, ks, s, H, W, B])
    # PixelShuffle permutation (bit-exact gate) and PositionEncoding in both dtype paths
    xs = torch.randn(2, 3 * 25, 4, 5)
    blocks["ps5/x"], blocks["ps5/y"] = xs.numpy(), torch.nn.PixelShuffle(5)(xs).numpy()
    pe = mb.PositionEncoding("pe_1.25_80", "pi")
    idx64 = torch.tensor([(i + 1) / 132 for i in (0, 17, 131)], dtype=torch.float64)
    blocks["pe/idx"] = idx64.numpy()
    blocks["pe/f64_then_f32"] = pe(idx64[:, None]).float().numpy()          # HNeRV_Boost path, model_hnerv.py:267
    blocks["pe/f32"] = pe(idx64[:, None].float()).numpy()                   # NeRV / E-NeRV path, model_nerv.py:47
    np.savez(os.path.join(HERE, "blocks.npz"), **blocks)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
