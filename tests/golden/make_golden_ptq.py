"""Mint PTQ golden vectors by executing the UNMODIFIED reference quant_tensor / dequant_tensor (hnerv_utils.py:101-134,
185-188) on CPU.  Build container only (the reference tree does not exist on the GPU box):
    python tests/golden/make_golden_ptq.py
Writes tests/golden/ptq.npz: per case the seeded input, the uint8 codes, the min / scale tables (f32 for the whole-tensor
candidate, f16 for a per-axis one) and the dequantised tensor, as produced by the reference.

hnerv_utils.py imports pytorch_msssim at module scope (:8); it is not installed and not used by these two functions, so
an empty sys.modules stub is registered first.  torchvision and PIL (its other imports) are installed.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"

# (name, shape, bits, generator): shapes of the tensors train_nerv_all.py:627-637 / :542 feed through quant_tensor
CASES = [
    ("conv_up", (96, 24, 3, 3), 8, "uniform"),          # OIHW up-conv weight: axis 0 is a candidate (96 > 50)
    ("conv_wide", (64, 72, 3, 3), 8, "uniform"),        # axes 0 and 1 both candidates
    ("conv_small", (12, 12, 3, 3), 8, "uniform"),       # no per-axis candidate
    ("bias", (135,), 8, "normal"),                      # 1-D: the axis-0 table is a single value (1/135 < 0.02)
    ("bias_small", (12,), 8, "normal"),
    ("sft", (60, 32, 1, 1), 8, "normal"),               # SFT 1x1 conv
    ("embed6", (64, 16, 2, 4), 6, "normal"),            # frame embeddings, --quant_embed_bit 6 (train_nerv_all.py:91)
    ("embed8", (132, 16, 3, 4), 8, "skewed"),           # per-frame ranges differ a lot -> per-axis wins
    ("linear", (280, 160), 8, "normal"),
    ("bits4", (70, 8, 3, 3), 4, "uniform"),
]


def make_input(shape, kind, seed):
    g = torch.Generator().manual_seed(seed)
    if kind == "uniform":
        return (torch.rand(shape, generator=g) - 0.5) * 0.2
    if kind == "normal":
        return torch.randn(shape, generator=g) * 0.3
    t = torch.randn(shape, generator=g)
    scale = torch.exp(torch.randn((shape[0],) + (1,) * (len(shape) - 1), generator=g) * 1.5)
    return t * scale


def main():
    pm = types.ModuleType("pytorch_msssim")
    pm.ms_ssim = pm.ssim = None
    sys.modules["pytorch_msssim"] = pm
    sys.path.insert(0, REF)
    import hnerv_utils as hu

    out = {}
    for i, (name, shape, bits, kind) in enumerate(CASES):
        t = make_input(shape, kind, 100 + i)
        q, new_t = hu.quant_tensor(t, bits)
        out[f"{name}/t"] = t.numpy()
        out[f"{name}/bits"] = np.int64(bits)
        out[f"{name}/quant"] = q["quant"].numpy()
        out[f"{name}/min"] = q["min"].numpy()
        out[f"{name}/scale"] = q["scale"].numpy()
        out[f"{name}/new_t"] = new_t.numpy()
        out[f"{name}/dequant"] = hu.dequant_tensor(q).numpy()
        print(name, shape, bits, "min table", tuple(q["min"].shape), q["min"].dtype)
    np.savez_compressed(os.path.join(HERE, "ptq.npz"), **out)


if __name__ == "__main__":
    main()
