"""Run the UNMODIFIED reference caller - `train_nerv_all.py main()` with --eval_only --eval_fps - on synthetic PNG frames,
either with the reference's own model modules or with the drop-in (`boosting-nerv_b200/model_*.py`) first on sys.path
(train_nerv_all.py:16-18 imports the classes by module name: that is the plug-in point).

    python tools/run_reference_caller.py {reference|dropin} WORKDIR [--model NeRV_Boost|ENeRV_Boost|HNeRV_Boost] [--frames N]

Writes WORKDIR/output/.../eval.csv (the reference's own Dump2CSV) and prints its path.  On a CUDA box the drop-in decodes
natively; without CUDA it is switched to its plain-torch wiring (`backend = "torch"`), which is what the CPU test compares
with the reference's modules.  Third-party imports of the caller that are not installed here are stubbed: imageio (GIF
dump, unused), dahuffman (Huffman table: the restatement of oracle/ptq_oracle.py), pytorch_msssim (oracle/msssim_oracle.py),
timm / decord (imported by the reference's model_blocks.py).  TEST / TOOL code: the oracle is used as a stand-in for absent
packages of the CALLER, never inside the product path.
"""
import argparse
import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("BNERV_REFERENCE", "/root/reference")


def install_stubs():
    sys.path.insert(0, ROOT)
    from oracle import msssim_oracle, ptq_oracle
    timm, tm, tl = types.ModuleType("timm"), types.ModuleType("timm.models"), types.ModuleType("timm.models.layers")
    tl.trunc_normal_ = torch.nn.init.trunc_normal_
    tl.DropPath = torch.nn.Identity
    timm.models, tm.layers = tm, tl
    dec = types.ModuleType("decord")
    dec.bridge = SimpleNamespace(set_bridge=lambda *a, **k: None)
    imageio = types.ModuleType("imageio")
    pm = types.ModuleType("pytorch_msssim")
    pm.ms_ssim, pm.ssim = msssim_oracle.ms_ssim, msssim_oracle.ssim
    dh = types.ModuleType("dahuffman")

    class HuffmanCodec:
        def __init__(self, table):
            self._table = table

        @classmethod
        def from_data(cls, data):
            vals, counts = np.unique(np.asarray(data), return_counts=True)
            lengths = ptq_oracle.huffman_code_lengths({int(v): int(c) for v, c in zip(vals, counts)})
            return cls({k: (int(l), 0) for k, l in lengths.items()})

        def get_code_table(self):
            return self._table
    dh.HuffmanCodec = HuffmanCodec
    sys.modules.update({"timm": timm, "timm.models": tm, "timm.models.layers": tl, "decord": dec, "imageio": imageio,
                        "pytorch_msssim": pm, "dahuffman": dh})


def write_frames(path, n, h, w):
    """Smooth moving pattern + a little noise (SURVEY.md §8d 'Bunny-synthetic'), 8-bit PNGs named 0001.png ..."""
    from PIL import Image
    os.makedirs(path, exist_ok=True)
    g = torch.Generator().manual_seed(1)
    yy, xx = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    for i in range(n):
        ch = [0.5 + 0.5 * torch.sin(2 * np.pi * ((1 + c) * xx / w + (2 - 0.5 * c) * yy / h + 0.13 * (c + 1) * i)) for c in range(3)]
        img = (torch.stack(ch, -1) + 0.05 * (torch.rand(h, w, 3, generator=g) - 0.5)).clamp(0, 1)
        Image.fromarray((img * 255).round().to(torch.uint8).numpy()).save(os.path.join(path, f"{i + 1:04d}.png"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", choices=["reference", "dropin"])
    ap.add_argument("workdir")
    ap.add_argument("--model", default="NeRV_Boost")
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--fps_off", action="store_true", help="omit --eval_fps (100 extra forwards per frame)")
    a = ap.parse_args()
    os.makedirs(a.workdir, exist_ok=True)
    install_stubs()
    data = os.path.join(a.workdir, "frames")
    if not os.path.isdir(data):
        write_frames(data, a.frames, 180, 320)
    sys.path.insert(0, REF)
    if a.which == "dropin":
        sys.path.insert(0, os.path.join(ROOT, "boosting-nerv_b200"))
        import bnerv_b200.models as M
        if not torch.cuda.is_available():
            M._BoostBase.backend = "torch"          # documented opt-out: the product path has no CPU fallback of its own
    import model_nerv
    print("model modules from:", os.path.dirname(os.path.abspath(model_nerv.__file__)), flush=True)
    os.chdir(a.workdir)
    argv = ["train_nerv_all.py", "--data_path", data, "--vid", "synthetic", "--outf", a.which, "--model", a.model,
            "--crop_list", "180_320", "--embed", "pe_1.25_80", "--fc_hw", "9_16", "--dec_strds", "5", "2", "2", "--dec_blks", "1", "1", "1",
            "--conv_type", "convnext", "pshuffel_3x3", "--act", "sin", "--norm", "none", "--sft_block", "res_sft", "--ch_t", "32",
            "--lower_width", "6", "--reduce", "2", "--modelsize", "0.05", "-b", "1", "-j", "0", "--eval_only", "--not_resume",
            "--manualSeed", "1", "--overwrite"]
    if a.model == "HNeRV_Boost":
        argv += ["--enc_strds", "5", "2", "2", "--enc_dim", "16_16", "--ks", "0_1_5", "--reduce", "1.2"]
    if not a.fps_off:
        argv.append("--eval_fps")
    sys.argv = argv
    import train_nerv_all
    train_nerv_all.main()
    csv = os.path.join(a.workdir, "output", a.which, "synthetic", "Size0.05", "eval.csv")
    print("CSV:", csv, flush=True)


if __name__ == "__main__":
    main()
