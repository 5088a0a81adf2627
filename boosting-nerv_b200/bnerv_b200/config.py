"""Argument namespaces for the reference's shipped presets and the model-size solver.

The reference builds every model from one flat argparse namespace (train_nerv_all.py:28-112) and
derives ``fc_dim`` — hence every channel width — from ``--modelsize`` by solving a quadratic
(train_nerv_all.py:194-217).  Both are restated here so configs can be built without the training
script.  Preset values come from scripts/regression/{bunny,UVG}/*.sh.
"""
from types import SimpleNamespace

import numpy as np

_COMMON = dict(
    embed="pe_1.25_80", lfreq="pi", fc_hw="9_16", ch_t=32, ks="0_3_3", dec_strds=[5, 2, 2, 2, 2],
    dec_blks=[1, 1, 2, 2, 2], reduce=2.0, lower_width=12, conv_type=["convnext", "pshuffel_3x3"], norm="none",
    act="sin", sft_block="res_sft", out_bias="tanh", outf="unify", quant=False, enc_strds=[], enc_dim="64_16",
    enc_blks=1, block_dim=128, saturate_stages=-1, interpolation=False, fc_dim=None,
)


def solve_fc_dim(args, final_size, full_data_length):
    """train_nerv_all.py:194-217: largest root of a*x^2 + b*x + (c - decoder_size) = 0, truncated."""
    if ("pe" in args.embed or "le" in args.embed) and "HNeRV_Boost" not in args.model:
        embed_param = 0.0
        embed_dim = int(args.embed.split("_")[-1]) * 2
        fc_param = np.prod([int(v) for v in args.fc_hw.split("_")])
    else:
        total_enc = np.prod(args.enc_strds)
        embed_hw = final_size / total_enc ** 2
        enc_dim1, ratio = [float(v) for v in args.enc_dim.split("_")]
        embed_dim = int(ratio * args.modelsize * 1e6 / full_data_length / embed_hw) if ratio < 1 else int(ratio)
        embed_param = float(embed_dim) / total_enc ** 2 * final_size * full_data_length
        if args.interpolation:
            embed_param /= 2
        args.enc_dim = f"{int(enc_dim1)}_{embed_dim}"
        fc_param = (np.prod(args.enc_strds) // np.prod(args.dec_strds)) ** 2 * 9
    decoder_size = args.modelsize * 1e6 - embed_param
    r = 1.0 / args.reduce
    k1, k2 = [int(v) for v in args.ks.split("_")[1:]]
    n_fix = len(args.dec_strds) if args.saturate_stages == -1 else args.saturate_stages
    a = r * sum(r ** (2 * i) * s ** 2 * min(2 * i + k1, k2) ** 2 for i, s in enumerate(args.dec_strds[:n_fix]))
    b = embed_dim * fc_param
    c = args.lower_width ** 2 * sum(s ** 2 * min(2 * (n_fix + i) + k1, k2) ** 2
                                    for i, s in enumerate(args.dec_strds[n_fix:]))
    return int(np.roots([a, b, c - decoder_size]).max())


def make_args(model, modelsize, height, width, n_frames, **overrides):
    a = SimpleNamespace(**{k: (list(v) if isinstance(v, list) else v) for k, v in _COMMON.items()})
    a.model, a.modelsize = model, modelsize
    a.__dict__.update(overrides)
    a.final_size, a.full_data_length = height * width, n_frames
    if a.fc_dim is None:
        a.fc_dim = solve_fc_dim(a, a.final_size, n_frames)
    return a


# BASELINE.json configs -> (model class name, args).  Sizes per SURVEY.md §0 item 2.
def preset(name):
    if name == "nerv_xs":      # C1: NeRV-Boost 0.75M, Bunny 720x1280 (scripts/regression/bunny/nerv_boost.sh)
        return make_args("NeRV_Boost", 0.375, 720, 1280, 64)
    if name == "nerv_xs_640":  # C1 at the literal 640x1280 (needs --fc_hw 8_16)
        return make_args("NeRV_Boost", 0.375, 640, 1280, 64, fc_hw="8_16")
    if name == "nerv_s":       # C2: NeRV-Boost 1.5M, 132 frames
        return make_args("NeRV_Boost", 0.8, 720, 1280, 132)
    if name == "nerv_s_640":   # C2 at the literal 640x1280 of BASELINE.json (--fc_hw 8_16)
        return make_args("NeRV_Boost", 0.8, 640, 1280, 132, fc_hw="8_16")
    if name == "enerv_m":      # C3: E-NeRV-Boost 10M, UVG 1080p (scripts/regression/UVG/enerv_boost.sh)
        return make_args("ENeRV_Boost", 4.3, 1080, 1920, 600, dec_strds=[5, 3, 2, 2, 2])
    if name == "hnerv_l":      # C4/C5: HNeRV-Boost 15M, UVG 1080p (scripts/regression/UVG/hnerv_boost.sh)
        return make_args("HNeRV_Boost", 13.6, 1080, 1920, 600, ks="0_1_5", reduce=1.2,
                         dec_strds=[5, 3, 2, 2, 2], enc_strds=[5, 3, 2, 2, 2])
    if name == "hnerv_m":      # 10M alternative
        return make_args("HNeRV_Boost", 9.1, 1080, 1920, 600, ks="0_1_5", reduce=1.2,
                         dec_strds=[5, 3, 2, 2, 2], enc_strds=[5, 3, 2, 2, 2])
    if name == "hnerv_bunny":  # HNeRV-Boost 1.5M on Bunny 720p (scripts/regression/bunny/hnerv_boost.sh)
        return make_args("HNeRV_Boost", 0.64, 720, 1280, 132, ks="0_1_5", reduce=1.2,
                         dec_strds=[5, 2, 2, 2, 2], enc_strds=[5, 2, 2, 2, 2])
    raise KeyError(name)


def tiny_args(model, **overrides):
    """Small configs with the full block structure, for tests and golden fixtures."""
    base = dict(fc_hw="2_4", dec_strds=[5, 2, 2], dec_blks=[1, 2, 1], lower_width=6, fc_dim=10, block_dim=32)
    if model == "HNeRV_Boost":
        base.update(ks="0_1_5", reduce=1.2, enc_strds=[5, 2, 2], enc_dim="16_16", fc_dim=13)
    base.update(overrides)
    strd = int(np.prod(base["dec_strds"]))
    fh, fw = [int(v) for v in base["fc_hw"].split("_")]
    return make_args(model, 0.0, fh * strd, fw * strd, 8, **base)
