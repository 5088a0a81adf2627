"""GPU: the higher-precision ("split", hi + lo f16) conv form - bnerv_conv_fused_split and DecoderEngine.set_precise().

f16 operands keep 11 significant bits; on reference-TRAINED weights the intermediate maps of the plain decode sit 1.5e-3..3.7e-3
from the f32 reference (tests/test_gpu_models.py).  The split form carries activations and weights as f16 pairs (~22 bits) through
the same tcgen05 kernel - 3x the tensor work of the selected blocks - and must bring those maps far inside north_star's 1e-3."""
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, max_rel
from oracle import nerv_oracle as orc
from bnerv_b200 import ENeRV_Boost, HNeRV_Boost, NeRV_Boost, ops, tiny_args

pytestmark = pytest.mark.gpu


def _split_map(x):
    """[B,C,H,W] f32 -> split C8 map [B, 3*Cp/8, H, W, 8] (hi | lo | hi)."""
    xp = F.pad(x, (0, 0, 0, 0, 0, ops.round_up(x.shape[1], 16) - x.shape[1]))
    hi = xp.half().float()
    return ops.nchw_to_c8(torch.cat([hi, xp - hi, hi], dim=1))


def _unsplit(m, c):
    g = m.shape[1] // 3
    assert torch.equal(m[:, :g], m[:, 2 * g:])                       # the two hi blocks are the same map
    return ops.c8_to_nchw(m[:, :g].contiguous(), c).double() + ops.c8_to_nchw(m[:, g:2 * g].contiguous(), c).double()


@pytest.mark.parametrize("cin,cout,k,s,act", [(72, 40, 3, 1, "none"), (72, 40, 3, 1, "gelu"), (40, 24, 3, 2, "sin"), (24, 12, 3, 3, "none"),
                                              (12, 12, 3, 1, "none"), (160, 136, 1, 1, "none")])
def test_split_conv_against_f64(cin, cout, k, s, act):
    """One conv in both forms against an f64 reference: the split form is >= 50x closer than the plain one and inside 2e-5 of the
    f64 result (pre-activation: ~2^-21 operands, f32 accumulation); its outputs, split again, feed a second split conv with a
    split residual (the ResBlock_SFT shape) to the same accuracy."""
    torch.manual_seed(cin * 7 + s)
    B, H, W = 2, 19, 27
    x = (torch.randn(B, cin, H, W) * 2).cuda()
    w = (torch.randn(cout * s * s, cin, k, k) / (cin * k * k) ** 0.5).cuda()
    b = torch.randn(cout * s * s).cuda()
    g1p, beta = (1 + 0.3 * torch.randn(B, ops.round_up(cout, 16))).cuda(), torch.randn(B, ops.round_up(cout, 16)).cuda()
    ref_pre = F.conv2d(x.double(), w.double(), b.double(), padding=k // 2)
    if s > 1:
        ref_pre = F.pixel_shuffle(ref_pre, s)
    ref = {"none": ref_pre, "gelu": F.gelu(ref_pre), "sin": torch.sin(ref_pre)}[act]
    ref_aff = ref * g1p[:, :cout, None, None].double() + beta[:, :cout, None, None].double()
    Ho, Wo, cp = H * s, W * s, ops.round_up(cout, 16)
    # plain form
    pc = ops.PackedConv(w, b, s)
    pre = torch.zeros(ops.c8_shape(B, cout, Ho, Wo), dtype=torch.float16, device="cuda")
    ops.conv_fused(ops.nchw_to_c8(x), pc, cin, H, W, act=act, out_pre=pre)
    err_plain = max_rel(ops.c8_to_nchw(pre, cout).double(), ref)
    # split form
    ps = ops.PackedConv(w, b, s, split_in=True)
    assert ps.cin == 3 * ops.round_up(cin, 16)
    pre_s = torch.zeros(ops.c8_shape(B, 3 * cp, Ho, Wo), dtype=torch.float16, device="cuda")
    aff_s = torch.zeros_like(pre_s)
    ops.conv_fused(_split_map(x), ps, ps.cin, H, W, act=act, g1p=g1p, beta=beta, out_pre=pre_s, out_aff=aff_s, split=1)
    err_split, err_aff = max_rel(_unsplit(pre_s, cout), ref), max_rel(_unsplit(aff_s, cout), ref_aff)
    print(f"cin {cin} cout {cout} k {k} s {s} {act}: plain {err_plain:.2e}, split {err_split:.2e} (affine output {err_aff:.2e})")
    assert err_split < 2e-5 and err_aff < 2e-5 and err_split * 50 < err_plain
    # second conv over the split affine map, split residual = the first conv's pre map, plain and split outputs
    w2 = (torch.randn(cout, cout, 3, 3) / (cout * 9) ** 0.5).cuda()
    b2 = torch.randn(cout).cuda()
    p2 = ops.PackedConv(w2, b2, 1, split_in=True)
    ref2 = F.conv2d(ref_aff, w2.double(), b2.double(), padding=1) + ref
    out_plain = torch.zeros(ops.c8_shape(B, cout, Ho, Wo), dtype=torch.float16, device="cuda")
    out_split = torch.zeros_like(pre_s)
    ops.conv_fused(aff_s, p2, 3 * cp, Ho, Wo, act="none", resid=pre_s, out_pre=out_plain, split=2)
    ops.conv_fused(aff_s, p2, 3 * cp, Ho, Wo, act="none", resid=pre_s, out_pre=out_split, split=3)
    e_plain, e_split = max_rel(ops.c8_to_nchw(out_plain, cout).double(), ref2), max_rel(_unsplit(out_split, cout), ref2)
    assert torch.equal(out_plain, out_split[:, :cp // 8])            # the hi block IS the plain f16 output
    assert e_split < 3e-5 and e_plain < 6e-4, (e_split, e_plain)      # plain output: only its own f16 rounding (2^-11 of max)


def test_split_conv_argument_checks():
    from bnerv_b200 import _capi
    x = torch.zeros(ops.c8_shape(1, 16, 8, 8), dtype=torch.float16, device="cuda")
    pc = ops.PackedConv(torch.zeros(16, 16, 3, 3).cuda(), None, 1)
    out = torch.zeros(ops.c8_shape(1, 48, 8, 8), dtype=torch.float16, device="cuda")
    img = torch.zeros(1, 16, 8, 8, device="cuda")
    lib, ptr = _capi.lib, ops.ptr
    call = lambda resid, o, nchw, split: lib.bnerv_conv_fused_split(ptr(x), 1, 16, 8, 8, ptr(pc.w), ptr(pc.b), 16, 3, 1, 0, ptr(resid), None,
                                                                    None, ptr(o), None, ptr(nchw), split, None)
    assert call(None, out, None, 2) == _capi.E_BADARG               # split residual without a residual
    assert call(None, None, img, 1) == _capi.E_BADARG               # split output without a C8 output
    assert call(None, out, None, 4) == _capi.E_BADARG
    assert call(None, out, None, 1) == 0
    torch.cuda.synchronize()


@pytest.mark.parametrize("model,gold", [("HNeRV_Boost", "hnerv_tiny_trained.npz"), ("NeRV_Boost", "nerv_tiny_trained.npz"),
                                        ("ENeRV_Boost", "enerv_tiny_trained.npz")])
def test_precise_decode_brings_trained_golden_intermediates_inside_1e_3(model, gold):
    """VERDICT r1 2b.  The reference-trained goldens' block outputs, 1.5e-3..3.7e-3 from f32 in the plain decode, are inside 1e-3
    (measured: ~1e-5 where the whole path is split) with every block in the split form; switching it off restores the plain path
    bit for bit."""
    sd, g = load_golden(gold)
    a = tiny_args(model)
    m = NeRV_Boost(1, a) if model == "NeRV_Boost" else ENeRV_Boost(3, a) if model == "ENeRV_Boost" else HNeRV_Boost(a)
    m.load_state_dict(sd, strict=True)
    m = m.eval().cuda()
    m.keep_intermediates = True
    run = lambda: m.forward_decoder(g["emb"].cuda(), g["t"].cuda()) if model == "HNeRV_Boost" else m(g["t"].cuda())
    with torch.no_grad():
        img0, outs0, _ = run()
        img0, outs0 = img0.clone(), [o.clone() for o in outs0]
        plain = [max_rel(o.cpu(), g[f"out{i}"]) for i, o in enumerate(outs0)]
        m.engine().set_precise("all")
        img1, outs1, _ = run()
        img1, outs1 = img1.clone(), [o.clone() for o in outs1]
        precise = [max_rel(o.cpu(), g[f"out{i}"]) for i, o in enumerate(outs1)]
        m.engine().set_precise(None)
        img2, outs2, _ = run()
    print(f"{gold}: block outputs vs f32 reference, plain {['%.1e' % v for v in plain]} -> precise {['%.1e' % v for v in precise]}; "
          f"image {max_rel(img0.cpu(), g['img']):.1e} -> {max_rel(img1.cpu(), g['img']):.1e}")
    # every map inside 1e-3; ~1e-5 where producer and consumer are both split.  Plain-form exceptions: E-NeRV's stage-0 pre-conv,
    # and the last block's output in front of the CUDA-core 1x1 head (one f16 rounding: <= 2^-12 of the map's maximum)
    lo = 1 if model == "ENeRV_Boost" else 0
    hi = len(precise) if model == "HNeRV_Boost" else len(precise) - 1
    assert max(precise) < 6e-4 and max(precise[lo:hi]) < 2e-4, precise
    assert max_rel(img1.cpu(), g["img"]) < 2e-4
    assert torch.equal(img2, img0) and all(torch.equal(p, q) for p, q in zip(outs2, outs0))


def test_precise_block_selection_and_errors():
    m = HNeRV_Boost(tiny_args("HNeRV_Boost")).eval().cuda()
    eng = m.engine()
    with pytest.raises(ValueError):
        eng.set_precise([99])
    eng.set_precise("1,2,head")
    assert eng.precise == frozenset({1, 2, "head"})
    emb = torch.rand(1, 16, *[int(v) for v in tiny_args("HNeRV_Boost").fc_hw.split("_")]).cuda()
    t = torch.tensor([0.25], dtype=torch.float64).cuda()
    sd = {k: v.detach().float().cpu() for k, v in m.state_dict().items()}
    ref, _ = orc.hnerv_boost_decode(sd, orc.cfg_from_args(tiny_args("HNeRV_Boost")), emb.cpu(), t.cpu())
    with torch.no_grad():
        img = m.forward_decoder(emb, t)[0]
    assert max_rel(img.cpu(), ref) < 1e-3
