"""GPU: the one-kernel NeRVBlock (bnerv_nerv_block_fused / bnerv_resblock_fused, csrc/block_fused.cu) against the
three-launch form (bnerv_nerv_block_fwd) it replaces, and against the reference block goldens.

The fused kernel performs the same arithmetic operation by operation (same K order of the tensor-core accumulation, same
epilogue functions, f16 rounding of x0 / u / w at the same places), so the gate is BIT-EXACT equality with the three-launch
path - whose parity with the reference (model_blocks.py:34-46, 83-89, 101-105) is pinned by the block goldens and the
model tests.  Shapes cover every narrow stage of the benchmarked presets (12, 15, 21, 30, 43 channels; PixelShuffle 1 and 2;
Cin != C), ragged sizes that are not multiples of the region / row-block sizes, a map smaller than one region and B > 1."""
import pytest
import torch

from conftest import load_block_golden, max_rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from bnerv_b200 import ops as o
    return o


def make_block(ops, B, cin, C, H, W, s, seed=0, k_up=3):
    g = torch.Generator(device="cuda").manual_seed(seed)
    dev = "cuda"
    rn = lambda *shape: torch.randn(*shape, device=dev, generator=g)
    x = rn(B, cin, H, W)
    w_up = rn(C * s * s, cin, k_up, k_up) / (cin * k_up * k_up) ** 0.5 * 2.0
    b_up = rn(C * s * s) * 0.3
    w_c0, b_c0 = rn(C, C, 3, 3) / (C * 9) ** 0.5 * 2.0, rn(C) * 0.3
    w_c1, b_c1 = rn(C, C, 3, 3) / (C * 9) ** 0.5, rn(C) * 0.1
    cp = ops.round_up(C, 16)
    tabs = []
    for i in range(4):
        t = torch.zeros(B, cp, device=dev)
        t[:, :C] = (1.0 if i % 2 == 0 else 0.0) + 0.3 * rn(B, C)
        tabs.append(t)
    return (ops.nchw_to_c8(x), ops.PackedConv(w_up, b_up, s), ops.PackedConv(w_c0, b_c0, 1), ops.PackedConv(w_c1, b_c1, 1), tabs)


CASES = [  # B, cin, C, H, W, s
    (1, 12, 12, 64, 96, 1),       # NeRV-S / XS s = 1 stage
    (1, 12, 12, 45, 80, 2),       # NeRV up stage: 12 -> 48 + PixelShuffle(2)
    (1, 30, 15, 45, 80, 2),       # NeRV-S layers.1: Cin_p = 32 -> Cp = 16
    (1, 15, 12, 33, 47, 2),       # ragged: odd sizes
    (2, 12, 12, 37, 51, 1),       # B > 1, ragged
    (1, 12, 12, 7, 9, 1),         # smaller than one region
    (1, 21, 21, 56, 72, 1),       # E-NeRV-M 1080p stage width (Cp = 32)
    (1, 43, 21, 30, 44, 2),       # E-NeRV-M layers.6
    (1, 43, 43, 40, 56, 1),       # E-NeRV-M 540p stage width (Cp = 48)
    (3, 15, 15, 20, 28, 1),       # NeRV-XS fc_dim stage
    (1, 12, 12, 180, 320, 2),     # many regions per CTA
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "B%d_%dto%d_%dx%d_s%d" % c)
def test_fused_block_is_bit_identical_to_three_launches(ops, case):
    B, cin, C, H, W, s = case
    x, up, c0, c1, (g0, b0, g1, b1) = make_block(ops, *case)
    ref, _ = ops.nerv_block_fwd(x, up, c0, c1, cin, H, W, "sin", "gelu", g0, b0, g1, b1)
    out = ops.nerv_block_fused(x, up, c0, c1, cin, H, W, "sin", "gelu", g0, b0, g1, b1)
    assert out is not None, "shape unexpectedly outside the fused kernel's range"
    torch.cuda.synchronize()
    if not torch.equal(out, ref):
        d = (out.float() - ref.float()).abs()
        bad = d > 0
        idx = bad.nonzero()[:8].tolist()
        raise AssertionError(f"{int(bad.sum())} of {bad.numel()} values differ, max |diff| {d.max().item():.3e} "
                             f"(max |ref| {ref.float().abs().max().item():.3e}); first [b, group, y, x, c]: {idx}")


STREAM_CASES = [  # B, cin, C, H, W, s : the row-streaming form (<= 16 channels, no PixelShuffle)
    (1, 12, 12, 64, 96, 1),
    (1, 12, 12, 180, 320, 1),     # several strips (320 = 2 * 122 + 76) and row segments
    (2, 12, 12, 37, 51, 1),       # B > 1, ragged
    (1, 12, 12, 7, 9, 1),         # smaller than one strip / segment
    (3, 15, 15, 20, 28, 1),       # NeRV-XS fc_dim stage
    (1, 16, 16, 33, 123, 1),      # full 16 channels, one column more than a strip
    (1, 12, 12, 360, 640, 1),     # one CTA per SM, long segments
    (1, 9, 12, 50, 245, 1),       # Cin != C, W = 2 strips + 1
    (1, 12, 12, 45, 80, 2),       # PixelShuffle(2) up-conv inside the kernel: NeRV up stages
    (1, 15, 12, 33, 47, 2),       # ragged input, Cin != C
    (2, 12, 12, 20, 24, 2),       # B > 1
    (1, 12, 12, 7, 9, 2),         # smaller than one strip / segment
    (1, 12, 12, 180, 320, 2),     # several strips and segments (output 360 x 640)
    (1, 16, 16, 31, 61, 2),       # 16 channels: no pad channels to skip; output width 122 = exactly one strip
]


@pytest.mark.parametrize("case", STREAM_CASES, ids=lambda c: "B%d_%dto%d_%dx%d_s%d" % c)
def test_stream_block_is_bit_identical_to_three_launches(ops, case):
    B, cin, C, H, W, s = case
    x, up, c0, c1, (g0, b0, g1, b1) = make_block(ops, *case)
    ref, _ = ops.nerv_block_fwd(x, up, c0, c1, cin, H, W, "sin", "gelu", g0, b0, g1, b1)
    out = ops.nerv_block_fused(x, up, c0, c1, cin, H, W, "sin", "gelu", g0, b0, g1, b1, form="stream")
    assert out is not None, "shape unexpectedly outside the streaming kernel's range"
    torch.cuda.synchronize()
    if not torch.equal(out, ref):
        d = (out.float() - ref.float()).abs()
        bad = d > 0
        raise AssertionError(f"{int(bad.sum())} of {bad.numel()} values differ, max |diff| {d.max().item():.3e} "
                             f"(max |ref| {ref.float().abs().max().item():.3e}); first [b, group, y, x, c]: {bad.nonzero()[:8].tolist()}")
    # the ResBlock_SFT half alone, fed with the first launch's outputs
    mk = lambda: torch.empty_like(ref)
    x0, u = mk(), mk()
    ops.conv_fused(x, up, cin, H, W, act="sin", g1p=g0, beta=b0, out_pre=x0, out_aff=u)
    out2 = ops.resblock_fused(u, x0, c0, c1, C, H * s, W * s, "gelu", g1, b1, form="stream")
    assert out2 is not None and torch.equal(out2, ref)


RES32_CASES = [  # B, C, H, W : ResBlock_SFT half, 17..32 channels (block_stream32.cu)
    (1, 21, 64, 96),              # E-NeRV-M's 1080p stages: 21 channels, 11 live pairs
    (1, 21, 135, 250),            # three strips (250 = 2 * 122 + 6), several row segments
    (2, 21, 37, 51),              # B > 1, ragged
    (1, 21, 5, 9),                # fewer rows than middle warpgroups
    (1, 30, 45, 80),              # NeRV-S stage 0
    (1, 24, 33, 123),             # 12 live pairs, one column more than a strip
    (1, 32, 40, 122),             # all 32 channels, exactly one strip
    (1, 17, 21, 30),
    (1, 21, 270, 480),            # one CTA per SM and more
]


@pytest.mark.parametrize("case", RES32_CASES, ids=lambda c: "B%d_C%d_%dx%d" % c)
@pytest.mark.parametrize("act", ["gelu", "relu"])
def test_stream_resblock_32_channels_is_bit_identical_to_two_launches(ops, case, act):
    B, C, H, W = case
    x, up, c0, c1, (g0, b0, g1, b1) = make_block(ops, B, C, C, H, W, 1)
    mk = lambda: torch.empty(ops.c8_shape(B, C, H, W), dtype=torch.float16, device="cuda")
    x0, u, wmap, ref = mk(), mk(), mk(), mk()
    ops.conv_fused(x, up, C, H, W, act="sin", g1p=g0, beta=b0, out_pre=x0, out_aff=u)
    ops.conv_fused(u, c0, C, H, W, act=act, g1p=g1, beta=b1, out_aff=wmap)
    ops.conv_fused(wmap, c1, C, H, W, act="none", resid=x0, out_pre=ref)
    out = ops.resblock_fused(u, x0, c0, c1, C, H, W, act, g1, b1, form="stream")
    assert out is not None, "shape unexpectedly outside the streaming kernel's range"
    torch.cuda.synchronize()
    if not torch.equal(out, ref):
        d = (out.float() - ref.float()).abs()
        bad = d > 0
        raise AssertionError(f"{int(bad.sum())} of {bad.numel()} values differ, max |diff| {d.max().item():.3e} "
                             f"(max |ref| {ref.float().abs().max().item():.3e}); first [b, group, y, x, c]: {bad.nonzero()[:8].tolist()}")


@pytest.mark.parametrize("case", [(1, 21, 64, 96, 3), (2, 21, 37, 131, 3), (1, 30, 45, 80, 3), (1, 32, 20, 122, 4), (1, 24, 9, 250, 1)],
                         ids=lambda c: "B%d_C%d_%dx%d_head%d" % c)
def test_stream_resblock_with_fused_head_gives_the_two_launch_image_bit_for_bit(ops, case):
    """bnerv_resblock_stream_head: ResBlock_SFT half + 1x1 head conv + OutImg in one kernel against bnerv_resblock_stream followed
    by bnerv_head_conv1 - the image must be identical (same f16 rounding of the block output, same FMA order in the head)."""
    B, C, H, W, cout = case
    x, up, c0, c1, (g0, b0, g1, b1) = make_block(ops, B, C, C, H, W, 1)
    mk = lambda: torch.empty(ops.c8_shape(B, C, H, W), dtype=torch.float16, device="cuda")
    x0, u = mk(), mk()
    ops.conv_fused(x, up, C, H, W, act="sin", g1p=g0, beta=b0, out_pre=x0, out_aff=u)
    gen = torch.Generator(device="cuda").manual_seed(5)
    head = ops.PackedHead1(torch.randn(cout, C, 1, 1, device="cuda", generator=gen) / C ** 0.5, torch.randn(cout, device="cuda", generator=gen) * 0.1)
    out = ops.resblock_fused(u, x0, c0, c1, C, H, W, "gelu", g1, b1, form="stream")
    ref = torch.empty(B, cout, H, W, device="cuda")
    ops.conv_fused(out, head, C, H, W, act="tanh01", out_nchw=ref)
    img = torch.full_like(ref, float("nan"))
    assert ops.resblock_head_fused(u, x0, c0, c1, C, H, W, "gelu", g1, b1, head, img) is not None
    torch.cuda.synchronize()
    assert torch.equal(img, ref), f"max |diff| {(img - ref).abs().max().item():.3e}"
    small = ops.PackedHead1(torch.zeros(3, 12, 1, 1, device="cuda"), None)
    x12, up12, c012, c112, (_, _, g12, b12) = make_block(ops, 1, 12, 12, 16, 16, 1)
    u12 = torch.zeros(ops.c8_shape(1, 12, 16, 16), dtype=torch.float16, device="cuda")
    assert ops.resblock_head_fused(u12, u12, c012, c112, 12, 16, 16, "gelu", g12, b12, small, torch.zeros(1, 3, 16, 16, device="cuda")) is None


@pytest.mark.parametrize("case", [(1, 12, 12, 64, 96, 3), (2, 12, 12, 37, 131, 3), (1, 15, 15, 20, 28, 3), (1, 16, 16, 33, 123, 4),
                                  (1, 9, 12, 180, 320, 1)], ids=lambda c: "B%d_%dto%d_%dx%d_head%d" % c)
def test_stream_block_with_fused_head_gives_the_two_launch_image_bit_for_bit(ops, case):
    """bnerv_nerv_block_stream_head (a whole NeRVBlock + 1x1 head conv + OutImg, the tail of a NeRV-Boost frame) against
    bnerv_nerv_block_stream followed by bnerv_head_conv1: identical image."""
    B, cin, C, H, W, cout = case
    x, up, c0, c1, (g0, b0, g1, b1) = make_block(ops, B, cin, C, H, W, 1)
    gen = torch.Generator(device="cuda").manual_seed(7)
    head = ops.PackedHead1(torch.randn(cout, C, 1, 1, device="cuda", generator=gen) / C ** 0.5, torch.randn(cout, device="cuda", generator=gen) * 0.1)
    out = ops.nerv_block_fused(x, up, c0, c1, cin, H, W, "sin", "gelu", g0, b0, g1, b1, form="stream")
    ref = torch.empty(B, cout, H, W, device="cuda")
    ops.conv_fused(out, head, C, H, W, act="tanh01", out_nchw=ref)
    img = torch.full_like(ref, float("nan"))
    assert ops.nerv_block_head_fused(x, up, c0, c1, cin, H, W, g0, b0, g1, b1, head, img) is not None
    torch.cuda.synchronize()
    assert torch.equal(img, ref), f"max |diff| {(img - ref).abs().max().item():.3e}"


@pytest.mark.parametrize("case", [(1, 21, 21, 64, 96), (2, 21, 21, 37, 131), (1, 30, 17, 45, 80), (1, 32, 32, 20, 122), (1, 24, 21, 9, 250),
                                  (1, 21, 21, 270, 480)], ids=lambda c: "B%d_%dto%d_%dx%d" % c)
@pytest.mark.parametrize("act", ["sin", "relu"])
def test_stream_upconv_32_channels_is_bit_identical_to_the_fused_conv(ops, case, act):
    """bnerv_upconv_stream (3x3 up-conv + activation + TAT affine, two outputs, 17..32 channels) against bnerv_conv_fused."""
    B, cin, C, H, W = case
    x, up, c0, c1, (g0, b0, g1, b1) = make_block(ops, B, cin, C, H, W, 1)
    mk = lambda: torch.full(ops.c8_shape(B, C, H, W), float("nan"), dtype=torch.float16, device="cuda")
    x0r, ur, x0, u = mk(), mk(), mk(), mk()
    ops.conv_fused(x, up, cin, H, W, act=act, g1p=g0, beta=b0, out_pre=x0r, out_aff=ur)
    assert ops.upconv_stream(x, up, cin, H, W, act, g0, b0, x0, u) is not None
    torch.cuda.synchronize()
    assert torch.equal(x0, x0r), f"x0: max |diff| {(x0.float() - x0r.float()).abs().max().item():.3e}"
    assert torch.equal(u, ur), f"u: max |diff| {(u.float() - ur.float()).abs().max().item():.3e}"
    x12, up12, _, _, (g12, b12, _, _) = make_block(ops, 1, 12, 12, 16, 16, 1)
    o = torch.zeros(ops.c8_shape(1, 12, 16, 16), dtype=torch.float16, device="cuda")
    assert ops.upconv_stream(x12, up12, 12, 16, 16, "sin", g12, b12, o, o.clone()) is None


@pytest.mark.parametrize("case", [(1, 43, 43, 64, 96), (2, 43, 43, 37, 131), (1, 33, 48, 20, 126), (1, 48, 40, 9, 253), (1, 43, 43, 270, 480),
                                  (1, 21, 21, 40, 130), (1, 30, 17, 5, 9)], ids=lambda c: "B%d_%dto%d_%dx%d" % c)
@pytest.mark.parametrize("form", ["up_sin", "c0_gelu", "c1_resid", "relu_both"])
def test_conv_stream_is_bit_identical_to_the_fused_conv(ops, case, form):
    """bnerv_conv_stream (one 3x3 conv of 17..48 channels in the row-streaming form) against bnerv_conv_fused for the epilogue
    shapes of a NeRVBlock: up-conv (sin, x0 + u), conv0 (GELU, affine output only), conv1 (residual), and a run-time activation."""
    B, cin, C, H, W = case
    x, up, c0, c1, (g0, b0, g1, b1) = make_block(ops, B, cin, C, H, W, 1)
    mk = lambda: torch.full(ops.c8_shape(B, C, H, W), float("nan"), dtype=torch.float16, device="cuda")
    res = ops.nchw_to_c8(torch.randn(B, C, H, W, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3)))
    kw = {"up_sin": dict(act="sin", g1p=g0, beta=b0, out_pre=True, out_aff=True), "c0_gelu": dict(act="gelu", g1p=g1, beta=b1, out_aff=True),
          "c1_resid": dict(act="none", resid=res, out_pre=True), "relu_both": dict(act="relu", resid=res, g1p=g1, beta=b1, out_pre=True, out_aff=True)}[form]
    outs_r = {k: mk() for k in ("out_pre", "out_aff") if kw.get(k)}
    outs_s = {k: mk() for k in outs_r}
    base = {k: v for k, v in kw.items() if k not in ("out_pre", "out_aff")}
    ops.conv_fused(x, up, cin, H, W, **base, **outs_r)
    assert ops.conv_stream(x, up, cin, H, W, **base, **outs_s) is True
    torch.cuda.synchronize()
    for k in outs_r:
        assert torch.equal(outs_s[k], outs_r[k]), f"{k}: max |diff| {(outs_s[k].float() - outs_r[k].float()).abs().max().item():.3e}"
    x12, up12, _, _, _ = make_block(ops, 1, 12, 12, 16, 16, 1)
    assert ops.conv_stream(x12, up12, 12, 16, 16, out_pre=torch.zeros(ops.c8_shape(1, 12, 16, 16), dtype=torch.float16, device="cuda")) is None


def test_stream_block_refuses_what_it_does_not_implement(ops):
    x, up, c0, c1, (g0, b0, g1, b1) = make_block(ops, 1, 12, 12, 20, 24, 3)           # PixelShuffle(3)
    assert ops.nerv_block_fused(x, up, c0, c1, 12, 20, 24, "sin", "gelu", g0, b0, g1, b1, form="stream") is None
    x, up, c0, c1, (g0, b0, g1, b1) = make_block(ops, 1, 30, 15, 20, 24, 2)           # 30 input channels
    assert ops.nerv_block_fused(x, up, c0, c1, 30, 20, 24, "sin", "gelu", g0, b0, g1, b1, form="stream") is None
    x, up, c0, c1, (g0, b0, g1, b1) = make_block(ops, 1, 21, 21, 20, 24, 1)           # 21 channels
    assert ops.nerv_block_fused(x, up, c0, c1, 21, 20, 24, "sin", "gelu", g0, b0, g1, b1, form="stream") is None


@pytest.mark.parametrize("case", [(1, 12, 12, 64, 96, 1), (2, 30, 30, 45, 80, 1), (1, 43, 43, 37, 53, 1), (1, 21, 21, 7, 5, 1)],
                         ids=lambda c: "B%d_C%d_%dx%d" % (c[0], c[2], c[3], c[4]))
def test_fused_resblock_is_bit_identical_to_two_launches(ops, case):
    B, _, C, H, W, _ = case
    x, up, c0, c1, (g0, b0, g1, b1) = make_block(ops, B, C, C, H, W, 1)
    dev = x.device
    mk = lambda: torch.empty(ops.c8_shape(B, C, H, W), dtype=torch.float16, device=dev)
    x0, u, wmap, ref = mk(), mk(), mk(), mk()
    ops.conv_fused(x, up, C, H, W, act="sin", g1p=g0, beta=b0, out_pre=x0, out_aff=u)
    ops.conv_fused(u, c0, C, H, W, act="gelu", g1p=g1, beta=b1, out_aff=wmap)
    ops.conv_fused(wmap, c1, C, H, W, act="none", resid=x0, out_pre=ref)
    out = ops.resblock_fused(u, x0, c0, c1, C, H, W, "gelu", g1, b1)
    assert out is not None
    torch.cuda.synchronize()
    assert torch.equal(out, ref), f"max |diff| {(out.float() - ref.float()).abs().max().item():.3e}"


def test_fused_block_generic_activations_and_unsupported_shapes(ops):
    # run-time activation codes (the generic instantiation): relu / none
    x, up, c0, c1, (g0, b0, g1, b1) = make_block(ops, 1, 12, 12, 40, 48, 1)
    ref, _ = ops.nerv_block_fwd(x, up, c0, c1, 12, 40, 48, "relu", "relu", g0, b0, g1, b1)
    out = ops.nerv_block_fused(x, up, c0, c1, 12, 40, 48, "relu", "relu", g0, b0, g1, b1)
    assert torch.equal(out, ref)
    # outside the range: nothing is launched, the wrapper reports None and the caller takes the three-launch path
    from bnerv_b200 import _capi
    x, up, c0, c1, (g0, b0, g1, b1) = make_block(ops, 1, 12, 12, 9, 16, 3)           # PixelShuffle(3)
    n0 = _capi.launch_count()
    assert ops.nerv_block_fused(x, up, c0, c1, 12, 9, 16, "sin", "gelu", g0, b0, g1, b1) is None
    x64 = make_block(ops, 1, 64, 64, 16, 16, 1)                                     # 64 channels
    n1 = _capi.launch_count()
    assert ops.nerv_block_fused(x64[0], x64[1], x64[2], x64[3], 64, 16, 16, "sin", "gelu", *x64[4]) is None
    assert _capi.launch_count() == n1 and n1 > n0


@pytest.mark.parametrize("name", ["s1_k3", "s2_k3"])
def test_fused_block_against_reference_block_golden(ops, name):
    """The narrow reference block goldens (tests/golden/blocks.npz, minted from the unmodified NeRVBlock,
    model_blocks.py:34-46) through the one-kernel form: 1e-3 (north_star) and bit-identity with the three launches."""
    c = load_block_golden()[name]
    sd = {k[3:]: torch.from_numpy(v).cuda() for k, v in c.items() if k.startswith("sd/")}
    ngf, new_ngf, ks, s, H, W, B = [int(v) for v in c["meta"]]
    x, e = torch.from_numpy(c["x"]).cuda(), torch.from_numpy(c["e"]).cuda()
    layers = []
    for sft in ("sft0", "sft1"):
        layers.append(tuple(sd[f"sft_block.{sft}.SFT_{br}_conv{i}.{p}"].reshape(-1, 32).contiguous() if p == "weight"
                            else sd[f"sft_block.{sft}.SFT_{br}_conv{i}.{p}"]
                            for br in ("scale", "shift") for i in (0, 1) for p in ("weight", "bias")))
    tab = ops.SftTable(layers, B, x.device)
    tab.run(e.flatten(1))
    up = ops.PackedConv(sd["conv.upconv.0.weight"], sd["conv.upconv.0.bias"], s)
    c0 = ops.PackedConv(sd["sft_block.conv0.weight"], sd["sft_block.conv0.bias"], 1)
    c1 = ops.PackedConv(sd["sft_block.conv1.weight"], sd["sft_block.conv1.bias"], 1)
    args = (ops.nchw_to_c8(x), up, c0, c1, ngf, H, W, "sin", "gelu", tab.g1p[0], tab.beta[0], tab.g1p[1], tab.beta[1])
    out = ops.nerv_block_fused(*args)
    assert out is not None
    assert max_rel(ops.c8_to_nchw(out, new_ngf).cpu(), torch.from_numpy(c["y"])) < 1e-3
    assert torch.equal(out, ops.nerv_block_fwd(*args)[0])
