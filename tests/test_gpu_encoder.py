"""GPU: native forward of HNeRV_Boost's ConvNeXt encoder (csrc/encoder_ops.cu through bnerv_b200.encoder) against the golden
minted from the unmodified reference (tests/golden/hnerv_tiny.npz: frame -> enc) and against the torch module in strict f32
(conftest disables TF32) with NON-initial weights - at initialisation the layer scale gamma = 1e-6 hides the whole MLP
branch of every block.  Gate: 1e-5 max|diff|/max|ref| (f32 arithmetic, summation order only)."""
import pytest
import torch

from conftest import load_golden, max_rel
from bnerv_b200 import HNeRV_Boost, preset, tiny_args
from bnerv_b200.layers import ConvNeXt

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.mark.parametrize("gold", ["hnerv_tiny.npz", "hnerv_tiny_trained.npz"])
def test_encoder_matches_reference_golden_through_forward_encoder(gold):
    from bnerv_b200 import _capi
    sd, g = load_golden(gold)
    m = HNeRV_Boost(tiny_args("HNeRV_Boost")).eval()
    m.load_state_dict(sd)
    m = m.cuda()
    n0 = _capi.launch_count()
    with torch.no_grad():
        enc = m.forward_encoder(g["frame"].cuda())
    assert _capi.launch_count() - n0 >= 10                       # native kernels, not the torch module
    assert enc.shape == g["enc"].shape and max_rel(enc.cpu(), g["enc"]) < TOL
    m.backend = "torch"
    with torch.no_grad():
        assert max_rel(enc, m.forward_encoder(g["frame"].cuda())) < TOL
    m.backend = "b200"
    with pytest.raises(RuntimeError, match="CUDA"), torch.no_grad():
        m.forward_encoder(g["frame"])                            # CPU tensor: no fallback
    frame = g["frame"].cuda().requires_grad_(False)
    out = m.forward_encoder(frame)                               # autograd on + trainable parameters -> torch module (training)
    assert out.requires_grad


def _randomise(enc, seed):
    """Trained-like parameters: O(1) layer scale (the reference initialises gamma to 1e-6 and biases to 0)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in enc.named_parameters():
            if name.endswith("gamma"):
                p.copy_(torch.rand(p.shape, generator=g) * 2 - 1)
            elif p.dim() == 1 and name.endswith("weight"):                      # LayerNorm weights
                p.copy_(torch.rand(p.shape, generator=g) + 0.5)
            elif p.dim() == 1:                                                  # every bias
                p.copy_(torch.randn(p.shape, generator=g) * 0.2)
            else:                                                               # conv / linear weights, fan-in scaled
                p.copy_(torch.randn(p.shape, generator=g) * (1.5 / max(1, p[0].numel()) ** 0.5))


@pytest.mark.parametrize("strds,dims,blocks,hw,batch", [
    ([5, 2, 2], [16, 16, 16], 1, (40, 80), 2),          # the tiny golden's structure
    ([5, 3, 2, 2, 2], [64, 64, 64, 64, 16], 1, (360, 480), 1),      # the shipped encoder (enc_dim 64_16) at reduced resolution
    ([4, 2], [24, 40], 2, (37, 53), 3),                 # two blocks per stage, sizes that are no multiple of the stride or the tiles
    ([3], [70], 1, (30, 33), 1),                        # channel count that is no multiple of 4 / 32
])
def test_encoder_matches_torch_module_with_trained_like_weights(strds, dims, blocks, hw, batch):
    from bnerv_b200.encoder import convnext_forward
    enc = ConvNeXt(stage_blocks=blocks, strds=strds, dims=dims).eval()
    _randomise(enc, 17 + len(strds))
    enc = enc.cuda()
    x = torch.rand(batch, 3, *hw, generator=torch.Generator().manual_seed(3)).cuda()
    with torch.no_grad():
        ref = enc(x)
        got = convnext_forward(enc, x)
    assert got.shape == ref.shape
    assert float(ref.abs().max()) > 1e-2
    assert max_rel(got, ref) < TOL


def test_encoder_full_size_1080p_matches_torch():
    from bnerv_b200.encoder import convnext_forward
    torch.manual_seed(1)
    m = HNeRV_Boost(preset("hnerv_l")).eval()
    _randomise(m.encoder, 5)
    m = m.cuda()
    x = torch.rand(1, 3, 1080, 1920, generator=torch.Generator().manual_seed(4)).cuda()
    with torch.no_grad():
        ref = m.encoder(x)
        got = convnext_forward(m.encoder, x)
    assert tuple(got.shape) == (1, 16, 9, 16)
    assert max_rel(got, ref) < TOL
