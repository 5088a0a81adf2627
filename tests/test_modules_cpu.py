"""CPU: the drop-in modules' API, checkpoint layout and torch wiring against the reference goldens."""
import copy

import pytest
import torch

from conftest import load_golden, max_rel
from bnerv_b200 import ENeRV_Boost, HNeRV_Boost, NeRV_Boost, preset, tiny_args


def _build(model):
    a = tiny_args(model)
    if model == "NeRV_Boost":
        return NeRV_Boost(1, a)
    if model == "ENeRV_Boost":
        return ENeRV_Boost(3, a)
    return HNeRV_Boost(a)


@pytest.mark.parametrize("model,gold", [("NeRV_Boost", "nerv_tiny.npz"), ("ENeRV_Boost", "enerv_tiny.npz"), ("HNeRV_Boost", "hnerv_tiny.npz"),
                                        ("HNeRV_Boost", "hnerv_tiny_trained.npz"), ("NeRV_Boost", "nerv_tiny_trained.npz"),
                                        ("ENeRV_Boost", "enerv_tiny_trained.npz")])
def test_state_dict_layout_and_torch_wiring(model, gold):
    sd, g = load_golden(gold)
    m = _build(model).eval()
    own = m.state_dict()
    assert list(own.keys()) == list(sd.keys())                     # same keys in the same order as the reference
    assert all(own[k].shape == sd[k].shape for k in sd)
    m.load_state_dict(sd, strict=True)
    m.backend = "torch"
    with torch.no_grad():
        if model == "HNeRV_Boost":
            img, outs, dt = m.forward_decoder(g["emb"], g["t"])
            img2, outs2, _ = m(None, g["emb"], norm_idx=g["t"])      # input_embed short-circuits the encoder
            assert torch.equal(img, img2) and torch.equal(outs[0], g["emb"])
            assert max_rel(m.forward_encoder(g["frame"]), g["enc"]) < 1e-5
            img_full, _, _ = m(g["frame"], norm_idx=g["t"])
            assert max_rel(img_full, g["img_full"]) < 1e-5
        else:
            img, outs, dt = m(g["t"])
    assert isinstance(dt, float) and dt >= 0
    assert max_rel(img, g["img"]) < 5e-6
    for i, o in enumerate(outs):
        assert max_rel(o, g[f"out{i}"]) < 5e-6


def test_reference_key_counts_at_shipped_presets():
    # 186 keys NeRV_Boost, 206 ENeRV_Boost, 269 HNeRV_Boost (65 of them encoder)
    assert len(NeRV_Boost(1, preset("nerv_xs")).state_dict()) == 186
    hn = HNeRV_Boost(preset("hnerv_bunny"))
    assert len(hn.state_dict()) == 269
    assert sum(k.startswith("encoder.") for k in hn.state_dict()) == 65


def test_model_size_solver_matches_reference_values():
    # SURVEY.md appendix B sanity values, obtained by running train_nerv_all.py:194-217 unmodified
    assert preset("nerv_xs").fc_dim == 15 and preset("nerv_s").fc_dim == 30
    assert preset("enerv_m").fc_dim == 115
    assert preset("hnerv_l").fc_dim == 280 and preset("hnerv_m").fc_dim == 222 and preset("hnerv_bunny").fc_dim == 50


def test_hnerv_l_channel_schedule():
    m = HNeRV_Boost(preset("hnerv_l"))
    widths = [blk.sft_block.conv0.out_channels for blk in m.decoder]
    assert widths == [280, 233, 194, 162, 162, 135, 135, 112, 112]
    assert m.decoder[1].conv.upconv[0].kernel_size == (1, 1) and m.decoder[2].conv.upconv[0].kernel_size == (3, 3)
    assert m.head_layer.kernel_size == (3, 3)
    assert abs(m.decoder_params() - 13.66) < 0.02


def test_b200_backend_refuses_cpu_tensors():
    m = _build("NeRV_Boost").eval()
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU|CUDA"):
        m(torch.tensor([0.5], dtype=torch.float64))
    m.train()
    with pytest.raises(RuntimeError, match="CUDA"):            # no silent CPU fallback under autograd either
        m(torch.tensor([0.5], dtype=torch.float64))
    m.train_backend = "bogus"
    with pytest.raises(ValueError):
        m(torch.tensor([0.5], dtype=torch.float64))


def test_training_mode_uses_autograd_path_and_deepcopy_is_clean():
    m = _build("NeRV_Boost")
    m.train_backend = "torch"                                  # explicit opt-out: plain torch autograd (runs on CPU)
    img, _, _ = m(torch.tensor([0.5], dtype=torch.float64))
    img.mean().backward()
    assert m.head_layer.weight.grad is not None
    m2 = copy.deepcopy(m)
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))


def test_out_of_scope_baselines_are_importable_but_refuse():
    import model_hnerv
    with pytest.raises(NotImplementedError):
        model_hnerv.HNeRV(None)
    import model_blocks, model_enerv, model_nerv  # noqa: F401  (drop-in module names of train_nerv_all.py:16-18)
    with pytest.raises(KeyError):
        model_blocks.ActivationLayer("nope")           # model_blocks.py:156
    with pytest.raises(NotImplementedError):
        model_blocks.NormLayer("ln", 4)                # model_blocks.py:169


def test_batched_sft_tables_equal_per_layer_affine():
    """train._Layout.sft_tables (all TAT MLPs of the cascade in a few batched ops) == SFTLayer.affine per layer
    (model_blocks.py:101-104), and autograd reaches every SFT parameter through it."""
    from bnerv_b200 import ENeRV_Boost, HNeRV_Boost, tiny_args
    from bnerv_b200.engine import DecoderEngine
    from bnerv_b200.train import _Layout
    for cls, fam in ((HNeRV_Boost, "HNeRV_Boost"), (ENeRV_Boost, "ENeRV_Boost")):
        torch.manual_seed(0)
        a = tiny_args(fam)
        m = cls(a) if fam == "HNeRV_Boost" else cls(3, a)
        eng = DecoderEngine(m)
        lay = _Layout(eng, 3)
        cond = torch.randn(3, a.ch_t, 1, 1)
        sc, sh = lay.sft_tables(cond)
        i = 0
        for blk in eng.blocks:
            for layer in blk.sfts:
                s_ref, h_ref = layer.affine(cond)
                off, c = lay.sft_cols[i]
                assert torch.allclose(sc[:, off:off + c], s_ref.flatten(1), atol=1e-6)
                assert torch.allclose(sh[:, off:off + c], h_ref.flatten(1), atol=1e-6)
                i += 1
        (sc.sum() + sh.sum()).backward()
        for blk in eng.blocks:
            for layer in blk.sfts:
                assert all(p.grad is not None for p in layer.parameters())


def test_hnerv_utils_shim_re_exports_the_reference_module_and_overrides_loss_fn(tmp_path):
    """boosting-nerv_b200/shims/hnerv_utils.py: executes the next hnerv_utils.py on sys.path, keeps every symbol, and
    routes loss_fn to the device losses for CUDA tensors only."""
    import importlib
    import os
    import sys
    from conftest import ROOT
    fake = tmp_path / "hnerv_utils.py"
    fake.write_text("MARK = 41\n\ndef loss_fn(pred, target, loss_type='L2', batch_average=True):\n    return ('reference', loss_type)\n\n"
                    "def psnr_fn_single(a, b):\n    return 'psnr'\n\ndef quant_tensor(t, bits=8):\n    return ('reference', bits)\n")
    shim_dir = os.path.join(ROOT, "boosting-nerv_b200", "shims")
    old_path, old_mod = list(sys.path), sys.modules.pop("hnerv_utils", None)
    try:
        sys.path[:0] = [shim_dir, str(tmp_path)]
        mod = importlib.import_module("hnerv_utils")
        assert mod.MARK == 41 and mod.psnr_fn_single(0, 0) == "psnr"
        assert mod.loss_fn(torch.zeros(1), torch.zeros(1), "L1") == ("reference", "L1")       # CPU tensors: reference code
        assert mod.loss_fn is not mod._reference_loss_fn
        assert mod.quant_tensor(torch.zeros(4), 6) == ("reference", 6)                          # CPU tensors: reference code
    finally:
        sys.path[:] = old_path
        sys.modules.pop("hnerv_utils", None)
        if old_mod is not None:
            sys.modules["hnerv_utils"] = old_mod
