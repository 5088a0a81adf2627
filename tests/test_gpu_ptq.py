"""GPU: post-training quantisation + Huffman statistics on the device kernels (csrc/ptq_ops.cu through bnerv_b200.ptq)
against the goldens minted from the unmodified reference quant_tensor and against the CPU oracle.  Gate: BIT-EXACT
(codes, min / scale tables incl. dtype and keepdim shape, reconstruction, code histogram, code lengths)."""
import types

import numpy as np
import pytest
import torch

from conftest import load_golden, max_rel
from oracle import nerv_oracle as orc
from oracle import ptq_oracle as po
from test_ptq_cpu import CASES, load_ptq_golden
from bnerv_b200 import HNeRV_Boost, tiny_args

pytestmark = pytest.mark.gpu


def _same(q, new_t, want_q, want_new_t):
    assert q["quant"].dtype == torch.uint8 and q["quant"].shape == want_q["quant"].shape
    assert torch.equal(q["quant"].cpu(), want_q["quant"])
    for key in ("min", "scale"):
        assert q[key].dtype == want_q[key].dtype and q[key].shape == want_q[key].shape, key
        assert torch.equal(q[key].cpu(), want_q[key]), key
    assert new_t.dtype == torch.float32 and torch.equal(new_t.cpu(), want_new_t)


@pytest.mark.parametrize("name", CASES)
def test_quant_tensor_is_bit_identical_to_the_reference_goldens(name):
    from bnerv_b200 import _capi, ptq
    g = load_ptq_golden()[name]
    n0 = _capi.launch_count()
    q, new_t = ptq.quant_tensor(torch.from_numpy(g["t"]).cuda(), int(g["bits"]))
    assert _capi.launch_count() - n0 >= 5
    want = {k: torch.from_numpy(g[k]) for k in ("quant", "min", "scale")}
    _same(q, new_t, want, torch.from_numpy(g["new_t"]))
    assert torch.equal(ptq.reconstruct_tensor(q).cpu(), torch.from_numpy(g["new_t"]))      # decode side: stored form -> new_t
    d = ptq.dequant_tensor(q)
    assert d.dtype == torch.from_numpy(g["dequant"]).dtype
    if d.dtype == torch.float32:                 # f16 tables: the reference's dequant_tensor is f16 arithmetic, device-dependent rounding
        assert torch.equal(d.cpu(), torch.from_numpy(g["dequant"]))


@pytest.mark.parametrize("shape,bits,kind", [
    ((448, 135, 3, 3), 8, "uniform"),      # HNeRV-L decoder.7 up-conv: thread-per-group min/max (1215 and 4032 groups)
    ((5825, 280, 1, 1), 8, "normal"),      # decoder.1 1x1 up-conv
    ((6120, 256), 8, "normal"),            # NeRV stem: last-axis candidate with 6120 groups
    ((600, 16, 9, 16), 6, "skewed"),       # UVG frame embeddings, 6 bits
    ((1746,), 8, "normal"),                # bias of decoder.2's up-conv
    ((3, 112, 3, 3), 8, "uniform"),        # head
    ((2000001,), 5, "normal"),             # odd length, split whole-tensor reduction
    ((7,), 1, "normal"), ((2, 1, 1, 3), 8, "normal"),
])
def test_quant_tensor_matches_oracle_at_model_sizes(shape, bits, kind):
    from bnerv_b200 import ptq
    g = torch.Generator().manual_seed(len(shape) * 1000 + bits)
    if kind == "uniform":
        t = (torch.rand(shape, generator=g) - 0.5) * 0.1
    elif kind == "normal":
        t = torch.randn(shape, generator=g) * 0.05
    else:
        t = torch.randn(shape, generator=g) * torch.exp(torch.randn((shape[0],) + (1,) * (len(shape) - 1), generator=g) * 1.5)
    want_q, want_new = po.quant_tensor(t, bits)
    q, new_t = ptq.quant_tensor(t.cuda(), bits)
    _same(q, new_t, want_q, want_new)
    assert torch.equal(ptq.reconstruct_tensor(q), new_t)


def test_quant_tensors_multi_call_is_bitwise_the_single_tensor_call():
    """One bnerv_ptq_quant_tensors call over a model's worth of mixed shapes (five launches in total) returns, tensor by
    tensor, exactly what the per-tensor call returns: codes, tables (dtype, keepdim shape), reconstruction."""
    from bnerv_b200 import _capi, ptq
    g = torch.Generator().manual_seed(11)
    shapes = [(448, 135, 3, 3), (1746,), (3, 112, 3, 3), (6120, 256), (7,), (2, 1, 1, 3), (600, 16, 9, 16), (51, 60), (2,), (64, 52, 1, 1)]
    ts = [(torch.randn(s, generator=g) * 0.05 * (1 + i)).cuda() for i, s in enumerate(shapes)]
    ts.append(ts[3].t())                                            # non-contiguous input
    n0 = _capi.launch_count()
    multi = ptq.quant_tensors(ts, 8)
    assert _capi.launch_count() - n0 == 5
    for t, (q, new_t) in zip(ts, multi):
        q1, new_1 = ptq.quant_tensor(t, 8)
        _same(q, new_t, {k: v.cpu() for k, v in q1.items()}, new_1.cpu())
        want_q, want_new = po.quant_tensor(t.cpu(), 8)
        _same(q, new_t, want_q, want_new)
    assert ptq.quant_tensors([], 8) == []
    with pytest.raises(TypeError):
        ptq.quant_tensors([ts[0], ts[1].double()], 8)


def test_constant_tensor_is_reconstructed_exactly():
    """max == min makes the reference divide 0 by 0 (NaN codes, NaN weights); here the clamp absorbs the NaN: code 0 and
    min + 0 * 0 = the constant.  Defined behaviour where the reference has none - not a parity case."""
    from bnerv_b200 import ptq
    for shape in [(), (1, 1, 1, 1), (12,), (64, 3)]:
        t = torch.full(shape, 0.375).cuda()
        q, new_t = ptq.quant_tensor(t, 8)
        assert int(q["quant"].max()) == 0 and torch.equal(new_t, t) and q["min"].dim() == 0


def test_non_contiguous_input_and_refusals():
    from bnerv_b200 import ptq
    t = torch.randn(60, 70, generator=torch.Generator().manual_seed(3))
    want_q, want_new = po.quant_tensor(t.t().contiguous(), 8)
    q, new_t = ptq.quant_tensor(t.cuda().t(), 8)
    _same(q, new_t, want_q, want_new)
    with pytest.raises(RuntimeError):
        ptq.quant_tensor(t, 8)                     # CPU tensor: no fallback
    with pytest.raises(TypeError):
        ptq.quant_tensor(t.cuda().half(), 8)
    with pytest.raises(ValueError):
        ptq.quant_tensor(t.cuda(), 9)


def test_code_histogram_matches_bincount_for_every_alignment():
    from bnerv_b200 import ptq
    g = torch.Generator().manual_seed(11)
    base = torch.clamp(torch.randn(1 << 20, generator=g) * 12 + 128, 0, 255).to(torch.uint8).cuda()
    for off, n in [(0, 1 << 20), (1, 1000003), (2, 77), (3, 5), (0, 1), (1, 2), (3, 4), (2, 4097)]:
        q = base[off:off + n]
        layer = {"quant": q, "min": torch.zeros(3), "scale": torch.zeros(3)}
        counts, nt = ptq.code_histogram({"a": layer, "b": layer})
        assert nt == 12
        assert torch.equal(counts.cpu(), 2 * torch.bincount(q.cpu().long(), minlength=256)), (off, n)


def test_quant_model_huffman_bits_and_quantised_decode():
    """The whole evaluation-side chain of train_nerv_all.py:620-641, 542, 581-607 on a tiny HNeRV_Boost: quantise the
    decoder, quantise the embeddings, count the Huffman bits, decode with the quantised weights on the native kernels."""
    from bnerv_b200 import _capi, ptq
    sd, g = load_golden("hnerv_tiny.npz")
    a = tiny_args("HNeRV_Boost")
    m = HNeRV_Boost(a).eval()
    m.load_state_dict(sd)
    m = m.cuda()
    args = types.SimpleNamespace(quant_model_bit=8)
    models, quant_ckt = ptq.quant_model(m, args)
    assert len(models) == 2 and set(quant_ckt) == {k for k in sd if "encoder" not in k}
    want_ckt, want_sd = po.quant_model_state(sd, 8)
    qsd = models[1].state_dict()
    for k in sd:
        assert torch.equal(qsd[k].cpu(), want_sd[k]), k                                 # dequantised weights, encoder untouched
    for k in quant_ckt:
        _same(quant_ckt[k], qsd[k], want_ckt[k], want_sd[k])
    assert ptq.quant_model(m, types.SimpleNamespace(quant_model_bit=-1))[1] is None
    # decode side of the format: a fresh model filled from the stored codes + tables equals the quantised model bit for bit
    fresh = ptq.load_quant_ckt(HNeRV_Boost(a).eval().cuda(), quant_ckt)
    fsd = fresh.state_dict()
    assert all(torch.equal(fsd[k], qsd[k]) for k in quant_ckt)
    with pytest.raises(KeyError):
        ptq.load_quant_ckt(fresh, {k: v for k, v in list(quant_ckt.items())[1:]})

    emb = torch.rand(64, 16, 2, 4, generator=torch.Generator().manual_seed(5))
    q_emb, deq_emb = ptq.quant_tensor(emb.cuda(), 6)
    want_qe, want_de = po.quant_tensor(emb, 6)
    _same(q_emb, deq_emb, want_qe, want_de)

    bits = ptq.huffman_bits(quant_ckt, q_emb)
    assert bits == po.huffman_bits(want_ckt, want_qe)
    assert 0 < bits["bits_per_param"] <= 8 and bits["total_bits"] > bits["code_bits"]

    t = g["t"].cuda()
    n0 = _capi.launch_count()
    with torch.no_grad():
        img, _, _ = models[1].forward_decoder(deq_emb[:2], t)
    assert _capi.launch_count() - n0 >= 10
    ref, _ = orc.hnerv_boost_decode(want_sd, orc.cfg_from_args(a), want_de[:2], g["t"])
    assert max_rel(img.cpu(), ref) < 1e-3


def test_hnerv_utils_shim_routes_cuda_quant_tensor_to_the_device_kernels(tmp_path):
    import importlib
    import os
    import sys
    from conftest import ROOT
    from bnerv_b200 import _capi
    (tmp_path / "hnerv_utils.py").write_text("def loss_fn(*a, **k):\n    return 'reference'\n\ndef quant_tensor(t, bits=8):\n    return ('reference', bits)\n")
    old_path, old_mod = list(sys.path), sys.modules.pop("hnerv_utils", None)
    try:
        sys.path[:0] = [os.path.join(ROOT, "boosting-nerv_b200", "shims"), str(tmp_path)]
        mod = importlib.import_module("hnerv_utils")
        t = torch.randn(96, 24, 3, 3, generator=torch.Generator().manual_seed(2))
        n0 = _capi.launch_count()
        q, new_t = mod.quant_tensor(t.cuda(), 8)
        assert _capi.launch_count() > n0
        want_q, want_new = po.quant_tensor(t, 8)
        _same(q, new_t, want_q, want_new)
        assert mod.quant_tensor(t, 8) == ("reference", 8)               # CPU tensors stay with the reference code
    finally:
        sys.path[:] = old_path
        sys.modules.pop("hnerv_utils", None)
        if old_mod is not None:
            sys.modules["hnerv_utils"] = old_mod
