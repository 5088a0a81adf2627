"""CPU: pin oracle/ptq_oracle.py against vectors minted from the unmodified reference quant_tensor / dequant_tensor
(tests/golden/ptq.npz), and check the host half of the C-ABI (candidate planning, Huffman code lengths) without a GPU."""
import ctypes
import heapq
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import ptq_oracle as po

CASES = ["conv_up", "conv_wide", "conv_small", "bias", "bias_small", "sft", "embed6", "embed8", "linear", "bits4"]


def load_ptq_golden():
    z = np.load(os.path.join(GOLDEN, "ptq.npz"))
    return {name: {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(name + "/")} for name in CASES}


@pytest.mark.parametrize("name", CASES)
def test_oracle_quant_tensor_is_bit_identical_to_the_reference(name):
    g = load_ptq_golden()[name]
    q, new_t = po.quant_tensor(torch.from_numpy(g["t"]), int(g["bits"]))
    assert q["quant"].dtype == torch.uint8 and np.array_equal(q["quant"].numpy(), g["quant"])
    for key in ("min", "scale"):
        assert q[key].numpy().dtype == g[key].dtype and q[key].shape == g[key].shape          # f32 scalar or f16 keepdim table
        assert np.array_equal(q[key].numpy(), g[key])
    assert np.array_equal(new_t.numpy(), g["new_t"])
    d = po.dequant_tensor(q)
    assert d.numpy().dtype == g["dequant"].dtype and np.array_equal(d.numpy(), g["dequant"])


def test_goldens_cover_whole_tensor_and_two_different_axes():
    g = load_ptq_golden()
    kinds = set()
    for name in CASES:
        m = g[name]["min"]
        kinds.add("whole" if m.ndim == 0 else tuple(i for i, n in enumerate(m.shape) if n == 1 and g[name]["t"].shape[i] > 1))
    assert "whole" in kinds and (0,) in kinds and (1,) in kinds


def _plan(shape):
    from bnerv_b200 import _capi
    arr = (ctypes.c_int64 * max(1, len(shape)))(*shape)
    plan = _capi.PtqPlan()
    rc = _capi.lib.bnerv_ptq_plan_tensor(arr, len(shape), ctypes.byref(plan))
    return rc, plan


def test_plan_lists_the_reference_candidates():
    # axis a is a candidate iff (numel / shape[a]) / numel < 0.02, i.e. shape[a] > 50 (hnerv_utils.py:108-113)
    for shape in [(448, 135, 3, 3), (12, 12, 3, 3), (135,), (50,), (51,), (280, 160), (600, 16, 9, 16), (64, 72, 1, 1), ()]:
        rc, p = _plan(shape)
        assert rc == 0
        numel = int(np.prod(shape)) if shape else 1
        want = [-1] + [a for a, n in enumerate(shape) if (numel // n) / numel < 0.02]
        assert p.n_cand == len(want) and list(p.axis[:p.n_cand]) == want
        assert [p.groups[c] for c in range(p.n_cand)] == [1] + [numel // shape[a] for a in want[1:]]
        off = 0
        for c in range(p.n_cand):
            assert p.table_offset[c] == off
            off += 2 * p.groups[c]
        assert p.table_floats == off and p.scratch_doubles > 0
    from bnerv_b200 import _capi
    assert _plan((2, 2, 2, 2, 2))[0] == -2 and b"dimensions" in _capi.lib.bnerv_last_error()
    assert _plan((4, 0))[0] == -1
    one = ctypes.c_void_p(16)
    arr = (ctypes.c_int64 * 1)(8)
    assert _capi.lib.bnerv_ptq_quant_tensor(one, arr, 1, 9, one, None, one, one, one, one, None) == -2      # > 8 bits: codes are uint8
    assert _capi.lib.bnerv_ptq_quant_tensor(None, arr, 1, 8, one, None, one, one, one, one, None) == -1
    assert _capi.lib.bnerv_histogram_u8(None, 4, one, None) == -1
    assert _capi.lib.bnerv_ptq_dequant_tensor(one, arr, 1, 1, one, one, 1, one, None) == -1 and b"axis" in _capi.lib.bnerv_last_error()
    assert _capi.lib.bnerv_launch_count() == 0


def _native_lengths(counts):
    from bnerv_b200 import _capi
    c = np.ascontiguousarray(counts, dtype=np.uint64)
    out = np.zeros(len(c), dtype=np.int32)
    rc = _capi.lib.bnerv_huffman_code_lengths(c.ctypes.data, len(c), out.ctypes.data)
    assert rc == 0, _capi.lib.bnerv_last_error()
    return out


def _optimal_cost_with_eof(counts):
    """Total cost of ANY Huffman tree over the counts plus one EOF leaf of weight 1 = sum of the merged weights."""
    h = [int(c) for c in counts if c] + [1]
    heapq.heapify(h)
    cost = 0
    while len(h) > 1:
        a, b = heapq.heappop(h), heapq.heappop(h)
        cost += a + b
        heapq.heappush(h, a + b)
    return cost


def test_native_huffman_lengths_equal_the_restated_dahuffman_construction():
    rng = np.random.default_rng(7)
    for trial in range(120):
        n = int(rng.integers(1, 257))
        kind = trial % 5
        if kind == 0:
            counts = rng.integers(0, 4, size=n)                     # many ties, many absent symbols
        elif kind == 1:
            counts = rng.integers(0, 100000, size=n)
        elif kind == 2:                                             # the shape real weight codes have: a peak around mid-range
            counts = np.round(np.exp(-0.5 * ((np.arange(n) - n / 2) / (n / 8 + 1)) ** 2) * 1e6).astype(np.int64)
        elif kind == 3:
            counts = np.full(n, int(rng.integers(1, 4)))            # all equal: order decided by the symbol value only
        else:
            counts = np.array([2 ** min(k, 40) for k in range(n)])  # maximally skewed: a chain, lengths up to n
        if counts.sum() == 0:
            counts[0] = 1
        got = _native_lengths(counts)
        ref = po.huffman_code_lengths({s: int(c) for s, c in enumerate(counts) if c})
        want = np.zeros(n, dtype=np.int32)
        for s, l in ref.items():
            want[s] = l
        assert np.array_equal(got, want), (trial, n)
        # invariants of every Huffman code over {symbols, EOF}: the real symbols leave exactly the EOF leaf's share of the
        # Kraft sum, and the cost including the EOF leaf is the optimum whatever the tie-breaking
        present = got[counts > 0]
        assert (got[counts == 0] == 0).all() and (present >= 1).all()
        kraft = sum(2.0 ** -int(l) for l in present)
        eof_len = -np.log2(1.0 - kraft) if kraft < 1 else None
        assert eof_len is not None and abs(eof_len - round(eof_len)) < 1e-9
        cost = int((counts[counts > 0].astype(object) * present.astype(object)).sum()) + int(round(eof_len))
        assert cost == _optimal_cost_with_eof(counts)


def test_single_symbol_and_error_paths():
    assert _native_lengths([0, 0, 9, 0]).tolist() == [0, 0, 1, 0]      # {symbol, EOF}: one bit
    from bnerv_b200 import _capi
    z = np.zeros(4, dtype=np.uint64)
    out = np.zeros(4, dtype=np.int32)
    assert _capi.lib.bnerv_huffman_code_lengths(z.ctypes.data, 4, out.ctypes.data) == -1
    assert _capi.lib.bnerv_huffman_code_lengths(z.ctypes.data, 0, out.ctypes.data) == -2


def test_oracle_huffman_bits_accounting():
    g = load_ptq_golden()
    ckt = {n: {k: torch.from_numpy(g[n][k]) for k in ("quant", "min", "scale")} for n in ("conv_up", "bias", "conv_small")}
    emb = {k: torch.from_numpy(g["embed6"][k]) for k in ("quant", "min", "scale")}
    r = po.huffman_bits(ckt, emb)
    n = sum(g[k]["quant"].size for k in ("conv_up", "bias", "conv_small", "embed6"))
    tables = sum(g[k]["min"].size + g[k]["scale"].size for k in ("conv_up", "bias", "conv_small", "embed6"))
    assert r["total_symbols"] == n and r["tmin_scale_len"] == tables
    assert r["total_bits"] == r["code_bits"] + 16 * tables
    allv = np.concatenate([g[k]["quant"].ravel() for k in ("embed6", "conv_up", "bias", "conv_small")])
    p = np.bincount(allv, minlength=256) / n
    entropy = -(p[p > 0] * np.log2(p[p > 0])).sum()
    assert entropy <= r["bits_per_param"] < entropy + 1.0 + 1e-3           # Huffman bound (the EOF leaf costs < 1e-3 bit here)
