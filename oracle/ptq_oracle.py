"""CPU restatement of the reference's post-training quantisation and Huffman bit accounting (SURVEY.md §8f rank 4,
the on-disk format either side of the decoder).  TEST INFRASTRUCTURE ONLY: imported by tests/ and nothing else; the
product path is csrc/ptq_ops.cu behind bnerv_b200/ptq.py.

  quant_tensor / dequant_tensor   hnerv_utils.py:101-134, 185-188   PINNED: tests/golden/ptq.npz holds outputs of the
                                  unmodified reference functions (tests/golden/make_golden_ptq.py), checked bit for bit.
  quant_model                     train_nerv_all.py:620-641         (state_dict walk; 'encoder' keys are left alone)
  huffman_code_lengths            dahuffman==0.4.1 (requirements.txt), HuffmanCodec.from_data -> from_frequencies.
                                  NOT vendored and NOT installed here: **parity unpinned** for the tie-breaking of equal
                                  frequencies; the package's published algorithm is restated below and anchored on the
                                  reference's call site (train_nerv_all.py:596-600) and on the invariants every Huffman
                                  code satisfies (Kraft equality, optimal total cost), which tests/ check.
  huffman_bits                    train_nerv_all.py:581-610         bits_per_param, full_bits_per_param, total bits
"""
import heapq

import numpy as np
import torch


def quant_tensor(t, bits=8):
    """hnerv_utils.py:101-134.  Candidates: one (min, scale) for the whole tensor (f32), plus one per axis whose
    reduced table is < 2 % of the tensor (i.e. t.shape[axis] > 50), stored as f16; the candidate with the smallest
    mean absolute reconstruction error wins (first one on ties: list.index of min)."""
    levels = 2 ** bits - 1
    cands = []
    t_min, t_max = t.min(), t.max()
    cands.append((t_min, (t_max - t_min) / levels))
    for axis in range(t.dim()):
        a_min, a_max = t.min(axis, keepdim=True)[0], t.max(axis, keepdim=True)[0]
        if a_min.nelement() / t.nelement() < 0.02:
            cands.append((a_min.to(torch.float16), ((a_max - a_min) / levels).to(torch.float16)))
    best = None
    for c_min, c_scale in cands:
        m, s = c_min.expand_as(t), c_scale.expand_as(t)
        q = ((t - m) / s).round().clamp(0, levels)
        new_t = m + s * q
        err = (t - new_t).abs().mean()
        if best is None or err < best[0]:
            best = (err, q, new_t, c_min, c_scale)
    _, q, new_t, c_min, c_scale = best
    return {"quant": q.to(torch.uint8), "min": c_min, "scale": c_scale}, new_t


def dequant_tensor(quant_t):
    """hnerv_utils.py:185-188."""
    q, tmin, scale = quant_t["quant"], quant_t["min"], quant_t["scale"]
    return tmin.expand_as(q) + scale.expand_as(q) * q


def quant_model_state(state_dict, bits):
    """train_nerv_all.py:627-637 on a state_dict: every non-encoder tensor goes through quant_tensor.
    -> (quant_ckt {key: {'quant','min','scale'}}, dequantised state_dict incl. the untouched encoder tensors)."""
    quant_ckt, cur_ckt = {}, {}
    for k, v in state_dict.items():
        if "encoder" in k:
            cur_ckt[k] = v
        else:
            quant_ckt[k], cur_ckt[k] = quant_tensor(v, bits)
    return quant_ckt, cur_ckt


class _EOF:
    """dahuffman's end-of-stream symbol: compares smaller than every real symbol."""

    def __lt__(self, other):
        return True

    def __gt__(self, other):
        return False

    def __eq__(self, other):
        return other.__class__ is self.__class__

    def __hash__(self):
        return hash(self.__class__)


def huffman_code_lengths(freq):
    """{symbol: count} -> {symbol: code length in bits}, dahuffman 0.4.1 HuffmanCodec.from_frequencies restated:
    an extra EOF leaf of frequency 1 joins the alphabet; heap items are (frequency, [(symbol, (bitsize, value))...]) tuples,
    so equal frequencies are ordered by their leaf lists (Python tuple/list comparison); the two smallest items merge, the
    first popped taking bit 0.  The EOF leaf's own length is not returned (the reference never charges it,
    train_nerv_all.py:603-605), but it does lengthen the codes around it."""
    eof = _EOF()
    heap = [(int(f), [(s, (0, 0))]) for s, f in freq.items()]
    heap.append((1, [(eof, (0, 0))]))
    heapq.heapify(heap)
    while len(heap) > 1:
        a = heapq.heappop(heap)
        b = heapq.heappop(heap)
        merged = (a[0] + b[0], [(s, (n + 1, v)) for (s, (n, v)) in a[1]] + [(s, (n + 1, (1 << n) + v)) for (s, (n, v)) in b[1]])
        heapq.heappush(heap, merged)
    table = dict(heap[0][1])
    return {s: nv[0] for s, nv in table.items() if not isinstance(s, _EOF)}


def huffman_bits(quant_ckt, quant_embed=None):
    """train_nerv_all.py:581-607.  -> dict(total_symbols, code_bits, bits_per_param, tmin_scale_len, total_bits,
    full_bits_per_param); total_bpp = total_bits / final_size / full_data_length is the caller's division (:610)."""
    parts, tmin_scale_len = [], 0
    if quant_embed is not None:
        parts.append(quant_embed["quant"].flatten().numpy())
        tmin_scale_len += quant_embed["min"].nelement() + quant_embed["scale"].nelement()
    for _, layer in quant_ckt.items():
        parts.append(layer["quant"].flatten().numpy())
        tmin_scale_len += layer["min"].nelement() + layer["scale"].nelement()
    allv = np.concatenate(parts)
    unique, counts = np.unique(allv, return_counts=True)
    freq = {int(u): int(c) for u, c in zip(unique, counts)}
    lengths = huffman_code_lengths(freq)
    code_bits = sum(freq[s] * lengths[s] for s in freq)
    total_bits = code_bits + tmin_scale_len * 16
    n = int(allv.size)
    return {"total_symbols": n, "code_bits": code_bits, "bits_per_param": code_bits / n, "tmin_scale_len": tmin_scale_len,
            "total_bits": total_bits, "full_bits_per_param": total_bits / n}
