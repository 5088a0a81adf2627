"""Post-training quantisation and Huffman bit accounting on the device kernels.

Same names, arguments and return values as the reference's helpers, so an evaluation script can swap them in:

  quant_tensor(t, bits=8)        hnerv_utils.py:101-134  -> ({'quant': uint8, 'min', 'scale'}, new_t)
  dequant_tensor(quant_t)        hnerv_utils.py:185-188
  quant_model(model, args)       train_nerv_all.py:620-641 -> ([model, quantised copy], quant_ckt)
  huffman_bits(quant_ckt, quant_embed=None)   the statistics of train_nerv_all.py:581-607

Everything is computed by csrc/ptq_ops.cu through the C-ABI (bnerv_ptq_quant_tensor, bnerv_histogram_u8,
bnerv_huffman_code_lengths); CPU tensors raise, there is no fallback.  Where the reference pulls every code to the host
(`.flatten().tolist()`, np.unique, a Python Huffman build over 15 M symbols) this reads back 256 counters.
"""
import ctypes
from copy import deepcopy

import torch

from . import _capi
from ._capi import PtqPlan, check, lib, ptr
from .ops import _need_cuda, _stream, require_current_device


def _plan(shape):
    arr = (ctypes.c_int64 * max(1, len(shape)))(*shape)
    plan = PtqPlan()
    check("bnerv_ptq_plan_tensor", lib.bnerv_ptq_plan_tensor(arr, len(shape), ctypes.byref(plan)))
    return arr, plan


class _Job:
    """One launched quant_tensor: device buffers + the plan needed to slice the winning tables once `best` is known."""

    def __init__(self, t, bits, want_new_t=True):
        _need_cuda(t)
        require_current_device(t.device)
        if t.dtype != torch.float32:
            raise TypeError(f"quant_tensor expects float32, got {t.dtype}")
        if not 1 <= bits <= 8:
            raise ValueError(f"bits = {bits}: codes are uint8 (1..8)")
        self.shape = tuple(t.shape)
        t = t.contiguous()
        arr, self.plan = _plan(self.shape)
        dev = t.device
        self.quant = torch.empty(self.shape, dtype=torch.uint8, device=dev)
        self.new_t = torch.empty(self.shape, dtype=torch.float32, device=dev) if want_new_t else None
        self.tables = torch.empty(self.plan.table_floats, dtype=torch.float32, device=dev)
        self.err = torch.empty(_capi.PTQ_MAX_CAND, dtype=torch.float64, device=dev)
        self.best = torch.empty(1, dtype=torch.int32, device=dev)
        scratch = torch.empty(self.plan.scratch_doubles, dtype=torch.float64, device=dev)
        check("bnerv_ptq_quant_tensor",
              lib.bnerv_ptq_quant_tensor(ptr(t), arr, len(self.shape), bits, ptr(self.quant), ptr(self.new_t), ptr(self.tables),
                                         ptr(self.err), ptr(self.best), ptr(scratch), _stream()))

    def result(self, best):
        """-> the reference's dict for candidate `best` (tables as views of the right keepdim shape and dtype)."""
        p = self.plan
        off, G, axis = p.table_offset[best], p.groups[best], p.axis[best]
        tmin, scale = self.tables[off:off + G], self.tables[off + G:off + 2 * G]
        if axis < 0:
            tmin, scale = tmin.reshape(()), scale.reshape(())                       # 0-dim f32, like t.min()
        else:
            keep = tuple(1 if d == axis else n for d, n in enumerate(self.shape))
            tmin, scale = tmin.reshape(keep).to(torch.float16), scale.reshape(keep).to(torch.float16)   # exact: already f16 values
        return {"quant": self.quant, "min": tmin, "scale": scale}


def quant_tensor(t, bits=8):
    job = _Job(t, bits)
    return job.result(int(job.best.item())), job.new_t


def dequant_tensor(quant_t):
    """hnerv_utils.py:185-188 verbatim semantics (dtype promotion included: f16 tables give an f16 result); not used by
    the reference's own scripts, which keep quant_tensor's f32 reconstruction."""
    q, tmin, scale = quant_t["quant"], quant_t["min"], quant_t["scale"]
    return tmin.expand_as(q) + scale.expand_as(q) * q


def reconstruct_tensor(quant_t):
    """Stored form ({'quant': uint8, 'min', 'scale'} as quant_tensor returns it) -> the f32 tensor quant_model loads into the
    quantised model (train_nerv_all.py:634-638), bit-identical to quant_tensor's second return value."""
    q, tmin, scale = quant_t["quant"], quant_t["min"], quant_t["scale"]
    _need_cuda(q, tmin, scale)
    require_current_device(q.device)
    if q.dtype != torch.uint8 or tmin.dtype != scale.dtype or tmin.shape != scale.shape:
        raise TypeError("reconstruct_tensor: expected uint8 codes and min / scale tables of one dtype and shape")
    shape = tuple(q.shape)
    if tmin.dim() == 0:
        axis = -1                                   # whole-tensor candidate
    else:                                           # keepdim table of the one reduced axis (extent > 50, hnerv_utils.py:110)
        red = [d for d, (a, b) in enumerate(zip(tmin.shape, shape)) if a == 1 and b != 1]
        if tmin.dim() != len(shape) or len(red) != 1 or any(a != b for d, (a, b) in enumerate(zip(tmin.shape, shape)) if d != red[0]):
            raise ValueError(f"reconstruct_tensor: tables of shape {tuple(tmin.shape)} are not a keepdim reduction of {shape}")
        axis = red[0]
    if tmin.dtype not in (torch.float16, torch.float32):
        raise TypeError(f"reconstruct_tensor: tables must be float16 or float32, got {tmin.dtype}")
    arr = (ctypes.c_int64 * max(1, len(shape)))(*shape)
    out = torch.empty(shape, dtype=torch.float32, device=q.device)
    q, tmin, scale = q.contiguous(), tmin.contiguous(), scale.contiguous()
    check("bnerv_ptq_dequant_tensor",
          lib.bnerv_ptq_dequant_tensor(ptr(q), arr, len(shape), axis, ptr(tmin), ptr(scale), int(tmin.dtype == torch.float16), ptr(out), _stream()))
    return out


def load_quant_ckt(model, quant_ckt):
    """Decode side of quant_model: fill `model`'s non-encoder tensors from the stored codes + tables (strict on keys)."""
    sd = model.state_dict()
    missing = [k for k in sd if "encoder" not in k and k not in quant_ckt]
    extra = [k for k in quant_ckt if k not in sd]
    if missing or extra:
        raise KeyError(f"load_quant_ckt: missing {missing[:3]}, unexpected {extra[:3]}")
    with torch.no_grad():
        for k, qt in quant_ckt.items():
            sd[k].copy_(reconstruct_tensor(qt))
    return model


def quant_state_dict(state_dict, bits):
    """quant_tensor over every non-encoder tensor of a state_dict (the loop of train_nerv_all.py:630-636), all launches
    issued before the single read-back of the winning candidate indices.
    -> (quant_ckt {key: {'quant','min','scale'}}, {key: dequantised tensor} incl. the untouched encoder tensors)."""
    jobs, cur = {}, {}
    for k, v in state_dict.items():
        if "encoder" in k:
            cur[k] = v
        else:
            jobs[k] = _Job(v, bits)
    if jobs:
        best = torch.cat([j.best for j in jobs.values()]).tolist()
        quant_ckt = {}
        for (k, j), b in zip(jobs.items(), best):
            quant_ckt[k] = j.result(b)
            cur[k] = j.new_t
    else:
        quant_ckt = {}
    return quant_ckt, {k: cur[k] for k in state_dict}


def quant_model(model, args):
    """train_nerv_all.py:620-641: -> ([copy of model, copy with dequantised decoder weights], quant_ckt), or
    ([copy], None) when args.quant_model_bit == -1."""
    model_list = [deepcopy(model)]
    if args.quant_model_bit == -1:
        return model_list, None
    cur_model = deepcopy(model)
    quant_ckt, cur_ckt = quant_state_dict(cur_model.state_dict(), args.quant_model_bit)
    cur_model.load_state_dict(cur_ckt)
    model_list.append(cur_model)
    return model_list, quant_ckt


def code_histogram(quant_ckt, quant_embed=None):
    """-> (u64[256] counts of all codes on the device, number of min + scale entries stored beside them)."""
    layers = ([quant_embed] if quant_embed is not None else []) + list(quant_ckt.values())
    if not layers:
        raise ValueError("nothing to count")
    dev = layers[0]["quant"].device
    require_current_device(dev)
    counts = torch.zeros(256, dtype=torch.int64, device=dev)
    tmin_scale_len = 0
    for layer in layers:
        q = layer["quant"]
        _need_cuda(q)
        if q.dtype != torch.uint8:
            raise TypeError("codes must be uint8")
        q = q.contiguous()
        check("bnerv_histogram_u8", lib.bnerv_histogram_u8(ptr(q), q.numel(), ptr(counts), _stream()))
        tmin_scale_len += layer["min"].nelement() + layer["scale"].nelement()
    return counts, tmin_scale_len


def huffman_code_lengths(counts):
    """counts: sequence / tensor of non-negative ints per symbol -> list of code lengths (0 for absent symbols)."""
    c = torch.as_tensor(counts, dtype=torch.int64).cpu().contiguous()
    lengths = torch.zeros(c.numel(), dtype=torch.int32)
    check("bnerv_huffman_code_lengths",
          lib.bnerv_huffman_code_lengths(ctypes.c_void_p(c.data_ptr()), c.numel(), ctypes.c_void_p(lengths.data_ptr())))
    return lengths.tolist()


def huffman_bits(quant_ckt, quant_embed=None):
    """train_nerv_all.py:581-607: Huffman-coded size of all codes (+ 16 bits per stored min / scale entry).
    -> dict(total_symbols, code_bits, bits_per_param, tmin_scale_len, total_bits, full_bits_per_param); the caller's
    bits per pixel is total_bits / final_size / full_data_length (:610)."""
    counts, tmin_scale_len = code_histogram(quant_ckt, quant_embed)
    counts = counts.cpu()
    lengths = huffman_code_lengths(counts)
    n = int(counts.sum())
    code_bits = sum(int(c) * l for c, l in zip(counts.tolist(), lengths))
    total_bits = code_bits + tmin_scale_len * 16
    return {"total_symbols": n, "code_bits": code_bits, "bits_per_param": code_bits / n, "tmin_scale_len": tmin_scale_len,
            "total_bits": total_bits, "full_bits_per_param": total_bits / n}
