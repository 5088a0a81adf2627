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


class _Jobs:
    """quant_tensor over a LIST of tensors as one bnerv_ptq_quant_tensors call: five launches over a descriptor table, and six
    flat device buffers (codes, reconstructions, f32 + f16 tables, errors, winners) that the per-tensor results are views of."""

    def __init__(self, tensors, bits, want_new_t=True):
        if not 1 <= bits <= 8:
            raise ValueError(f"bits = {bits}: codes are uint8 (1..8)")
        _need_cuda(*tensors)
        dev = tensors[0].device
        require_current_device(dev)
        self.shapes, self.plans, src = [], [], []
        n_tot = tab_tot = 0
        self.offs = []                                  # (element offset, table offset) per tensor
        for t in tensors:
            if t.dtype != torch.float32:
                raise TypeError(f"quant_tensor expects float32, got {t.dtype}")
            if t.device != dev:
                raise ValueError("all tensors of one call must live on one device")
            shape = tuple(t.shape)
            _, plan = _plan(shape)
            self.shapes.append(shape)
            self.plans.append(plan)
            src.append(t.contiguous())
            self.offs.append((n_tot, tab_tot))
            n_tot += t.numel()                          # back to back: the codes of a model are ONE contiguous byte string
            tab_tot += plan.table_floats
        n = len(tensors)
        self.quant = torch.empty(n_tot, dtype=torch.uint8, device=dev)
        self.new_t = torch.empty(n_tot, dtype=torch.float32, device=dev) if want_new_t else None
        self.tables = torch.empty(tab_tot, dtype=torch.float32, device=dev)
        self.tables16 = torch.empty(tab_tot, dtype=torch.float16, device=dev)
        self.err = torch.empty(n * _capi.PTQ_MAX_CAND, dtype=torch.float64, device=dev)
        self.best = torch.empty(n, dtype=torch.int32, device=dev)
        jobs = (_capi.PtqJob * n)()
        q0, nt0, tb0, th0, e0, b0 = (self.quant.data_ptr(), self.new_t.data_ptr() if want_new_t else 0, self.tables.data_ptr(),
                                     self.tables16.data_ptr(), self.err.data_ptr(), self.best.data_ptr())
        for i, (t, shape, (eo, to)) in enumerate(zip(src, self.shapes, self.offs)):
            j = jobs[i]
            j.t, j.ndim = t.data_ptr(), len(shape)
            for d, v in enumerate(shape):
                j.shape[d] = v
            j.quant, j.new_t = q0 + eo, (nt0 + 4 * eo) if want_new_t else None
            j.tables, j.tables_f16 = tb0 + 4 * to, th0 + 2 * to
            j.err, j.best = e0 + 8 * _capi.PTQ_MAX_CAND * i, b0 + 4 * i
        need = lib.bnerv_ptq_quant_tensors_scratch_bytes(jobs, n)
        if need == 0:
            check("bnerv_ptq_quant_tensors", lib.bnerv_ptq_quant_tensors(jobs, n, bits, None, 0, _stream()))   # reports the bad job
        scratch = torch.empty(need, dtype=torch.uint8, device=dev)
        self._keep = src
        check("bnerv_ptq_quant_tensors", lib.bnerv_ptq_quant_tensors(jobs, n, bits, ptr(scratch), need, _stream()))

    def results(self):
        """-> [(the reference's dict {'quant','min','scale'}, new_t or None)] - ONE read-back (the winning candidates)."""
        out = []
        for i, best in enumerate(self.best.tolist()):
            shape, p, (eo, to) = self.shapes[i], self.plans[i], self.offs[i]
            numel = 1
            for v in shape:
                numel *= v
            off, G, axis = to + p.table_offset[best], p.groups[best], p.axis[best]
            if axis < 0:                                # 0-dim f32, like t.min()
                tmin, scale = self.tables[off:off + G].reshape(()), self.tables[off + G:off + 2 * G].reshape(())
            else:                                       # keepdim f16 tables (hnerv_utils.py:113); the f16 copy is exact
                keep = tuple(1 if d == axis else v for d, v in enumerate(shape))
                tmin, scale = self.tables16[off:off + G].reshape(keep), self.tables16[off + G:off + 2 * G].reshape(keep)
            q = {"quant": self.quant[eo:eo + numel].view(shape), "min": tmin, "scale": scale}
            out.append((q, None if self.new_t is None else self.new_t[eo:eo + numel].view(shape)))
        return out


def quant_tensor(t, bits=8):
    return _Jobs([t], bits).results()[0]


def quant_tensors(tensors, bits=8):
    """quant_tensor for every tensor of a list in one multi-tensor call -> [(quant dict, new_t)]."""
    return _Jobs(list(tensors), bits).results() if len(tensors) else []


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
    in ONE multi-tensor call (five launches) and one read-back of the winning candidate indices.
    -> (quant_ckt {key: {'quant','min','scale'}}, {key: dequantised tensor} incl. the untouched encoder tensors)."""
    keys = [k for k in state_dict if "encoder" not in k]
    cur = {k: v for k, v in state_dict.items() if "encoder" in k}
    quant_ckt = {}
    for k, (q, new_t) in zip(keys, quant_tensors([state_dict[k] for k in keys], bits)):
        quant_ckt[k] = q
        cur[k] = new_t
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


def _flat_codes(qs):
    """If (a subset of) the uint8 code tensors tile one storage back to back - the layout quant_tensors produces - return
    (address, bytes, membership test) of that byte string, else None."""
    groups = {}
    for q in qs:
        if q.is_cuda and q.dtype == torch.uint8 and q.is_contiguous():
            groups.setdefault(q.untyped_storage().data_ptr(), []).append(q)
    best = max(groups.values(), key=len, default=[])
    if len(best) < 2:
        return None
    best = sorted(best, key=lambda q: q.data_ptr())
    pos = best[0].data_ptr()
    for q in best:
        if q.data_ptr() != pos:
            return None
        pos += q.numel()
    ids = {id(q) for q in best}
    return ctypes.c_void_p(best[0].data_ptr()), pos - best[0].data_ptr(), (lambda q: id(q) in ids)


def code_histogram(quant_ckt, quant_embed=None):
    """-> (u64[256] counts of all codes on the device, number of min + scale entries stored beside them)."""
    layers = ([quant_embed] if quant_embed is not None else []) + list(quant_ckt.values())
    if not layers:
        raise ValueError("nothing to count")
    dev = layers[0]["quant"].device
    require_current_device(dev)
    counts = torch.zeros(256, dtype=torch.int64, device=dev)
    tmin_scale_len = 0
    flat = _flat_codes([l["quant"] for l in layers])
    if flat is not None:                           # codes produced by one quant_tensors call: one launch over the whole string
        check("bnerv_histogram_u8", lib.bnerv_histogram_u8(flat[0], flat[1], ptr(counts), _stream()))
        for layer in layers:
            if flat[2](layer["quant"]):
                tmin_scale_len += layer["min"].nelement() + layer["scale"].nelement()
        layers = [l for l in layers if not flat[2](l["quant"])]
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
