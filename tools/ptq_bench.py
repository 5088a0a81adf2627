"""Evaluation-side quantisation chain of a full-size model (train_nerv_all.py:620-641, 542, 581-607): native kernels
(bnerv_b200.ptq) vs the reference's formulation in torch ops on the same GPU tensors, wall-clock incl. device sync.
Usage: python tools/ptq_bench.py [hnerv_l]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "boosting-nerv_b200"))
import bench  # noqa: E402
from bnerv_b200 import ptq  # noqa: E402


def torch_quant_tensor(t, bits=8):
    """The torch formulation hnerv_utils.quant_tensor executes (inlined: tools do not import oracle/)."""
    lv = 2 ** bits - 1
    cands = [(t.min(), (t.max() - t.min()) / lv)]
    for ax in range(t.dim()):
        lo, hi = t.min(ax, keepdim=True)[0], t.max(ax, keepdim=True)[0]
        if lo.nelement() / t.nelement() < 0.02:
            cands.append((lo.half(), ((hi - lo) / lv).half()))
    outs = []
    for lo, sc in cands:
        q = ((t - lo.expand_as(t)) / sc.expand_as(t)).round().clamp(0, lv)
        nt = lo.expand_as(t) + sc.expand_as(t) * q
        outs.append(((t - nt).abs().mean(), q, nt, lo, sc))
    errs = [o[0] for o in outs]
    b = errs.index(min(errs))
    return {"quant": outs[b][1].to(torch.uint8), "min": outs[b][3], "scale": outs[b][4]}, outs[b][2]


def wall(fn, reps=3):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3, out


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "hnerv_l"
    model, args = bench.build_model(name)
    model = model.cuda()
    sd = {k: v for k, v in model.state_dict().items() if "encoder" not in k}
    n_par = sum(v.numel() for v in sd.values())
    fh, fw = [int(v) for v in args.fc_hw.split("_")]
    emb = torch.rand(bench.N_FRAMES, 16, fh, fw, device="cuda")
    print(f"{name}: {len(sd)} decoder tensors, {n_par / 1e6:.2f} M parameters + {emb.numel() / 1e6:.2f} M embedding values")

    def native():
        ckt, _ = ptq.quant_state_dict(sd, 8)
        qe, _ = ptq.quant_tensor(emb, 6)
        return ckt, qe

    def reference():
        ckt = {k: torch_quant_tensor(v, 8)[0] for k, v in sd.items()}
        return ckt, torch_quant_tensor(emb, 6)[0]

    t_nat, (ckt, qe) = wall(native)
    t_ref, (ckt_r, qe_r) = wall(reference)
    print(f"quantise decoder (8 bit) + embeddings (6 bit): native {t_nat:.2f} ms, torch ops on this GPU {t_ref:.2f} ms")
    # torch's CUDA kernel divides by the Python scalar 2^bits-1 as a multiplication by its f32 reciprocal, torch's CPU kernel (the
    # goldens, the oracle, this library) as a true division: scales can differ by 1 ulp, so a few codes on rounding boundaries move
    pairs = [(ckt[k], ckt_r[k]) for k in ckt] + [(qe, qe_r)]
    diff = sum(int((a["quant"] != b["quant"]).sum()) for a, b in pairs)
    worst = max(int((a["quant"].int() - b["quant"].int()).abs().max()) for a, b in pairs)
    ident = sum(torch.equal(a["quant"], b["quant"]) for a, b in pairs)
    print(f"vs torch-on-GPU codes: {ident}/{len(pairs)} tensors identical, {diff} of {n_par + emb.numel()} codes differ, by at most {worst} level "
          f"(scalar division as reciprocal multiply on CUDA; bit-exactness vs the CPU reference is what tests/test_gpu_ptq.py gates)")
    t_dec, _ = wall(lambda: [ptq.reconstruct_tensor(v) for v in ckt.values()])
    print(f"decode side (codes + tables -> f32 weights, {len(ckt)} tensors): {t_dec:.2f} ms")

    t_bits, bits = wall(lambda: ptq.huffman_bits(ckt, qe))

    def reference_stats():      # train_nerv_all.py:583-593 literally (dahuffman's own pass over the list comes on top)
        v = qe["quant"].flatten().tolist()
        for layer in ckt.values():
            v.extend(layer["quant"].flatten().tolist())
        return np.unique(v, return_counts=True)

    t_stats, (_, counts) = wall(reference_stats, reps=1)
    assert int(counts.sum()) == bits["total_symbols"]
    pix = fh * fw * int(np.prod(args.dec_strds)) ** 2
    print(f"Huffman statistics: native {t_bits:.2f} ms, reference .tolist() + np.unique alone {t_stats:.0f} ms")
    print(f"bits per parameter {bits['bits_per_param']:.3f} (with tables {bits['full_bits_per_param']:.3f}), "
          f"bits per pixel over {bench.N_FRAMES} frames {bits['total_bits'] / pix / bench.N_FRAMES:.5f}")


if __name__ == "__main__":
    main()
