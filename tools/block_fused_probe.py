"""Bring-up / timing probe of the one-kernel NeRVBlock (csrc/block_fused.cu): every case against the three-launch path
(bit identity expected; mismatch statistics are printed, nothing asserts), then device timings of both forms at the
narrow stages of the benchmarked presets.  Usage: python tools/block_fused_probe.py [--time-only] [--check-only]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "boosting-nerv_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from bnerv_b200 import ops  # noqa: E402
from test_gpu_block_fused import CASES, make_block  # noqa: E402


def describe(out, ref, C):
    if torch.equal(out, ref):
        return "bit-identical"
    o, r = out.float(), ref.float()
    d = (o - r).abs()
    bad = d > 0
    B, G, H, W, _ = out.shape
    msg = f"{int(bad.sum())}/{bad.numel()} differ, max|d| {d.max().item():.3e} (max|ref| {r.abs().max().item():.3e}), nan out {int(torch.isnan(o).sum())}"
    by = bad.any(dim=4).any(dim=1).any(dim=0)                # [H, W]
    ys, xs = by.any(dim=1).nonzero().flatten(), by.any(dim=0).nonzero().flatten()
    msg += f"; rows {ys.min().item()}..{ys.max().item()} ({len(ys)}), cols {xs.min().item()}..{xs.max().item()} ({len(xs)})"
    gs = bad.any(dim=4).any(dim=3).any(dim=2).any(dim=0).nonzero().flatten().tolist()
    msg += f"; groups {gs}; first {bad.nonzero()[:4].tolist()}"
    big = d > 1e-2 * r.abs().max()
    msg += f"; >1% of range: {int(big.sum())}"
    return msg


def check_stream():
    from test_gpu_block_fused import STREAM_CASES
    for case in STREAM_CASES:
        B, cin, C, H, W, s = case
        x, up, c0, c1, (g0, b0, g1, b1) = make_block(ops, *case)
        ref, x0 = ops.nerv_block_fwd(x, up, c0, c1, cin, H, W, "sin", "gelu", g0, b0, g1, b1)
        mk = lambda: torch.empty_like(ref)
        x0b, u = mk(), mk()
        ops.conv_fused(x, up, cin, H, W, act="sin", g1p=g0, beta=b0, out_pre=x0b, out_aff=u)
        for name, fn in (("resblock", lambda o: ops.resblock_fused(u, x0b, c0, c1, C, H * s, W * s, "gelu", g1, b1, out=o, form="stream")),
                         ("block", lambda o: ops.nerv_block_fused(x, up, c0, c1, cin, H, W, "sin", "gelu", g0, b0, g1, b1, out=o, form="stream"))):
            try:
                out = torch.full_like(ref, float("nan"))
                res = fn(out)
                torch.cuda.synchronize()
                print(f"stream {name} {case}: " + ("UNSUPPORTED" if res is None else describe(out, ref, C)), flush=True)
            except Exception as ex:
                print(f"stream {name} {case}: EXCEPTION {ex}", flush=True)
                return


def check():
    for case in CASES:
        B, cin, C, H, W, s = case
        x, up, c0, c1, (g0, b0, g1, b1) = make_block(ops, *case)
        ref, x0 = ops.nerv_block_fwd(x, up, c0, c1, cin, H, W, "sin", "gelu", g0, b0, g1, b1)
        try:
            out = torch.full_like(ref, float("nan"))
            res = ops.nerv_block_fused(x, up, c0, c1, cin, H, W, "sin", "gelu", g0, b0, g1, b1, out=out)
            torch.cuda.synchronize()
            print(f"block {case}: " + ("UNSUPPORTED" if res is None else describe(out, ref, C)), flush=True)
        except Exception as ex:          # a trapped kernel poisons the context: report and stop
            print(f"block {case}: EXCEPTION {ex}", flush=True)
            return
        if s == 1 and cin == C:
            mk = lambda: torch.empty_like(ref)
            u, wm, ref2 = mk(), mk(), mk()
            x0b = mk()
            ops.conv_fused(x, up, C, H, W, act="sin", g1p=g0, beta=b0, out_pre=x0b, out_aff=u)
            ops.conv_fused(u, c0, C, H, W, act="gelu", g1p=g1, beta=b1, out_aff=wm)
            ops.conv_fused(wm, c1, C, H, W, act="none", resid=x0b, out_pre=ref2)
            try:
                out = torch.full_like(ref, float("nan"))
                res = ops.resblock_fused(u, x0b, c0, c1, C, H, W, "gelu", g1, b1, out=out)
                torch.cuda.synchronize()
                print(f"  resblock: " + ("UNSUPPORTED" if res is None else describe(out, ref2, C)), flush=True)
            except Exception as ex:
                print(f"  resblock: EXCEPTION {ex}", flush=True)
                return


def timed(fn, iters=20):
    """Device time per call: `iters` calls captured in one CUDA graph (no host launch cost between them), replayed 3 times."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters)
    return best


def timing():
    shapes = [  # B, cin, C, H, W, s   (input resolution)
        (1, 12, 12, 720, 1280, 1), (1, 12, 12, 360, 640, 2), (1, 12, 12, 360, 640, 1), (1, 12, 12, 180, 320, 2),
        (1, 12, 12, 180, 320, 1), (1, 15, 12, 90, 160, 2), (1, 30, 15, 45, 80, 2),
        (1, 21, 21, 1080, 1920, 1), (1, 43, 21, 540, 960, 2), (1, 43, 43, 540, 960, 1), (8, 12, 12, 720, 1280, 1),
    ]
    print(f"{'case':36s} {'3 launches ms':>14s} {'fused ms':>10s} {'x':>6s} {'GB/s in+out':>12s}")
    for case in shapes:
        B, cin, C, H, W, s = case
        x, up, c0, c1, (g0, b0, g1, b1) = make_block(ops, *case)
        cp = ops.round_up(C, 16)
        mk = lambda: torch.empty(ops.c8_shape(B, C, H * s, W * s), dtype=torch.float16, device="cuda")
        x0, u, wm, out = mk(), mk(), mk(), mk()

        def three():
            ops.conv_fused(x, up, cin, H, W, act="sin", g1p=g0, beta=b0, out_pre=x0, out_aff=u)
            ops.conv_fused(u, c0, C, H * s, W * s, act="gelu", g1p=g1, beta=b1, out_aff=wm)
            ops.conv_fused(wm, c1, C, H * s, W * s, act="none", resid=x0, out_pre=out)

        def fused():
            ops.nerv_block_fused(x, up, c0, c1, cin, H, W, "sin", "gelu", g0, b0, g1, b1, out=out)

        def stream():
            ops.nerv_block_fused(x, up, c0, c1, cin, H, W, "sin", "gelu", g0, b0, g1, b1, out=out, form="stream")

        def up_stream():
            ops.conv_fused(x, up, cin, H, W, act="sin", g1p=g0, beta=b0, out_pre=x0, out_aff=u)
            ops.resblock_fused(u, x0, c0, c1, C, H * s, W * s, "gelu", g1, b1, out=out, form="stream")

        t3, tf = timed(three), timed(fused)
        bytes_io = 2.0 * B * (ops.round_up(cin, 16) * H * W + cp * H * s * W * s)
        msg = f"{str(case):36s} {t3:14.4f} {tf:10.4f} {t3 / tf:6.2f} {bytes_io / tf / 1e6:12.1f}"
        if cp == 16:
            if cin <= 16:
                ts = timed(stream)
                msg += f" | stream {ts:.4f} ms ({t3 / ts:.2f}x, {bytes_io / ts / 1e6:.0f} GB/s)"
            tu = timed(up_stream)
            msg += f" | up + resblock-stream {tu:.4f} ms ({t3 / tu:.2f}x)"
        print(msg, flush=True)


def stamps(case, n_ctas=300, show=(0, 1, 148, 149, 295)):
    """Phase durations (clock cycles) of the first regions of the first CTAs: see BF_STAMP in csrc/block_fused.cu."""
    from bnerv_b200._capi import lib, ptr
    B, cin, C, H, W, s = case
    x, up, c0, c1, (g0, b0, g1, b1) = make_block(ops, *case)
    out = torch.empty(ops.c8_shape(B, C, H * s, W * s), dtype=torch.float16, device="cuda")
    for _ in range(2):
        ops.nerv_block_fused(x, up, c0, c1, cin, H, W, "sin", "gelu", g0, b0, g1, b1, out=out)
    buf = torch.zeros(n_ctas * 4 * 12, dtype=torch.int64, device="cuda")
    lib.bnerv_debug_set_buffer(ptr(buf), n_ctas)
    ops.nerv_block_fused(x, up, c0, c1, cin, H, W, "sin", "gelu", g0, b0, g1, b1, out=out)
    torch.cuda.synchronize()
    lib.bnerv_debug_set_buffer(None, 0)
    t = buf.cpu().view(n_ctas, 4, 12)
    names = ["up issue", "up mma done", "up epi+sync", "c0 issue", "c0 mma done", "c0 epi+sync", "c1 issue", "c1 epi", "c1 sync"]
    print(f"stamps {case}: cycles per phase [cta][region]; columns: " + " | ".join(names) + " | region total | smid")
    t0c, t0g = int(t[0, 0, 0]), int(t[0, 0, 10])
    for c in show:
        for r in range(4):
            st = t[c, r]
            if st[9] == 0:
                continue
            print(f"    [cta {c} reg {r}] start: clock64 - cta0 = {int(st[0]) - t0c}, globaltimer - cta0 = {int(st[10]) - t0g} ns")
            d = [int(st[k + 1] - st[k]) if st[k + 1] and st[k] else 0 for k in range(9)]
            print(f"  cta {c} reg {r}: " + " ".join(f"{v:6d}" for v in d) + f" | {int(st[9] - st[0]):7d} | sm {int(st[11])}"
                  + (f" | gap to next {int(t[c, r + 1, 0] - st[9])}" if r < 3 and t[c, r + 1, 0] else ""))


def stream_stamps(case=(1, 12, 12, 720, 1280, 1)):
    """Per-role clock64 stamps of CTA 0 of the streaming kernel (BS_STAMP in csrc/block_stream.cu)."""
    from bnerv_b200._capi import lib, ptr
    B, cin, C, H, W, s = case
    x, up, c0, c1, (g0, b0, g1, b1) = make_block(ops, *case)
    out = torch.empty(ops.c8_shape(B, C, H, W), dtype=torch.float16, device="cuda")
    run = lambda: ops.nerv_block_fused(x, up, c0, c1, cin, H, W, "sin", "gelu", g0, b0, g1, b1, out=out, form="stream")
    run(); run()
    buf = torch.zeros(9 * 32 * 4, dtype=torch.int64, device="cuda")
    lib.bnerv_debug_set_buffer(ptr(buf), 1)
    run()
    torch.cuda.synchronize()
    lib.bnerv_debug_set_buffer(None, 0)
    t = buf.cpu().view(9, 32, 4)
    t0 = int(t[t > 0].min())
    names = ["front0", "front1", "front2", "mid0", "mid1", "back", "issue up", "issue c0", "issue c1"]
    print(f"stream stamps {case} (cycles since the first stamp; per iteration: 4 stamps)")
    for r in range(9):
        print(f" {names[r]}:")
        for it in range(4, 14):
            st = [int(v) - t0 if v else -1 for v in t[r, it]]
            print(f"   it {it:2d}: {st}   deltas {[st[i + 1] - st[i] for i in range(3)]}" + (f"  period {st[0] - (int(t[r, it - 1, 0]) - t0)}" if t[r, it - 1, 0] else ""))


if __name__ == "__main__":
    torch.manual_seed(0)
    if "--stream-stamps" in sys.argv:
        stream_stamps()
        sys.exit(0)
    if "--stream-check" in sys.argv:
        check_stream()
        sys.exit(0)
    if "--time-only" not in sys.argv:
        check()
    if "--stamps" in sys.argv:
        stamps((1, 12, 12, 720, 1280, 1))
        stamps((1, 12, 12, 360, 640, 2))
        stamps((1, 21, 21, 1080, 1920, 1))
    elif "--check-only" not in sys.argv:
        timing()
