"""Per-launch device time of every fused conv of one decoded frame (CUDA events on the launching stream, averaged over
steps), with algorithmic TFLOP/s.  Usage: python tools/layer_times.py [config] [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "boosting-nerv_b200"))
import bench  # noqa: E402
from bnerv_b200 import ops  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "hnerv_l"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
model, args = bench.build_model(cfg)
model = model.cuda()
is_h = args.model == "HNeRV_Boost"
fh, fw = [int(v) for v in args.fc_hw.split("_")]
emb = torch.rand(1, 16, fh, fw, device="cuda")
t = torch.tensor([0.5], dtype=torch.float64, device="cuda")
run = (lambda: model.decode(emb, t)) if is_h else (lambda: model.decode(t))
with torch.no_grad():
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(steps):
        run()
    g1.record()
    torch.cuda.synchronize()
    print(f"{cfg}: CUDA-graph replay {g0.elapsed_time(g1) / steps:.3f} ms/frame ({1e3 * steps / g0.elapsed_time(g1):.1f} frames/s)")
    model.engine().use_graph = False          # per-launch events need eager launches
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    ops.TIMING = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        run()
    e1.record()
    torch.cuda.synchronize()
timing, ops.TIMING = ops.TIMING, None
per = len(timing) // steps
tot_ms = e0.elapsed_time(e1) / steps
print(f"{cfg}: eager with per-launch events {tot_ms:.3f} ms/frame ({1e3 / tot_ms:.1f} frames/s), {per} fused-conv launches per frame")
print(f"{'#':>3s} {'cin':>4s} {'cout':>4s} k s {'H':>5s} {'W':>5s} {'act':>6s} {'ms':>8s} {'share':>6s} {'TFLOP/s':>8s}")
acc = 0.0
for i in range(per):
    rows = [timing[i + j * per] for j in range(steps)]
    ms = sum(r[1].elapsed_time(r[2]) for r in rows) / steps
    fl, shp = rows[0][0], rows[0][3]
    acc += ms
    print(f"{i:3d} {shp[0]:4d} {shp[1]:4d} {shp[2]} {shp[3]} {shp[4]:5d} {shp[5]:5d} {shp[6]:>6s} {ms:8.4f} {100 * ms / tot_ms:5.1f}% {fl / ms / 1e9:8.1f}")
print(f"sum of conv launches {acc:.3f} ms = {100 * acc / tot_ms:.1f}% of the frame")
