"""Turn an ncu CSV (`--metrics dram__bytes_read.sum,dram__bytes_write.sum --kernel-name-base demangled -k regex:bnerv`: only this
library's kernels are captured; the CSV prints their names without the namespace) of decoded frames into
profiles/frame_traffic_<config>.json (what bench.py reports as roofline.traffic), next to the bytes the same launches must
move by design (C8 f16 input (+ residual) + output map(s) + packed weights per launch; a fused block: input + output + weights;
the head writes NCHW f32).  `launches per frame` = this library's kernels in one decoded frame (the last frame is used).
Usage: python tools/frame_traffic.py <csv> <config> <launches per frame> "<command>" """
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "boosting-nerv_b200"))

path, cfg, per_frame, cmd = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
rows = [r for r in csv.DictReader(l for l in open(path) if not l.startswith("==")) if r.get("Kernel Name") and "at::" not in r["Kernel Name"] and "cutlass" not in r["Kernel Name"]]
unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
rd = [float(r["Metric Value"].replace(",", "")) * unit[r["Metric Unit"]] for r in rows if r["Metric Name"] == "dram__bytes_read.sum"]
wr = [float(r["Metric Value"].replace(",", "")) * unit[r["Metric Unit"]] for r in rows if r["Metric Name"] == "dram__bytes_write.sum"]
if per_frame <= 0:                               # the capture holds exactly one frame (ncu --profile-from-start off + tools/frame_once.py)
    per_frame = len(rd)
assert len(rd) == len(wr) and len(rd) >= per_frame, (len(rd), per_frame)
rd, wr = rd[-per_frame:], wr[-per_frame:]        # the last frame's launches (the first frame includes cold weights)


def algorithmic_bytes(cfg):
    """Walk the preset's conv list the way engine.run_cascade launches it."""
    import torch
    import bench
    from bnerv_b200.engine import DecoderEngine
    model, args = bench.build_model(cfg)
    eng = DecoderEngine(model)
    r16 = lambda c: (c + 15) // 16 * 16
    fh, fw = [int(v) for v in args.fc_hw.split("_")]
    H, W, tot = fh, fw, 0
    cin = eng.blocks[0].pre.cin if eng.blocks[0].pre is not None else eng.blocks[0].up.cin
    for blk in eng.blocks:
        if blk.pre is not None:
            s = blk.pre
            tot += 2 * r16(s.cin) * H * W + 2 * r16(s.cout) * H * W * s.s ** 2 + 2 * s.k ** 2 * r16(s.cin) * r16(s.cout) * s.s ** 2
            H, W = H * s.s, W * s.s
        s = blk.up
        Ho, Wo = H * s.s, W * s.s
        m = 2 * r16(blk.cout) * Ho * Wo                       # one C8 f16 map at the block's output resolution
        w_up, w_c = 2 * s.k ** 2 * r16(s.cin) * r16(s.cout) * s.s ** 2, 2 * 9 * r16(blk.cout) ** 2
        if blk.fuse is not None and blk.fuse[1] == "block":   # one kernel per block: input, output, three weight sets
            tot += 2 * r16(s.cin) * H * W + m + w_up + 2 * w_c
        elif blk.fuse is not None:                            # up-conv launch (x0 + u) + one kernel for the ResBlock_SFT half
            tot += 2 * r16(s.cin) * H * W + 2 * m + w_up
            tot += 3 * m + 2 * w_c                            # u, x0 -> out
        else:
            tot += 2 * r16(s.cin) * H * W + 2 * m + w_up      # up: in, x0 + u
            tot += 2 * m + w_c                                # conv0: u -> w
            tot += 3 * m + w_c                                # conv1: w, x0 -> out
        H, W = Ho, Wo
    s = eng.head
    tot += 2 * r16(s.cin) * H * W + 4 * 3 * H * W + 2 * s.k ** 2 * r16(s.cin) * 32
    return tot


out = {"config": cfg, "what": f"dram__bytes_read.sum + dram__bytes_write.sum summed over the {per_frame} bnerv:: kernel launches of one decoded frame (ncu)",
       "dram_read_bytes": sum(rd), "dram_write_bytes": sum(wr), "launches": per_frame, "command": cmd,
       "algorithmic_bytes_per_frame": algorithmic_bytes(cfg),
       "algorithmic_bytes_note": "sum over the launches of C8 f16 input (+ residual) + output map(s) + packed weights, i.e. what a launch-fused conv must move"}
json.dump(out, open(os.path.join(ROOT, "profiles", f"frame_traffic_{cfg}.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
