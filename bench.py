#!/usr/bin/env python
"""bench.py — decode frames/s of the Boosting-NeRV conditional decoder on B200 (driver contract).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port), rank 0 only

Workload (config.workload): HNeRV-Boost L (15M, fc_dim 280) decoder at 1920x1080 — BASELINE.json configs[3],
the configuration the metric "decode frames/sec @1920x1080" is quoted on; it fits one GPU, and at N GPUs the
frames are sharded round-robin with no data-path collective (weak scaling: one frame per rank per step).
A step = one decoded frame per rank: PE -> stem_t MLP -> 9 NeRV blocks (27 fused convs) -> head conv.
Synthetic data: reference-architecture random-init weights under torch.manual_seed(1); embeddings U(0,1)
[16,9,16] per frame; norm_idx=(i+1)/600 (hnerv_utils.py:47).

value : frames/s with embeddings + indices resident in HBM, CUDA-event timed, max over ranks.
e2e   : the same through model.forward_decoder() (the call the reference's evaluate() makes,
        train_nerv_all.py:482-486) with pinned-host inputs copied H2D and the decoded f32 frame copied D2H
        inside the timed region.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "boosting-nerv_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

N_FRAMES = 600          # UVG sequence length; fixes norm_idx = (i+1)/600
METRIC = "decode frames/sec @1920x1080"
ALG_GFLOP = {"hnerv_l": 4429.9, "enerv_m": 443.3, "nerv_s": 19.33, "nerv_xs": 19.04, "hnerv_m": 2782.0}
# SURVEY.md §8d algorithmic bytes per frame: block-fused, fp32 I/O (block input + block output + weights, summed over blocks)
ALG_BYTES = {"hnerv_l": 4.35e9, "enerv_m": 1.15e9, "nerv_s": 190e6, "nerv_xs": 190e6}
# which roofline bounds the preset: configs 3-5 tensor cores, configs 1-2 HBM bandwidth + launch latency
BOUND = {"hnerv_l": "tensor", "hnerv_m": "tensor", "enerv_m": "tensor", "nerv_s": "hbm", "nerv_xs": "hbm"}


def preset_roofline(name, ms_per_frame, pk, pk_src, alg_gflop):
    """Step-level roofline of one decoded frame: §8d algorithmic work / the timed ms per frame, against the measured peak of
    the bounding resource; the other resource's fraction is reported beside it."""
    tf = alg_gflop / ms_per_frame                                   # GFLOP / ms = TFLOP/s
    out = {"tensor": {"achieved": tf, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": tf / pk["bf16_tflops_sustained"]}}
    if name in ALG_BYTES:
        gbs = ALG_BYTES[name] / ms_per_frame / 1e6                  # bytes / ms -> GB/s
        out["hbm"] = {"achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                      "algorithmic_bytes": ALG_BYTES[name]}
    bound = BOUND.get(name, "tensor")
    if bound not in out:
        bound = "tensor"
    return {"bound": bound, **out[bound], "other": {k: v for k, v in out.items() if k != bound}, "peak_source": pk_src,
            "definition": "SURVEY.md 8d algorithmic FLOPs (unpadded channels) / block-fused fp32-I/O bytes per frame, divided by the timed ms per frame"}


def algorithmic_gflop(name, model=None, args=None):
    """SURVEY.md §8d figure for the tabulated presets, else the same sum (2*Cout*s^2*Cin*k^2*H*W over the cascade's
    convs, unpadded channels) computed from the model's conv list."""
    if name in ALG_GFLOP:
        return ALG_GFLOP[name]
    from bnerv_b200.engine import DecoderEngine
    if model is None:
        model, args = build_model(name)
    eng = DecoderEngine(model)
    H, W = [int(v) for v in args.fc_hw.split("_")]
    tot = 0.0
    for blk in eng.blocks:
        for slot in ([blk.pre] if blk.pre is not None else []) + [blk.up]:
            tot += 2.0 * slot.cout * slot.s ** 2 * slot.cin * slot.k ** 2 * H * W
            H, W = H * slot.s, W * slot.s
        tot += 2 * (2.0 * blk.cout * blk.cout * 9 * H * W)
    tot += 2.0 * eng.head.cout * eng.head.cin * eng.head.k ** 2 * H * W
    return tot / 1e9


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=250)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="hnerv_l", help="preset name in bnerv_b200.config (default: the metric's workload)")
    ap.add_argument("--batch", type=int, default=1, help="frames per launch (reference scripts use -b 1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the extra training-step measurement (N=1 only)")
    ap.add_argument("--no-others", action="store_true", help="skip the short decode runs of the other BASELINE presets (N=1 only)")
    ap.add_argument("--no-ptq", action="store_true", help="skip the post-training-quantisation extra (N=1 only)")
    ap.add_argument("--cpu-budget-s", type=float, default=20.0, help="wall-clock budget of the cpu_baseline sample")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t_begin, t_end):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t_begin - 0.05 <= t <= t_end + 0.15 and len(r) >= 7] or [r for _, r in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = [float(r[0]) for r in rows]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(r[3 + j].lower().startswith("active") for r in rows)]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(rows[0][1]), "power_w_max": max(float(r[2]) for r in rows),
                "samples": len(rows), "reasons": reasons}


# ------------------------------------------------------------------------------------------------
# model / data
# ------------------------------------------------------------------------------------------------
def build_model(name):
    from bnerv_b200 import ENeRV_Boost, HNeRV_Boost, NeRV_Boost, preset
    args = preset(name)
    torch.manual_seed(1)
    if args.model == "HNeRV_Boost":
        m = HNeRV_Boost(args)
    elif args.model == "ENeRV_Boost":
        m = ENeRV_Boost(3, args)
    else:
        m = NeRV_Boost(1, args)
    return m.eval(), args


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# CPU leg: the oracle port of the reference forward on the host cores (reference arm / cpu_baseline)
# ------------------------------------------------------------------------------------------------
def cpu_decode_fn(model, args, crop):
    """Returns (fn(i) decoding frame i on CPU with the oracle, fraction of a full frame that fn computes).
    The decoder is fully convolutional, so a bounded sample is a spatial crop of the 9x16 stem grid:
    crop=(h,w) decodes h*w/(9*16) of the frame's pixels with identical per-pixel work."""
    from oracle import nerv_oracle as orc
    sd = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
    cfg = orc.cfg_from_args(args)
    fh, fw = cfg["fc_hw"]
    ch, cw = crop
    g = torch.Generator().manual_seed(1234)
    emb = torch.rand(1, 16, fh, fw, generator=g)

    def fn(i):
        t = torch.tensor([(i + 1) / N_FRAMES], dtype=torch.float64)
        with torch.no_grad():
            if args.model == "HNeRV_Boost":
                return orc.hnerv_boost_decode(sd, cfg, emb[:, :, :ch, :cw], t)[0]
            if (ch, cw) != (fh, fw):
                raise RuntimeError("spatial-crop sampling is only defined for the HNeRV decoder")
            return orc.forward(args.model, sd, cfg, t)[0]
    return fn, (ch * cw) / float(fh * fw)


def pick_crop(model, args, n_steps, budget_s):
    """Largest crop of the stem grid such that n_steps steps fit the wall-clock budget (calibrated on a 3x4 probe)."""
    fh, fw = [int(v) for v in args.fc_hw.split("_")]
    if args.model != "HNeRV_Boost":
        return (fh, fw)
    fn, frac = cpu_decode_fn(model, args, (3, 4))
    fn(0)
    t0 = time.perf_counter()
    fn(1)
    per_full = (time.perf_counter() - t0) / frac
    for crop in [(fh, fw), (fh, fw // 2), (fh // 2 + 1, fw // 2), (3, 4), (2, 3)]:
        if per_full * crop[0] * crop[1] / (fh * fw) * n_steps <= budget_s:
            return crop
    return (2, 3)


def run_reference(opt):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    model, args = build_model(opt.config)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    crop = pick_crop(model, args, opt.steps + opt.warmup, 150.0)
    fn, frac = cpu_decode_fn(model, args, crop)
    for i in range(opt.warmup):
        fn(i)
    t0 = time.perf_counter()
    for i in range(opt.steps):
        fn(opt.warmup + i)
    dt = time.perf_counter() - t0
    fps = opt.steps * frac / dt
    fh_, fw_ = [int(v) for v in args.fc_hw.split("_")]
    sample = (f"{opt.steps} steps, each a {crop[0]}x{crop[1]} crop of the {fh_}x{fw_} stem grid = {frac:.3f} of a frame "
              "(fully convolutional decoder), oracle port of the reference forward (torch CPU f32, oneDNN)")
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": opt.gpus, "steps": opt.steps,
            "warmup": opt.warmup, "ms_per_step": dt / opt.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(opt.config, args), "batch": 1, "host_threads": torch.get_num_threads()},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_name(cfg_name, args):
    fh, fw = [int(v) for v in args.fc_hw.split("_")]
    up = 1
    for s_ in args.dec_strds:
        up *= s_
    return (f"{args.model} {cfg_name} (modelsize {args.modelsize}, fc_dim {args.fc_dim}) decode @{fh * up}x{fw * up}, "
            f"{N_FRAMES}-frame synthetic sequence")


# ------------------------------------------------------------------------------------------------
# GPU leg
# ------------------------------------------------------------------------------------------------
def run_b200(opt):
    import torch.distributed as dist
    from bnerv_b200 import _capi, ops
    from bnerv_b200.shard import frame_indices, norm_index

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    model, args = build_model(opt.config)
    model = model.to(dev)
    is_h = args.model == "HNeRV_Boost"
    B, K, W = opt.batch, opt.steps, opt.warmup
    fh, fw = [int(v) for v in args.fc_hw.split("_")]

    # this rank's shard of the frame sequence (round-robin, no collective on the data path)
    mine = frame_indices(N_FRAMES, rank, world)
    need = (K + W) * B
    idx = [mine[j % len(mine)] for j in range(need)]
    g = torch.Generator().manual_seed(1234 + rank)
    emb_host = torch.rand(need, 16, fh, fw, generator=g).pin_memory() if is_h else None
    t_host = torch.tensor([norm_index(i, N_FRAMES) for i in idx], dtype=torch.float64).pin_memory()
    emb_dev = emb_host.to(dev) if is_h else None
    t_dev = t_host.to(dev)

    def step_dev(j):
        sl = slice(j * B, (j + 1) * B)
        return model.decode(emb_dev[sl], t_dev[sl]) if is_h else model.decode(t_dev[sl])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        # ---------------- device-resident throughput (CUDA-graph replay per frame) ----------------
        for j in range(W):
            step_dev(j)
        barrier()
        sampler = ClockSampler(local) if rank == 0 else None
        n0 = _capi.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_begin = time.time()
        e0.record()
        for j in range(K):
            img = step_dev(W + j)
        e1.record()
        barrier()
        t_end = time.time()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop(t_begin, t_end) if sampler else None
        assert torch.isfinite(img).all()

        # ---------------- per-kernel timing: the same K steps launched eagerly with CUDA events around every
        # fused-conv launch on the launching stream (events cannot be recorded inside a replayed graph) ----------------
        eng = model.engine()
        eng.use_graph = False
        step_dev(0)
        torch.cuda.synchronize()
        ops.TIMING = []
        n1 = _capi.launch_count()
        for j in range(K):
            step_dev(W + j)
        torch.cuda.synchronize()
        launches = (_capi.launch_count() - n1) // K * K     # kernels of this library per K steps (same count replayed by the graph)
        timing, ops.TIMING = ops.TIMING, None
        eng.use_graph = True
        # per launch shape: median over the K steps (a host hiccup between the two event records of one eager
        # launch would otherwise show up as a multi-ms "launch"); a step's conv time = sum of the shape medians
        by_shape = {}
        for f, a, b, shp in timing:
            by_shape.setdefault(shp, []).append((a.elapsed_time(b), f))
        per_step = K if K > 0 else 1
        shapes = []
        for shp, rows in by_shape.items():
            med = statistics.median(r[0] for r in rows)
            shapes.append({"shape": shp, "ms": med, "flops": rows[0][1], "per_step": len(rows) / per_step})
        conv_ms = sum(r["ms"] * r["per_step"] for r in shapes) * K
        conv_flops = sum(r["flops"] * r["per_step"] for r in shapes) * K
        top = max(shapes, key=lambda r: r["ms"] * r["per_step"]) if shapes else None

        # ---------------- end to end through the reference-facing API ----------------
        out_host = torch.empty((B, 3, img.shape[-2], img.shape[-1]), dtype=torch.float32).pin_memory()

        def step_e2e(j):
            sl = slice(j * B, (j + 1) * B)
            t = t_host[sl].to(dev, non_blocking=True)
            if is_h:
                o, _, _ = model.forward_decoder(emb_host[sl].to(dev, non_blocking=True), t)
            else:
                o, _, _ = model(t)
            out_host.copy_(o, non_blocking=True)
            torch.cuda.synchronize()

        for j in range(W):
            step_e2e(j)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for j in range(K):
            step_e2e(W + j)
        f1.record()
        barrier()
        ms_percall = f0.elapsed_time(f1)

        # ---------------- end to end, pipelined: bnerv_b200.decode_to_host (the package's host-to-host throughput API):
        # the same per-step H2D of the inputs and D2H of the decoded frame, the read-back of frame i overlapping the
        # decode of frame i+1 on a second stream ----------------
        from bnerv_b200 import decode_to_host
        ring = 16 * B                                  # pinned host ring buffer the frames land in (consumer side not modelled)
        frames_host = torch.empty((ring, 3, img.shape[-2], img.shape[-1]), dtype=torch.float32).pin_memory()
        emb_k = emb_host[W * B:(W + K) * B] if is_h else None
        t_k = t_host[W * B:(W + K) * B]
        decode_to_host(model, t_host[:W * B], frames_host, emb_host[:W * B] if is_h else None, batch=B, ring=True)
        barrier()
        t_w0 = time.perf_counter()
        f0.record()
        decode_to_host(model, t_k, frames_host, emb_k, batch=B, ring=True)
        f1.record()
        barrier()
        wall_e2e = (time.perf_counter() - t_w0) * 1e3
        ms_e2e = max(f0.elapsed_time(f1), wall_e2e)      # the call returns only after the last D2H: wall clock covers the copy stream
        assert torch.isfinite(frames_host).all() and frames_host.max() > 0

    tmax = torch.tensor([ms, ms_e2e, ms_percall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms, ms_e2e, ms_percall = tmax.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk, pk_src = peaks()
    frames = world * K * B
    value = frames / (ms / 1e3)
    alg_gflop = algorithmic_gflop(opt.config, model, args)
    peak_tf = pk["bf16_tflops_sustained"]       # kernel timed inside a long step -> sustained figure
    ach_tf = conv_flops / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else None
    traffic, traffic_note, alg_bytes = None, None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", f"frame_traffic_{opt.config}.json")))
        traffic = (tj["dram_read_bytes"] + tj["dram_write_bytes"]) * B
        traffic_note = "per step (one frame): " + tj["what"]
        alg_bytes = tj.get("algorithmic_bytes_per_frame", None)
        alg_bytes = None if alg_bytes is None else alg_bytes * B
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 operands, f32 accumulate (tcgen05 kind::f16); f16 activations between kernels",
        "data": "synthetic",
        "config": {"workload": workload_name(opt.config, args), "batch": B, "frames_per_step_per_gpu": B,
                   "sharding": f"frames round-robin over {world} rank(s), no data-path collective",
                   "launch": "one CUDA-graph replay per frame (PE, stem MLP, SFT table, the fused-conv / fused-block launches chained by programmatic dependent launch)",
                   "l2": "per-step activation traffic (>=4 GB at 1080p, every map 0.4-0.9 GB) exceeds the 126 MB L2; no explicit flush",
                   "algorithmic_gflop_per_frame": alg_gflop},
        "gpu_launches": launches,
        "clocks": clocks,
        "e2e": {"value": frames / (ms_e2e / 1e3), "unit": "frames/s",
                "h2d_bytes_per_step": B * ((16 * fh * fw * 4 if is_h else 0) + 8), "d2h_bytes_per_step": out_host.numel() * 4,
                "api": "bnerv_b200.decode_to_host(model, norm_idx_host, out_host, embed_host): pinned host inputs -> pinned host frames, "
                       "H2D + CUDA-graph decode + D2H per step, read-back of step i overlapped with the decode of step i+1",
                "ms_per_step": ms_e2e / K,
                "per_call": {"value": frames / (ms_percall / 1e3), "ms_per_step": ms_percall / K,
                             "api": ("model.forward_decoder(img_embed, norm_idx)" if is_h else "model(t)") +
                                    " per frame with the reference's semantics (device sync inside the call, model_nerv.py:58-59) + synchronous D2H"}},
        "roofline": None,
    }
    # step-level roofline: SURVEY.md 8d algorithmic work per frame / the TIMED graph-replay region, per GPU (every rank decodes
    # B frames per step); the bounding resource per preset (tensor cores for configs 3-5, HBM for the 12-channel NeRV presets)
    rl = preset_roofline(opt.config, ms / K / B, pk, pk_src, alg_gflop)
    rl.update({
        "kernel": "the decode step (all launches of the captured graph; the fused-conv / fused-block kernels are > 90 % of it)",
        "conv_only": {"achieved": ach_tf, "unit": "TFLOP/s", "frac": (ach_tf / peak_tf) if ach_tf else None, "conv_ms_per_step": conv_ms / K,
                      "kernel_timing": "separate eager pass of the same K steps, CUDA events around each fused launch (events cannot be "
                                       "recorded inside a replayed graph); host-bound for the small presets"},
        "traffic": traffic, "traffic_note": traffic_note,
        "algorithmic_bytes": None if opt.config not in ALG_BYTES else ALG_BYTES[opt.config] * B,
        "algorithmic_bytes_note": "SURVEY.md 8d: block-fused, fp32 I/O (block in + block out + weights)",
        "launch_fused_bytes": alg_bytes,
        "launch_fused_bytes_note": "what this design's launches must move: C8 f16 input (+ residual) + output map(s) + weights per launch",
        "measured_hbm": None if (traffic is None) else {"achieved": traffic / (ms / K) / 1e6, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                                        "frac": traffic / (ms / K) / 1e6 / pk["hbm_gbs"]},
        "top_launch": None if top is None else {"shape(cin,cout,k,s,H,W,act)": list(top["shape"]), "ms": top["ms"],
                                                 "launches_per_step": top["per_step"], "tflops": top["flops"] / top["ms"] / 1e9},
    })
    line["roofline"] = rl
    if world == 1 and not opt.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        cpu_model = model.cpu()
        crop = pick_crop(cpu_model, args, 3, opt.cpu_budget_s)
        fn, frac = cpu_decode_fn(cpu_model, args, crop)
        fn(0)
        ts = []
        for i in range(2):
            t0 = time.perf_counter()
            fn(1 + i)
            ts.append(time.perf_counter() - t0)
        line["cpu_baseline"] = {"value": frac / statistics.median(ts), "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"median of 2 decodes (1 warm-up) of a {crop[0]}x{crop[1]} crop of the {fh}x{fw} stem grid = {frac:.3f} of a frame, oracle port (torch CPU f32, oneDNN)"}
    if world == 1 and not opt.no_train:
        # extra, outside the metric: one training step (forward + MSE + backward, batch 1) of the same workload on the
        # native backward kernels vs torch autograd with cuDNN TF32 allowed (the reference's default GPU arithmetic)
        try:
            line["train_step"] = train_step_times(opt.config, dev)
        except Exception as ex:  # never let the extra measurement take the metric line down
            line["train_step"] = {"error": f"{type(ex).__name__}: {str(ex)[:200]}"}
    if world == 1 and not opt.no_train:
        # extra: the "library" baseline SURVEY.md §8d asks for - stock PyTorch on this B200 (cuDNN), same module, same frame
        try:
            line["library_baseline"] = torch_decode_fps(opt.config, dev)
        except Exception as ex:
            line["library_baseline"] = {"error": f"{type(ex).__name__}: {str(ex)[:200]}"}
    if world == 1 and not opt.no_others:
        # extra, outside the metric: device-resident decode frames/s of the other BASELINE.json presets (configs 1-2),
        # same timing rules (CUDA-graph replay per frame, CUDA events, 10 warm-up + 200 timed frames)
        others = {}
        for name in ("enerv_m", "nerv_s", "nerv_xs"):
            if name == opt.config:
                continue
            try:
                others[name] = other_preset_fps(name, dev, pk, pk_src)
            except Exception as ex:
                others[name] = {"error": f"{type(ex).__name__}: {str(ex)[:200]}"}
        line["other_presets"] = others
    if world == 1 and not opt.no_ptq:
        # extra, outside the metric (BASELINE.json configs[4] asks for decode FPS + bpp of the quantised model): 8-bit PTQ of
        # the decoder + 6-bit embeddings, Huffman bits, and the decode of the quantised model - all on the native kernels
        try:
            line["ptq"] = ptq_extra(opt.config, dev)
        except Exception as ex:
            line["ptq"] = {"error": f"{type(ex).__name__}: {str(ex)[:200]}"}
        try:
            line["compression_path"] = compression_extra(opt.config, dev)
        except Exception as ex:
            line["compression_path"] = {"error": f"{type(ex).__name__}: {str(ex)[:200]}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _gaussian_bits(code, quant):
    """DiffEntropyModel.cal_global_bitrate, eval branch, restated (lib/entropy_model.py:20-44): the codes' mean / std
    parametrise a Gaussian, bits = sum max(0, -log2(cdf(q + .5) - cdf(q - .5) + 1e-5))."""
    mean, std = code.mean(), code.std().clamp(1e-5, 1e10)
    cdf = lambda v: 0.5 * (1 + torch.erf((v - mean) * std.reciprocal() / math.sqrt(2.0)))
    probs = cdf(quant + 0.5) - cdf(quant - 0.5)
    return torch.clamp(-torch.log(probs + 1e-5) / math.log(2.0), min=0).sum()


def compression_extra(cfg_name, dev, frames=50):
    """BASELINE.json configs[4] (compression path, scripts/compression/hnerv_boost.sh: --quant, `scale` quantisers for weights and
    biases, `scalebeta` for the embedding, 8 bits, DiffEntropyModel): what `cal_params` leaves behind - dequant_w / dequant_b on
    every decoder layer (Scale_T.forward, lib/transform_ops.py:246-250, with the scales of init_data(), :221-237) - decoded on the
    native kernels, and the entropy model's bit estimate for weights + biases + embeddings -> bits per pixel.  The reference's own
    quantiser / entropy-model classes are not on the GPU box; they are exercised against the drop-in in
    tests/test_compression_path_cpu.py - here their arithmetic is restated in torch on random-init weights."""
    from bnerv_b200 import ops
    model, args = build_model(cfg_name)
    model = model.to(dev).eval()
    ref_model, _ = build_model(cfg_name)
    ref_model = ref_model.to(dev).eval()
    ref_model.load_state_dict(model.state_dict())
    is_h = args.model == "HNeRV_Boost"
    fh, fw = [int(v) for v in args.fc_hw.split("_")]
    emb = torch.rand(N_FRAMES, 16, fh, fw, device=dev, generator=torch.Generator(device=dev).manual_seed(7)) if is_h else None
    enc = getattr(model, "encoder", None)
    skip = set() if enc is None else {id(m) for m in enc.modules()}
    bits = torch.zeros((), device=dev)
    n_sym = 0
    with torch.no_grad():
        for m in model.modules():
            if id(m) in skip or not hasattr(m, "dequant_w") or getattr(m, "weight", None) is None:
                continue
            for name in ("weight", "bias"):
                w = getattr(m, name, None)
                if w is None:
                    continue
                scale = (w.max() - w.min()) / 255.0                      # Scale_T.init_data, 8 bits
                code = w / scale
                quant = torch.round(code)
                setattr(m, "dequant_w" if name == "weight" else "dequant_b", quant * scale)
                bits = bits + _gaussian_bits(code, quant)
                n_sym += w.numel()
        deq_emb = None
        if is_h:                                                         # ScaleBeta_T on the embeddings (:264-286)
            beta, scale = emb.min(), (emb.max() - emb.min()) / 255.0
            code = (emb - beta) / scale
            quant = torch.round(code)
            deq_emb = quant * scale + beta
            bits = bits + _gaussian_bits(code, quant)
            n_sym += emb.numel()
    model.engine().invalidate()
    t = torch.tensor([norm_index_(i) for i in range(frames + 5)], dtype=torch.float64, device=dev)
    psnr = []
    with torch.no_grad():
        for i in range(5):
            model.decode(deq_emb[i:i + 1], t[i:i + 1]) if is_h else model.decode(t[i:i + 1])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(5, frames + 5):
            model.decode(deq_emb[i:i + 1], t[i:i + 1]) if is_h else model.decode(t[i:i + 1])
        e1.record()
        torch.cuda.synchronize()
        for i in (5, 6, 7):
            a = (model.decode(deq_emb[i:i + 1], t[i:i + 1]) if is_h else model.decode(t[i:i + 1])).clone()
            b = ref_model.decode(emb[i:i + 1], t[i:i + 1]) if is_h else ref_model.decode(t[i:i + 1])
            psnr.append(float(ops.frame_metrics(a, b)[0, 2]))
    pixels = fh * fw
    for s_ in args.dec_strds:
        pixels *= s_ * s_
    total_bits = float(bits)
    out = {"what": "BASELINE configs[4]: decoder with dequant_w / dequant_b as cal_params() sets them (8-bit `scale` quantisers at their "
                   "init_data() scales, `scalebeta` embeddings), decoded on the native kernels; bits = the Gaussian entropy model's "
                   "estimate (lib/entropy_model.py:20-44 restated); random-init weights",
           "decode_frames_per_s": 1e3 * frames / e0.elapsed_time(e1), "estimated_bits_per_param": total_bits / n_sym,
           "estimated_bpp": total_bits / pixels / N_FRAMES, "psnr_vs_unquantised_db": sum(psnr) / len(psnr)}
    del model, ref_model
    torch.cuda.empty_cache()
    return out


def ptq_extra(cfg_name, dev, frames=50):
    """quant_model (train_nerv_all.py:620-641) + embedding quantisation (:542) + Huffman accounting (:581-610) on the device
    kernels, then decode frames/s and PSNR of the quantised model against the unquantised one on the same frames."""
    from types import SimpleNamespace
    from bnerv_b200 import ops, ptq
    model, args = build_model(cfg_name)
    model = model.to(dev)
    is_h = args.model == "HNeRV_Boost"
    fh, fw = [int(v) for v in args.fc_hw.split("_")]
    emb = torch.rand(N_FRAMES, 16, fh, fw, device=dev, generator=torch.Generator(device=dev).manual_seed(7)) if is_h else None
    torch.cuda.synchronize()
    tc = time.perf_counter()
    models, quant_ckt = ptq.quant_model(model, SimpleNamespace(quant_model_bit=8))      # first call: allocator growth, kernel load
    torch.cuda.synchronize()
    first_ms = (time.perf_counter() - tc) * 1e3
    ptq.huffman_bits(quant_ckt)                               # (warm-up of the histogram kernel too)
    del models, quant_ckt
    sd = model.state_dict()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    quant_ckt, _ = ptq.quant_state_dict(sd, 8)                # the quantisation itself: ONE multi-tensor call (five launches)
    q_emb, deq_emb = ptq.quant_tensor(emb, 6) if is_h else (None, None)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    bits = ptq.huffman_bits(quant_ckt, q_emb)
    t2 = time.perf_counter()
    models, quant_ckt = ptq.quant_model(model, SimpleNamespace(quant_model_bit=8))      # the reference-shaped call
    torch.cuda.synchronize()
    model_ms = (time.perf_counter() - t2) * 1e3
    qmodel = models[1].eval()
    t = torch.tensor([norm_index_(i) for i in range(frames + 5)], dtype=torch.float64, device=dev)
    psnr = []
    with torch.no_grad():
        for i in range(5):
            qmodel.decode(deq_emb[i:i + 1], t[i:i + 1]) if is_h else qmodel.decode(t[i:i + 1])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(5, frames + 5):
            img = qmodel.decode(deq_emb[i:i + 1], t[i:i + 1]) if is_h else qmodel.decode(t[i:i + 1])
        e1.record()
        torch.cuda.synchronize()
        for i in (5, 6, 7):
            a = (qmodel.decode(deq_emb[i:i + 1], t[i:i + 1]) if is_h else qmodel.decode(t[i:i + 1])).clone()
            b = model.decode(emb[i:i + 1], t[i:i + 1]) if is_h else model.decode(t[i:i + 1])
            psnr.append(float(ops.frame_metrics(a, b)[0, 2]))
    pixels = fh * fw
    for s_ in args.dec_strds:
        pixels *= s_ * s_
    out = {"what": "8-bit PTQ of the decoder (+ 6-bit embeddings for HNeRV), Huffman-coded size, decode of the quantised model; random-init "
                   "weights, so bits/param ~ 8 and the PSNR only says how far 8-bit weights move the output",
           "quantise_ms": (t1 - t0) * 1e3, "quantise_note": "quant_state_dict over all decoder tensors (one multi-tensor call) + the "
           "embedding tensor, second call", "quant_model_ms": model_ms, "quant_model_note": "quant_model() as the reference shapes it "
           "(train_nerv_all.py:620-641): the same quantisation plus TWO deepcopy(model) and a load_state_dict, which are host-side "
           "Python and dominate", "quant_model_first_call_ms": first_ms, "huffman_ms": (t2 - t1) * 1e3, "bits_per_param": bits["bits_per_param"],
           "full_bits_per_param": bits["full_bits_per_param"], "total_bpp": bits["total_bits"] / pixels / N_FRAMES,
           "quantised_decode_frames_per_s": 1e3 * frames / e0.elapsed_time(e1),
           "psnr_quantised_vs_unquantised_db": sum(psnr) / len(psnr)}
    del model, models, qmodel
    torch.cuda.empty_cache()
    return out


def norm_index_(i):
    return (i + 1) / N_FRAMES


def other_preset_fps(name, dev, pk, pk_src, steps=200, warm=10):
    """Device-resident decode frames/s at batch 1 (the reference's -b 1) and batch 8 (frames are independent: batching
    amortises the per-launch latency that dominates the narrow presets)."""
    model, args = build_model(name)
    model = model.to(dev)
    out = {"workload": workload_name(name, args)}
    for B in (1, 8):
        n = steps // B + warm
        t = torch.tensor([[(i * B + j + 1) / N_FRAMES % 1.0 + 1.0 / N_FRAMES for j in range(B)] for i in range(n)], dtype=torch.float64, device=dev)
        with torch.no_grad():
            for j in range(warm):
                model.decode(t[j])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for j in range(warm, n):
                img = model.decode(t[j])
            e1.record()
            torch.cuda.synchronize()
        assert torch.isfinite(img).all()
        ms = e0.elapsed_time(e1) / ((n - warm) * B)
        key = "" if B == 1 else f"_batch{B}"
        out["frames_per_s" + key] = 1e3 / ms
        out["ms_per_frame" + key] = ms
        out["algorithmic_tflops" + key] = algorithmic_gflop(name, model, args) / ms
        out["roofline" + key] = preset_roofline(name, ms, pk, pk_src, algorithmic_gflop(name, model, args))
    out["frames_timed"] = steps
    del model
    torch.cuda.empty_cache()
    return out


def torch_decode_fps(cfg_name, dev, steps=5):
    """Decode frames/s of the plain torch forward (model.backend = 'torch': F.conv2d / pixel_shuffle / sin / gelu, i.e. the
    reference's own ops through cuDNN) on this GPU, TF32 convs allowed (PyTorch's default) and strict fp32."""
    model, args = build_model(cfg_name)
    model = model.to(dev).eval()
    model.backend = "torch"
    is_h = args.model == "HNeRV_Boost"
    fh, fw = [int(v) for v in args.fc_hw.split("_")]
    emb = torch.rand(1, 16, fh, fw, device=dev) if is_h else None
    t = torch.tensor([0.5], dtype=torch.float64, device=dev)
    out = {"what": "plain torch forward of the same module on this GPU (cuDNN), batch 1, frames/s"}
    for tf32, key in ((True, "torch_cudnn_tf32"), (False, "torch_fp32")):
        torch.backends.cudnn.allow_tf32 = tf32
        with torch.no_grad():
            for _ in range(2):
                model.forward_decoder(emb, t) if is_h else model(t)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                model.forward_decoder(emb, t) if is_h else model(t)
            e1.record()
            torch.cuda.synchronize()
        out[key] = 1e3 * steps / e0.elapsed_time(e1)
    torch.backends.cudnn.allow_tf32 = True
    del model
    torch.cuda.empty_cache()
    return out


def train_step_times(cfg_name, dev, steps=5):
    out = {"what": "forward + MSE loss + backward, batch 1, CUDA-event timed, ms per step (SURVEY.md §8f rank 1)"}
    for mode, tf32 in (("b200", False), ("torch", True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        model, args = build_model(cfg_name)
        model = model.to(dev).train()
        model.train_backend = mode
        is_h = args.model == "HNeRV_Boost"
        fh, fw = [int(v) for v in args.fc_hw.split("_")]
        emb = torch.rand(1, 16, fh, fw, device=dev) if is_h else None
        t = torch.tensor([0.5], dtype=torch.float64, device=dev)
        target = None

        def step():
            nonlocal target
            model.zero_grad(set_to_none=True)
            img = (model.forward_decoder(emb, t) if is_h else model(t))[0]
            if target is None:
                target = torch.rand_like(img)
            ((img - target) ** 2).mean().backward()

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        out["native_ms" if mode == "b200" else "torch_autograd_cudnn_tf32_ms"] = e0.elapsed_time(e1) / steps
        del model
        torch.cuda.empty_cache()
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = False
    return out


if __name__ == "__main__":
    o = parse()
    if o.impl == "reference":
        run_reference(o)
    else:
        run_b200(o)
