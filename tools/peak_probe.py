"""Sustained cuBLAS GEMM throughput of this box for f16 vs bf16 operands (8192^3, back to back for ~3 s each) with
the SM clock / power sampled during the run: the power-capped ceiling the f16 tcgen05 convs compete with.
Usage: python tools/peak_probe.py"""
import subprocess
import threading
import time

import torch


def sample(stop, rows):
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"],
                         stdout=subprocess.PIPE, text=True)
    for line in p.stdout:
        rows.append([float(v) for v in line.split(",")])
        if stop.is_set():
            break
    p.terminate()


for dt in (torch.bfloat16, torch.float16, torch.bfloat16, torch.float16):
    a = torch.randn(8192, 8192, device="cuda", dtype=dt)
    b = torch.randn(8192, 8192, device="cuda", dtype=dt)
    for _ in range(5):
        a @ b
    torch.cuda.synchronize()
    rows, stop = [], threading.Event()
    th = threading.Thread(target=sample, args=(stop, rows), daemon=True)
    th.start()
    n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    while time.time() - t0 < 3.0:
        for _ in range(20):
            a @ b
        n += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    stop.set()
    ms = e0.elapsed_time(e1)
    tail = rows[len(rows) // 2:] or [[0, 0]]
    print(f"{str(dt):16s} {2 * 8192 ** 3 * n / ms / 1e9:8.1f} TFLOP/s sustained over {ms / 1e3:.1f} s; "
          f"SM clock median (2nd half) {sorted(r[0] for r in tail)[len(tail) // 2]:.0f} MHz, power max {max(r[1] for r in tail):.0f} W", flush=True)
