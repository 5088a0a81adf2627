"""GPU: `python bench.py` prints ONE JSON line with every key of the driver contract (run on the small NeRV-S preset so the
test stays seconds long; the headline workload is the same code path)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_b200_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--config", "nerv_s", "--steps", "6", "--warmup", "3",
                          "--no-train", "--no-others", "--cpu-budget-s", "2"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "gpu_launches", "clocks", "e2e", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["unit"] == "frames/s" and d["n_gpus"] == 1 and d["steps"] == 6 and d["warmup"] == 3 and d["scaling"] == "weak"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic" and d["value"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["gpu_launches"] > 0
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(d["clocks"])
    e = d["e2e"]
    assert e["unit"] == "frames/s" and e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] == 3 * 720 * 1280 * 4
    assert e["value"] != d["value"] and e["per_call"]["value"] > 0
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s") and r["peak"] > 0 and r["achieved"] > 0
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and "traffic" in r
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and c["unit"] == "frames/s" and c["sample"]
    q = d["ptq"]                     # extra: PTQ + Huffman accounting + quantised decode on the native kernels
    assert "error" not in q and 0 < q["bits_per_param"] <= 9 and q["total_bpp"] > 0 and q["quantised_decode_frames_per_s"] > 0
    assert q["psnr_quantised_vs_unquantised_db"] > 25
