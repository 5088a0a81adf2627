"""CPU: caller-level drop-in proof (SURVEY.md §4 test-pyramid item 3).  The UNMODIFIED reference caller
`train_nerv_all.py main()` (--eval_only --eval_fps; :27-148, :220-231, :451-519) is run twice on the same synthetic PNG frames -
once importing the reference's own model_*.py, once with `boosting-nerv_b200/` first on sys.path - through
tools/run_reference_caller.py, and the CSV the caller itself writes (Dump2CSV, :434-448) is compared: every PSNR / MS-SSIM
column within 0.01 dB / 1e-4, bits per pixel identical (8-bit PTQ + Huffman accounting of evaluate()), and the FPS column
populated from the `dec_time` the model returns (:518-519).  Needs /root/reference (the caller is the reference's code):
skipped where it is absent.  On CPU the drop-in runs its plain-torch wiring; on a CUDA box with the reference present the same
tool decodes natively (python tools/run_reference_caller.py dropin WORKDIR)."""
import os
import subprocess
import sys

import pandas as pd
import pytest

from conftest import ROOT

REF = "/root/reference"
TOOL = os.path.join(ROOT, "tools", "run_reference_caller.py")


def _run(which, workdir, model, extra=()):
    res = subprocess.run([sys.executable, TOOL, which, workdir, "--model", model, "--frames", "3", *extra], capture_output=True, text=True,
                         timeout=900, cwd=workdir)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("CSV:")][-1]
    src = [l for l in res.stdout.splitlines() if l.startswith("model modules from:")][-1]
    return pd.read_csv(line.split("CSV:", 1)[1].strip()), src


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference tree (train_nerv_all.py is the reference's caller)")
@pytest.mark.parametrize("model,extra", [("NeRV_Boost", ()), ("HNeRV_Boost", ("--fps_off",))])
def test_unmodified_caller_gives_the_same_csv_with_the_drop_in(tmp_path, model, extra):
    wd_ref, wd_new = str(tmp_path / "ref"), str(tmp_path / "new")
    os.makedirs(wd_ref), os.makedirs(wd_new)
    ref, src_ref = _run("reference", wd_ref, model, extra)
    new, src_new = _run("dropin", wd_new, model, extra)
    assert src_ref.endswith("/root/reference") and src_new.endswith("boosting-nerv_b200")
    cols = [c for c in ref.columns if "psnr" in c or "ssim" in c]
    assert len(cols) == 16
    for c in cols:
        tol = 0.01 if "psnr" in c else 1e-4
        assert abs(float(ref[c][0]) - float(new[c][0])) <= tol, (c, ref[c][0], new[c][0])
    assert float(ref["pred_seen_psnr"][0]) > 5.0                                   # the metric columns are populated
    assert float(new["FPS"][0]) > 0 and float(ref["FPS"][0]) > 0                    # from the returned dec_time
    for c in ("bits/pixel", "bits/param", "Size (M)"):
        assert str(ref[c][0]) == str(new[c][0]), (c, ref[c][0], new[c][0])
