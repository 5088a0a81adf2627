"""CPU: the compression path (BASELINE.json config 5) through the drop-in modules ON the reference's own lib/ quantisers and
entropy model - state_dict keys, init_data / cal_params / forward_encoder / forward_embed_quant / forward_decoder /
get_bitrate_sum against a golden minted from the unmodified reference (tests/golden/make_golden_quant.py).  Needs
/root/reference (the quantisers are the reference's code, not this repo's): skipped where it is absent (GPU box)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT

REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference tree for lib/quant_ops.py, lib/transform_ops.py, lib/entropy_model.py")
def test_drop_in_reproduces_the_reference_compression_call_sequence():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "compression_path_check.py"), "cpu"], capture_output=True, text=True,
                         timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "ok" in res.stdout and "quantised layers" in res.stdout


def test_quant_golden_is_self_consistent():
    """The committed golden itself (no reference needed): Scale_T semantics dequant = round(w / scale) * scale
    (lib/transform_ops.py:239-251) hold for every layer, the embedding codes are integers in the 8-bit range."""
    z = np.load(os.path.join(GOLDEN, "hnerv_tiny_quant.npz"))
    n = 0
    for k in z.files:
        if k.startswith("dq/") and k.endswith(".weight"):
            name = k[3:-len(".weight")]
            w, scale = torch.from_numpy(z["sd/" + name + ".weight"]), torch.from_numpy(z["sd/" + name + ".weight_quantizer.scale"])
            assert torch.equal(torch.round(w / scale) * scale, torch.from_numpy(z[k])), name
            n += 1
    assert n > 20
    q = z["quant_e"]
    assert np.array_equal(q, np.round(q)) and q.min() >= 0 and q.max() <= 255
    deq = z["quant_e"] * z["sd/embed_quantizer.scale"] + z["sd/embed_quantizer.beta"]
    assert np.abs(deq - z["deq_e"]).max() < 1e-6
