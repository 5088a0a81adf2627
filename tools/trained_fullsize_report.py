"""Where the f16-operand error enters a TRAINED full-size preset: trains it natively for N Adam steps on synthetic frames (as
tests/test_gpu_models.py::test_benchmarked_presets_after_training_against_oracle), then compares every block output and the
image of the native decode with the CPU oracle in f32 and with the oracle's f16-operand emulation (orc.EMULATE).
Usage: python tools/trained_fullsize_report.py [config] [steps] [precise_blocks]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "boosting-nerv_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import bench  # noqa: E402
from conftest import elementwise_rel, max_rel  # noqa: E402
from oracle import nerv_oracle as orc  # noqa: E402  (checker only)

name = sys.argv[1] if len(sys.argv) > 1 else "hnerv_l"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
model, a = bench.build_model(name)
cfg = orc.cfg_from_args(a)
fh, fw = [int(v) for v in a.fc_hw.split("_")]
up = 1
for s_ in a.dec_strds:
    up *= s_
H, W = fh * up, fw * up
is_h = a.model == "HNeRV_Boost"
n = 2
t = torch.tensor([(i + 1) / 600 for i in range(n)], dtype=torch.float64, device="cuda")
emb = torch.rand(n, 16, fh, fw, generator=torch.Generator().manual_seed(9)).cuda() if is_h else None
yy, xx = torch.meshgrid(torch.linspace(0, 1, H, device="cuda"), torch.linspace(0, 1, W, device="cuda"), indexing="ij")
frames = torch.stack([torch.stack([0.5 + 0.45 * torch.sin(6.2832 * ((1 + c) * xx + (2 - 0.5 * c) * yy + 0.13 * (c + 1) * i)) for c in range(3)])
                      for i in range(n)])
model = model.cuda().train()
opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-4)
for it in range(steps):
    i = it % n
    opt.zero_grad(set_to_none=True)
    out = (model.forward_decoder(emb[i:i + 1], t[i:i + 1]) if is_h else model(t[i:i + 1]))[0]
    loss = ((out - frames[i:i + 1]) ** 2).mean()
    loss.backward()
    opt.step()
print(f"{name}: {steps} native Adam steps, final loss {loss.item():.5f}, train_backend {model.train_backend}")
model.eval()
model.keep_intermediates = True
sd = {k: v.detach().float().cpu().clone() for k, v in model.state_dict().items()}
torch.set_num_threads(max(torch.get_num_threads(), 8))
if len(sys.argv) > 3:
    model.engine().set_precise(sys.argv[3])
    print(f"  precise blocks: {sorted(map(str, model.engine().precise))}")
with torch.no_grad():
    img, outs, _ = model.forward_decoder(emb[:1], t[:1]) if is_h else model(t[:1])
    args_ = (sd, cfg, emb[:1].cpu(), t[:1].cpu()) if is_h else (a.model, sd, cfg, t[:1].cpu())
    fn = orc.hnerv_boost_decode if is_h else orc.forward
    ref, ref_outs = fn(*args_)
    orc.EMULATE = torch.float16
    emu, emu_outs = fn(*args_)
    orc.EMULATE = None
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
model.keep_intermediates = False
with torch.no_grad():
    for it in range(25):
        if it == 5:
            e0.record()
        _ = model.forward_decoder(emb[:1], t[:1]) if is_h else model(t[:1])
    e1.record()
torch.cuda.synchronize()
print(f"  decode: {e0.elapsed_time(e1) / 20:.3f} ms per frame in this mode")
skip = 1 if (is_h or a.model == "ENeRV_Boost") else 0
print(f"  image: vs f32 oracle {max_rel(img.cpu(), ref):.2e} (element-wise {elementwise_rel(img.cpu(), ref):.2e}), vs f16 emulation "
      f"{max_rel(img.cpu(), emu):.2e}; emulation vs f32 {max_rel(emu, ref):.2e}; PSNR vs frame ours {orc.psnr(img.cpu(), frames[:1].cpu()):.4f} "
      f"oracle {orc.psnr(ref, frames[:1].cpu()):.4f}")
for i, o in enumerate(outs[skip:]):
    r, e = ref_outs[skip + i], emu_outs[skip + i]
    print(f"  block {i}: {tuple(o.shape[1:])} max|ref| {float(r.abs().max()):.2f}  native vs f32 {max_rel(o.cpu(), r):.2e}  vs emulation "
          f"{max_rel(o.cpu(), e):.2e}  emulation vs f32 {max_rel(e, r):.2e}")
