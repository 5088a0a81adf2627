import os, sys, torch
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "boosting-nerv_b200"))
import bench
cfg = sys.argv[1]
model, args = bench.build_model(cfg)
model = model.cuda()
t = torch.tensor([0.5], dtype=torch.float64, device="cuda")
with torch.no_grad():
    for _ in range(3):
        model.decode(t)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    model.decode(t)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
