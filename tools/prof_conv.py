"""Launch a few HNeRV-L conv shapes back to back (for ncu).  Usage: python tools/prof_conv.py [case ...]"""
import sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gpu_probe as gp
names = sys.argv[1:] or ["L_dec8_c1", "L_dec8_c0"]
for n in names:
    kw = dict(gp.CASES[n]); kw["time_it"] = False
    for _ in range(3):
        gp.run_case(n, **kw)
torch.cuda.synchronize()
