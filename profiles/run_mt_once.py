"""Profiling driver: fused multi-threshold kernel, config 3 with T thresholds (argv[1], default 4)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from improver_b200 import engine
T = int(sys.argv[1]) if len(sys.argv) > 1 else 4
R, L, NY, NX = 18, 36, 970, 1042
g = torch.Generator(device="cuda").manual_seed(3)
cube = torch.empty((R, L, NY, NX), device="cuda")
for r in range(R):
    u = torch.rand((2, L, NY, NX), generator=g, device="cuda").clamp_min_(1e-12)
    cube[r] = -1.5 * (torch.log(u[0]) + torch.log(u[1]))
    del u
vals = [0.5, 1.0, 2.0, 4.0, 6.0, 8.0, 10.0, 12.0][:T]
thr = [(v, 0.7 * v, 1.3 * v, ">") for v in vals]
for _ in range(3):
    engine.neighbourhood(cube, 5, thresholds=thr, collapse_members=True)
torch.cuda.synchronize()
print("ok")
