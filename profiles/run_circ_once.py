"""Profiling driver: config-2 circular land+sea launch (240 UK slices), radius from argv."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from improver_b200 import engine, workloads as W
r = int(sys.argv[1]) if len(sys.argv) > 1 else 9
n = int(sys.argv[2]) if len(sys.argv) > 2 else 240
g = torch.Generator(device="cuda").manual_seed(1)
d = torch.rand((n, 970, 1042), generator=g, device="cuda")
land = torch.from_numpy((W.land_sea_mask() > 0).astype(np.uint8)).cuda()
for _ in range(2):
    engine.neighbourhood_land_sea(d, r, land, method="circular")
torch.cuda.synchronize()
print("ok")
