"""Profiling driver: un-fused square neighbourhood (240 UK slices), r and mask from argv."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from improver_b200 import engine, workloads as W
r = int(sys.argv[1]) if len(sys.argv) > 1 else 5
masked = len(sys.argv) > 2 and sys.argv[2] == "mask"
g = torch.Generator(device="cuda").manual_seed(1)
d = torch.rand((240, 970, 1042), generator=g, device="cuda")
m = torch.from_numpy((W.land_sea_mask() > 0).astype(np.uint8)).cuda() if masked else None
for _ in range(3):
    engine.neighbourhood(d, r, method="square", mask2d=m)
torch.cuda.synchronize()
print("ok")
