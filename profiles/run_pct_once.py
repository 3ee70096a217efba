"""Profiling driver: config 4 (percentiles of a circular neighbourhood r=10 over 18 UK slices)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from improver_b200 import engine
g = torch.Generator(device="cuda").manual_seed(4)
d = torch.rand((18, 970, 1042), generator=g, device="cuda")
for _ in range(2):
    out, st = engine.neighbourhood_percentiles(d, 10, (5, 25, 50, 75, 95))
torch.cuda.synchronize()
print("ok")
