"""Profiling driver: un-fused square (240 UK slices, r=5) and un-fused threshold (T=4)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from improver_b200 import engine
g = torch.Generator(device="cuda").manual_seed(1)
d = torch.rand((120, 970, 1042), generator=g, device="cuda")
thr = [(t, 0.7 * t, 1.3 * t, ">") for t in (0.5, 1.0, 2.0, 4.0)]
for _ in range(3):
    engine.neighbourhood(d, 5, method="square")
    engine.threshold(d, thr)
torch.cuda.synchronize()
print("ok")
