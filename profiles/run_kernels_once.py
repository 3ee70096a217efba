#!/usr/bin/env python
"""One launch of every secondary kernel on UK-sized inputs, for an `ncu --set full` capture:
un-fused square (rows mode of square_rs_kernel), circular r=9 with a land mask (prefix + gather),
percentiles r=10, un-fused thresholds (T=4) and threshold-in-vicinity (r=2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from improver_b200 import engine

NY, NX = 970, 1042
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(5)
prob = torch.rand((48, NY, NX), generator=g, device=dev)
mask = (torch.rand((NY, NX), generator=g, device=dev) < 0.6).to(torch.uint8)
rain = -1.5 * (torch.log(torch.rand((6, 8, NY, NX), generator=g, device=dev).clamp_min_(1e-12)) +
               torch.log(torch.rand((6, 8, NY, NX), generator=g, device=dev).clamp_min_(1e-12)))
thr = [(t, 0.7 * t, 1.3 * t, ">") for t in (0.5, 1.0, 2.0, 4.0)]
_, _, st = engine.neighbourhood(prob, 5, method="square"); engine.check_status(st)
_, _, st = engine.neighbourhood(prob, 9, method="circular", mask2d=mask); engine.check_status(st)
_, st = engine.neighbourhood_percentiles(prob[:2], 10, [5, 25, 50, 75, 95]); engine.check_status(st)
_, _, st = engine.threshold(rain, thr); engine.check_status(st)
_, _, st = engine.threshold_vicinity(rain, thr[:2], [2], landmask=mask, collapse_leading=True); engine.check_status(st)
torch.cuda.synchronize()
print("done")
