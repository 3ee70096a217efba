#!/usr/bin/env python
"""Kernel time of the config-3 fused call as a function of the padding of the member stride
(device layout of the cube: members at (L*Y*X + pad) floats)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from improver_b200 import engine, pipeline

R, L, NY, NX = (int(v) for v in os.environ.get("SWEEP_SHAPE", "18,36,970,1042").split(","))
dev = torch.device("cuda", 0)
thr = [(2.0, 1.4, 2.6, ">")]
out = torch.empty((L, 1, NY, NX), dtype=torch.float32, device=dev)
block = L * NY * NX
for pad in [int(a) for a in sys.argv[1:]] or [0]:
    flat = torch.empty(R * (block + pad) + 64, dtype=torch.float32, device=dev)
    cube = flat.as_strided((R, L, NY, NX), (block + pad, NY * NX, NX, 1))
    g = torch.Generator(device=dev).manual_seed(3)
    for r in range(R):
        u = torch.rand((2, L, NY, NX), generator=g, device=dev).clamp_min_(1e-12)
        cube[r] = -1.5 * (torch.log(u[0]) + torch.log(u[1]))
        del u
    def step():
        return engine.neighbourhood(cube, 5, method="square", thresholds=thr, collapse_members=True, out=out)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    ms = pipeline.time_main_kernel(step, reps=10)
    print(f"pad {pad:8d} floats ({pad*4:9d} B)  kernel {ms:.4f} ms  {pipeline.last_profiled_kernel()[:24]}", flush=True)
    del cube, flat
