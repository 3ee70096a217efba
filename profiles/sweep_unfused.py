#!/usr/bin/env python
"""Time the un-fused square neighbourhood (single-member slices) under env settings.
usage: SWEEP_SHAPE=240,970,1042 python profiles/sweep_unfused.py "NBH_SQ_TILE=1" "NBH_SQ_RS_ROWS=0" ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from improver_b200 import engine, pipeline

S, NY, NX = (int(v) for v in os.environ.get("SWEEP_SHAPE", "240,970,1042").split(","))
cells = int(os.environ.get("SWEEP_CELLS", "5"))
wrap = bool(int(os.environ.get("SWEEP_WRAP", "0")))
use_mask = bool(int(os.environ.get("SWEEP_MASK", "0")))
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
cube = torch.rand((S, NY, NX), generator=g, device=dev)
mask = (torch.rand((NY, NX), generator=g, device=dev) < 0.6).to(torch.uint8) if use_mask else None
out = torch.empty((S, NY, NX), dtype=torch.float32, device=dev)
ref = None
for spec in sys.argv[1:] or [""]:
    keys = []
    for kv in spec.split():
        k, v = kv.split("=")
        os.environ[k] = v
        keys.append(k)
    def step():
        return engine.neighbourhood(cube, cells, method="square", mask2d=mask, wrap_x=wrap, out=out)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    ms = pipeline.time_main_kernel(step, reps=10)
    name = pipeline.last_profiled_kernel()
    if ref is None:
        ref = out.clone()
    err = float((out - ref).abs().nan_to_num().max())
    evals = S * NY * NX
    print(f"{spec or '(default)':40s} {name[:28]:28s} kernel {ms:.4f} ms  {evals/ms/1e6:.1f} G evals/s  {8*evals/ms/1e6:.0f} GB/s algorithmic  maxdiff {err:.1e}", flush=True)
    for k in keys:
        del os.environ[k]
