#!/usr/bin/env python
"""Time the dominant kernel of the config-3 fused call under several env settings.
usage: python profiles/sweep_env.py "NBH_SQ_RS_MS=6 NBH_SQ_RS_STAGES=6" "NBH_SQ_RS=0" ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from improver_b200 import engine, pipeline

R, L, NY, NX = 18, 36, 970, 1042
if os.environ.get("SWEEP_SHAPE"):
    R, L, NY, NX = (int(v) for v in os.environ["SWEEP_SHAPE"].split(","))
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(3)
cube = torch.empty((R, L, NY, NX), dtype=torch.float32, device=dev)
for r in range(R):
    u = torch.rand((2, L, NY, NX), generator=g, device=dev).clamp_min_(1e-12)
    cube[r] = -1.5 * (torch.log(u[0]) + torch.log(u[1]))
T = int(os.environ.get("SWEEP_T", "1"))
out = torch.empty((L, T, NY, NX), dtype=torch.float32, device=dev)
thr = [(v, 0.7 * v, 1.3 * v, ">") for v in (2.0, 1.0, 4.0, 0.5, 6.0, 8.0, 3.0, 5.0)[:T]]
cells = int(os.environ.get("SWEEP_CELLS", "5"))
ref = None
for spec in sys.argv[1:] or [""]:
    keys = []
    for kv in spec.split():
        k, v = kv.split("=")
        os.environ[k] = v
        keys.append(k)
    def step():
        return engine.neighbourhood(cube, cells, method="square", thresholds=thr, collapse_members=True, out=out)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    ms = pipeline.time_main_kernel(step, reps=10)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step()
    e1.record(); torch.cuda.synchronize()
    if ref is None:
        ref = out.clone()
    err = float((out - ref).abs().max())
    print(f"{spec or '(default)':60s} kernel {ms:.4f} ms  call {e0.elapsed_time(e1)/10:.4f} ms  maxdiff-vs-first {err:.2e}", flush=True)
    for k in keys:
        del os.environ[k]
