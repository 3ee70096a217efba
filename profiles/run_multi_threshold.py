"""Config-3 variant with T thresholds (fused threshold -> square r=5 -> mean over 18 members):
time per pass with the multi-threshold kernel (NBH_SQ_MT=1, default) and without (one pass per threshold)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from improver_b200 import engine

R, L, NY, NX = 18, 36, 970, 1042
g = torch.Generator(device="cuda").manual_seed(3)
cube = torch.empty((R, L, NY, NX), device="cuda")
for r in range(R):
    u = torch.rand((2, L, NY, NX), generator=g, device="cuda").clamp_min_(1e-12)
    cube[r] = -1.5 * (torch.log(u[0]) + torch.log(u[1]))
    del u
for T in (1, 2, 4, 8):
    vals = [0.5, 1.0, 2.0, 4.0, 6.0, 8.0, 10.0, 12.0][:T]
    thr = [(v, 0.7 * v, 1.3 * v, ">") for v in vals]
    for _ in range(3):
        engine.neighbourhood(cube, 5, thresholds=thr, collapse_members=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        _, _, st = engine.neighbourhood(cube, 5, thresholds=thr, collapse_members=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    algo = 4 * NY * NX * L * (R + T)
    print(json.dumps(dict(T=T, mt=os.environ.get("NBH_SQ_MT", "1"), ms=round(ms, 4), evals_per_s=R * T * L * NY * NX / ms * 1e3,
                          algo_GBps=algo / ms / 1e6)), flush=True)
