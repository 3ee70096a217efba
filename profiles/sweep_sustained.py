#!/usr/bin/env python
"""Kernel time of the config-3 fused call under SUSTAINED load (power-capped clocks): 2 s of
back-to-back launches, then 50 timed launches, with nvidia-smi clocks sampled during the load.
usage: python profiles/sweep_sustained.py "NBH_SQ_RS_SLEEP_NS=0" "NBH_SQ_RS_SLEEP_NS=100" ..."""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from improver_b200 import engine, pipeline

R, L, NY, NX = 18, 36, 970, 1042
dev = torch.device("cuda", 0)
cube = pipeline.member_padded_empty((R, L, NY, NX), dev)
g = torch.Generator(device=dev).manual_seed(3)
for r in range(R):
    u = torch.rand((2, L, NY, NX), generator=g, device=dev).clamp_min_(1e-12)
    cube[r] = -1.5 * (torch.log(u[0]) + torch.log(u[1]))
    del u
out = torch.empty((L, 1, NY, NX), dtype=torch.float32, device=dev)
thr = [(2.0, 1.4, 2.6, ">")]
def step():
    return engine.neighbourhood(cube, 5, method="square", thresholds=thr, collapse_members=True, out=out)
def smi():
    o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-i", "0"],
                       capture_output=True, text=True).stdout.strip()
    return o
for spec in sys.argv[1:] or [""]:
    keys = []
    for kv in spec.split():
        k, v = kv.split("="); os.environ[k] = v; keys.append(k)
    t_end = time.perf_counter() + 2.0
    while time.perf_counter() < t_end:
        for _ in range(50): step()
        torch.cuda.synchronize()
    for _ in range(200): step()
    state = smi()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100): step()
    e1.record(); torch.cuda.synchronize()
    print(f"{spec or '(default)':50s} call {e0.elapsed_time(e1)/100:.4f} ms   smi[sm MHz, W, power_cap] = {state}", flush=True)
    for k in keys: del os.environ[k]
    time.sleep(3)
