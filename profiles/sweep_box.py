#!/usr/bin/env python
"""Sweep of the box-kernel experiment knobs (env) on 240 UK slices.
usage: python profiles/sweep_box.py "NBH_BOX_WAIT_NS=0" "NBH_BOX_WAIT_NS=0 NBH_BOX_PER_SM=1" ..."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.environ["NBH_ROOT"])
from improver_b200 import engine, workloads as W
def timed(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
g = torch.Generator(device="cuda").manual_seed(1)
d = torch.rand((240, 970, 1042), generator=g, device="cuda")
m = torch.from_numpy((W.land_sea_mask() > 0).astype(np.uint8)).cuda()
thr = [(0.5, 0.35, 0.65, ">")]
res = {}
for r in (int(x) for x in os.environ.get("SWEEP_R", "5").split(",")):
    res[f"r{r}"] = round(timed(lambda: engine.neighbourhood(d, r)), 4)
    res[f"r{r}_mask"] = round(timed(lambda: engine.neighbourhood(d, r, mask2d=m)), 4)
    res[f"r{r}_thr"] = round(timed(lambda: engine.neighbourhood(d, r, thresholds=thr)), 4)
print(json.dumps(res))
'''
for spec in sys.argv[1:] or [""]:
    env = dict(os.environ, NBH_ROOT=ROOT)
    for kv in spec.split():
        k, v = kv.split("=")
        env[k] = v
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:]
    print(json.dumps({"env": spec, "ms": line}), flush=True)
