#!/usr/bin/env python
"""Side benchmark of the other BASELINE.json configs (not the driver's bench.py):
config 1 (latency), config 2 (circular, land/sea masks, lead-time radii),
config 4 (percentiles), un-fused square, un-fused threshold.  Prints one JSON
line per case: device time, gridpoint-evals/s, algorithmic GB/s."""
import json
import sys
import os
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from improver_b200 import engine  # noqa: E402

NY, NX = 970, 1042


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def report(name, ms, evals, algo_bytes, **extra):
    print(json.dumps(dict(case=name, ms=round(ms, 4), evals_per_s=evals / ms * 1e3,
                          algo_GBps=algo_bytes / ms / 1e6, **extra)), flush=True)


def main():
    quick = "--quick" in sys.argv
    g = torch.Generator(device="cuda").manual_seed(1)
    # config 1: 200x200, r=5, one slice (launch-bound)
    d1 = torch.rand((1, 200, 200), generator=g, device="cuda")
    ms = timed(lambda: engine.neighbourhood(d1, 5, method="square"), reps=50)
    report("config1_square_200x200_r5", ms, 200 * 200, 8 * 200 * 200)

    # un-fused square on probability cube, 240 slices, r=5 and r=27, mask or not
    n = 60 if quick else 240
    d2 = torch.rand((n, NY, NX), generator=g, device="cuda")
    mask = (torch.rand((NY, NX), generator=g, device="cuda") < 0.6).to(torch.uint8)
    for r in (5, 27):
        for m in (None, mask):
            ms = timed(lambda: engine.neighbourhood(d2, r, method="square", mask2d=m, want_mask=m is not None))
            report(f"square_unfused_r{r}_{'mask' if m is not None else 'nomask'}", ms, n * NY * NX,
                   8 * n * NY * NX, slices=n)
    # config 2: circular, lead-time radii 18/54/90/162 km -> r = 9, 27, 45, 81 ; land + sea passes
    for r in (5, 9, 27) if quick else (5, 9, 27, 45, 81):
        nn = n if r <= 27 else 24
        dd = d2[:nn]
        ms = timed(lambda: (engine.neighbourhood(dd, r, method="circular", mask2d=mask, want_mask=True),
                            engine.neighbourhood(dd, r, method="circular", mask2d=1 - mask, want_mask=True)),
                   reps=2, warm=1)
        report(f"config2_circular_r{r}_land+sea", ms, 2 * nn * NY * NX, 2 * 8 * nn * NY * NX, slices=nn)
    # config 4: percentiles r=10, 18 realizations
    d4 = torch.rand((6 if quick else 18, NY, NX), generator=g, device="cuda")
    ms = timed(lambda: engine.neighbourhood_percentiles(d4, 10, (5, 25, 50, 75, 95)), reps=2, warm=1)
    report("config4_percentiles_r10", ms, d4.shape[0] * NY * NX, (4 + 20) * d4.shape[0] * NY * NX,
           slices=int(d4.shape[0]))
    # un-fused threshold, T=4
    thr = [(t, 0.7 * t, 1.3 * t, ">") for t in (0.5, 1.0, 2.0, 4.0)]
    ms = timed(lambda: engine.threshold(d2, thr))
    report("threshold_unfused_T4", ms, 4 * n * NY * NX, (4 + 16) * n * NY * NX)


if __name__ == "__main__":
    main()
