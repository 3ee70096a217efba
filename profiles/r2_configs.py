#!/usr/bin/env python
"""Round-2 side benchmark: config 2 (circular, land + sea in one launch vs two passes),
un-fused square, percentiles.  One JSON line per case.  usage: python profiles/r2_configs.py [cases...]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from improver_b200 import engine, workloads as W  # noqa: E402

NY, NX = W.UK


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
    want = set(sys.argv[1:])
    on = lambda k: not want or k in want
    g = torch.Generator(device="cuda").manual_seed(1)
    n = 240
    d2 = torch.rand((n, NY, NX), generator=g, device="cuda")
    land = torch.from_numpy((W.land_sea_mask() > 0).astype(np.uint8)).cuda()
    sea = 1 - land
    if on("circ"):
        for r in (5, 9, 27, 45, 81):
            nn = n if r <= 27 else 48
            dd = d2[:nn]
            ms = timed(lambda: engine.neighbourhood_land_sea(dd, r, land, method="circular"), reps=3, warm=1)
            report(f"config2_circular_r{r}_land_sea_one_launch", ms, nn * NY * NX, 8 * nn * NY * NX, slices=nn)
            ms = timed(lambda: (engine.neighbourhood(dd, r, method="circular", mask2d=land, want_mask=True),
                                engine.neighbourhood(dd, r, method="circular", mask2d=sea, want_mask=True)),
                       reps=2, warm=1)
            report(f"config2_circular_r{r}_two_passes", ms, nn * NY * NX, 8 * nn * NY * NX, slices=nn)
            ms = timed(lambda: engine.neighbourhood(dd, r, method="circular"), reps=3, warm=1)
            report(f"circular_r{r}_nomask", ms, nn * NY * NX, 8 * nn * NY * NX, slices=nn)
    if on("square"):
        for r in (5, 27):
            for m, nm in ((None, "nomask"), (land, "mask")):
                for ex in (False, True):
                    ms = timed(lambda: engine.neighbourhood(d2, r, method="square", mask2d=m, exact=ex))
                    report(f"square_unfused_r{r}_{nm}{'_exact' if ex else ''}", ms, n * NY * NX, 8 * n * NY * NX, slices=n)
            ms = timed(lambda: engine.neighbourhood_land_sea(d2, r, land, method="square"))
            report(f"square_unfused_r{r}_land_sea", ms, n * NY * NX, 8 * n * NY * NX, slices=n)
        thr = [(0.5, 0.35, 0.65, ">")]
        ms = timed(lambda: engine.neighbourhood(d2, 5, method="square", thresholds=thr))
        report("square_unfused_r5_thresholded", ms, n * NY * NX, 8 * n * NY * NX, slices=n)
    if on("pct"):
        d4 = torch.rand((18, NY, NX), generator=g, device="cuda")
        ms = timed(lambda: engine.neighbourhood_percentiles(d4, 10, (5, 25, 50, 75, 95)), reps=2, warm=1)
        report("config4_percentiles_r10", ms, 18 * NY * NX, (4 + 20) * 18 * NY * NX, slices=18)


if __name__ == "__main__":
    main()
