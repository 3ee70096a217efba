import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from improver_b200 import engine, workloads as W
dev = torch.device("cuda", 0)
for shape, n in (((1920, 2560), 20), ((970, 1042), 70)):
    d = W.counter_uniform(0, n, shape, 5, dev)
    for r in (5, 27):
        out, _, st = engine.neighbourhood(d, r, method="circular"); engine.check_status(st)
        bad = []
        for k in (0, 5, 6, 7, 11, 12, n - 1):
            one, _, st = engine.neighbourhood(d[k:k + 1].contiguous(), r, method="circular")
            diff = (out[k] - one[0]).abs()
            bad.append((k, float(diff.max()), int((diff > 0).sum())))
        print(shape, "r", r, bad, flush=True)
