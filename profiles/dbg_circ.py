"""Debug: circular wrap_x timing / correctness on many global slices."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from improver_b200 import engine, workloads as W
dev = torch.device("cuda", 0)
d = W.counter_uniform(0, 96, (1920, 2560), 5, dev)
out = torch.empty_like(d)
for method in ("circular", "square"):
    for wrap in (True, False):
        fn = lambda: engine.neighbourhood(d, 5, method=method, wrap_x=wrap, out=out)
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): _, _, st = fn()
        e1.record(); torch.cuda.synchronize(); engine.check_status(st)
        ms = e0.elapsed_time(e1) / 5
        one, _, st = engine.neighbourhood(d[95:96].contiguous(), 5, method=method, wrap_x=wrap)
        print(method, "wrap" if wrap else "nowrap", "96 slices ms", round(ms, 4), "GB/s algo", round(8 * d.numel() / ms / 1e6, 1),
              "last slice equal to single launch:", bool(torch.equal(out[95], one[0])), flush=True)
