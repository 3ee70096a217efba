import os, sys
sys.path.insert(0, "/root/repo")
import torch
from improver_b200 import engine, pipeline
S, NY, NX = 240, 970, 1042
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
cube = torch.rand((S, NY, NX), generator=g, device=dev)
for name, thr in (("raw", None), ("thr_det", [(0.5, 0.5, 0.5, ">")]), ("thr_fuzzy", [(0.5, 0.35, 0.65, ">")])):
    for env in ("1", "0"):
        os.environ["NBH_SQ_RS_ROWS"] = env
        def step():
            return engine.neighbourhood(cube, 5, method="square", thresholds=thr)
        for _ in range(3): step()
        torch.cuda.synchronize()
        ms = pipeline.time_main_kernel(step, reps=10)
        print(name, "RS_ROWS=" + env, pipeline.last_profiled_kernel()[:26], f"{ms:.4f} ms", flush=True)
