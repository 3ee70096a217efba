"""Debug: plugin-chain (lazy, host source) vs engine on the device, per lead time; circular wrap check."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from improver_b200 import engine, workloads as W
dev = torch.device("cuda", 0)
cube = bench._make_cube(dev, 3)
thr = [(2.0, 1.4, 2.6, ">")]
ref, _, st = engine.neighbourhood(cube, 5, thresholds=thr, collapse_members=True)
engine.check_status(st)
host = torch.empty((18, 36, 970, 1042), dtype=torch.float32, pin_memory=True)
host.copy_(cube); torch.cuda.synchronize()
print("host==cube", bool((host.cuda() == cube).all()))
res = np.asarray(bench._plugin_chain(host))
r = ref.cpu().numpy()
print("shapes", res.shape, r.shape)
err = np.abs(res - r).reshape(36, -1).max(axis=1)
print("per-lead max err", np.round(err, 4))
from improver_b200 import pipeline
out = pipeline.threshold_neighbourhood_mean_host(host, [2.0], [(1.4, 2.6)], ">", 5)
print("pipeline direct err", np.abs(out.numpy() - r).reshape(36, -1).max(axis=1).round(4))
# circular wrap at the global size
d = W.counter_uniform(0, 2, (1920, 2560), 5, dev)
o, _, st = engine.neighbourhood(d, 5, method="circular", wrap_x=True); engine.check_status(st)
o2, _, st = engine.neighbourhood(d, 5, method="circular", wrap_x=False); engine.check_status(st)
print("circ wrap vs nowrap interior equal:", float((o[:, :, 10:-10] - o2[:, :, 10:-10]).abs().max()), "edge differs:", float((o[:, :, :3] - o2[:, :, :3]).abs().max()))
import time
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(5): engine.neighbourhood(d, 5, method="circular", wrap_x=True)
torch.cuda.synchronize(); print("circ 2 slices ms", (time.perf_counter()-t0)/5*1e3)
