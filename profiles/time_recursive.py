"""Time the recursive filter on 240 UK slices (970 x 1042, edge_width 15, 1 iteration)
and report algorithmic GB/s: per iteration every padded point is read and written once by each
of the four directional sweeps (32 B; the shared coefficient planes stay in L2)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from improver_b200 import engine  # noqa: E402

S = int(os.environ.get("RF_SLICES", 240))
IT = int(os.environ.get("RF_ITER", 1))
ny, nx, ew = 970, 1042, 15
g = torch.Generator(device="cuda").manual_seed(1)
data = torch.rand((S, ny, nx), device="cuda", generator=g)
cx = torch.rand((ny, nx - 1), device="cuda", generator=g) * 0.5
cy = torch.rand((ny - 1, nx), device="cuda", generator=g) * 0.5
for label, kw in [("plain", {}), ("mask_zeros", {"mask_zeros": True})]:
    d = data.clone()
    if kw:
        d[d < 0.3] = 0.0
    for _ in range(3):
        out = engine.recursive_filter(d, cx, cy, IT, ew, **kw)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        out = engine.recursive_filter(d, cx, cy, IT, ew, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    NY, NX = ny + 4 * ew, nx + 4 * ew
    alg = S * NY * NX * 32.0 * IT
    print(f"{label}: {ms:.3f} ms per call, {S * ny * nx / ms / 1e6:.1f} G points/s, "
          f"{alg / ms / 1e6:.0f} GB/s algorithmic", flush=True)
# CPU reference timing of the same slice shape (one slice, numpy row loops)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import recursive_filter_oracle as rfo  # noqa: E402

h = data[0].cpu().numpy()
t0 = time.perf_counter()
rfo.filter_slice(h, cx.cpu().numpy(), cy.cpu().numpy(), IT, ew)
print(f"cpu oracle: {time.perf_counter() - t0:.3f} s per slice")
