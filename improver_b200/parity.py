"""Multi-rank self-checks: slice-sharded and y-band-tiled results of a fixed seeded cube against
the single-rank result, bit for bit.  Used by `bench.py --gpus N` (the `parity` object of its
JSON line) and by tests/test_gpu_multi.py; needs an initialised NCCL process group."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import distributed as D
from . import engine, workloads


def _max_abs(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.nan_to_num(-1.0)
    b = b.nan_to_num(-1.0)
    return float((a - b).abs().max()) if a.numel() else 0.0


def quirk_field(n: int, ny: int, nx: int, device) -> torch.Tensor:
    """All-ones slices with a few zeros: the non-one bounding box is smaller than the domain, so
    the reference trims with ones padding (nbhood.py:353-409) and window cells outside the
    domain count as ones next to the edges the box touches.  Slice 0: zeros near the top edge
    and in the lowest band (the box spans bands); slice 1: a single zero on the bottom edge."""
    f = torch.ones((n, ny, nx), dtype=torch.float32, device=device)
    f[0, 0, nx // 3] = 0.0
    f[0, ny - 2, 5] = 0.0
    if n > 1:
        f[1, ny - 1, nx - 1] = 0.0
    for k in range(2, n):
        f[k, (k * 37) % ny, (k * 91) % nx] = 0.25
    return f


def multi_rank_parity(device, ny: int = 192, nx: int = 256, n_slices: int = 6, seed: int = 11) -> dict:
    """Returns max-abs differences (0.0 = bit-identical) between the single-rank result computed
    redundantly on this rank and (a) the slice-sharded + gathered result, (b) the y-band tiled
    result of this rank's band, for square (plain, masked, periodic x), circular, the
    ones-trimming quirk case and sharded percentiles.  Max over ranks."""
    world, rank = dist.get_world_size(), dist.get_rank()
    n_slices = max(n_slices, world + 1)                   # every rank a share, the shares uneven
    full = workloads.counter_uniform(0, n_slices, (ny, nx), seed, device)
    g = torch.Generator(device="cpu").manual_seed(seed)
    mask = (torch.rand((ny, nx), generator=g) < 0.8).to(torch.uint8).to(device)
    res = {"sharded_max_abs": 0.0, "tiled_max_abs": 0.0, "tiled_quirk_case": 0.0, "tiled_quirk_fixed_slices": 0,
           "percentiles_sharded_max_abs": 0.0, "ranks": world, "grid": [n_slices, ny, nx]}
    y0, y1 = D.band_range(ny, world, rank)
    cases = [("square", 4, False, None), ("square", 4, False, mask), ("square", 3, True, None),
             ("circular", 5, False, None), ("circular", 5, False, mask)]
    for method, cells, wrap, m in cases:
        ref, _, st = engine.neighbourhood(full, cells, method=method, mask2d=m, wrap_x=wrap,
                                          exact=(method == "square"))
        engine.check_status(st)
        out, _, st = D.neighbourhood_y_tiled(full[:, y0:y1].contiguous(), cells, ny, y0,
                                             mask_band=m[y0:y1].contiguous() if m is not None else None,
                                             method=method, wrap_x=wrap)
        engine.check_status(st)
        res["tiled_max_abs"] = max(res["tiled_max_abs"], _max_abs(out, ref[:, y0:y1]))
        ref_d, _, st = engine.neighbourhood(full, cells, method=method, mask2d=m, wrap_x=wrap)
        engine.check_status(st)
        got, _, st, _ = D.neighbourhood_sharded(full, cells, gather=True, method=method, mask2d=m, wrap_x=wrap)
        engine.check_status(st)
        res["sharded_max_abs"] = max(res["sharded_max_abs"], _max_abs(got, ref_d))
    # the ones-trimming quirk: global bounding boxes through the 10-scalar all-reduce
    q = quirk_field(3, ny, nx, device)
    ref, _, st = engine.neighbourhood(q, 4, method="square", exact=True)
    fixed = engine.check_status(st)
    out, _, st = D.neighbourhood_y_tiled(q[:, y0:y1].contiguous(), 4, ny, y0, method="square")
    engine.check_status(st)
    res["tiled_quirk_case"] = _max_abs(out, ref[:, y0:y1])
    res["tiled_quirk_fixed_slices"] = int(fixed)
    # percentiles: slices over ranks
    pd = full[:max(2, min(n_slices, world + 1)), :64, :96].contiguous()
    refp, st = engine.neighbourhood_percentiles(pd, 3, (10, 50, 90))
    engine.check_status(st)
    gotp, st, _ = D.percentiles_sharded(pd, 3, (10, 50, 90), gather=True)
    engine.check_status(st)
    res["percentiles_sharded_max_abs"] = _max_abs(gotp, refp)
    # max over ranks
    keys = ["sharded_max_abs", "tiled_max_abs", "tiled_quirk_case", "percentiles_sharded_max_abs"]
    t = torch.tensor([res[k] for k in keys], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    for k, v in zip(keys, t.tolist()):
        res[k] = float(v)
    return res
