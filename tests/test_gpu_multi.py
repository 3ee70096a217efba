"""Two-GPU checks over NCCL: slice sharding with gather must reproduce the
single-GPU result bit for bit; y-band tiling with halo exchange must reproduce, bit for bit, the
single-GPU result of the exact-sum kernels it runs on (which the default float32 tile kernel matches
to 1e-6).  Skipped on boxes with fewer
than two GPUs (the CPU-side logic is covered by tests/test_distributed_gloo.py)."""
import os
import socket
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["NBH_ROOT"])
import numpy as np
import torch
import torch.distributed as dist
from improver_b200 import distributed as D, engine

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=dev)
g = torch.Generator(device="cpu").manual_seed(7)
full = torch.rand((5, 90, 128), generator=g).to(dev)                 # same on every rank
mask = (torch.rand((90, 128), generator=g) < 0.8).to(torch.uint8).to(dev)
fails = []
for method, cells, wrap in (("square", 4, False), ("square", 3, True), ("circular", 5, False)):
    for m in (None, mask):
        ref, _, st = engine.neighbourhood(full, cells, method=method, mask2d=m, wrap_x=wrap)
        engine.check_status(st)
        with engine.exact_square_sums():       # window sums that do not depend on the tile origin
            ref_exact, _, st = engine.neighbourhood(full, cells, method=method, mask2d=m, wrap_x=wrap)
        engine.check_status(st)
        if not torch.allclose(ref_exact.nan_to_num(-1.0), ref.nan_to_num(-1.0), rtol=0.0, atol=1e-6):
            fails.append(("exact-vs-default", method, cells, wrap, m is not None))
        # y-band tiling: halo rows over NCCL send/recv, kernels on the extended band, crop
        y0, y1 = D.band_range(full.shape[-2], world, rank)
        out, _, st = D.neighbourhood_y_tiled(full[:, y0:y1].contiguous(), cells,
                                             mask_band=m[y0:y1].contiguous() if m is not None else None,
                                             method=method, wrap_x=wrap)
        engine.check_status(st)
        if not torch.equal(out.nan_to_num(-1.0), ref_exact[:, y0:y1].nan_to_num(-1.0)):
            fails.append(("tiled", method, cells, wrap, m is not None))
        # slice sharding + all-gather
        got, _, st, _ = D.neighbourhood_sharded(full, cells, gather=True, method=method, mask2d=m, wrap_x=wrap)
        engine.check_status(st)
        if not torch.equal(got.nan_to_num(-1.0), ref.nan_to_num(-1.0)):
            fails.append(("sharded", method, cells, wrap, m is not None))
bad = torch.tensor([len(fails)], device=dev)
dist.all_reduce(bad)
if fails:
    print(f"rank {rank} failures: {fails}", flush=True)
if rank == 0:
    print("MULTI_GPU_OK" if int(bad.item()) == 0 else f"MULTI_GPU_FAIL {fails}")
dist.destroy_process_group()
'''


def test_two_gpu_tiling_and_sharding_match_single_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, NBH_ROOT=ROOT)
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, env=env, timeout=600)
    assert "MULTI_GPU_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
