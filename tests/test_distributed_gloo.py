"""World-size-2 (and 3) gloo tests of the multi-GPU host logic: slice
sharding and the y-band halo exchange."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from improver_b200 import distributed as D


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, cells, ny, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(3 * ny * 7, dtype=torch.float32).reshape(3, ny, 7)
        y0, y1 = D.band_range(ny, world, rank)
        ext = D.exchange_halo_rows(full[:, y0:y1].contiguous(), cells)
        e0 = max(0, y0 - cells) if rank > 0 else y0
        e1 = min(ny, y1 + cells) if rank < world - 1 else y1
        ok = torch.equal(ext, full[:, e0:e1])
        own = D.crop_own_rows(ext, cells)
        ok = ok and torch.equal(own, full[:, y0:y1])
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,cells,ny", [(2, 3, 20), (3, 2, 17)])
def test_halo_exchange_matches_global_slices(world, cells, ny):
    port = _free_port()
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_worker, args=(world, port, cells, ny, results), nprocs=world, join=True)
        assert dict(results) == {r: True for r in range(world)}


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 36, 648):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        D.shard_range(4, 2, 2)
