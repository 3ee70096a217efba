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


def _sharded_worker(rank, world, port, n_slices, results):
    """neighbourhood_sharded with a stand-in for the native call (a 3-point mean along x on the CPU):
    the host logic under test is the share of every rank, the rank without a share when there are
    more ranks than slices, and the gather of uneven shares."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from improver_b200 import engine

        def fake_neighbourhood(data, cells, **kwargs):
            assert data.shape[0] >= 1, "the native entry point rejects empty input"
            pad = torch.nn.functional.pad(data, (1, 1), mode="replicate")
            out = (pad[..., :-2] + pad[..., 1:-1] + pad[..., 2:]) / 3
            return out, None, torch.zeros(2, dtype=torch.int32)

        engine.neighbourhood = fake_neighbourhood
        full = torch.arange(n_slices * 5 * 6, dtype=torch.float32).reshape(n_slices, 5, 6)
        ref, _, _ = fake_neighbourhood(full, 1)
        got, _, _, (s0, s1) = D.neighbourhood_sharded(full, 1, gather=True)
        ok = torch.equal(got, ref) and (s0, s1) == D.shard_range(n_slices, world, rank)
        part, _, _, _ = D.neighbourhood_sharded(full, 1, gather=False)
        ok = ok and torch.equal(part, ref[s0:s1])
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_slices", [(2, 5), (3, 2), (3, 1)])
def test_sharded_neighbourhood_uneven_and_empty_shares(world, n_slices):
    port = _free_port()
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_sharded_worker, args=(world, port, n_slices, results), nprocs=world, join=True)
        assert dict(results) == {r: True for r in range(world)}
