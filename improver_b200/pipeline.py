"""Host-facing pipeline of the fused chain and measurement helpers.

`threshold_neighbourhood_mean_host` is the call a user of the reference makes
when the cube lives in host memory: it streams lead-time chunks through pinned
buffers (H2D on a copy stream, kernels on the compute stream, D2H of the much
smaller result), so the PCIe copy overlaps the kernels.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from typing import Sequence

import numpy as np
import torch

from . import _lib, engine

LAUNCHES_PER_FUSED_CALL = 5     # init_stats, square kernel, collect_suspects, square_stats, square_fixup


MEMBER_PAD_BYTES = 65536


def member_padded_empty(shape, device, pad_bytes: int = MEMBER_PAD_BYTES) -> torch.Tensor:
    """Uninitialised float32 device cube (members, ...) whose member blocks are contiguous but
    `pad_bytes` apart from each other.  The device layout of a fused cube is ours to choose (the
    host side hands over contiguous members one by one): with the plain stride of the 18 x 36
    UK cube (145 546 560 B) the 18 member streams of a row alias in the DRAM channel hash and
    the fused kernel runs 4-5 % slower (profiles/sweep_pad.py)."""
    shape = tuple(int(s) for s in shape)
    block = int(np.prod(shape[1:]))
    stride = block + pad_bytes // 4
    flat = torch.empty(shape[0] * stride, dtype=torch.float32, device=device)
    inner = []
    acc = 1
    for s in reversed(shape[1:]):
        inner.append(acc)
        acc *= s
    return flat.as_strided(shape, (stride,) + tuple(reversed(inner)))


def pinned_like(t: torch.Tensor) -> torch.Tensor:
    return torch.empty(tuple(t.shape), dtype=t.dtype, pin_memory=True)


def _pinned_out(shape):
    """Pinned result buffer from PyTorch's caching host allocator: a fresh tensor per call (the
    caller owns it), cheap after the first call because released pinned blocks are reused."""
    return torch.empty(shape, dtype=torch.float32, pin_memory=True)


def threshold_neighbourhood_mean_host(host_in, thresholds: Sequence[float], bounds, comparison: str,
                                      cells: int, method: str = "square", mask2d=None,
                                      chunk_leads: int = 6, device=None, out=None):
    """Fuzzy threshold -> neighbourhood -> realization mean for a HOST cube.

    host_in   float32 (R, L, ny, nx) numpy array or (preferably pinned) CPU tensor
    out       optional pinned float32 (L, T, ny, nx) CPU tensor to receive the result
              (default: a new pinned tensor owned by the caller)
    returns   float32 (L, T, ny, nx) pinned CPU tensor
    Raises ValueError("Error: NaN detected in input cube data") like the
    reference plugins (nbhood.py:167-168)."""
    if isinstance(host_in, np.ndarray):
        host_in = torch.from_numpy(np.ascontiguousarray(host_in, dtype=np.float32))
    if host_in.is_cuda or host_in.dtype != torch.float32 or host_in.dim() != 4:
        raise TypeError("host_in must be a float32 CPU array of shape (R, L, ny, nx)")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    R, L, ny, nx = (int(s) for s in host_in.shape)
    T = len(thresholds)
    thr = [(t, b[0], b[1], comparison) for t, b in zip(thresholds, bounds)]
    out_host = out if out is not None else _pinned_out((L, T, ny, nx))
    if tuple(out_host.shape) != (L, T, ny, nx) or out_host.dtype != torch.float32 or out_host.is_cuda:
        raise ValueError("out must be a float32 CPU tensor of shape (L, T, ny, nx)")
    m2 = None
    if mask2d is not None:
        m2 = torch.as_tensor(np.asarray(mask2d) != 0).to(torch.uint8).to(dev)
    chunk = max(c for c in range(1, max(1, min(chunk_leads, L)) + 1) if L % c == 0)
    bufs = [member_padded_empty((R, chunk, ny, nx), dev) for _ in range(2)]
    outs = [torch.empty((chunk, T, ny, nx), dtype=torch.float32, device=dev) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)      # host -> device
    back_stream = torch.cuda.Stream(device=dev)      # device -> host (own DMA engine, own queue)
    compute = torch.cuda.current_stream(dev)
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    drained = [torch.cuda.Event() for _ in range(2)]
    statuses = []
    n_chunks = L // chunk
    for i in range(n_chunks):
        b = i & 1
        l0, l1 = i * chunk, (i + 1) * chunk
        with torch.cuda.stream(copy_stream):
            if i >= 2:
                copy_stream.wait_event(consumed[b])     # kernel that read bufs[b] has finished
            for r in range(R):
                bufs[b][r].copy_(host_in[r, l0:l1], non_blocking=True)
            copied[b].record(copy_stream)
        compute.wait_event(copied[b])
        if i >= 2:
            compute.wait_event(drained[b])              # outs[b] has been copied back
        _, _, st = engine.neighbourhood(bufs[b], cells, method=method, mask2d=m2, thresholds=thr,
                                        collapse_members=True, out=outs[b])
        statuses.append(st)
        consumed[b].record(compute)
        with torch.cuda.stream(back_stream):
            back_stream.wait_event(consumed[b])
            out_host[l0:l1].copy_(outs[b], non_blocking=True)
            drained[b].record(back_stream)
    status = torch.stack(statuses).amax(dim=0)
    compute.wait_stream(copy_stream)
    compute.wait_stream(back_stream)
    engine.check_status(status)          # D2H read of the status words: synchronises
    torch.cuda.current_stream(dev).synchronize()
    return out_host


class GraphedCall:
    """A native call sequence captured once into a CUDA graph and replayed.

    The fused call is one memset and five kernels, four of them tiny (trimming fix-up
    bookkeeping): replayed from a graph they run back to back without per-launch gaps
    (~0.03 ms of a 0.5 ms step).  `fn` must only use tensors that stay alive and must not
    synchronise; its return value (e.g. the status tensor) is kept as `result`."""

    def __init__(self, fn, warmup: int = 2):
        for _ in range(warmup):                # allocate workspaces / module state outside the capture
            fn()
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.result = fn()

    def __call__(self):
        self.graph.replay()
        return self.result


def time_main_kernel(fn, reps: int = 5) -> float:
    """Average duration (ms) of the dominant kernel of `fn()`'s native call,
    measured with CUDA events recorded inside the library on the launching
    stream (nbh_profile_enable / nbh_profile_last_ms)."""
    lib = _lib.load()
    lib.nbh_profile_enable(1)
    total = 0.0
    try:
        for _ in range(reps):
            fn()
            ms = C.c_float()
            _lib.check(lib.nbh_profile_last_ms(C.byref(ms)))
            total += ms.value
    finally:
        lib.nbh_profile_enable(0)
    return total / reps


def last_profiled_kernel() -> str:
    """Name of the kernel bracketed by the last profiled native call."""
    return _lib.load().nbh_profile_last_kernel().decode("utf-8", "replace")


def recorded_traffic_bytes():
    """DRAM bytes per launch of the dominant kernel from the committed ncu
    capture (profiles/roofline.json), or None if no capture has been made."""
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles",
                        "roofline.json")
    try:
        with open(path) as fh:
            return json.load(fh).get("square_fused_traffic_bytes")
    except (OSError, ValueError):
        return None
