#!/usr/bin/env python
"""Benchmark of the neighbourhood hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Headline workload (config.workload): BASELINE.json configs[2], the configuration the metric is
quoted on — fused fuzzy Threshold -> square neighbourhood (radius 10 km = 5 cells) ->
realization mean on an 18-member x 36-lead-time UK 2 km (970 x 1042) float32 cube, synthetic
gamma(2, 1.5) "rain rate".  The job is a GLOBAL BATCH of CYCLES (8) such cubes ("forecast
cycles"), sharded over the ranks with distributed.shard_range (strong scaling: the batch does
not grow with N; lead-time slices are independent, no data-path collective).  One step = one
pass of the hot path over the whole batch; `cube_ms` is the time of one cube (= configs[2]).
Unit: gridpoint-evaluations/s, one evaluation = one (realization, threshold, lead-time, y, x)
neighbourhood result.

Besides the headline the JSON line carries
  e2e      the same job through the plugin API (Threshold -> NeighbourhoodProcessing ->
           collapse_realizations on deferred cubes) from pinned HOST buffers, copies inside;
  roofline the dominant kernel, timed in-run with CUDA events inside the library;
  configs  (N = 1) the other BASELINE.json configurations: time, roofline against the unit that
           bounds them, in-run max-abs error against the oracle, 1-core / all-core CPU figures;
  scale    a BASELINE.json configs[4] subset (global 1920 x 2560 grid, periodic x, slices
           generated on the device from a counter-based generator) slice-sharded over the ranks,
           and y-band tiled with the NCCL halo exchange inside the timed region;
  parity   (N > 1) sharded / tiled results against the single-rank result, bit for bit.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

R, L, NY, NX = 18, 36, 970, 1042
CYCLES = 8
CELLS = 5
THRESHOLD, FUZZY = 2.0, 0.7
BOUNDS = (THRESHOLD * FUZZY, THRESHOLD * (2.0 - FUZZY))
T = 1
EVALS_PER_CUBE = R * T * L * NY * NX
ALGO_BYTES_PER_CUBE = 4 * NY * NX * L * (R + T)
METRIC = "neighbourhood gridpoint-evals/s"
UNIT = "gridpoint-evals/s"
WORKLOAD = ("fused fuzzy Threshold(2.0, ff=0.7) -> square nbhood r=5 cells (10 km) -> mean over "
            "18 realizations; 36 lead times; UK 2 km grid 970x1042; float32")
GLOBAL_NY, GLOBAL_NX = 1920, 2560


def base_config():
    """Identical for both arms (`--impl ours` / `--impl reference`)."""
    return {"workload": WORKLOAD, "shape": [R, L, NY, NX], "cells": CELLS, "thresholds": T,
            "global_batch": f"{CYCLES} forecast cycles of the cube above, sharded over the ranks (shard_range)"}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._thr = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                     "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thr.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active")
                                                          for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]),
                "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------- #
# CPU side: the reference's own functions (oracle/_ref) or the oracle port
# ----------------------------------------------------------------------------- #
def synth_host(n_lead: int, seed: int) -> np.ndarray:
    from improver_b200 import workloads as W

    return W.config3_rain((R, n_lead, NY, NX), seed)


def _cpu_one_lead(args):
    seed, n_thr, cells = args
    from oracle import ref_runtime as rr
    data = synth_host(1, seed)
    thr = [THRESHOLD * (0.5 + 0.5 * k) for k in range(1, n_thr + 1)] if n_thr > 1 else [THRESHOLD]
    bounds = [(t * FUZZY, t * (2.0 - FUZZY)) for t in thr]
    t0 = time.perf_counter()
    rr.threshold_square_mean(data, thr, bounds, ">", cells)
    return time.perf_counter() - t0


def cpu_reference_rate(n_lead: int, procs: int, n_thr: int = T, cells: int = CELLS):
    """The reference chain on `n_lead` lead times spread over `procs` processes (slices are
    independent: how operational suites run the reference).  Returns evals/s (wall clock of the
    pool), the 1-process rate of one lead time, and the sample text."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        pool.map(_cpu_one_lead, [(100 + i, n_thr, cells) for i in range(procs)][:1])     # start-up / imports
        t0 = time.perf_counter()
        per = pool.map(_cpu_one_lead, [(200 + i, n_thr, cells) for i in range(n_lead)], chunksize=1)
        wall = time.perf_counter() - t0
    evals = R * n_thr * n_lead * NY * NX
    return evals / wall, f"{n_lead} of {L} lead times x {R} members, {procs} processes", per


def cpu_kind():
    from oracle import ref_runtime as rr

    k = rr.kind()
    note = ("oracle/_ref: the reference's own NeighbourhoodProcessing._calculate_neighbourhood / "
            "Threshold._calculate_truth_value, numpy/scipy" if k == "reference" else
            "oracle/nbhood_oracle.py: numpy/scipy restatement of the reference path (oracle/_ref not built)")
    return k, note


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    cores = max(1, min(cores, int(os.environ.get("NBH_BENCH_REF_PROCS", cores))))   # tests shrink the sample
    kind, note = cpu_kind()
    vals = []
    sample = ""
    for _ in range(args.warmup):
        cpu_reference_rate(max(1, cores // 4), cores)
    for _ in range(args.steps):
        v, sample, _ = cpu_reference_rate(cores, cores)
        vals.append(v)
    value = float(np.mean(vals))
    step_ms = (R * T * cores * NY * NX) / value * 1e3
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32/f64-accum",
        "data": "synthetic", "config": base_config(),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": sample + " per step (" + note + ")"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---- CPU figures of the secondary configurations (bounded samples) -------------------------- #
def _cpu_config_case(case):
    """Run one bounded CPU sample of a secondary configuration; returns (seconds, evals, sample)."""
    from improver_b200 import workloads as W
    from oracle import ref_runtime as rr

    name = case["name"]
    if name == "c1":
        d = np.random.default_rng(0).random((1, 200, 200), dtype=np.float32)
        t0 = time.perf_counter()
        for _ in range(20):
            rr.neighbourhood_slices(d, None, "square", 5)
        return (time.perf_counter() - t0) / 20, 200 * 200, "the 200 x 200 slice itself"
    if name.startswith("c2_r"):
        r = case["cells"]
        side = 16 + 2 * r
        y0, x0 = 400, 400
        d = W.config2_probabilities(1)[:, y0:y0 + side, x0:x0 + side]
        land = W.land_sea_mask()[y0:y0 + side, x0:x0 + side]
        t0 = time.perf_counter()
        rr.land_sea(np.ascontiguousarray(d), np.ascontiguousarray(land), "circular", r)
        return time.perf_counter() - t0, side * side, f"one {side} x {side} crop, land + sea passes (cost is per point)"
    if name.startswith("unfused"):
        r = case["cells"]
        d = W.config2_probabilities(2)
        m = W.land_sea_mask() if case.get("mask") else None
        t0 = time.perf_counter()
        rr.neighbourhood_slices(d, m, "square", r)
        return time.perf_counter() - t0, 2 * NY * NX, "2 of 240 UK slices"
    if name == "c4":
        d = np.random.default_rng(4).random((140, 140), dtype=np.float32)
        t0 = time.perf_counter()
        rr.percentiles(d, 10, (5, 25, 50, 75, 95))
        return time.perf_counter() - t0, 140 * 140, "one 140 x 140 crop (cost is per point)"
    raise ValueError(name)


def cpu_config_rates(cases, cores):
    """1-core rate of every case (run alone) and an all-core rate (the same sample on every core
    at once: slices are independent, this is how the reference scales on a host)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = {}
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_config_case, [cases[0]])                  # imports
        for case in cases:
            sec, evals, sample = pool.map(_cpu_config_case, [case])[0]
            t0 = time.perf_counter()
            pool.map(_cpu_config_case, [case] * cores, chunksize=1)
            wall = time.perf_counter() - t0
            out[case["name"]] = {"one_core": evals / sec, "all_cores": evals * cores / wall, "cores": cores,
                                 "unit": UNIT, "sample": sample}
    return out


# ----------------------------------------------------------------------------- #
# GPU side
# ----------------------------------------------------------------------------- #
def _numa_bind(local_rank: int, world: int):
    """Run this rank (and therefore allocate its pinned buffers: first touch) on the CPUs of one
    NUMA node: the node of its GPU when the platform reports one, else ranks are spread evenly
    over the nodes.  Returns a description for the JSON line."""
    policy = os.environ.get("NBH_BENCH_NUMA", "auto")
    try:
        nodes = sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
    except OSError:
        nodes = []
    if policy == "off" or len(nodes) < 2:
        return {"policy": "none", "nodes": len(nodes)}
    import torch

    node = None
    if policy in ("auto", "local"):
        try:
            bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
            dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
            dev = torch.cuda.get_device_properties(local_rank).pci_device_id
            path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
            with open(path) as fh:
                node = int(fh.read().strip())
        except Exception:
            node = None
        if node is not None and node < 0:
            node = None
    if node is None or policy == "spread":
        node = nodes[(local_rank * len(nodes)) // max(world, 1)]
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            spec = fh.read().strip()
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, cpus)
        return {"policy": policy, "node": node, "nodes": len(nodes), "cpus": len(cpus)}
    except Exception as exc:
        return {"policy": "failed", "error": type(exc).__name__}


def _log(msg):
    if os.environ.get("RANK", "0") == "0":
        print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def _events():
    import torch

    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def _timed(fn, min_s=0.25, warm=2, sync=None):
    """Mean device time (ms) of fn(), over at least `min_s` seconds of back-to-back calls."""
    import torch

    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = _events()
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    one = max(e0.elapsed_time(e1), 1e-3)
    reps = int(max(3, min(2000, min_s * 1e3 / one)))
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, reps


def _make_cube(dev, seed):
    """Synthetic gamma(2, 1.5) cube on the device (sum of two exponentials); member blocks
    64 KiB further apart than the contiguous stride (pipeline.member_padded_empty)."""
    import torch

    from improver_b200 import pipeline

    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    cube = pipeline.member_padded_empty((R, L, NY, NX), dev)
    for r in range(R):
        u = torch.rand((2, L, NY, NX), generator=gen, device=dev).clamp_min_(1e-12)
        cube[r] = -1.5 * (torch.log(u[0]) + torch.log(u[1]))
        del u
    return cube


def _plugin_chain(host_cube_tensor, n_thr=1):
    """The call a user of the reference makes, on a deferred cube over a pinned host buffer."""
    from improver_b200 import lazy
    from improver_b200.cube import Coord, Cube
    from improver_b200.nbhood.nbhood import NeighbourhoodProcessing
    from improver_b200.threshold import Threshold
    from improver_b200.utilities.cube_manipulation import collapse_realizations

    y = Coord(np.arange(NY, dtype=np.float32) * 2000.0, "projection_y_coordinate", "m", "y")
    x = Coord(np.arange(NX, dtype=np.float32) * 2000.0, "projection_x_coordinate", "m", "x")
    cube = Cube(lazy.Source(host_cube_tensor), "lwe_precipitation_rate", "mm h-1",
                [(Coord(np.arange(R, dtype=np.int32), "realization", "1"), 0),
                 (Coord(np.arange(L, dtype=np.int64) * 3600, "time", "seconds"), 1), (y, 2), (x, 3)],
                [Coord([3600], "forecast_period", "seconds")])
    thr = [THRESHOLD] if n_thr == 1 else [THRESHOLD * (0.5 + 0.5 * k) for k in range(1, n_thr + 1)]
    prob = Threshold(threshold_values=thr, fuzzy_factor=FUZZY, comparison_operator=">")(cube)
    nbh = NeighbourhoodProcessing("square", CELLS * 2000.0)(prob)
    mean = collapse_realizations(nbh)
    return mean.data                                        # realises: streamed H2D, fused kernel, D2H


def secondary_configs(dev, cube, sm_clock_mhz):
    """configs of the JSON line (rank 0, N = 1): the other BASELINE.json configurations."""
    import torch

    from improver_b200 import engine, pipeline, workloads as W
    from oracle import nbhood_oracle as orc

    peak_hbm, _ = peaks()
    smem_peak = 148 * 128 * (sm_clock_mhz or 1965.0) * 1e6 / 1e9      # GB/s: 128 B / clk / SM
    res = {}

    def roof(bound, achieved, peak, unit="GB/s", **kw):
        return dict({"bound": bound, "achieved": achieved, "peak": peak, "unit": unit,
                     "frac": achieved / peak if peak else None}, **kw)

    def max_abs(got, exp):
        got, exp = np.asarray(got), np.asarray(exp)
        ok = ~np.isnan(exp)
        if not np.array_equal(np.isnan(got), np.isnan(exp)):
            return float("inf")
        return float(np.max(np.abs(got[ok] - exp[ok]))) if ok.any() else 0.0

    # ---- config 1: 200 x 200, square r = 5 (launch-latency bound) ----
    h1 = np.random.default_rng(0).random((1, 200, 200), dtype=np.float32)
    d1 = torch.from_numpy(h1).to(dev)
    ms, _ = _timed(lambda: engine.neighbourhood(d1, 5), min_s=0.05)
    out, _, st = engine.neighbourhood(d1, 5)
    engine.check_status(st)
    exp = np.ma.getdata(orc.neighbourhood(h1[0], None, method="square", cells=5, re_mask=False))
    res["c1"] = {"workload": "configs[0]: square r=5 on one 200x200 slice", "ms": ms, "value": 200 * 200 / ms * 1e3,
                 "roofline": roof("launch latency", 8 * 200 * 200 / ms / 1e6, peak_hbm, note="160 KB: one wave, not a bandwidth case"),
                 "max_abs": max_abs(out[0].cpu().numpy(), exp)}

    # ---- un-fused square on 240 UK slices (NeighbourhoodProcessing.process itself) ----
    n = 240
    gen = torch.Generator(device=dev).manual_seed(1)
    d2 = torch.rand((n, NY, NX), generator=gen, device=dev)
    land_h = W.land_sea_mask()
    land = torch.from_numpy((land_h > 0).astype(np.uint8)).to(dev)
    crop = (slice(0, 2), slice(300, 360), slice(0, 90))
    for r in (5, 27):
        for m, tag in ((None, ""), (land, "_mask")):
            ms, _ = _timed(lambda: engine.neighbourhood(d2, r, mask2d=m))
            kernel_ms = pipeline.time_main_kernel(lambda: engine.neighbourhood(d2, r, mask2d=m), reps=3)
            out, _, st = engine.neighbourhood(d2, r, mask2d=m)           # the timed launch: last two slices checked
            engine.check_status(st)
            out = out[-2:]
            ya, yb = 300 - r, 360 + r
            sub = d2[-2:, ya:yb, :90 + r].cpu().numpy()
            subm = land_h[ya:yb, :90 + r] if m is not None else None
            exp = np.stack([np.ma.getdata(orc.neighbourhood(s_, subm, method="square", cells=r, re_mask=False))
                            for s_ in sub])[:, r:r + 60, :90]
            res[f"unfused_r{r}{tag}"] = {
                "workload": f"square r={r} on {n} UK slices" + (", land mask weighted" if m is not None else ""),
                "ms": ms, "value": n * NY * NX / ms * 1e3,
                "roofline": roof("hbm", 8 * n * NY * NX / kernel_ms / 1e6, peak_hbm, kernel=pipeline.last_profiled_kernel(),
                                 kernel_ms=kernel_ms, algorithmic_bytes=8 * n * NY * NX),
                "max_abs": max_abs(out[crop].cpu().numpy(), exp)}

    # ---- config 2: circular, land + sea in one launch, lead-time radii ----
    for r in (9, 27, 45, 81):
        nn = n if r <= 27 else 48
        dd = d2[:nn]
        ms, _ = _timed(lambda: engine.neighbourhood_land_sea(dd, r, land, method="circular"), min_s=0.1, warm=1)
        out, st = engine.neighbourhood_land_sea(dd, r, land, method="circular")     # the timed launch: last slice checked
        engine.check_status(st)
        out = out[-1:]
        side = 12
        y0, x0 = 500, 480
        ya, yb, xa, xb = y0 - r, y0 + side + r, x0 - r, x0 + side + r
        exp = orc.land_sea_neighbourhood(dd[-1:, ya:yb, xa:xb].cpu().numpy(), land_h[ya:yb, xa:xb],
                                         method="circular", cells=r)[:, r:r + side, r:r + side]
        gathered = nn * NY * NX * 2 * (2 * r + 1) * 4
        res[f"c2_r{r}"] = {
            "workload": f"configs[1]: circular r={r} cells, land-sea mask weighted, re_mask, {nn} UK slices (one launch for land + sea)",
            "ms": ms, "value": nn * NY * NX / ms * 1e3,
            "roofline": roof("shared memory / L1 (2(2r+1) prefix reads of 4 B per result)", gathered / ms / 1e6, smem_peak,
                             hbm_algorithmic_GBs=8 * nn * NY * NX / ms / 1e6, peak_source="148 SMs x 128 B/clk x SM clock"),
            "max_abs": max_abs(out[:, y0:y0 + side, x0:x0 + side].cpu().numpy(), exp)}

    # ---- config 3 variants: four thresholds, radius 27 ----
    thr4 = [(t, t * FUZZY, t * (2.0 - FUZZY), ">") for t in (1.0, 2.0, 3.0, 4.0)]
    out4 = torch.empty((L, 4, NY, NX), dtype=torch.float32, device=dev)
    fn4 = lambda: engine.neighbourhood(cube, CELLS, thresholds=thr4, collapse_members=True, out=out4)
    ms, _ = _timed(fn4)
    host = cube[:, 3:4].cpu().numpy()
    exp = orc.threshold_square_mean(host, [t[0] for t in thr4], [(t[1], t[2]) for t in thr4], ">", CELLS)
    res["c3_T4"] = {"workload": "configs[2] with four thresholds", "ms": ms, "value": 4 * EVALS_PER_CUBE / ms * 1e3,
                    "roofline": roof("hbm", 4 * NY * NX * L * (R + 4) / ms / 1e6, peak_hbm,
                                     algorithmic_bytes=4 * NY * NX * L * (R + 4)),
                    "max_abs": max_abs(out4[3:4].cpu().numpy(), exp)}
    out1 = torch.empty((L, 1, NY, NX), dtype=torch.float32, device=dev)
    thr1 = [(THRESHOLD, BOUNDS[0], BOUNDS[1], ">")]
    fn27 = lambda: engine.neighbourhood(cube, 27, thresholds=thr1, collapse_members=True, out=out1)
    ms, _ = _timed(fn27)
    exp = orc.threshold_square_mean(host, [THRESHOLD], [BOUNDS], ">", 27)
    fn27()
    res["c3_r27"] = {"workload": "configs[2] with a radius of 27 cells (54 km)", "ms": ms, "value": EVALS_PER_CUBE / ms * 1e3,
                     "roofline": roof("hbm", ALGO_BYTES_PER_CUBE / ms / 1e6, peak_hbm, algorithmic_bytes=ALGO_BYTES_PER_CUBE),
                     "max_abs": max_abs(out1[3:4].cpu().numpy(), exp)}

    # ---- config 4: neighbourhood percentiles ----
    gen4 = torch.Generator(device=dev).manual_seed(4)
    d4 = torch.rand((18, NY, NX), generator=gen4, device=dev)
    pcts = (5, 25, 50, 75, 95)
    ms, _ = _timed(lambda: engine.neighbourhood_percentiles(d4, 10, pcts), min_s=0.1, warm=1)
    out, st = engine.neighbourhood_percentiles(d4[:1], 10, pcts)
    engine.check_status(st)
    exp = orc.neighbourhood_percentiles(d4[0, 90:150, 190:260].cpu().numpy(), 10, pcts)
    res["c4"] = {"workload": "configs[3]: percentiles 5/25/50/75/95 of a circular r=10 neighbourhood, 18 realizations",
                 "ms": ms, "value": 18 * NY * NX / ms * 1e3,
                 "roofline": roof("instruction issue (selection over 317 keys per point)", 18 * NY * NX / ms * 1e3,
                                  None, unit="points/s", hbm_algorithmic_GBs=(4 + 20) * 18 * NY * NX / ms / 1e6),
                 "max_abs": max_abs(out[0, :, 100:140, 200:250].cpu().numpy(), exp[:, 10:50, 10:60])}
    return res


def scale_workload(dev, rank, world):
    """BASELINE.json configs[4] subset, strong scaling: S slices of the global 1920 x 2560 grid
    (periodic x) generated on the device from the counter-based generator.  (a) slice sharding
    (shard_range, no collective), (b) y-band tiling of a smaller stack with the halo exchange,
    the statistics all-reduce and the crop inside the timed region."""
    import torch
    import torch.distributed as dist

    from improver_b200 import distributed as D, engine, workloads as W

    S = 384
    s0, s1 = D.shard_range(S, world, rank)
    mine = W.counter_uniform(s0, s1 - s0, (GLOBAL_NY, GLOBAL_NX), 5, dev)
    out = torch.empty_like(mine)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxrank(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    res = {"workload": f"configs[4] subset: {S} slices of the global {GLOBAL_NY}x{GLOBAL_NX} grid, periodic x, "
                       "counter-based generator (seed 5), square and circular r=5", "scaling": "strong",
           "slices": S, "slices_per_rank": s1 - s0}
    for method in ("square", "circular"):
        fn = lambda: engine.neighbourhood(mine, 5, method=method, wrap_x=True, out=out)
        for _ in range(3):
            fn()
        barrier()
        e0, e1 = _events()
        reps = 10
        e0.record()
        for _ in range(reps):
            _, _, st = fn()
        e1.record()
        torch.cuda.synchronize()
        engine.check_status(st)
        ms = maxrank(e0.elapsed_time(e1) / reps)
        # the last slice of the share against a launch of that slice alone (bit for bit)
        one, _, st = engine.neighbourhood(mine[-1:].contiguous(), 5, method=method, wrap_x=True)
        engine.check_status(st)
        res[f"sharded_{method}"] = {"ms": ms, "value": S * GLOBAL_NY * GLOBAL_NX / ms * 1e3,
                                    "hbm_algorithmic_GBs": 8 * S * GLOBAL_NY * GLOBAL_NX / ms / 1e6,
                                    "last_slice_max_abs_vs_single_launch": float((out[-1] - one[0]).abs().max())}
    del out
    # y-band tiling: Sb slices, every rank a band of rows of all of them
    Sb = 48
    y0, y1 = D.band_range(GLOBAL_NY, world, rank)
    full_rows = W.counter_uniform(0, Sb, (GLOBAL_NY, GLOBAL_NX), 5, dev)[:, y0:y1].contiguous()
    halo_ms = 0.0

    def tiled():
        return D.neighbourhood_y_tiled(full_rows, 5, GLOBAL_NY, y0, method="square", wrap_x=True)

    for _ in range(2):
        tiled()
    barrier()
    e0, e1 = _events()
    reps = 5
    e0.record()
    for _ in range(reps):
        _, _, st = tiled()
    e1.record()
    torch.cuda.synchronize()
    engine.check_status(st)
    ms = maxrank(e0.elapsed_time(e1) / reps)
    # the exchange alone (pack, ncclSend / ncclRecv, unpack), same stream
    barrier()
    h0, h1 = _events()
    h0.record()
    for _ in range(reps):
        D.exchange_halo_rows(full_rows, 5)
    h1.record()
    torch.cuda.synchronize()
    halo_ms = maxrank(h0.elapsed_time(h1) / reps)
    res["y_band"] = {"slices": Sb, "rows_per_rank": y1 - y0, "ms": ms, "value": Sb * GLOBAL_NY * GLOBAL_NX / ms * 1e3,
                     "halo_exchange_ms": halo_ms,
                     "halo_bytes_per_rank": int(Sb * 5 * GLOBAL_NX * 4 * (int(rank > 0) + int(rank < world - 1))),
                     "includes": "halo exchange (nbh_halo_exchange: pack, ncclSend/ncclRecv, unpack), statistics "
                                 "kernel + min/max all-reduce, kernels on the extended band, crop"}
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist

    from improver_b200 import distributed as D, engine, pipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = _numa_bind(local, world)
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner there while the
        # communicator is created, so file descriptor 1 points at stderr until that is done
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxrank(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- this rank's share of the global batch of forecast cycles ----
    c0, c1 = D.shard_range(CYCLES, world, rank)
    cubes = [_make_cube(dev, 3 + c) for c in range(c0, c1)]
    thr = [(THRESHOLD, BOUNDS[0], BOUNDS[1], ">")]
    outs = [torch.empty((L, T, NY, NX), dtype=torch.float32, device=dev) for _ in cubes]

    def eager_step():
        st = None
        for cube, out in zip(cubes, outs):
            _, _, st = engine.neighbourhood(cube, CELLS, method="square", thresholds=thr,
                                            collapse_members=True, out=out)
        return st

    _log(f"{len(cubes)} cube(s) generated")
    warm = max(args.warmup, 3)
    for _ in range(warm):
        status = eager_step()
    if status is not None:
        engine.check_status(status)
    step = eager_step
    launch_mode = "eager"
    if not args.no_graph and cubes:
        try:
            graphed = pipeline.GraphedCall(eager_step)
            step = graphed
            launch_mode = "cuda graph replay"
            engine.check_status(step())
        except Exception as exc:                      # keep measuring eagerly, but say so
            step = eager_step
            launch_mode = f"eager (graph capture failed: {type(exc).__name__})"

    # ---- device-resident timing.  `value` is the SUSTAINED rate: >= 2 s of back-to-back steps
    # bring the part to the clocks / power state it holds under load, then exactly K steps are
    # timed.  The burst figure (K steps after 2 s of idling) is reported beside it. ----
    ev0, ev1 = _events()
    with ClockSampler(local) as clocks:
        t_end = time.perf_counter() + float(os.environ.get("NBH_BENCH_PRELOAD_S", "2.0"))
        n_pre = 0
        while time.perf_counter() < t_end:
            for _ in range(10):
                status = step()
            torch.cuda.synchronize()
            n_pre += 10
        barrier()
        ev0.record()
        for _ in range(args.steps):
            status = step()
        ev1.record()
        torch.cuda.synchronize()
        ms_sustained = maxrank(ev0.elapsed_time(ev1) / args.steps)
        # dominant kernel, same state (events recorded inside the library on the launching stream)
        kernel_ms = pipeline.time_main_kernel(lambda: engine.neighbourhood(
            cubes[0], CELLS, method="square", thresholds=thr, collapse_members=True, out=outs[0]),
            reps=max(5, args.steps)) if cubes else float("nan")
        kernel_name = pipeline.last_profiled_kernel()
    if status is not None:
        engine.check_status(status)
    time.sleep(2.0)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        status = step()
    ev1.record()
    torch.cuda.synchronize()
    ms_burst = maxrank(ev0.elapsed_time(ev1) / args.steps)
    value = EVALS_PER_CUBE * CYCLES / (ms_sustained * 1e-3)
    peak, peak_src = peaks()
    achieved = ALGO_BYTES_PER_CUBE / (kernel_ms * 1e-3) / 1e9

    _log(f"device timing done: sustained {ms_sustained:.4f} ms/step, burst {ms_burst:.4f}, kernel {kernel_ms:.4f}")
    # ---- in-run check of the timed output against the oracle (one lead time of the first cycle) ----
    max_abs = None
    if rank == 0 and cubes and not args.no_cpu:
        from oracle import nbhood_oracle as orc
        lead = 7
        exp = orc.threshold_square_mean(cubes[0][:, lead:lead + 1].cpu().numpy(), [THRESHOLD], [BOUNDS], ">", CELLS)
        max_abs = float(np.max(np.abs(outs[0][lead:lead + 1].cpu().numpy() - exp)))

    # ---- end to end through the plugin API with HOST buffers ----
    e2e_steps = max(1, min(args.steps, 2))
    e2e = {"value": None, "unit": UNIT}
    if not args.no_e2e and cubes:
        host_in = torch.empty((R, L, NY, NX), dtype=torch.float32, pin_memory=True)   # NUMA-local: this rank's CPUs
        host_in.copy_(cubes[0])
        torch.cuda.synchronize()
        res = _plugin_chain(host_in)
        ref = outs[0].cpu().numpy()
        e2e_err = float(np.max(np.abs(np.asarray(res).reshape(ref.shape) - ref)))
        del res
        barrier()
        ev0.record()
        for _ in range(e2e_steps):
            for _c in cubes:
                res = _plugin_chain(host_in)
                del res
        ev1.record()
        torch.cuda.synchronize()
        mine_ms = ev0.elapsed_time(ev1) / e2e_steps
        e2e_ms = maxrank(mine_ms)
        h2d = R * L * NY * NX * 4 * CYCLES
        d2h = (L * T * NY * NX * 4 + 8) * CYCLES
        gbs = torch.tensor([len(cubes) * R * L * NY * NX * 4 / (mine_ms * 1e-3) / 1e9], dtype=torch.float64, device=dev)
        if world > 1:
            all_gbs = [torch.zeros_like(gbs) for _ in range(world)]
            dist.all_gather(all_gbs, gbs)
            per_rank = [round(float(g.item()), 2) for g in all_gbs]
        else:
            per_rank = [round(float(gbs.item()), 2)]
        e2e = {"value": EVALS_PER_CUBE * CYCLES / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": e2e_ms,
               "api": "Threshold(...)(cube) -> NeighbourhoodProcessing('square', 10000)(cube) -> "
                      "collapse_realizations(cube) -> .data on deferred cubes (improver_b200.lazy) over a pinned host "
                      "buffer; one chain per forecast cycle",
               "max_abs_vs_device_path": e2e_err, "h2d_GBs_per_rank": per_rank, "numa": numa,
               "limit": "PCIe host-to-device copy of the 2.62 GB cube per cycle"}
        del host_in

    _log(f"e2e done: {e2e.get('value')}")
    # ---- multi-rank parity and the sharded configs[4] subset ----
    parity = None
    if world > 1 and not args.no_scale:
        from improver_b200 import parity as P
        parity = P.multi_rank_parity(dev)
    scale = None
    if not args.no_scale:
        for c in cubes[1:]:
            del c
        cubes_keep = cubes[:1]
        cubes, outs = cubes_keep, outs[:1]
        torch.cuda.empty_cache()
        scale = scale_workload(dev, rank, world)

    _log("scale / parity done")
    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms_sustained, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32 (exact fixed-point / f64 window sums)",
            "data": "synthetic",
            "config": dict(base_config(), cycles_per_rank=c1 - c0,
                           l2="inputs (2.6 GB per cube) larger than L2; no flush needed",
                           sharding="forecast cycles over ranks (distributed.shard_range), lead-time slices "
                                    "independent, no collective",
                           device_layout="(R, L, y, x) float32, member stride padded by 64 KiB",
                           launch=launch_mode, timing=f"sustained: {n_pre} untimed steps (>= 2 s) then {args.steps} timed"),
            "cube_ms": ms_sustained / max(1, c1 - c0), "burst": {"ms_per_step": ms_burst,
                                                                 "value": EVALS_PER_CUBE * CYCLES / (ms_burst * 1e-3),
                                                                 "note": "same K steps after 2 s of idling"},
            "max_abs": max_abs, "max_abs_check": "lead time 7 of the first cycle's timed output vs oracle",
            "clocks": clocks.summary(),
            "e2e": e2e,
            "gpu_launches": pipeline.LAUNCHES_PER_FUSED_CALL * args.steps * (c1 - c0),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "frac_of_8TBs": achieved / 8000.0,
                         "traffic": pipeline.recorded_traffic_bytes(),
                         "kernel": kernel_name,
                         "kernel_ms": kernel_ms, "state": "sustained (measured right after the preload + timed steps)",
                         "algorithmic_bytes": ALGO_BYTES_PER_CUBE, "peak_source": peak_src},
            "hbm_gbs": achieved,
            "parity": parity,
            "scale": scale,
        }
    # ---- secondary configurations and CPU figures beside them (rank 0, N == 1 only) ----
    if rank == 0 and world == 1 and not args.no_configs:
        sm_clock = line["clocks"].get("sm_mhz")
        line["configs"] = secondary_configs(dev, cubes[0], sm_clock)
        _log("secondary configs done")
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        kind, note = cpu_kind()
        n_lead = max(1, min(cores, 16))
        v, sample, per = cpu_reference_rate(n_lead, min(cores, n_lead))
        one = cpu_reference_rate(1, 1)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": min(cores, n_lead), "kind": kind,
                                "sample": sample + " (" + note + ")",
                                "one_core": {"value": one[0], "sample": "1 lead time x 18 members, 1 process"}}
        if "configs" in line:
            cases = ([{"name": "c1"}] + [{"name": f"c2_r{r}", "cells": r} for r in (9, 27, 45, 81)] +
                     [{"name": f"unfused_r{r}{'_mask' if m else ''}", "cells": r, "mask": m} for r in (5, 27) for m in (False, True)] +
                     [{"name": "c4"}])
            rates = cpu_config_rates(cases, cores)
            for name, rate in rates.items():
                line["configs"][name]["cpu_baseline"] = dict(rate, kind=kind)
            for name, n_thr, cells in (("c3_T4", 4, CELLS), ("c3_r27", 1, 27)):
                v4, s4, per4 = cpu_reference_rate(min(cores, 8), min(cores, 8), n_thr=n_thr, cells=cells)
                line["configs"][name]["cpu_baseline"] = {
                    "all_cores": v4, "one_core": R * n_thr * NY * NX / float(np.mean(per4)), "cores": min(cores, 8),
                    "unit": UNIT, "sample": s4, "kind": kind,
                    "note": "one_core from the per-process times of the pool (contended memory bandwidth)"}
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs and the in-run oracle check")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host end-to-end leg (profiling runs)")
    ap.add_argument("--no-configs", action="store_true", help="skip the secondary configurations")
    ap.add_argument("--no-scale", action="store_true", help="skip the configs[4] subset and the multi-rank parity leg")
    ap.add_argument("--no-graph", action="store_true", help="launch the timed steps eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
