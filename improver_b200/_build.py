"""Build libimprover_b200.so (hand-written sm_100a CUDA + the C ABI) in-tree.

nvcc cross-compiles without a GPU; the resulting shared object travels to the
GPU box with the repository snapshot (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIBPATH = os.path.join(LIBDIR, "libimprover_b200.so")
SOURCES = ["nbh_capi.cu", "nbh_square.cu", "nbh_square_v1.cu", "nbh_square_v2.cu", "nbh_square_v4.cu", "nbh_square_w1.cu", "nbh_square_w2.cu", "nbh_square_w4.cu", "nbh_square_rs.cu", "nbh_circular.cu", "nbh_percentile.cu", "nbh_threshold.cu", "nbh_vicinity.cu", "nbh_recursive.cu", "nbh_tiling.cu", "nbh_square_box.cu"] + [f"nbh_square_box_r{r}.cu" for r in range(1, 9)] + ["nbh_square_mt.cu", "nbh_square_vh.cu"] + [f"nbh_square_mt_r{r}.cu" for r in range(1, 9)]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
]


def _nvcc() -> str:
    cand = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    return cand if os.path.exists(cand) else "nvcc"


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    headers = [os.path.join(CSRC, "nbh_common.cuh"), os.path.join(CSRC, "nbh_square.cuh"), os.path.join(CSRC, "nbh_square_warp.cuh"), os.path.join(CSRC, "nbh_async.cuh"), os.path.join(CSRC, "nbh_square_rs.cuh"), os.path.join(CSRC, "nbh_square_box.cuh"), os.path.join(CSRC, "nbh_square_box.cu"), os.path.join(CSRC, "nbh_square_mt.cuh"), os.path.join(CSRC, "nbh_square_mt.cu"), 
               os.path.join(HERE, "..", "include", "improver_b200.h")]
    objs = []
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            jobs.append(cmd)

    def run(cmd):
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        return res.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for out in ex.map(run, jobs):
                if verbose and out:
                    print(out, file=sys.stderr)
    if jobs or force or _stale(LIBPATH, objs):
        run([_nvcc(), "-shared", "-o", LIBPATH] + objs)
    return LIBPATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
