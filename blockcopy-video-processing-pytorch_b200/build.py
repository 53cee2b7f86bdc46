"""Builds libblockcopy_sm100.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python blockcopy-video-processing-pytorch_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU.  The shared object lands in
blockcopy/_lib/ next to the Python package so that it travels to the GPU box with the tree
(it is git-ignored, not gpurun-ignored).  The CUDA runtime is linked statically; the driver API
(cuTensorMapEncodeTiled) is resolved at run time through cudaGetDriverEntryPoint, so the
library loads -- and exports all of include/blockcopy_b200.h -- on a machine without libcuda.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "blockcopy", "_lib")
OUT = os.path.join(OUT_DIR, "libblockcopy_sm100.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-shared", "-cudart", "static",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libblockcopy_sm100.so (there is no CPU fallback)")
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-I", CSRC]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", OUT] + sources()
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libblockcopy_sm100.so")
    return OUT


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
