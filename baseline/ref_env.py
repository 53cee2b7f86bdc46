"""baseline/ref_env.py -- REFERENCE-ARM INFRASTRUCTURE: makes the UNMODIFIED reference importable.

The reference package is also called ``blockcopy``, so it can never live in the same process as
the product package; everything here is meant for scripts that run in their own interpreter: the
golden-vector generators (oracle/make_golden_cpu.py here, oracle/make_golden_gpu.py on the B200 box)
and the reference-arm timing script baseline/time_reference_gpu.py that bench.py spawns.  Nothing of
the product imports this file.

Two modes:
  * CPU  (no GPU): ``cupy`` is a stub exposing ``memoize`` (the only attribute the reference touches
    at import time, utils/cuda.py:6,25).  oracle/ref_env.py additionally rebinds the reference's four
    kernel entry points to the C oracle for the CPU fixture generator.
  * GPU  (B200 box): ``cupy`` is a shim over cuda-python (NVRTC + driver launch) that provides
    exactly ``memoize`` and ``cuda.compile_with_cache(code, options).get_function(name)(grid=,
    block=, args=, stream=)``, so the reference's own CUDA C strings are compiled for sm_100a
    and launched unchanged.

Reference location: /root/reference when present (the build container), else baseline/_ref (staged,
git-ignored copy that travels with gpurun).
"""
from __future__ import annotations

import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)


def reference_roots():
    """(blockcopy package parent dir, semantic_segmentation dir) or None if not available."""
    for base in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        bc = os.path.join(base, "blockcopy")
        ss = os.path.join(base, "semantic_segmentation")
        if os.path.isdir(os.path.join(bc, "blockcopy")) and os.path.isdir(ss):
            return bc, ss
    return None


def stage_reference() -> str | None:
    """Copy the reference files the GPU-side generator needs into baseline/_ref (git-ignored).
    Called from __graft_entry__.build() in the container where /root/reference exists."""
    import shutil

    src = "/root/reference"
    if not os.path.isdir(src):
        return None
    dst = os.path.join(ROOT, "baseline", "_ref")
    wanted = [
        ("blockcopy/blockcopy", "blockcopy/blockcopy"),
        ("semantic_segmentation/lib/__init__.py", "semantic_segmentation/lib/__init__.py"),
        ("semantic_segmentation/lib/models", "semantic_segmentation/lib/models"),
        ("semantic_segmentation/lib/utils/bn_fusion.py", "semantic_segmentation/lib/utils/bn_fusion.py"),
    ]
    for a, b in wanted:
        s, d = os.path.join(src, a), os.path.join(dst, b)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if os.path.isdir(s):
            shutil.copytree(s, d, dirs_exist_ok=True, ignore=shutil.ignore_patterns("__pycache__"))
        else:
            shutil.copy2(s, d)
    open(os.path.join(dst, "semantic_segmentation", "lib", "utils", "__init__.py"), "a").close()
    return dst


# --------------------------------------------------------------------------------------------- cupy stand-ins
def _install_cupy_stub():
    m = types.ModuleType("cupy")

    def memoize(for_each_device=False):
        return lambda f: f

    m.memoize = memoize
    sys.modules["cupy"] = m


def _install_cupy_nvrtc_shim():
    """cupy.cuda.compile_with_cache on top of cuda-python; enough for utils/cuda.py:25-31."""
    from cuda.bindings import driver, nvrtc
    import torch

    def _ok(res):
        err = res[0]
        if int(err) != 0:
            raise RuntimeError(f"CUDA/NVRTC error {err}")
        return res[1:] if len(res) > 2 else (res[1] if len(res) == 2 else None)

    class _Function:
        def __init__(self, fn):
            self.fn = fn

        def __call__(self, grid, block, args, stream=None, shared_mem=0):
            import ctypes
            import numpy as np

            grid = tuple(grid) + (1,) * (3 - len(grid))
            block = tuple(block) + (1,) * (3 - len(block))
            # Python ints: device pointers are 64-bit, everything else the reference passes is an int
            vals, types_ = [], []
            for a in args:
                if a > 0xFFFFFFFF or a < -0x80000000:
                    vals.append(ctypes.c_uint64(a)); types_.append(ctypes.c_void_p)
                else:
                    vals.append(ctypes.c_int32(a)); types_.append(ctypes.c_int)
            # reference argument lists are (pointers..., npixels): pointers may be small only if NULL
            n = len(args)
            for i in range(n - 1):
                vals[i] = ctypes.c_uint64(args[i]); types_[i] = ctypes.c_void_p
            ptrs = (ctypes.c_void_p * n)(*[ctypes.addressof(v) for v in vals])
            st = stream.ptr if stream is not None else 0
            _ok(driver.cuLaunchKernel(self.fn, grid[0], grid[1], grid[2], block[0], block[1], block[2],
                                      shared_mem, st, ctypes.addressof(ptrs), 0))

    class _Module:
        def __init__(self, code, options):
            torch.cuda.init()
            prog = _ok(nvrtc.nvrtcCreateProgram(code.encode(), b"k.cu", 0, [], []))
            major, minor = torch.cuda.get_device_capability()
            opts = [o.encode() for o in options if not o.startswith("-I") or os.path.isdir(o[2:])]
            opts.append(f"--gpu-architecture=sm_{major}{minor}a".encode() if major >= 9
                        else f"--gpu-architecture=sm_{major}{minor}".encode())
            res = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
            if int(res[0]) != 0:
                size = _ok(nvrtc.nvrtcGetProgramLogSize(prog))
                log = b" " * size
                nvrtc.nvrtcGetProgramLog(prog, log)
                raise RuntimeError("NVRTC failed:\n" + log.decode(errors="replace"))
            size = _ok(nvrtc.nvrtcGetCUBINSize(prog))
            cubin = b" " * size
            _ok(nvrtc.nvrtcGetCUBIN(prog, cubin))
            self.module = _ok(driver.cuModuleLoadData(cubin))

        def get_function(self, name):
            return _Function(_ok(driver.cuModuleGetFunction(self.module, name.encode())))

    cache = {}

    def compile_with_cache(code, options=()):
        key = (code, tuple(options))
        if key not in cache:
            cache[key] = _Module(code, options)
        return cache[key]

    def memoize(for_each_device=False):
        def deco(f):
            memo = {}

            def wrapper(*a, **kw):
                k = (a, tuple(sorted(kw.items())))
                if k not in memo:
                    memo[k] = f(*a, **kw)
                return memo[k]

            return wrapper

        return deco

    m = types.ModuleType("cupy")
    m.memoize = memoize
    m.cuda = types.ModuleType("cupy.cuda")
    m.cuda.compile_with_cache = compile_with_cache
    sys.modules["cupy"] = m
    sys.modules["cupy.cuda"] = m.cuda


# --------------------------------------------------------------------------------------------- import
def import_reference(mode: str):
    """mode in {'cpu', 'gpu'}; returns the reference ``blockcopy`` module."""
    roots = reference_roots()
    if roots is None:
        raise RuntimeError("reference not found (neither /root/reference nor baseline/_ref)")
    assert "blockcopy" not in sys.modules, "the product package is already imported in this process"
    bc_root, ss_root = roots
    if mode == "cpu":
        _install_cupy_stub()
    else:
        _install_cupy_nvrtc_shim()
    sys.path.insert(0, ss_root)
    sys.path.insert(0, bc_root)
    import warnings

    warnings.filterwarnings("ignore", message=".*__torch_function__.*")
    import blockcopy  # noqa: the reference

    assert os.path.realpath(blockcopy.__file__).startswith(os.path.realpath(bc_root)), blockcopy.__file__
    return blockcopy
