"""oracle/cpu_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front-end of oracle/libbc_oracle.so (the C restatement of the reference's
block kernels) on CPU torch tensors, plus fp32 torch restatements of the
floating-point steps of the path (padded op on a plane crop, per-tile bilinear,
information gain).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.

Pinning status (details in oracle/README.md):
  * grid mappings  -> pinned to the reference's get_grid_mappings (tests/golden/index_kat.json)
  * four kernels   -> pinned to the reference's own CUDA C run via NVRTC on a B200
                      (tests/golden/ref_kernels_*.npz, made by oracle/make_golden_gpu.py)
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    """Compile libbc_oracle.so with gcc (oracle/Makefile)."""
    so = os.path.join(_HERE, "libbc_oracle.so")
    src = os.path.join(_HERE, "bc_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libbc_oracle.so"])
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.oracle_grid_mappings.restype = ctypes.c_int
        _LIB.oracle_transfer_idx.restype = ctypes.c_int
    return _LIB


def _p(t: torch.Tensor) -> ctypes.c_void_p:
    assert t.device.type == "cpu" and t.is_contiguous(), (t.device, t.stride())
    return ctypes.c_void_p(t.data_ptr())


def _es(t: torch.Tensor) -> int:
    assert t.dtype in (torch.float16, torch.float32), t.dtype  # utils/cuda.py:17-23
    return t.element_size()


# ----------------------------------------------------------------------------- index tensors
def grid_mappings(grid: torch.Tensor):
    """grid bool (N,1,GH,GW) -> (grid_idx int32 like grid, mapping_exec int32 (E,)).
    Follows core/tensorwrapper.py:108-128."""
    g = grid.to("cpu", torch.uint8).contiguous()
    G = g.numel()
    grid_idx = torch.empty(g.shape, dtype=torch.int32)
    mapping = torch.empty(G, dtype=torch.int32)
    e = lib().oracle_grid_mappings(_p(g), G, _p(grid_idx), _p(mapping))
    return grid_idx, mapping[:e].clone()


def transfer_idx(grid: torch.Tensor, prev_grid_idx: torch.Tensor) -> torch.Tensor:
    """core/tensorwrapper.py:175-178."""
    g = grid.to("cpu", torch.uint8).contiguous()
    out = torch.empty(g.numel(), dtype=torch.int32)
    k = lib().oracle_transfer_idx(_p(g), g.numel(), _p(prev_grid_idx.contiguous()), _p(out))
    return out[:k].clone()


# ----------------------------------------------------------------------------- the four kernels
def split(image: torch.Tensor, mapping_exec: torch.Tensor, block_size: int) -> torch.Tensor:
    N, C, H, W = image.shape
    E = mapping_exec.numel()
    blocks = torch.empty((E, C, block_size, block_size), dtype=image.dtype)
    lib().oracle_split(_p(blocks), _p(image), _p(mapping_exec), E, N, C, H, W, block_size, _es(image))
    return blocks


def combine_(blocks: torch.Tensor, out: torch.Tensor, mapping_exec: torch.Tensor) -> torch.Tensor:
    N, C, H, W = out.shape
    E, _, BS, _ = blocks.shape
    lib().oracle_combine(_p(blocks), _p(out), _p(mapping_exec), mapping_exec.numel(), N, C, H, W, BS, _es(out))
    return out


def transfer(out: torch.Tensor, prev_exec: torch.Tensor, prev_transfer: torch.Tensor,
             transfer_map: torch.Tensor, G: int, padding: int) -> torch.Tensor:
    T, C, BS, _ = out.shape
    lib().oracle_transfer(_p(out), _p(prev_exec), _p(prev_transfer), _p(transfer_map),
                          transfer_map.numel(), G, C, BS, padding, _es(out))
    return out


def repad(features: torch.Tensor, transfer_t: torch.Tensor, grid_idx: torch.Tensor,
          mapping_exec: torch.Tensor, pad: int) -> torch.Tensor:
    N, _, GH, GW = grid_idx.shape
    E, C, BS, _ = features.shape
    out = torch.empty((E, C, BS + 2 * pad, BS + 2 * pad), dtype=features.dtype)
    lib().oracle_repad(_p(out), _p(features), _p(transfer_t), _p(grid_idx.contiguous()), _p(mapping_exec),
                       mapping_exec.numel(), N, C, GH, GW, BS, pad, _es(features))
    return out


def plane_halo(plane: torch.Tensor, mapping_exec: torch.Tensor, block_size: int, pad: int) -> torch.Tensor:
    N, C, H, W = plane.shape
    E = mapping_exec.numel()
    out = torch.empty((E, C, block_size + 2 * pad, block_size + 2 * pad), dtype=plane.dtype)
    lib().oracle_plane_halo(_p(out), _p(plane), _p(mapping_exec), E, N, C, H, W, block_size, pad, _es(plane))
    return out


# ----------------------------------------------------------------------------- stateful emulation
class RingProtocol:
    """Frame-to-frame emulation of ONE padded-op slot of the reference's FIFO
    protocol (core/tensorwrapper.py:180-209, :445-476, :529-575) on CPU with the
    oracle kernels: keeps (exec tiles, transfer tiles, grid_idx) of the previous
    frame, produces the padded tile batch of the current one.  Interiors of the
    transfer tensor are poisoned with NaN to prove they are never consumed."""

    def __init__(self):
        self.prev = None  # (exec_tiles, transfer_tiles, grid_idx)

    def step(self, tiles: torch.Tensor, grid: torch.Tensor, pad: int) -> torch.Tensor:
        grid_idx, mapping = grid_mappings(grid)
        G = grid.numel()
        E, C, BS, _ = tiles.shape
        if self.prev is None:
            assert E == G, "first frame must execute every block (tensorwrapper.py:164-165)"
            data_transfer = torch.empty((0, C, BS, BS), dtype=tiles.dtype)
        else:
            pe, pt, pgi = self.prev
            tmap = transfer_idx(grid, pgi)
            data_transfer = torch.full((tmap.numel(), C, BS, BS), float("nan"), dtype=tiles.dtype)
            transfer(data_transfer, pe, pt, tmap, G, pad)
        out = repad(tiles.contiguous(), data_transfer, grid_idx, mapping, pad)
        self.prev = (tiles.contiguous().clone(), data_transfer, grid_idx)
        return out


# ------------------------------------------------------------------------------------------------------------------
# Driver-side steps.  numpy restatements, pinned in tests/test_oracle_io.py against the very
# torchvision / torch CPU calls the reference's driver makes (ExtToTensor + ExtNormalize:
# semantic_segmentation/lib/ext_transforms.py:317-372; F.interpolate(..., mode='bilinear') + max(dim=1):
# test_swiftnet.py:196-197).
def frame_from_u8(src_u8: torch.Tensor, mean, std, dtype=torch.float16) -> torch.Tensor:
    """(N,H,W,3) uint8 -> (N,3,H,W) `dtype`: ((x / 255) - mean) / std in fp32 (IEEE), one final rounding."""
    import numpy as np

    x = src_u8.numpy().astype(np.float32) / np.float32(255.0)
    m = np.asarray(mean, dtype=np.float32).reshape(1, 1, 1, 3)
    s = np.asarray(std, dtype=np.float32).reshape(1, 1, 1, 3)
    y = ((x - m) / s).astype(np.float32).transpose(0, 3, 1, 2)
    return torch.from_numpy(np.ascontiguousarray(y)).to(dtype)


def upsample_argmax(logits: torch.Tensor, scale: int) -> torch.Tensor:
    """(N,K,h,w) fp16/fp32 -> (N, scale*h, scale*w) int64: argmax over K of the bilinear (align_corners=False)
    upsampling computed in fp32 as ATen does -- src = (dst + 0.5) / scale - 0.5 clamped at 0, taps (i0, min(i0+1,
    last)), value = h0*(w0*a + w1*b) + h1*(w0*c + w1*d) -- rounded to the logits' dtype; ties -> lowest class."""
    import numpy as np

    x = logits.float().numpy()
    N, K, h, w = x.shape

    def taps(n_in):
        dst = np.arange(n_in * scale, dtype=np.float32)
        src = np.float32(1.0 / scale) * (dst + np.float32(0.5)) - np.float32(0.5)
        src = np.maximum(src, np.float32(0.0)).astype(np.float32)
        i0 = src.astype(np.int64)
        i1 = np.minimum(i0 + 1, n_in - 1)
        l1 = (src - i0.astype(np.float32)).astype(np.float32)
        return i0, i1, l1, (np.float32(1.0) - l1).astype(np.float32)

    y0, y1, hy1, hy0 = taps(h)
    x0, x1, wx1, wx0 = taps(w)
    best = np.full((N, h * scale, w * scale), -np.inf, dtype=np.float32)
    arg = np.zeros((N, h * scale, w * scale), dtype=np.int64)
    for c in range(K):
        p = x[:, c]
        top = (wx0 * p[:, y0][:, :, x0] + wx1 * p[:, y0][:, :, x1]).astype(np.float32)
        bot = (wx0 * p[:, y1][:, :, x0] + wx1 * p[:, y1][:, :, x1]).astype(np.float32)
        v = (hy0[None, :, None] * top + hy1[None, :, None] * bot).astype(np.float32)
        if logits.dtype == torch.float16:
            v = v.astype(np.float16).astype(np.float32)
        upd = v > best
        best[upd] = v[upd]
        arg[upd] = c
    return torch.from_numpy(arg)
