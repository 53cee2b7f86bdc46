"""ctypes binding of libblockcopy_sm100.so (include/blockcopy_b200.h).

This is the only place where the Python package touches native code.  There is no fallback:
if the library is missing or a call fails, an exception is raised (the reference behaves the
same way -- its block path asserts CUDA tensors, core/tensorwrapper.py:35).

Every wrapper launches on torch's current stream, exactly as the reference does
(utils/block_funcs.py:48), never synchronises and never allocates.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libblockcopy_sm100.so")

BC_F16, BC_F32 = 0, 1
BC_NCHW, BC_NHWC = 0, 1

_lib: Optional[ctypes.CDLL] = None

_vp, _i, _ip = ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p
_SIGNATURES = {
    "bc_version": ([], _i),
    "bc_last_error_string": ([], ctypes.c_char_p),
    "bc_build_info": ([], ctypes.c_char_p),
    "bc_set_tma_enabled": ([_i], _i),
    "bc_compact_mask": ([_vp, _i, _ip, _ip, _ip, _ip, _ip, _vp], _i),
    "bc_gather": ([_vp, _vp, _ip, _i, _i, _i, _i, _i, _i, _i, _i, _vp], _i),
    "bc_scatter": ([_vp, _vp, _ip, _i, _i, _i, _i, _i, _i, _i, _i, _vp], _i),
    "bc_copy_blocks": ([_vp, _vp, _vp, _ip, _i, _i, _i, _i, _i, _i, _i, _vp], _i),
    "bc_transfer": ([_vp, _vp, _vp, _ip, _i, _i, _i, _i, _i, _i, _i, _vp], _i),
    "bc_gather_halo_tiles": ([_vp, _vp, _vp, _ip, _ip, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp], _i),
    "bc_gather_halo": ([_vp, _vp, _ip, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp], _i),
    "bc_conv_igemm": ([_vp, _vp, _vp, _vp, _vp, _ip] + [_i] * 11 + [_vp, _ip, _i, _i, _i, _i, _vp, ctypes.c_longlong, _vp],
                      _i),
    "bc_ew_fused": ([_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _ip] + [_i] * 8 + [_vp], _i),
    "bc_maxpool_halo": ([_vp, _vp, _vp, _ip] + [_i] * 9 + [_vp], _i),
    "bc_debug_trace": ([_vp], _i),
    "bc_head_1x1": ([_vp] * 10 + [_i, _ip, _ip] + [_i] * 9 + [_vp], _i),
    "bc_stem_pack": ([_vp, _vp, _ip] + [_i] * 5 + [_vp], _i),
    "bc_conv_stem": ([_vp, _vp, _vp, _vp, _ip] + [_i] * 7 + [_vp, _vp], _i),
    "bc_policy_features": ([_vp, _vp, _vp, _vp, _vp] + [_i] * 10 + [_vp, ctypes.c_float, ctypes.c_float, _i, _vp], _i),
    "bc_policy_features_nhwc16": ([_vp, _i, _vp, _vp, _vp, _vp] + [_i] * 10 + [_vp, ctypes.c_float, ctypes.c_float, _i, _vp], _i),
    "bc_info_gain": ([_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp], _i),
    "bc_bn_update_running": ([_vp, _i, _vp], _i),
    "bc_gn_stats": ([_vp, _vp, _vp, ctypes.c_longlong, _i, _i, ctypes.c_float, _vp, ctypes.c_longlong, _vp], _i),
    "bc_sample_grid": ([_vp, _vp, _vp, _vp, _i, _i, _i, _vp], _i),
    "bc_raster_boxes": ([_vp, _vp, _vp, _i, _i, _i, _i, _vp], _i),
    "bc_depth_to_space": ([_vp, _vp, _i, _i, _i, _i, _i, _vp], _i),
    "bc_bwd_mask_add": ([_vp, _vp, _vp, _vp, ctypes.c_longlong, _vp], _i),
    "bc_bn_bwd_reduce": ([_vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_longlong, _i, _vp, ctypes.c_longlong, _vp], _i),
    "bc_bn_bwd_apply": ([_vp] * 9 + [_i] * 4 + [_vp], _i),
    "bc_conv_wgrad": ([_vp] * 4 + [_i] * 9 + [_vp] * 5 + [ctypes.c_longlong, _vp], _i),
    "bc_bn_stats": ([_vp, _vp, _vp, ctypes.c_longlong, _i, ctypes.c_float, _vp, ctypes.c_longlong, _vp], _i),
    "bc_pack_params": ([_vp, _i, ctypes.c_longlong, _vp], _i),
    "bc_graph_record": ([_i], _i),
    "bc_graph_last_node": ([_vp, _vp], _i),
    "bc_graph_patch_next": ([_vp, _vp, _vp], _i),
    "bc_graph_memcpy": ([_vp, _vp, ctypes.c_longlong, _vp, _vp], _i),
    "bc_graph_patch_memcpy": ([_vp, _vp, _vp, _vp, ctypes.c_longlong], _i),
    "bc_bn_norm": ([_vp] * 6 + [ctypes.c_longlong, _i, ctypes.c_float, _i, _vp, ctypes.c_longlong, _vp], _i),
    "bc_rmsprop_step": ([_vp, _i, ctypes.c_longlong] + [ctypes.c_float] * 5 + [_vp], _i),
    "bc_conv_fewout": ([_vp, _vp, _vp, _vp] + [_i] * 9 + [_vp, _vp], _i),
    "bc_frame_from_u8": ([_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp], _i),
    "bc_blocks_from_u8": ([_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp], _i),
    "bc_upsample_argmax": ([_vp, _vp, _i, _i, _i, _i, _vp, _i, _i, _i, _vp], _i),
    "bc_upsample_argmax_blocks": ([_vp, _vp, _vp, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _vp], _i),
    "bc_spp_pool": ([_vp, _vp] + [_i] * 5 + [_vp, _vp, _vp], _i),
    "bc_spp_levels": ([_vp, _vp, _vp, _vp] + [_i] * 5 + [_vp, _vp, _i, _vp], _i),
    "bc_spp_prep": ([_vp, _vp, _vp, _vp] + [_i] * 5 + [_vp, _vp, _i, _i, _vp], _i),
}


class BlockCopyNativeError(RuntimeError):
    """A libblockcopy_sm100 entry point returned a non-zero status."""


def exported_symbols():
    """Names every build of the library must export (checked by the CPU test-suite)."""
    return sorted(_SIGNATURES)


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing. Build it with "
                "`python blockcopy-video-processing-pytorch_b200/build.py` (nvcc, sm_100a). "
                "blockcopy has no CPU / eager fallback."
            )
        l = ctypes.CDLL(LIB_PATH)
        for name, (argtypes, restype) in _SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError if the build is stale
            fn.argtypes = argtypes
            fn.restype = restype
        if l.bc_version() != 1:
            raise ImportError(f"ABI mismatch: library reports version {l.bc_version()}, binding expects 1")
        _lib = l
    return _lib


_n_launches = 0  # kernels handed to the GPU through this binding (bench.py: "gpu_launches")


def launch_count() -> int:
    return _n_launches


def add_launches(n: int):
    """Account for kernels replayed from a captured CUDA graph (no Python call per launch)."""
    global _n_launches
    _n_launches += n


def _check(rc: int, what: str):
    global _n_launches
    if rc == 0:
        _n_launches += 1
        return
    if rc != 0:
        msg = lib().bc_last_error_string().decode()
        if rc == -2:
            # the reference raises AttributeError for shapes that are not divisible by the block size
            # (core/tensorwrapper.py:351-354)
            raise AttributeError(f"{what}: {msg}")
        raise BlockCopyNativeError(f"{what} failed with status {rc}: {msg}")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _dtype(t: torch.Tensor) -> int:
    es = t.element_size()
    if t.dtype in (torch.float16, torch.bfloat16) or (es == 2 and not t.dtype.is_complex):
        return BC_F16
    if es == 4:
        return BC_F32
    raise NotImplementedError(t.dtype)  # utils/cuda.py:17-23 raises the same for other dtypes


def layout_of(t: torch.Tensor) -> int:
    """BC_NHWC for channels_last-dense 4-D tensors, BC_NCHW for contiguous ones."""
    if t.is_contiguous():
        # a (N,1,H,W) / (N,C,1,1) tensor is both; NCHW is the reference's reading
        return BC_NCHW
    if t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last):
        return BC_NHWC
    raise AssertionError(f"tensor must be dense NCHW or channels_last, got strides {t.stride()}")


def _dense_in(t: torch.Tensor, lay: int) -> bool:
    """Is `t` dense in layout `lay`?  (1x1 tiles and 1-channel tensors are dense in both.)"""
    return t.is_contiguous() if lay == BC_NCHW else t.is_contiguous(memory_format=torch.channels_last)


def set_tma_enabled(flag: bool):
    _check(lib().bc_set_tma_enabled(int(flag)), "bc_set_tma_enabled")


def build_info() -> str:
    return lib().bc_build_info().decode()


def _dev(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise AssertionError("blockcopy kernels need CUDA tensors (there is no CPU fallback)")


# ------------------------------------------------------------------------------------------- index
def compact_mask(grid_u8: torch.Tensor, grid_idx: torch.Tensor, mapping_exec: torch.Tensor,
                 counts: torch.Tensor, prev_grid_idx: Optional[torch.Tensor] = None,
                 transfer_idx: Optional[torch.Tensor] = None):
    _dev(grid_u8, grid_idx, mapping_exec, counts)
    assert grid_u8.dtype in (torch.uint8, torch.bool) and grid_u8.is_contiguous()
    G = grid_u8.numel()
    assert grid_idx.numel() >= G and mapping_exec.numel() >= G and counts.numel() >= 2
    _check(lib().bc_compact_mask(grid_u8.data_ptr(), G, grid_idx.data_ptr(), mapping_exec.data_ptr(),
                                 counts.data_ptr(),
                                 prev_grid_idx.data_ptr() if prev_grid_idx is not None else None,
                                 transfer_idx.data_ptr() if transfer_idx is not None else None, _stream()),
           "bc_compact_mask")


# ------------------------------------------------------------------------------------------- movement
def gather(blocks: torch.Tensor, image: torch.Tensor, mapping_exec: torch.Tensor, E: int):
    _dev(blocks, image, mapping_exec)
    N, C, H, W = image.shape
    BS = blocks.shape[-1]
    lay = layout_of(image)
    assert E == 0 or _dense_in(blocks, lay), "tiles and plane must share a layout"
    _check(lib().bc_gather(blocks.data_ptr(), image.data_ptr(), mapping_exec.data_ptr(), E, N, C, H, W, BS,
                           _dtype(image), lay, _stream()), "bc_gather")
    return blocks


def scatter(blocks: torch.Tensor, image: torch.Tensor, mapping_exec: torch.Tensor, E: int):
    _dev(blocks, image, mapping_exec)
    N, C, H, W = image.shape
    BS = blocks.shape[-1]
    lay = layout_of(image)
    assert E == 0 or _dense_in(blocks, lay), "tiles and plane must share a layout"
    _check(lib().bc_scatter(blocks.data_ptr(), image.data_ptr(), mapping_exec.data_ptr(), E, N, C, H, W, BS,
                            _dtype(image), lay, _stream()), "bc_scatter")
    return image


def copy_blocks(out: torch.Tensor, prev: torch.Tensor, blocks: torch.Tensor, grid_idx: torch.Tensor):
    _dev(out, prev, blocks, grid_idx)
    N, C, H, W = out.shape
    BS = blocks.shape[-1]
    lay = layout_of(out)
    assert _dense_in(prev, lay) and prev.shape == out.shape and (blocks.numel() == 0 or _dense_in(blocks, lay))
    _check(lib().bc_copy_blocks(out.data_ptr(), prev.data_ptr(), blocks.data_ptr() if blocks.numel() else None,
                                grid_idx.data_ptr(), N, C, H, W, BS, _dtype(out), lay, _stream()),
           "bc_copy_blocks")
    return out


def transfer(out: torch.Tensor, prev_exec: torch.Tensor, prev_transfer: torch.Tensor,
             transfer_idx: torch.Tensor, G: int, padding: int):
    _dev(out, prev_exec, prev_transfer, transfer_idx)
    T, C, BS, _ = out.shape
    lay = layout_of(out) if T > 0 else BC_NCHW
    _check(lib().bc_transfer(out.data_ptr(), prev_exec.data_ptr(),
                             prev_transfer.data_ptr() if prev_transfer.numel() else None,
                             transfer_idx.data_ptr(), transfer_idx.numel(), G, C, BS, padding, _dtype(out), lay,
                             _stream()), "bc_transfer")
    return out


def gather_halo_tiles(out: torch.Tensor, exec_t: torch.Tensor, transfer_t: torch.Tensor, grid_idx: torch.Tensor,
                      mapping_exec: torch.Tensor, E: int, pad: int):
    _dev(out, exec_t, transfer_t, grid_idx, mapping_exec)
    N, _, GH, GW = grid_idx.shape
    _, C, BS, _ = exec_t.shape
    lay = layout_of(out) if E > 0 else BC_NCHW  # the padded output is never layout-ambiguous (edge >= 3)
    assert E == 0 or _dense_in(exec_t, lay)
    _check(lib().bc_gather_halo_tiles(out.data_ptr(), exec_t.data_ptr(),
                                      transfer_t.data_ptr() if transfer_t.numel() else None,
                                      grid_idx.data_ptr(), mapping_exec.data_ptr(), E, N, C, GH, GW, BS, pad,
                                      _dtype(out), lay, _stream()), "bc_gather_halo_tiles")
    return out


def gather_halo(out: torch.Tensor, plane: torch.Tensor, mapping_exec: torch.Tensor, E: int, BS: int, pad: int):
    _dev(out, plane, mapping_exec)
    N, C, H, W = plane.shape
    lay = layout_of(plane)
    assert E == 0 or _dense_in(out, lay)
    _check(lib().bc_gather_halo(out.data_ptr(), plane.data_ptr(), mapping_exec.data_ptr(), E, N, C, H, W, BS, pad,
                                _dtype(plane), lay, _stream()), "bc_gather_halo")
    return out


# ------------------------------------------------------------------------------------------- convolution
def conv_supported(dtype, weight: torch.Tensor, BS_in: int, stride: int, padding: int, dilation: int = 1,
                   groups: int = 1) -> bool:
    """True when bc_conv_igemm covers this conv: fp16, 1x1 (pad 0) or 3x3 with pad == dilation in 1..4 (dilation > 1
    only with stride 1), stride 1/2, groups 1, Cin and Cout multiples of 64, output block edge a power of two in
    [2,128]."""
    Cout, Cin, kh, kw = weight.shape
    if dtype != torch.float16 or weight.dtype != torch.float16 or not weight.is_cuda:
        return False
    if kh != kw or kh not in (1, 3) or stride not in (1, 2) or groups != 1:
        return False
    if kh == 1 and (padding != 0 or dilation != 1):
        return False
    if kh == 3 and (padding != dilation or not 1 <= dilation <= 4 or (dilation > 1 and stride != 1)):
        return False
    if Cin % 64 or Cout % 64 or BS_in % stride:
        return False
    bo = BS_in // stride
    return 2 <= bo <= 128 and (bo & (bo - 1)) == 0


def lazy_supported(x: torch.Tensor) -> bool:
    """Can bc_ew_fused handle this block tensor (packed fp16 NHWC tiles on the GPU, C % 8 == 0)?"""
    return x.is_cuda and x.dtype == torch.float16 and x.dim() == 4 and x.shape[1] % 8 == 0


_SPLITK_WS = {}
SPLITK_WS_BYTES = 32 << 20
class _SplitKOwner(__import__("threading").local):
    ws = None  # tensor handed over by splitk_workspace_scope on THIS thread, or None


_SPLITK = _SplitKOwner()


def _splitk_workspace(device: torch.device, stream: int) -> torch.Tensor:
    """Scratch for the split-K partial sums of bc_conv_igemm.  Launches on one stream are ordered, so they may
    share one scratch per (device, stream); concurrent streams must not.  A captured CUDA graph bakes the
    scratch address in and may later be replayed on ANY stream, next to other graphs captured on the same
    capture stream: graph owners therefore bring their own scratch (splitk_workspace_scope)."""
    if _SPLITK.ws is not None and _SPLITK.ws.device == device:
        return _SPLITK.ws
    key = (device.index, stream)
    ws = _SPLITK_WS.get(key)
    if ws is None:
        ws = _SPLITK_WS[key] = torch.empty(SPLITK_WS_BYTES, dtype=torch.uint8, device=device)
    return ws


class splitk_workspace_scope:
    """`with splitk_workspace_scope(ws):` -- every bc_conv_igemm issued inside uses the uint8 tensor `ws` as its
    split-K scratch instead of the per-stream one (BlockCopyModel in CUDA-graph mode: one scratch per model)."""

    def __init__(self, ws: Optional[torch.Tensor]):
        self.ws, self.prev = ws, None

    def __enter__(self):
        self.prev, _SPLITK.ws = _SPLITK.ws, self.ws
        return self.ws

    def __exit__(self, *exc):
        _SPLITK.ws = self.prev
        return False


def conv_igemm(out: torch.Tensor, plane: torch.Tensor, weight_cl: torch.Tensor, bias: Optional[torch.Tensor],
               residual: Optional[torch.Tensor], mapping_exec: Optional[torch.Tensor], E: int, BS_in: int,
               stride: int, padding: int, relu: bool = False, plane_out: Optional[torch.Tensor] = None,
               out_mapping: Optional[torch.Tensor] = None, split_k: bool = True, write_tiles: bool = True):
    """out (E,Cout,BS_out,BS_out) channels_last <- conv(plane (N,Cin,H,W) channels_last) on the E executed
    blocks (+bias, +residual, ReLU).  weight_cl: channels_last (Cout,Cin,k,k) fp16.  plane_out: the next padded
    op's persistent plane (N,Cout,GH*BS_out,GW*BS_out) channels_last, written in the same epilogue.
    write_tiles=False (needs plane_out): `out` only describes the shape, the tile batch is not written."""
    _dev(out, plane, weight_cl, bias, residual, mapping_exec, plane_out, out_mapping)
    N, Cin, H, W = plane.shape
    Cout, _, k, _ = weight_cl.shape
    assert weight_cl.is_contiguous(memory_format=torch.channels_last) or k == 1
    oN = oGH = oGW = 0
    if plane_out is not None:
        BSo = BS_in // stride
        oN, _, oH, oW = plane_out.shape
        oGH, oGW = oH // BSo, oW // BSo
        assert plane_out.is_contiguous(memory_format=torch.channels_last) and plane_out.shape[1] == Cout
        if out_mapping is None:
            out_mapping = mapping_exec
    stream = _stream()
    ws = _splitk_workspace(out.device, int(stream)) if split_k and split_k != "dsmem" else None
    assert write_tiles or plane_out is not None
    _check(lib().bc_conv_igemm(out.data_ptr() if write_tiles else None, plane.data_ptr(), weight_cl.data_ptr(),
                               bias.data_ptr() if bias is not None else None,
                               residual.data_ptr() if residual is not None else None,
                               mapping_exec.data_ptr() if mapping_exec is not None else None,
                               E, N, Cin, H, W, BS_in, Cout, k, stride, padding, int(relu),
                               plane_out.data_ptr() if plane_out is not None else None,
                               out_mapping.data_ptr() if out_mapping is not None else None, oN, oGH, oGW,
                               int(bool(split_k)), ws.data_ptr() if ws is not None else None,
                               ws.numel() if ws is not None else 0, stream),
           "bc_conv_igemm")
    return out


def ew_fused(out: Optional[torch.Tensor], a: torch.Tensor, residual: Optional[torch.Tensor] = None, bn=None,
             relu: bool = False, up2x: bool = False, plane_out: Optional[torch.Tensor] = None,
             mapping_exec: Optional[torch.Tensor] = None):
    """y = relu?(bn?(up2x?(a) + residual?)) on packed channels_last fp16 tiles, written to `out` and/or scattered
    into `plane_out`.  bn = (mean, invstd, weight|None, shift|None), fp32 [C]."""
    _dev(out, a, residual, plane_out, mapping_exec)
    E, C, BSa, Wa = a.shape
    if not up2x and plane_out is None:
        E, BSa = E * BSa * Wa, 1  # purely per-pixel: any dense (N,C,H,W) channels_last tensor
    else:
        assert BSa == Wa, "tiles must be square"
    BS = BSa * 2 if up2x else BSa
    N = H = W = 0
    if plane_out is not None:
        N, _, H, W = plane_out.shape
        assert plane_out.is_contiguous(memory_format=torch.channels_last) and plane_out.shape[1] == C
    mean = invstd = weight = shift = None
    if bn is not None:
        mean, invstd, weight, shift = bn
    ptr = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
    _check(lib().bc_ew_fused(ptr(out), ptr(plane_out), a.data_ptr(), ptr(residual), ptr(mean), ptr(invstd), ptr(weight),
                             ptr(shift), ptr(mapping_exec), E, C, BS, N, H, W, int(up2x), int(relu), _stream()),
           "bc_ew_fused")
    return out


def maxpool_halo(out: torch.Tensor, plane: torch.Tensor, mapping_exec: torch.Tensor, E: int, BS_in: int, k: int,
                 stride: int, padding: int, plane_out: Optional[torch.Tensor] = None):
    """out (E,C,BS_in/stride,..) channels_last <- max-pool of the E executed blocks of `plane` (halo from the
    neighbouring cells, zeros outside the frame); optionally also scattered into `plane_out`."""
    _dev(out, plane, mapping_exec, plane_out)
    N, C, H, W = plane.shape
    assert plane.is_contiguous(memory_format=torch.channels_last) and out.is_contiguous(memory_format=torch.channels_last)
    _check(lib().bc_maxpool_halo(out.data_ptr(), plane_out.data_ptr() if plane_out is not None else None,
                                 plane.data_ptr(), mapping_exec.data_ptr(), E, N, C, H, W, BS_in, k, stride, padding,
                                 _stream()), "bc_maxpool_halo")
    return out


# ------------------------------------------------------------------------------------------- CUDA-graph node patching
def graph_record(on: bool):
    """While on, every launch of this library on a CAPTURING stream remembers the graph node it created
    (graph_last_node); see bc_graph_record."""
    _rc(lib().bc_graph_record(int(bool(on))), "bc_graph_record")


def graph_last_node():
    """(node handle, kernel handle) of the last recorded launch."""
    node, func = ctypes.c_void_p(), ctypes.c_void_p()
    _rc(lib().bc_graph_last_node(ctypes.byref(node), ctypes.byref(func)), "bc_graph_last_node")
    return node.value, func.value


def graph_patch_next(graph_exec: int, node, func):
    """The NEXT launch of this library on this thread re-points `node` of the instantiated graph instead of running."""
    _rc(lib().bc_graph_patch_next(graph_exec, node, func), "bc_graph_patch_next")


def graph_memcpy(dst: torch.Tensor, src: torch.Tensor):
    """dst <- src (same dense layout) as a device-to-device copy on the current stream; returns the memcpy node's handle
    when the stream is capturing, else None."""
    _dev(dst, src)
    assert dst.shape == src.shape and dst.dtype == src.dtype and dst.stride() == src.stride()
    node = ctypes.c_void_p()
    _rc(lib().bc_graph_memcpy(dst.data_ptr(), src.data_ptr(), src.numel() * src.element_size(), _stream(), ctypes.byref(node)),
        "bc_graph_memcpy")
    return node.value


def graph_patch_memcpy(graph_exec: int, node, dst: torch.Tensor, src: torch.Tensor):
    _rc(lib().bc_graph_patch_memcpy(graph_exec, node, dst.data_ptr(), src.data_ptr(), src.numel() * src.element_size()),
        "bc_graph_patch_memcpy")


def _rc(rc: int, what: str):
    if rc != 0:
        raise BlockCopyNativeError(f"{what} failed with status {rc}: {lib().bc_last_error_string().decode()}")


# ------------------------------------------------------------------------------------------- derived-parameter cache
class TensorCache:
    """Values derived from parameter tensors (packed weights, folded batch-norm vectors), keyed by the
    IDENTITY of the source tensors plus their in-place version counters.  Entries keep the sources alive,
    so neither an id() nor a data_ptr() can be recycled by another tensor while the entry exists."""

    def __init__(self, capacity: int):
        self._d, self._cap = {}, capacity

    @staticmethod
    def _versions(tensors):
        return tuple(None if t is None else t._version for t in tensors)

    def get(self, tensors, extra=()):
        e = self._d.get(tuple(id(t) for t in tensors) + tuple(extra))
        if e is not None and all(a is b for a, b in zip(e[0], tensors)) and e[1] == self._versions(tensors):
            return e[2]
        return None

    def put(self, tensors, extra, value):
        if len(self._d) >= self._cap:
            self._d.clear()
        self._d[tuple(id(t) for t in tensors) + tuple(extra)] = (tuple(tensors), self._versions(tensors), value)
        return value


# ------------------------------------------------------------------------------------------- stem (7x7 s2, 3 ch)
_STEM_W = TensorCache(64)


def pack_stem_weight(w: torch.Tensor) -> torch.Tensor:
    """(Cout,3,7,7) -> (Cout,256): the 7x7 stride-2 kernel as a zero-extended 4x4 kernel over the
    space-to-depth(2) image with 16 (12 used) channels; cached per weight version."""
    hit = _STEM_W.get((w,))
    if hit is None:
        Cout = w.shape[0]
        wd = w.detach()
        wp = torch.zeros(Cout, 4, 4, 16, dtype=w.dtype, device=w.device)
        for dy in range(2):
            for dx in range(2):
                ch = (dy * 2 + dx) * 3
                for kh in range(4):
                    t = 2 * kh + dy - 1
                    if not 0 <= t < 7:
                        continue
                    for kw in range(4):
                        u = 2 * kw + dx - 1
                        if 0 <= u < 7:
                            wp[:, kh, kw, ch:ch + 3] = wd[:, :, t, u]
        hit = _STEM_W.put((w,), (), wp.reshape(Cout, 256).contiguous())
    return hit


def stem_supported(dtype, weight: torch.Tensor, BS_in: int, stride, padding, dilation=1, groups=1) -> bool:
    Cout, Cin, kh, kw = weight.shape
    bo = BS_in // 2
    return (dtype == torch.float16 and weight.dtype == torch.float16 and weight.is_cuda and Cin == 3 and kh == kw == 7
            and stride == 2 and padding == 3 and dilation == 1 and groups == 1 and Cout % 64 == 0 and BS_in % 2 == 0
            and 16 <= bo <= 128 and (bo & (bo - 1)) == 0)


STEM_XPAD = 2  # BC_STEM_XPAD of include/blockcopy_b200.h


def stem_plane(N: int, Hs: int, Ws: int, dtype, device) -> torch.Tensor:
    """Zero-filled space-to-depth plane (N,16,Hs,Ws+2*STEM_XPAD) channels_last; the pad columns stay zero."""
    return torch.zeros((N, 16, Hs, Ws + 2 * STEM_XPAD), dtype=dtype, device=device).contiguous(
        memory_format=torch.channels_last)


def stem_pack(s2d_plane: torch.Tensor, tiles: torch.Tensor, mapping_exec: torch.Tensor, E: int):
    """Executed NCHW input tiles (E,3,BS,BS) -> cells of the space-to-depth plane (see stem_plane)."""
    _dev(s2d_plane, tiles, mapping_exec)
    N, C16, Hs, Ws = s2d_plane.shape
    Ws -= 2 * STEM_XPAD
    assert C16 == 16 and s2d_plane.is_contiguous(memory_format=torch.channels_last) and tiles.is_contiguous()
    assert tiles.shape[1] == 3
    _check(lib().bc_stem_pack(s2d_plane.data_ptr(), tiles.data_ptr(), mapping_exec.data_ptr(), E, N, Hs * 2, Ws * 2,
                              tiles.shape[-1], _stream()), "bc_stem_pack")
    return s2d_plane


def conv_stem(out: torch.Tensor, s2d_plane: torch.Tensor, weight_packed: torch.Tensor, bias: Optional[torch.Tensor],
              mapping_exec: torch.Tensor, E: int, relu: bool = False, plane_out: Optional[torch.Tensor] = None,
              write_tiles: bool = True):
    _dev(out, s2d_plane, weight_packed, bias, mapping_exec, plane_out)
    assert write_tiles or plane_out is not None
    N, _, Hs, Ws = s2d_plane.shape
    Ws -= 2 * STEM_XPAD
    Cout, BSo = out.shape[1], out.shape[-1]
    assert out.is_contiguous(memory_format=torch.channels_last) and weight_packed.shape == (Cout, 256)
    _check(lib().bc_conv_stem(out.data_ptr() if write_tiles else None, s2d_plane.data_ptr(), weight_packed.data_ptr(),
                              bias.data_ptr() if bias is not None else None, mapping_exec.data_ptr(), E, N, Hs, Ws, BSo,
                              Cout, int(relu), plane_out.data_ptr() if plane_out is not None else None, _stream()),
           "bc_conv_stem")
    return out


# ------------------------------------------------------------------------------------------- output head
HEAD_MAX_COUT = 32


def head_supported(dtype, weight: torch.Tensor, stride, padding, dilation=1, groups=1) -> bool:
    """1x1 conv with few output channels (class logits) on fp16 CUDA blocks: bc_head_1x1."""
    if weight.dim() != 4:
        return False
    Cout, Cin, kh, kw = weight.shape
    return (dtype == torch.float16 and weight.dtype == torch.float16 and weight.is_cuda and kh == kw == 1
            and stride == 1 and padding == 0 and dilation == 1 and groups == 1 and Cout <= HEAD_MAX_COUT
            and Cin % 8 == 0 and Cin <= 1024)


def head_1x1(tiles_in: torch.Tensor, weight2d: torch.Tensor, bias: Optional[torch.Tensor], bn, relu_in: bool,
             tiles_out: Optional[torch.Tensor] = None, dense_out: Optional[torch.Tensor] = None,
             dense_prev: Optional[torch.Tensor] = None, grid_idx: Optional[torch.Tensor] = None,
             mapping_exec: Optional[torch.Tensor] = None):
    """y = conv1x1(relu?(bn?(tiles_in))) + bias on channels_last tiles (E,Cin,BS,BS); y goes to `tiles_out`
    (E,Cout,BS,BS) and/or, combined with `dense_prev`, to `dense_out` (N,Cout,GH*BS,GW*BS).  bn = (mean, invstd,
    weight|None, shift|None) fp32 or None; weight2d fp16 (Cout,Cin) contiguous."""
    _dev(tiles_in, weight2d, bias, tiles_out, dense_out, dense_prev, grid_idx, mapping_exec)
    E, Cin, BS, _ = tiles_in.shape
    Cout = weight2d.shape[0]
    assert tiles_in.is_contiguous(memory_format=torch.channels_last) and weight2d.is_contiguous()
    assert weight2d.shape == (Cout, Cin)
    N = GH = GW = 1
    dl = BC_NCHW
    if dense_out is not None:
        N, _, GH, GW = grid_idx.shape
        assert tuple(dense_out.shape) == (N, Cout, GH * BS, GW * BS)
        dl = layout_of(dense_out)
        if dense_prev is not None:
            assert dense_prev.shape == dense_out.shape and layout_of(dense_prev) == dl \
                and dense_prev.data_ptr() != dense_out.data_ptr()
    tl = layout_of(tiles_out) if tiles_out is not None else BC_NCHW
    mean, invstd, w, sh = bn if bn is not None else (None, None, None, None)
    ptr = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
    _check(lib().bc_head_1x1(ptr(tiles_out), ptr(dense_out), ptr(dense_prev), tiles_in.data_ptr(), weight2d.data_ptr(),
                             ptr(bias), ptr(mean), ptr(invstd), ptr(w), ptr(sh), int(relu_in), ptr(grid_idx),
                             ptr(mapping_exec), E, N, GH, GW, BS, Cin, Cout, tl, dl, _stream()), "bc_head_1x1")
    return dense_out if dense_out is not None else tiles_out


# ------------------------------------------------------------------------------------------- policy features
def policy_features(frame: torch.Tensor, frame_state: torch.Tensor, output_repr: torch.Tensor, grid: torch.Tensor,
                    scale_factor: float) -> torch.Tensor:
    """(N, 3+3+K+1, H*s, W*s) fp32 input of the policy net (reference policy/net.py:84-113) in one kernel."""
    _dev(frame, frame_state, output_repr, grid)
    N, _, H, W = frame.shape
    K, h, w = output_repr.shape[1:]
    Ho, Wo = int(H * scale_factor), int(W * scale_factor)  # F.interpolate(scale_factor=...) floors
    assert frame.is_contiguous() and frame_state.is_contiguous() and frame_state.dtype == frame.dtype
    if output_repr.dtype != frame.dtype:
        output_repr = output_repr.to(frame.dtype)
    g = grid.to(torch.bool).contiguous()
    out = torch.empty((N, 7 + K, Ho, Wo), dtype=torch.float32, device=frame.device)
    strides = (ctypes.c_int64 * 4)(*output_repr.stride())
    _check(lib().bc_policy_features(out.data_ptr(), frame.data_ptr(), frame_state.data_ptr(), output_repr.data_ptr(),
                                    g.data_ptr(), N, K, H, W, h, w, g.shape[2], g.shape[3], Ho, Wo,
                                    ctypes.cast(strides, ctypes.c_void_p), 1.0 / scale_factor, 1.0 / scale_factor,
                                    _dtype(frame), _stream()), "bc_policy_features")
    return out


def policy_features_shape(frame: torch.Tensor, output_repr: torch.Tensor, scale_factor: float):
    N, _, H, W = frame.shape
    return (N, 7 + output_repr.shape[1], int(H * scale_factor), int(W * scale_factor))


def policy_features_nhwc16(out16: torch.Tensor, frame: torch.Tensor, frame_state: torch.Tensor, output_repr: torch.Tensor,
                           grid: torch.Tensor, scale_factor: float) -> torch.Tensor:
    """The policy net's input features written as fp16 channels_last into out16 (N, Cp, Ho, Wo), Cp >= 7+K padded
    (see bc_policy_features_nhwc16): the values of policy_features() rounded to fp16."""
    _dev(out16, frame, frame_state, output_repr, grid)
    N, _, H, W = frame.shape
    K, h, w = output_repr.shape[1:]
    Ho, Wo = int(H * scale_factor), int(W * scale_factor)
    assert out16.dtype == torch.float16 and out16.is_contiguous(memory_format=torch.channels_last)
    assert out16.shape[0] == N and tuple(out16.shape[2:]) == (Ho, Wo)
    assert frame.is_contiguous() and frame_state.is_contiguous() and frame_state.dtype == frame.dtype
    if output_repr.dtype != frame.dtype:
        output_repr = output_repr.to(frame.dtype)
    g = grid.to(torch.bool).contiguous()
    strides = (ctypes.c_int64 * 4)(*output_repr.stride())
    _check(lib().bc_policy_features_nhwc16(out16.data_ptr(), out16.shape[1], frame.data_ptr(), frame_state.data_ptr(),
                                           output_repr.data_ptr(), g.data_ptr(), N, K, H, W, h, w, g.shape[2], g.shape[3],
                                           Ho, Wo, ctypes.cast(strides, ctypes.c_void_p), 1.0 / scale_factor,
                                           1.0 / scale_factor, _dtype(frame), _stream()), "bc_policy_features_nhwc16")
    return out16


def info_gain(outputs: torch.Tensor, outputs_prev: torch.Tensor) -> torch.Tensor:
    """InformationGainSemSeg.forward for fp16 logits (N,K,h,w) in one kernel -> (N,1,h/4,w/4) fp16."""
    _dev(outputs, outputs_prev)
    assert outputs.dtype == torch.float16 and outputs_prev.dtype == torch.float16
    if outputs_prev.stride() != outputs.stride():
        outputs_prev = outputs_prev.contiguous(memory_format=torch.channels_last) if not outputs.is_contiguous() \
            else outputs_prev.contiguous()
        assert outputs_prev.stride() == outputs.stride()
    N, K, h, w = outputs.shape
    out = torch.empty((N, 1, h // 4, w // 4), dtype=torch.float16, device=outputs.device)
    strides = (ctypes.c_int64 * 4)(*outputs.stride())
    _check(lib().bc_info_gain(out.data_ptr(), outputs.data_ptr(), outputs_prev.data_ptr(), N, K, h, w,
                              ctypes.cast(strides, ctypes.c_void_p), _stream()), "bc_info_gain")
    return out


SAMPLE_GRID_MAX_CELLS = 8192


def sample_grid(probs: torch.Tensor, uniforms: torch.Tensor, multiple: int, at_least_one: bool = False):
    """Bernoulli draw + executed-count quantisation on the device (see bc_sample_grid): probs fp32 (any shape, G
    cells), uniforms fp32 (2G,) in [0,1) -> (grid bool like probs, counts int32 (2,) = executed after / before)."""
    _dev(probs, uniforms)
    G = probs.numel()
    assert probs.dtype == torch.float32 and probs.is_contiguous() and uniforms.dtype == torch.float32
    assert uniforms.is_contiguous() and uniforms.numel() >= 2 * G
    grid = torch.empty(probs.shape, dtype=torch.bool, device=probs.device)
    counts = torch.empty(2, dtype=torch.int32, device=probs.device)
    _check(lib().bc_sample_grid(grid.data_ptr(), counts.data_ptr(), probs.data_ptr(), uniforms.data_ptr(), G, int(multiple),
                                int(bool(at_least_one)), _stream()), "bc_sample_grid")
    return grid, counts


def raster_boxes(out: torch.Tensor, rects: torch.Tensor, values: torch.Tensor, shift: int = 0) -> torch.Tensor:
    """out (H,W) fp32 CUDA <- per pixel max(0, values of the boxes containing (x >> shift, y >> shift));
    rects int32 (n,4) = x1,y1,x2,y2 half-open, values fp32 (n,), both on the device (see bc_raster_boxes)."""
    _dev(out, rects, values)
    assert out.dtype == torch.float32 and out.dim() == 2 and out.is_contiguous()
    n = rects.shape[0]
    assert rects.dtype == torch.int32 and rects.is_contiguous() and (n == 0 or rects.shape[1] == 4)
    assert values.dtype == torch.float32 and values.is_contiguous() and values.numel() == n
    _check(lib().bc_raster_boxes(out.data_ptr(), rects.data_ptr() if n else None, values.data_ptr() if n else None, n,
                                 out.shape[0], out.shape[1], int(shift), _stream()), "bc_raster_boxes")
    return out


# ------------------------------------------------------------------------------------------- per-block ConvTranspose2d
_DECONV_W = TensorCache(64)


def deconv_supported(dtype, weight: torch.Tensor, BS_in: int, stride, padding, output_padding=0, dilation=1,
                     groups: int = 1) -> bool:
    """True when a per-block ConvTranspose2d can run as bc_conv_igemm + bc_depth_to_space: fp16, groups 1, no dilation /
    output padding, (k, s, p) = (4, 2, 1) [3x3 conv over the zero-bordered tile, four output phases] or k == s in
    {2, 4}, p = 0 [1x1 conv, s*s phases]; Cin and s*s*Cout multiples of 64, tile edge a power of two in [2, 128]."""
    Cin, Cout, kh, kw = weight.shape
    if dtype != torch.float16 or weight.dtype != torch.float16 or not weight.is_cuda or groups != 1:
        return False
    if dilation != 1 or output_padding != 0 or kh != kw:
        return False
    if not ((kh, stride, padding) == (4, 2, 1) or (kh == stride and padding == 0 and stride in (2, 4))):
        return False
    if Cin % 64 or (stride * stride * Cout) % 64 or Cout % 8:
        return False
    return 2 <= BS_in <= 128 and (BS_in & (BS_in - 1)) == 0


def pack_deconv_weight(w: torch.Tensor, bias: Optional[torch.Tensor], stride: int):
    """ConvTranspose2d weight (Cin,Cout,k,k) -> (conv weight (s*s*Cout, Cin, k', k') channels_last, bias repeated per
    phase | None), output channel (a*s + b)*Cout + co = output phase (a, b) of channel co; cached per version.
    k == s: k' = 1, W'[(a,b,co), ci] = w[ci, co, a, b].  (4, 2, 1): k' = 3 (cross-correlation, pad 1); output row
    2y + a reads input rows y-1, y (a = 0: taps 3, 1) or y, y+1 (a = 1: taps 2, 0); the other row of the 3x3 is zero."""
    hit = _DECONV_W.get((w, bias), (stride,))
    if hit is None:
        Cin, Cout, k, _ = w.shape
        s = stride
        wd = w.detach()
        if k == s:
            wp = wd.permute(2, 3, 1, 0).reshape(s * s * Cout, Cin, 1, 1).contiguous()
        else:
            taps = {0: {0: 3, 1: 1}, 1: {1: 2, 2: 0}}  # phase -> {3x3 row t: transposed-conv tap}
            wp = torch.zeros(2, 2, Cout, Cin, 3, 3, dtype=w.dtype, device=w.device)
            for a in (0, 1):
                for t, ky in taps[a].items():
                    for b in (0, 1):
                        for u, kx in taps[b].items():
                            wp[a, b, :, :, t, u] = wd[:, :, ky, kx].t()
            wp = wp.reshape(4 * Cout, Cin, 3, 3).contiguous(memory_format=torch.channels_last)
        bp = None if bias is None else bias.detach().repeat(s * s).contiguous()
        hit = _DECONV_W.put((w, bias), (stride,), (wp, bp))
    return hit


def depth_to_space(out: torch.Tensor, x: torch.Tensor, r: int) -> torch.Tensor:
    """out (E,C,r*h,r*w) channels_last <- x (E,r*r*C,h,w) channels_last, channel (a*r+b)*C + c -> pixel (r*y+a, r*x+b)."""
    _dev(out, x)
    E, C, H, W = out.shape
    h, w = x.shape[2], x.shape[3]
    assert x.dtype == out.dtype == torch.float16 and x.shape[0] == E and x.shape[1] == r * r * C and (H, W) == (r * h, r * w)
    assert x.is_contiguous(memory_format=torch.channels_last) and out.is_contiguous(memory_format=torch.channels_last)
    _check(lib().bc_depth_to_space(out.data_ptr(), x.data_ptr(), E, C, h, w, int(r), _stream()), "bc_depth_to_space")
    return out


# ------------------------------------------------------------------------------------------- policy CNN backward
WGRAD_WORKSPACE = 148 * 9 * 64 * 64 * 4


def _p(t):
    return t.data_ptr() if t is not None else None


def bwd_mask_add(dst: torch.Tensor, grad: torch.Tensor, out: Optional[torch.Tensor] = None, add: Optional[torch.Tensor] = None):
    """dst = (out > 0 ? grad : 0 when `out` is given, else grad) + (add or 0); dense fp16 tensors of one shape / layout."""
    _dev(dst, grad, out, add)
    assert dst.dtype == torch.float16 and all(t is None or (t.dtype == torch.float16 and t.shape == dst.shape) for t in (grad, out, add))
    _check(lib().bc_bwd_mask_add(dst.data_ptr(), grad.data_ptr(), _p(out), _p(add), dst.numel(), _stream()), "bc_bwd_mask_add")
    return dst


def bn_bwd_reduce(sums: torch.Tensor, g: torch.Tensor, out: Optional[torch.Tensor], z: torch.Tensor, mean: torch.Tensor,
                  invstd: torch.Tensor, workspace: torch.Tensor):
    """sums fp32 (2, C) <- [sum g, sum g * xhat] per channel over the NHWC fp16 planes g / z (N,C,H,W channels_last);
    g is masked by `out > 0` when `out` is given."""
    _dev(sums, g, out, z, mean, invstd, workspace)
    N, C, H, W = z.shape
    assert g.shape == z.shape and z.is_contiguous(memory_format=torch.channels_last) and g.is_contiguous(memory_format=torch.channels_last)
    assert sums.dtype == torch.float32 and sums.numel() >= 2 * C and sums.is_contiguous()
    _check(lib().bc_bn_bwd_reduce(sums.data_ptr(), g.data_ptr(), _p(out), z.data_ptr(), mean.data_ptr(), invstd.data_ptr(),
                                  N * H * W, C, workspace.data_ptr(), workspace.numel(), _stream()), "bc_bn_bwd_reduce")


def bn_bwd_apply(dz: Optional[torch.Tensor], dz_up: Optional[torch.Tensor], g: torch.Tensor, out: Optional[torch.Tensor],
                 z: torch.Tensor, mean: torch.Tensor, invstd: torch.Tensor, gamma: Optional[torch.Tensor], sums: torch.Tensor):
    """dz (and / or dz_up: the (N,C,2H,2W) zero-interleaved copy) <- gamma * invstd * (g - sum_g/P - xhat * sum_gx/P)."""
    _dev(dz, dz_up, g, out, z, mean, invstd, gamma, sums)
    N, C, H, W = z.shape
    assert dz is None or dz.shape == z.shape
    assert dz_up is None or tuple(dz_up.shape) == (N, C, 2 * H, 2 * W)
    _check(lib().bc_bn_bwd_apply(_p(dz), _p(dz_up), g.data_ptr(), _p(out), z.data_ptr(), mean.data_ptr(), invstd.data_ptr(),
                                 _p(gamma), sums.data_ptr(), N, H, W, C, _stream()), "bc_bn_bwd_apply")


def conv_wgrad(grad_w: torch.Tensor, dz: torch.Tensor, x: torch.Tensor, stride: int, inv_scale: Optional[torch.Tensor],
               workspace: torch.Tensor, bn_sums: Optional[torch.Tensor] = None, dgamma: Optional[torch.Tensor] = None,
               dbeta: Optional[torch.Tensor] = None):
    """grad_w fp32 (Cout,Cin,k,k) <- sum over pixels of dz (N,Cout_p,H/s,W/s) x x (N,Cin_p,H,W) (fp16 channels_last),
    times *inv_scale; optionally the unit's BatchNorm gradients from bn_sums (see bc_conv_wgrad)."""
    _dev(grad_w, dz, x, inv_scale, workspace, bn_sums, dgamma, dbeta)
    N, Cin_p, H, W = x.shape
    Cout_p = dz.shape[1]
    Cout, Cin, k, _ = grad_w.shape
    assert grad_w.dtype == torch.float32 and tuple(dz.shape) == (N, Cout_p, H // stride, W // stride)
    assert x.is_contiguous(memory_format=torch.channels_last) and dz.is_contiguous(memory_format=torch.channels_last)
    gs = (ctypes.c_longlong * 4)(*grad_w.stride())
    _check(lib().bc_conv_wgrad(grad_w.data_ptr(), ctypes.cast(gs, _vp), dz.data_ptr(), x.data_ptr(), N, H, W, Cin_p, Cout_p, Cin,
                               Cout, k, int(stride), _p(inv_scale), _p(dgamma), _p(dbeta), _p(bn_sums), workspace.data_ptr(),
                               workspace.numel(), _stream()), "bc_conv_wgrad")


# ------------------------------------------------------------------------------------------- train-mode BN statistics
BN_STATS_WORKSPACE = 16 + 2 * 148 * 2 * 128 * 4


def bn_stats(x: torch.Tensor, mean: torch.Tensor, invstd: torch.Tensor, eps: float, workspace: torch.Tensor):
    """mean / invstd (fp32 [C]) of a train-mode BatchNorm2d over all pixels of the dense channels_last fp16
    tensor x (N,C,H,W); workspace: uint8, >= BN_STATS_WORKSPACE bytes, zeroed once by the caller."""
    _dev(x, mean, invstd, workspace)
    assert x.dtype == torch.float16 and x.dim() == 4 and (x.is_contiguous(memory_format=torch.channels_last))
    N, C, H, W = x.shape
    assert mean.dtype == torch.float32 and invstd.dtype == torch.float32 and mean.numel() >= C and invstd.numel() >= C
    _check(lib().bc_bn_stats(mean.data_ptr(), invstd.data_ptr(), x.data_ptr(), N * H * W, C, float(eps),
                             workspace.data_ptr(), workspace.numel(), _stream()), "bc_bn_stats")


GN_STATS_WORKSPACE = 16 + 2 * 148 * 256 * 2 * 8


def gn_stats(x: torch.Tensor, groups: int, eps: float, mean: torch.Tensor, invstd: torch.Tensor, workspace: torch.Tensor):
    """Per-channel (mean, invstd) of each channel's GROUP over all pixels of the packed channels_last fp16 tile
    batch x (E,C,h,w): GroupNorm statistics with the executed blocks folded into one sample (bc_gn_stats)."""
    _dev(x, mean, invstd, workspace)
    assert x.dtype == torch.float16 and x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last)
    E, C, h, w = x.shape
    assert mean.dtype == torch.float32 and invstd.dtype == torch.float32 and mean.numel() >= C and invstd.numel() >= C
    _check(lib().bc_gn_stats(mean.data_ptr(), invstd.data_ptr(), x.data_ptr(), E * h * w, C, int(groups), float(eps),
                             workspace.data_ptr(), workspace.numel(), _stream()), "bc_gn_stats")


def gn_supported(x: torch.Tensor, groups: int) -> bool:
    C = x.shape[1]
    return (x.is_cuda and x.dtype == torch.float16 and x.dim() == 4 and C % 8 == 0 and C <= 2048 and 1 <= groups <= 256
            and C % groups == 0 and (C // groups) % 8 == 0)


def bn_norm(out: torch.Tensor, x: torch.Tensor, mean: torch.Tensor, invstd: torch.Tensor, weight: Optional[torch.Tensor],
            shift: Optional[torch.Tensor], eps: float, relu: bool, workspace: torch.Tensor):
    """out = relu?(weight * (x - mean) * invstd + shift) with the batch statistics of x (train-mode BatchNorm2d), one launch
    (bc_bn_norm = bn_stats + the normalisation behind a grid-wide barrier); mean / invstd (fp32 [C]) are written too."""
    _dev(out, x, mean, invstd, weight, shift, workspace)
    assert x.dtype == torch.float16 and x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last)
    assert out.dtype == torch.float16 and out.shape == x.shape and out.is_contiguous(memory_format=torch.channels_last)
    N, C, H, W = x.shape
    _check(lib().bc_bn_norm(out.data_ptr(), mean.data_ptr(), invstd.data_ptr(), x.data_ptr(),
                            weight.data_ptr() if weight is not None else None, shift.data_ptr() if shift is not None else None,
                            N * H * W, C, float(eps), int(bool(relu)), workspace.data_ptr(), workspace.numel(), _stream()),
           "bc_bn_norm")
    return out


def bn_update_running(table: torch.Tensor):
    """table: int64 CUDA tensor (n, 8), see bc_bn_update_running."""
    _dev(table)
    assert table.dtype == torch.int64 and table.dim() == 2 and table.shape[1] == 8 and table.is_contiguous()
    _check(lib().bc_bn_update_running(table.data_ptr(), table.shape[0], _stream()), "bc_bn_update_running")


def pack_params(table: torch.Tensor, total: int):
    """table: int64 CUDA tensor (n, 12), see bc_pack_params."""
    _dev(table)
    assert table.dtype == torch.int64 and table.dim() == 2 and table.shape[1] == 12 and table.is_contiguous()
    _check(lib().bc_pack_params(table.data_ptr(), table.shape[0], int(total), _stream()), "bc_pack_params")


def rmsprop_step(rows, device: torch.device, lr: float, alpha: float, eps: float, weight_decay: float, momentum: float):
    """rows: list of (param ptr, grad ptr, square_avg ptr, momentum buffer ptr | 0, cumulative offset, numel) -- a HOST
    table, see bc_rmsprop_step."""
    n = len(rows)
    flat = (ctypes.c_longlong * (6 * n))(*[int(v) for r in rows for v in r])
    total = rows[-1][4] + rows[-1][5]
    with torch.cuda.device(device):
        _check(lib().bc_rmsprop_step(ctypes.cast(flat, _vp), n, int(total), float(lr), float(alpha), float(eps),
                                     float(weight_decay), float(momentum), _stream()), "bc_rmsprop_step")


def conv_fewout(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], stride: int, padding: int,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """conv2d with <= 16 output channels: x (N,Cx,H,W) channels_last fp16 (the first weight.shape[1] channels are
    used), weight fp32 (Cout,C,k,k) any strides, bias fp32 -> (N,Cout,Ho,Wo) fp32."""
    _dev(x, weight, bias, out)
    assert x.dtype == torch.float16 and x.is_contiguous(memory_format=torch.channels_last) and weight.dtype == torch.float32
    N, Cx, H, W = x.shape
    Cout, C, k, _ = weight.shape
    Ho, Wo = (H + 2 * padding - k) // stride + 1, (W + 2 * padding - k) // stride + 1
    if out is None:
        out = torch.empty((N, Cout, Ho, Wo), dtype=torch.float32, device=x.device)
    ws = (ctypes.c_int64 * 4)(*weight.stride())
    _check(lib().bc_conv_fewout(out.data_ptr(), x.data_ptr(), weight.data_ptr(),
                                bias.data_ptr() if bias is not None else None, N, H, W, C, Cx, Cout, k, stride, padding,
                                ctypes.cast(ws, _vp), _stream()), "bc_conv_fewout")
    return out


# ------------------------------------------------------------------------------------------- driver-side steps
def frame_from_u8(src_u8: torch.Tensor, mean, std, dtype=torch.float16, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(N,H,W,3) or (H,W,3) uint8 CUDA frame -> normalised (N,3,H,W) `dtype` network input:
    ((x / 255) - mean) / std, bit-identical to to_tensor + normalize + .to(dtype) (see bc_frame_from_u8)."""
    _dev(src_u8, out)
    if src_u8.dim() == 3:
        src_u8 = src_u8.unsqueeze(0)
    assert src_u8.dtype == torch.uint8 and src_u8.dim() == 4 and src_u8.shape[3] == 3 and src_u8.is_contiguous(), \
        "frame must be a contiguous (N,H,W,3) uint8 tensor"
    N, H, W, _ = src_u8.shape
    if out is None:
        out = torch.empty((N, 3, H, W), dtype=dtype, device=src_u8.device)
    assert out.shape == (N, 3, H, W) and out.is_contiguous() and out.dtype in (torch.float16, torch.float32)
    m = (ctypes.c_float * 3)(*[float(v) for v in mean])
    sd = (ctypes.c_float * 3)(*[float(v) for v in std])
    _check(lib().bc_frame_from_u8(out.data_ptr(), src_u8.data_ptr(), ctypes.cast(m, _vp), ctypes.cast(sd, _vp), N, H, W,
                                  BC_F16 if out.dtype == torch.float16 else BC_F32, _stream()), "bc_frame_from_u8")
    return out


def blocks_from_u8(tiles: torch.Tensor, src_u8: torch.Tensor, mean, std, mapping_exec: torch.Tensor, E: int) -> torch.Tensor:
    """tiles (E,3,BS,BS) contiguous fp16/fp32 <- the normalised executed blocks of the (N,H,W,3) uint8 CUDA frame:
    frame_from_u8 + gather in one pass over the executed blocks only (see bc_blocks_from_u8)."""
    _dev(tiles, src_u8, mapping_exec)
    assert src_u8.dtype == torch.uint8 and src_u8.dim() == 4 and src_u8.shape[3] == 3 and src_u8.is_contiguous()
    N, H, W, _ = src_u8.shape
    BS = tiles.shape[2]
    assert tiles.is_contiguous() and tuple(tiles.shape[1:]) == (3, BS, BS) and tiles.shape[0] >= E
    assert mapping_exec.dtype == torch.int32 and mapping_exec.numel() >= E
    m = (ctypes.c_float * 3)(*[float(v) for v in mean])
    sd = (ctypes.c_float * 3)(*[float(v) for v in std])
    _check(lib().bc_blocks_from_u8(tiles.data_ptr(), src_u8.data_ptr(), ctypes.cast(m, _vp), ctypes.cast(sd, _vp),
                                   mapping_exec.data_ptr(), E, N, H, W, BS, BC_F16 if tiles.dtype == torch.float16 else BC_F32,
                                   _stream()), "bc_blocks_from_u8")
    return tiles


def upsample_argmax(logits: torch.Tensor, scale: int = 4, label_dtype=torch.uint8,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(N,K,h,w) fp16/fp32 logits -> (N, scale*h, scale*w) class map = argmax over classes of the bilinear
    (align_corners=False) upsampling, without materialising it (see bc_upsample_argmax)."""
    _dev(logits, out)
    assert logits.dim() == 4 and logits.dtype in (torch.float16, torch.float32)
    assert label_dtype in (torch.uint8, torch.int64)
    N, K, h, w = logits.shape
    if out is None:
        out = torch.empty((N, h * scale, w * scale), dtype=label_dtype, device=logits.device)
    assert out.shape == (N, h * scale, w * scale) and out.is_contiguous() and out.dtype == label_dtype
    strides = (ctypes.c_int64 * 4)(*logits.stride())
    _check(lib().bc_upsample_argmax(out.data_ptr(), logits.data_ptr(), N, K, h, w, ctypes.cast(strides, _vp), int(scale),
                                    BC_F16 if logits.dtype == torch.float16 else BC_F32, out.element_size(), _stream()),
           "bc_upsample_argmax")
    return out


def upsample_argmax_blocks(labels: torch.Tensor, logits: torch.Tensor, grid: torch.Tensor, scale: int = 4) -> torch.Tensor:
    """In-place block-sparse update of `labels` (the previous frame's label map, (N, scale*h, scale*w) uint8/int64)
    from this frame's dense `logits` (N,K,h,w): only executed cells of `grid` (bool (N,1,GH,GW)) + a one-logit-pixel
    ring are recomputed (see bc_upsample_argmax_blocks)."""
    _dev(labels, logits, grid)
    N, K, h, w = logits.shape
    assert logits.dtype in (torch.float16, torch.float32) and labels.dtype in (torch.uint8, torch.int64)
    assert tuple(labels.shape) == (N, h * scale, w * scale) and labels.is_contiguous()
    assert grid.dim() == 4 and grid.shape[0] == N and grid.shape[1] == 1 and grid.dtype in (torch.bool, torch.uint8)
    g = grid if grid.is_contiguous() else grid.contiguous()
    strides = (ctypes.c_int64 * 4)(*logits.stride())
    _check(lib().bc_upsample_argmax_blocks(labels.data_ptr(), logits.data_ptr(), g.data_ptr(), N, K, h, w,
                                           ctypes.cast(strides, _vp), int(scale),
                                           BC_F16 if logits.dtype == torch.float16 else BC_F32, labels.element_size(),
                                           g.shape[2], g.shape[3], _stream()), "bc_upsample_argmax_blocks")
    return labels


# ------------------------------------------------------------------------------------------- dense pyramid pooling
def _grids(gh, gw):
    L = len(gh)
    return (ctypes.c_int32 * L)(*gh), (ctypes.c_int32 * L)(*gw), L


def spp_pool(pooled, x0, gh, gw):
    _dev(pooled, x0)
    N, C, H, W = x0.shape
    a, b, L = _grids(gh, gw)
    _check(lib().bc_spp_pool(pooled.data_ptr(), x0.data_ptr(), N, C, H, W, L, ctypes.cast(a, _vp), ctypes.cast(b, _vp),
                             _stream()), "bc_spp_pool")
    return pooled


def spp_levels(out, pooled, bn, weights, x0_shape, gh, gw):
    _dev(out, pooled, bn, weights)
    N, C, H, W = x0_shape
    a, b, L = _grids(gh, gw)
    _check(lib().bc_spp_levels(out.data_ptr(), pooled.data_ptr(), bn.data_ptr(), weights.data_ptr(), N, C, H, W, L,
                               ctypes.cast(a, _vp), ctypes.cast(b, _vp), weights.shape[1], _stream()), "bc_spp_levels")
    return out


def spp_prep(y, x0, levels, bn, gh, gw):
    _dev(y, x0, levels, bn)
    N, C, H, W = x0.shape
    a, b, L = _grids(gh, gw)
    _check(lib().bc_spp_prep(y.data_ptr(), x0.data_ptr(), levels.data_ptr(), bn.data_ptr(), N, C, H, W, L,
                             ctypes.cast(a, _vp), ctypes.cast(b, _vp), levels.shape[1], y.shape[1], _stream()),
           "bc_spp_prep")
    return y
