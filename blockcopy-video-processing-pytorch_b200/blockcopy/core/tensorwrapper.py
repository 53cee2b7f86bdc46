"""TensorWrapper / BlockFeatures -- the block-sparse tensor type of the blockcopy API.

Public surface and semantics follow the reference (core/tensorwrapper.py:17-633): a
``torch.Tensor`` subclass that intercepts every torch call made by an unmodified CNN, keeps the
packed ``(E, C, BS, BS)`` tile batch of the executed blocks, and replaces zero padding by halos
taken from neighbouring blocks (this frame's value if the neighbour executed, the most recent
older value otherwise).

What is different underneath (B200 design):

* Temporal state is a list of dense, persistent feature PLANES -- one per padded op and one per
  combine point, identified by call order like the reference's FIFOs (tensorwrapper.py:180-224) --
  updated in place for executed blocks.  The reference instead keeps two tile tensors per layer
  and rebuilds halos with ``transfer`` (ring copy) + ``repad``; both formulations give bit-identical
  padded tiles (tests/test_oracle.py::test_ring_protocol_equals_plane).
* All index bookkeeping runs on the device (``bc_compact_mask``); the only host round trip left is
  the executed-block count, which the API exposes as a Python int anyway
  (policy.py:82-94 ``num_exec``) and which is reused here through ``grid._bc_num_exec``.
* Tiles may be NCHW or channels_last; kernels are layout-aware (``_C.layout_of``).
"""
from __future__ import annotations

import os
import warnings
from typing import Any, Callable, Dict, List, Optional, Tuple

import torch

from .. import _C
from ..utils.profiler import timings

FUSED_CONV = os.environ.get("BLOCKCOPY_FUSED_CONV", "1") != "0"  # route eligible convs on blocks to the tcgen05 implicit-GEMM kernel (bc_conv_igemm)
VERBOSE = False  # print a line per split / combine / grid
BLOCKPAD_WITH_ZEROES = False  # debugging: keep the op's own zero padding (wrong at block borders)


def is_tensorwrapper(x) -> bool:
    """True if ``x`` is a TensorWrapper."""
    return isinstance(x, TensorWrapper)


def is_block(x) -> bool:
    """True if ``x`` is a TensorWrapper holding packed blocks."""
    return isinstance(x, TensorWrapper) and x.is_blocks


def to_tensorwrapper(x: torch.Tensor) -> "TensorWrapper":
    """Reinterpret a CUDA tensor as TensorWrapper (no copy)."""
    assert x.is_cuda, "x must be on CUDA device!"
    return x.as_subclass(TensorWrapper)


def to_tensor(x):
    """TensorWrapper (or list / tuple / dict of them) -> plain dense torch.Tensor(s)."""
    if isinstance(x, TensorWrapper):
        return x.to_tensor()
    if isinstance(x, list):
        return [to_tensor(v) for v in x]
    if isinstance(x, tuple):
        return tuple(to_tensor(v) for v in x)
    if isinstance(x, dict):
        return {k: to_tensor(v) for k, v in x.items()}
    return x


# Operation classes (names as they reach __torch_function__), reference tensorwrapper.py:69-105.
OPS = {
    "PADDED": {"conv2d", "max_pool2d", "avg_pool2d", "lp_pool2d", "fractional_max_pool2d"},
    "INTERPOLATE": {"interpolate", "upsample_bilinear"},
    "BATCHED": {"group_norm"},
    "INCOMPATIBLE": {"adaptive_avg_pool2d", "adaptive_max_pool2d", "linear", "flip", "unsqueeze", "reshape", "view"},
    "CHANNELONLY": {"mean", "sum", "max", "min,", "std", "var", "argmax", "count_nonzero", "nonzero"},
    "WARNING": {""},
}
OPS_SPECIAL = set().union(*OPS.values())

# position of `padding` when a padded op is called positionally
_PADDING_ARG_INDEX = {"conv2d": 4, "max_pool2d": 3, "avg_pool2d": 3}


def get_grid_mappings(grid: torch.Tensor):
    """Device-side equivalent of the reference's TorchScript helper (tensorwrapper.py:108-128):
    bool grid (N,1,GH,GW) on CUDA -> (grid_idx int32 like grid, mapping_exec int32 (E,))."""
    g = grid.to(torch.bool).contiguous()
    G = g.numel()
    grid_idx = torch.empty(g.shape, dtype=torch.int32, device=g.device)
    mapping = torch.empty(G, dtype=torch.int32, device=g.device)
    counts = torch.empty(2, dtype=torch.int32, device=g.device)
    _C.compact_mask(g.view(torch.uint8), grid_idx, mapping, counts)
    return grid_idx, mapping[: int(counts[0])]


class BlockFeatures:
    """Temporal state of one stream for one frame (API name from the reference, :131-232).

    A new instance is created per frame by ``TensorWrapper.process_temporal_features(prev)``; it
    ADOPTS the persistent planes of ``prev`` instead of popping tile FIFOs.
    """

    def __init__(self, device, prev: Optional["BlockFeatures"] = None):
        self.device = device
        self._grid: Optional[torch.Tensor] = None          # bool (N,1,GH,GW)
        self._grid_idx: Optional[torch.Tensor] = None      # int32 (N,1,GH,GW)
        self._mapping_exec: Optional[torch.Tensor] = None  # int32 (E,)
        self._transfer_idx: Optional[torch.Tensor] = None  # int32 (T,) reference-protocol compat
        self.num_exec = 0
        self.num_total = 0
        self.has_history = prev is not None and not prev._was_reset
        self._was_reset = False
        self.track_transfer_idx = True  # also produce transfer_idx (reference tile protocol compat)
        # persistent planes; slot = call order within a frame
        self._planes: List[torch.Tensor] = prev._planes if prev is not None else []
        self._full: List[torch.Tensor] = prev._full if prev is not None else []
        self._plane_cursor = 0
        self._full_cursor = 0
        self._prev_grid_idx = prev._grid_idx if (prev is not None and self.has_history) else None
        if prev is not None:
            prev._planes, prev._full = [], []  # ownership moved

    # ------------------------------------------------------------------ grid
    def _process_grid(self, grid: torch.Tensor, meta_prev: Optional["BlockFeatures"] = None) -> None:
        with timings.env("tensorwrapper/process_grid", 10):
            assert grid.dim() == 4 and grid.shape[1] == 1, "grid must be (N,1,GH,GW)"
            hint = getattr(grid, "_bc_num_exec", None)
            g = grid.to(self.device, dtype=torch.bool).contiguous()
            G = g.numel()
            grid_idx = torch.empty(g.shape, dtype=torch.int32, device=self.device)
            buf = torch.empty(2 * G + 2, dtype=torch.int32, device=self.device)
            mapping, transfer, counts = buf[:G], buf[G:2 * G], buf[2 * G:]
            prev_idx = None
            if self.track_transfer_idx:
                prev_idx = self._prev_grid_idx if self._prev_grid_idx is not None else (
                    meta_prev._grid_idx if (meta_prev is not None and self.has_history) else None)
            _C.compact_mask(g.view(torch.uint8), grid_idx, mapping, counts, prev_idx,
                            transfer if prev_idx is not None else None)
            n_exec = int(hint) if hint is not None else int(counts[0])  # the single host round trip
            if not self.has_history:
                assert n_exec == G, "No previous features known, first run should execute all blocks!"
            self._grid, self._grid_idx = g, grid_idx
            self._mapping_exec = mapping[:n_exec]
            self._transfer_idx = transfer[: G - n_exec] if prev_idx is not None else None
            self.num_exec, self.num_total = n_exec, G
            if VERBOSE:
                print(f"TensorWrapper >> GRID {tuple(g.shape)} exec {n_exec}/{G}")

    # ------------------------------------------------------------------ planes
    def _next_plane(self, like: torch.Tensor, shape) -> torch.Tensor:
        """Plane of the next padded op in call order; allocated on the first frame."""
        i = self._plane_cursor
        self._plane_cursor += 1
        if i < len(self._planes):
            plane = self._planes[i]
            if plane.shape != tuple(shape) or plane.dtype != like.dtype:
                raise AssertionError(
                    f"padded op #{i}: plane {tuple(plane.shape)}/{plane.dtype} does not match this frame's "
                    f"{tuple(shape)}/{like.dtype}; the model must issue the same ops every frame")
            return plane
        assert not self.has_history or self.num_exec == self.num_total, \
            "No computed features to pop from stack, something seems wrong in the model."
        fmt = torch.channels_last if _C.layout_of(like) == _C.BC_NHWC else torch.contiguous_format
        plane = torch.empty(shape, dtype=like.dtype, device=like.device, memory_format=fmt)
        self._planes.append(plane)
        return plane

    def _next_full(self) -> Tuple[int, Optional[torch.Tensor]]:
        i = self._full_cursor
        self._full_cursor += 1
        return i, (self._full[i] if i < len(self._full) else None)

    def _set_full(self, i: int, plane: torch.Tensor):
        if i < len(self._full):
            self._full[i] = plane
        else:
            assert i == len(self._full)
            self._full.append(plane)

    def mark_reset(self):
        """Logical reset that keeps the planes allocated (CUDA-graph mode): the next frame must
        execute every block, which rewrites all of them."""
        self._was_reset = True

    def clear(self):
        """Drop all stored features."""
        self._planes.clear()
        self._full.clear()
        self._prev_grid_idx = None

    def state_bytes(self) -> int:
        return sum(p.numel() * p.element_size() for p in self._planes + self._full)


def _dense(t: torch.Tensor) -> torch.Tensor:
    """Plain, dense (NCHW or channels_last) view/copy of t for the kernels."""
    t = t.as_subclass(torch.Tensor)
    if t.is_contiguous() or (t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last)):
        return t
    return t.contiguous()


def _like_layout(shape, like: torch.Tensor) -> torch.Tensor:
    fmt = torch.channels_last if _C.layout_of(like) == _C.BC_NHWC else torch.contiguous_format
    return torch.empty(shape, dtype=like.dtype, device=like.device, memory_format=fmt)


def _match_layout(t: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    if t.numel() == 0 or _C.layout_of(t) == _C.layout_of(like):
        return t
    fmt = torch.channels_last if _C.layout_of(like) == _C.BC_NHWC else torch.contiguous_format
    return t.contiguous(memory_format=fmt)


class TensorWrapper(torch.Tensor):
    """Dense or block-sparse tensor with temporal feature propagation."""

    # class-level defaults: a freshly viewed wrapper is dense and has no state
    _is_blocks = False
    _features: Optional[BlockFeatures] = None
    _features_prev: Optional[BlockFeatures] = None

    # ------------------------------------------------------------------ metadata
    @property
    def is_init(self) -> bool:
        return self._features is not None

    def _inherit(self, other: "TensorWrapper") -> "TensorWrapper":
        self._is_blocks = other._is_blocks
        self._features = other._features
        self._features_prev = other._features_prev
        return self

    def process_temporal_features(self, features_prev: Optional[BlockFeatures] = None) -> BlockFeatures:
        """Start a new frame: returns the state object to hand in again for the next frame."""
        self._features_prev = features_prev
        self._features = BlockFeatures(device=self.device, prev=features_prev)
        self._is_blocks = False
        return self._features

    @property
    def data_shape(self) -> torch.Size:
        return self.shape

    @property
    def is_blocks(self) -> bool:
        return self._is_blocks

    @property
    def block_size(self) -> int:
        """Block edge in pixels at this tensor's resolution; -1 when dense."""
        return self.data_shape[-1] if self.is_blocks else -1

    def get_grid(self) -> torch.Tensor:
        return self._features._grid

    def get_grid_idx(self) -> torch.Tensor:
        return self._features._grid_idx

    def get_mapping_exec(self) -> torch.Tensor:
        return self._features._mapping_exec

    def get_features(self) -> BlockFeatures:
        return self._features

    # ------------------------------------------------------------------ dense <-> blocks
    def to_blocks(self, grid: torch.Tensor) -> "TensorWrapper":
        """Dense -> packed blocks of the True cells of ``grid`` (bool, (N,1,GH,GW))."""
        assert not self.is_blocks
        assert self.is_init, "need to call process_temporal_features before splitting in blocks!"
        self._features._process_grid(grid, self._features_prev)
        assert grid.dim() == 4 and self.dim() == 4
        assert self.shape[2] % grid.shape[2] == 0
        assert self.shape[3] % grid.shape[3] == 0
        return self._split(self.shape[2] // grid.shape[2])

    def to_blocks_like(self, other: "TensorWrapper") -> "TensorWrapper":
        """Dense -> packed blocks with the grid of ``other``."""
        self._inherit(other)
        self._is_blocks = False
        return self._split(self.shape[2] // self.get_grid().shape[2])

    def _split(self, block_size: int) -> "TensorWrapper":
        assert self.is_init, "need to call process_temporal_features before splitting in blocks!"
        with timings.env("tensorwrapper/split", 10):
            if self.is_blocks:
                raise AttributeError("TensorWrapper: already split in blocks! Cannot split again.")
            if self.dim() != 4:
                raise AttributeError("TensorWrapper only supports 4D NCHW tensors!")
            N, C, H, W = self.shape
            if H % block_size != 0 or W % block_size != 0:
                raise AttributeError(
                    f"TensorWrapper: Shape ({self.shape}) not divisibile by given block size ({block_size})!")
            feats = self._features
            _, _, GH, GW = feats._grid_idx.shape
            block_size = W // GW
            E = feats.num_exec
            image = _dense(self)
            out = _like_layout((E, C, block_size, block_size), image)
            _C.gather(out, image, feats._mapping_exec, E)
            out = out.as_subclass(TensorWrapper)._inherit(self)
            out._is_blocks = True
            return out

    def to_tensor(self) -> torch.Tensor:
        """Plain dense torch.Tensor (combines the blocks first if needed)."""
        out = self.combine() if self.is_blocks else self
        return out.as_subclass(torch.Tensor)

    def combine_(self) -> "TensorWrapper":
        """In-place combine: executed blocks are written into the previous frame's dense tensor."""
        return self.combine(inplace=True)

    def combine(self, inplace: bool = False) -> "TensorWrapper":
        """Packed blocks -> dense tensor; cells that were not executed keep the previous frame's
        values.  ``inplace=False`` leaves the previous dense tensor untouched (one fused
        copy+scatter pass instead of the reference's clone + scatter, tensorwrapper.py:421-434)."""
        with timings.env("tensorwrapper/combine", 4):
            if not self.is_blocks:
                raise AttributeError("TensorWrapper: Not split in blocks!")
            feats = self._features
            grid_idx, mapping = feats._grid_idx, feats._mapping_exec
            E, C, BS, _ = self.shape
            N, _, GH, GW = grid_idx.shape
            shape = (N, C, GH * BS, GW * BS)
            tiles = _dense(self)
            slot, prev = feats._next_full()
            if prev is not None:
                assert tuple(prev.shape) == shape, (shape, tuple(prev.shape))
                tiles = _match_layout(tiles, prev)
                if inplace:
                    out = _C.scatter(tiles, prev, mapping, E)
                else:
                    out = torch.empty_like(prev)
                    _C.copy_blocks(out, prev, tiles, grid_idx)
            else:
                assert E == grid_idx.numel(), "first frame must execute every block"
                out = _like_layout(shape, tiles)
                _C.scatter(tiles, out, mapping, E)
            feats._set_full(slot, out)
            out = out.as_subclass(TensorWrapper)._inherit(self)
            out._is_blocks = False
            return out

    # ------------------------------------------------------------------ halo
    def block_pad(self, padding: int) -> torch.Tensor:
        """Packed tiles (E,C,BS,BS) -> (E,C,BS+2p,BS+2p) with neighbour halos; also records this
        frame's tiles in the op's persistent plane (the reference does store_features + transfer
        + repad here, tensorwrapper.py:551-563)."""
        feats = self._features
        tiles = _dense(self)
        E, C, BS, _ = tiles.shape
        N, _, GH, GW = feats._grid_idx.shape
        plane = feats._next_plane(tiles, (N, C, GH * BS, GW * BS))
        tiles = _match_layout(tiles, plane)
        with timings.env("tensorwrapper/transfer", 10):
            _C.scatter(tiles, plane, feats._mapping_exec, E)
        with timings.env("tensorwrapper/pad", 10):
            out = _like_layout((E, C, BS + 2 * padding, BS + 2 * padding), plane)
            _C.gather_halo(out, plane, feats._mapping_exec, E, BS, padding)
        return out

    # ------------------------------------------------------------------ dispatch
    @classmethod
    def __torch_function__(cls, func: Callable, types: Tuple, args: Tuple = (), kwargs: Optional[Dict] = None) -> Any:
        kwargs = kwargs or {}
        src = _first_wrapper(args)
        if src is None and kwargs:
            src = _first_wrapper(tuple(kwargs.values()))
        if src is None or not src._is_blocks:
            out = cls._dense_dispatch(func, args, kwargs) if src is not None else None
            if out is None:
                out = super().__torch_function__(func, types, args, kwargs)
            if src is not None and isinstance(out, TensorWrapper) and out is not src and out._features is None:
                out._inherit(src)
            return out

        op = getattr(func, "__name__", "")
        if op in OPS_SPECIAL:
            if op in OPS["PADDED"]:
                out = src._func_replace_padding(func, types, args, kwargs)
            elif op in OPS["INTERPOLATE"]:
                out = src._func_interpolate(func, types, args, kwargs)
            elif op in OPS["BATCHED"]:
                out = src._func_batched(func, types, args, kwargs)
            elif op in OPS["CHANNELONLY"]:
                if kwargs.get("dim", None) != 1:
                    print(f"Operation {op} might behave differently with TensorWrapper when dim != 1!")
                out = super().__torch_function__(func, types, args, kwargs)
            elif op in OPS["INCOMPATIBLE"]:
                raise AttributeError(f"Operation {op} not supported for TensorWrapper!")
            else:
                warnings.warn(f"Operation {op} might behave differently with TensorWrapper!")
                out = super().__torch_function__(func, types, args, kwargs)
        else:
            out = super().__torch_function__(func, types, args, kwargs)

        if isinstance(out, TensorWrapper) and out is not src:
            out._inherit(src)
        return out

    @staticmethod
    def _dense_dispatch(func, args, kwargs):
        """Dense (combined) wrappers behave like plain tensors, except for ops whose stock CUDA
        kernel is pathological on the small channels_last planes of a ``blockcopy_noblocks`` region:
        adaptive_avg_pool2d to an output that divides the input is the same mean over the same
        windows as avg_pool2d but ~20x faster there."""
        if getattr(func, "__name__", "") != "adaptive_avg_pool2d" or len(args) < 2:
            return None
        x, size = args[0], args[1]
        if not isinstance(x, torch.Tensor) or x.dim() != 4:
            return None
        size = (size, size) if isinstance(size, int) else tuple(size)
        if len(size) != 2 or None in size or x.shape[2] % size[0] or x.shape[3] % size[1]:
            return None
        k = (x.shape[2] // size[0], x.shape[3] // size[1])
        out = torch.nn.functional.avg_pool2d(x.as_subclass(torch.Tensor), kernel_size=k, stride=k)
        return out.as_subclass(TensorWrapper)

    def _try_fused_conv(self, args, kwargs):
        """conv2d on blocks through bc_conv_igemm: the operand load reads the op's persistent plane
        by block index (halo included), bias is added in the epilogue; returns None if the conv is
        outside the kernel's envelope (then the generic gather-halo + torch path runs)."""
        names = ("input", "weight", "bias", "stride", "padding", "dilation", "groups")
        a = dict(bias=None, stride=1, padding=0, dilation=1, groups=1)
        a.update(zip(names, args))
        a.update(kwargs)
        one = lambda v: v if isinstance(v, int) else (v[0] if len(set(v)) == 1 else None)  # noqa: E731
        stride, padding, dilation = one(a["stride"]), one(a["padding"]), one(a["dilation"])
        if stride is None or padding is None or dilation is None or isinstance(padding, str):
            return None
        x, weight, bias = a["input"], a["weight"], a["bias"]
        if not isinstance(x, TensorWrapper) or not x._is_blocks or isinstance(weight, TensorWrapper):
            return None
        E, Cin, BS, _ = x.shape
        if E == 0 or weight.dim() != 4 or weight.shape[1] != Cin:
            return None
        if not _C.conv_supported(x.dtype, weight, BS, stride, padding, dilation, a["groups"]):
            return None
        if bias is not None and (bias.dtype != torch.float16 or not bias.is_contiguous()):
            return None
        feats = self._features
        if padding > 0 and feats._plane_cursor < len(feats._planes) and \
                _C.layout_of(feats._planes[feats._plane_cursor]) != _C.BC_NHWC:
            return None  # this op's plane was created NCHW on the first frame: stay on the generic path
        tiles = _dense(x).contiguous(memory_format=torch.channels_last)
        w = weight.detach()
        if not w.is_contiguous(memory_format=torch.channels_last):
            w = w.contiguous(memory_format=torch.channels_last)
        Cout, BSo = weight.shape[0], BS // stride
        out = torch.empty((E, Cout, BSo, BSo), dtype=tiles.dtype, device=tiles.device,
                          memory_format=torch.channels_last)
        if padding > 0:
            N, _, GH, GW = feats._grid_idx.shape
            plane = feats._next_plane(tiles, (N, Cin, GH * BS, GW * BS))
            with timings.env("tensorwrapper/transfer", 10):
                _C.scatter(tiles, plane, feats._mapping_exec, E)
            with timings.env("tensorwrapper/pad_func", 11):
                _C.conv_igemm(out, plane, w, bias, None, feats._mapping_exec, E, BS, stride, padding)
        else:
            with timings.env("tensorwrapper/pad_func0", 11):
                _C.conv_igemm(out, tiles, w, bias, None, None, E, BS, stride, 0)
        return out.as_subclass(TensorWrapper)

    def _func_replace_padding(self, func, types, args, kwargs):
        """Padded op: take the padding from neighbouring blocks instead of zeros, then run the op
        itself with padding 0 (reference: _func_replace_paddding, tensorwrapper.py:529-575)."""
        if BLOCKPAD_WITH_ZEROES:
            return super().__torch_function__(func, types, args, kwargs)
        op = func.__name__
        if FUSED_CONV and op == "conv2d":
            fused = self._try_fused_conv(args, kwargs)
            if fused is not None:
                return fused
        args = list(args)
        pos = _PADDING_ARG_INDEX.get(op, 4)
        padding = kwargs.get("padding", None)
        if padding is None:
            padding = args[pos] if len(args) > pos else 0
        zeros: Any = 0
        if isinstance(padding, str):
            if padding != "valid":
                raise NotImplementedError(f"Only support equal paddings, got {padding}")
            padding = 0
        if isinstance(padding, (tuple, list)):
            zeros = (0, 0)
            if len(padding) > 1 and padding[0] != padding[1]:
                raise NotImplementedError(f"Only support equal paddings, got {padding}")
            padding = padding[0]

        if padding > 0:
            args[0] = args[0].block_pad(int(padding)).as_subclass(TensorWrapper)._inherit(self)
            if "padding" in kwargs:
                kwargs = dict(kwargs, padding=zeros)
            else:
                args[pos] = zeros
            with timings.env("tensorwrapper/pad_func", 11):
                return super().__torch_function__(func, types, tuple(args), kwargs)
        with timings.env("tensorwrapper/pad_func0", 11):
            return super().__torch_function__(func, types, tuple(args), kwargs)

    def _func_interpolate(self, func, types, args, kwargs):
        """Per-block interpolation: every tile is one batch element, so bilinear taps clamp at the
        BLOCK border exactly like the reference's trilinear rewrite (tensorwrapper.py:577-598)."""
        return super().__torch_function__(func, types, args, kwargs)

    def _func_batched(self, func, types, args, kwargs):
        """Ops with per-sample statistics (group_norm): fold all executed blocks into ONE sample
        (reference tensorwrapper.py:600-633; batch size 1 only)."""
        args = list(args)
        x = args[0].as_subclass(torch.Tensor)
        E, C, h, w = x.shape
        folded = x.permute(1, 0, 2, 3).reshape(1, C, E * h * w, 1)
        args[0] = folded
        out = func(*args, **kwargs)
        out = out.reshape(C, E, h, w).permute(1, 0, 2, 3).contiguous()
        return out.as_subclass(TensorWrapper)


def _first_wrapper(items) -> Optional[TensorWrapper]:
    """The block TensorWrapper among the operands if there is one, else the first wrapper."""
    first = None
    for a in items:
        if isinstance(a, TensorWrapper):
            if a._is_blocks:
                return a
            if first is None:
                first = a
        elif isinstance(a, (list, tuple)):
            w = _first_wrapper(a)
            if w is not None:
                if w._is_blocks:
                    return w
                if first is None:
                    first = w
    return first
