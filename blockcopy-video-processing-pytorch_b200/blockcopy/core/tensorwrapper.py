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
  (policy.py:82-94 ``num_exec``) and which is reused here through the version-checked hint of utils/hints.py.
* Tiles may be NCHW or channels_last; kernels are layout-aware (``_C.layout_of``).
"""
from __future__ import annotations

import os
import threading
import warnings
from typing import Any, Callable, Dict, List, Optional, Tuple

import torch

from .. import _C
from ..utils.hints import get_num_exec_hint
from ..utils.profiler import timings

FUSED_CONV = os.environ.get("BLOCKCOPY_FUSED_CONV", "1") != "0"  # route eligible convs on blocks to the tcgen05 implicit-GEMM kernel (bc_conv_igemm)
# defer convs / elementwise ops on blocks and fuse them into epilogues (see _Pending); "0" = op-by-op execution
LAZY_FUSION = os.environ.get("BLOCKCOPY_LAZY", "1") != "0"
FUSED_HEAD = os.environ.get("BLOCKCOPY_FUSED_HEAD", "1") != "0"  # few-channel 1x1 output conv + combine as one kernel (bc_head_1x1)
# side branches (a pre-activation BN+ReLU on materialised tiles feeding a 1x1 conv: the skip bottlenecks of a
# ladder decoder) are issued on a second CUDA stream inside a side_stream_scope, see _SideState
SIDE_STREAM = os.environ.get("BLOCKCOPY_SIDE_STREAM", "1") != "0"
# a deferred elementwise result whose consumer is a padded op is written into that op's plane ONLY; the packed tile
# batch is gathered from the plane if (and when) somebody else reads it
TILELESS = os.environ.get("BLOCKCOPY_TILELESS", "1") != "0"
SIDE_DOWNSAMPLE = os.environ.get("BLOCKCOPY_SIDE_DS", "1") != "0"  # also 1x1 convs on materialised tiles (residual downsamples)
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
        self._index_buf: Optional[torch.Tensor] = None
        self.num_exec = 0
        self.num_total = 0
        self.has_history = prev is not None and not prev._was_reset
        self._was_reset = False
        self.track_transfer_idx = True  # also produce transfer_idx (reference tile protocol compat)
        # persistent planes; slot = call order within a frame
        self._planes: List[torch.Tensor] = prev._planes if prev is not None else []
        self._full: List[torch.Tensor] = prev._full if prev is not None else []
        # persistent tile-batch buffers of side-stream ops, sized for ALL blocks; slot = call order
        self._side_bufs: List[torch.Tensor] = prev._side_bufs if prev is not None else []
        self._plane_cursor = 0
        self._full_cursor = 0
        self._side_cursor = 0
        self._prev_grid_idx = prev._grid_idx if (prev is not None and self.has_history) else None
        if prev is not None:
            prev._planes, prev._full, prev._side_bufs = [], [], []  # ownership moved

    # ------------------------------------------------------------------ grid
    def _process_grid(self, grid: torch.Tensor, meta_prev: Optional["BlockFeatures"] = None) -> None:
        with timings.env("tensorwrapper/process_grid", 10):
            assert grid.dim() == 4 and grid.shape[1] == 1, "grid must be (N,1,GH,GW)"
            hint = get_num_exec_hint(grid)  # ignored when the grid was edited after the count was taken
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
            self._index_buf = buf  # [mapping_exec (G) | transfer_idx (G) | counts (2)]: re-filled in place by graph replays
            self._mapping_exec = mapping[:n_exec]
            self._transfer_idx = transfer[: G - n_exec] if prev_idx is not None else None
            self.num_exec, self.num_total = n_exec, G
            if VERBOSE:
                print(f"TensorWrapper >> GRID {tuple(g.shape)} exec {n_exec}/{G}")

    # ------------------------------------------------------------------ planes
    def _next_plane(self, like: Optional[torch.Tensor], shape, dtype=None, device=None, nhwc=None,
                    zero: bool = False) -> torch.Tensor:
        """Plane of the next padded op in call order; allocated on the first frame.  Layout / dtype
        come from `like` (packed tiles) or from the explicit descriptor."""
        if like is not None:
            dtype, device, nhwc = like.dtype, like.device, _C.layout_of(like) == _C.BC_NHWC
        i = self._plane_cursor
        self._plane_cursor += 1
        if i < len(self._planes):
            plane = self._planes[i]
            if plane.shape != tuple(shape) or plane.dtype != dtype:
                raise AssertionError(
                    f"padded op #{i}: plane {tuple(plane.shape)}/{plane.dtype} does not match this frame's "
                    f"{tuple(shape)}/{dtype}; the model must issue the same ops every frame")
            return plane
        assert not self.has_history or self.num_exec == self.num_total, \
            "No computed features to pop from stack, something seems wrong in the model."
        fmt = torch.channels_last if nhwc else torch.contiguous_format
        plane = torch.empty(shape, dtype=dtype, device=device, memory_format=fmt)
        if zero:
            plane.zero_()
        self._planes.append(plane)
        return plane

    def _next_side_buf(self, shape, dtype, device) -> torch.Tensor:
        """(E,C,BS,BS) channels_last view of the next side-stream buffer in call order.  These outputs must not
        come from the caching allocator: a block recycled there may still be read by kernels the main stream
        has queued, which a side-stream kernel does not wait for."""
        i = self._side_cursor
        self._side_cursor += 1
        E, tail = shape[0], tuple(shape[1:])
        if i < len(self._side_bufs):
            buf = self._side_bufs[i]
            if tuple(buf.shape[1:]) != tail or buf.dtype != dtype or buf.shape[0] < E:
                raise AssertionError(f"side op #{i}: buffer {tuple(buf.shape)}/{buf.dtype} does not match this "
                                     f"frame's {tuple(shape)}/{dtype}; the model must issue the same ops every frame")
        else:
            buf = torch.empty((max(self.num_total, E),) + tail, dtype=dtype, device=device,
                              memory_format=torch.channels_last)
            self._side_bufs.append(buf)
        return buf[:E]

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
        self._side_bufs.clear()
        self._prev_grid_idx = None

    def state_bytes(self) -> int:
        return sum(p.numel() * p.element_size() for p in self._planes + self._full + self._side_bufs)


REUSE_UNOBSERVED_PLANES = os.environ.get("BLOCKCOPY_REUSE_PLANES", "1") != "0"


def _only_the_store_holds(t: torch.Tensor) -> bool:
    """True when the previous frame's combined tensor `t` is referenced by nothing but the feature store (and
    this call chain): no Python variable, no view / alias / autograd node shares its storage.  A non-in-place
    combine may then write into it -- nobody can tell the difference from clone + scatter (reference
    tensorwrapper.py:421-434), and the copy of the unchanged blocks (a full pass over the plane: 67 MB per call for
    the 256-channel planes of Pedestron's head, csp_head.py:137,143,150) disappears.
    Reference counts: the store's list, the caller's local, this function's argument, getrefcount's own = 4; storage
    use count: the tensor + the temporary storage wrapper = 2."""
    import sys

    if not REUSE_UNOBSERVED_PLANES:
        return False
    try:
        return sys.getrefcount(t) <= 4 and torch._C._storage_Use_Count(t.untyped_storage()._cdata) <= 2
    except Exception:
        return False


def _raw(t: torch.Tensor) -> torch.Tensor:
    """Plain-tensor alias of t WITHOUT going through __torch_function__ (so it never materialises)."""
    with torch._C.DisableTorchFunctionSubclass():
        return t.as_subclass(torch.Tensor)


class _Pending:
    """Deferred producer of a block tensor (lazy epilogue fusion).

    A conv2d on blocks, and the elementwise ops the model applies to its result, are not launched
    when the model calls them: the returned TensorWrapper owns its (still unwritten) storage and
    this descriptor.  In-place ReLU / residual add are absorbed into the descriptor; eval-mode
    batch_norm, ReLU, add and per-block bilinear x2 on materialised tiles become an elementwise
    descriptor.  The kernel is launched when the tensor is first needed -- and if the consumer is a
    padded op, the same launch also writes the result into that op's persistent plane (the scatter
    of the north-star design, fused into the producer's epilogue)."""

    __slots__ = ("kind", "conv", "src", "up2x", "residual", "bn", "relu", "stage", "side", "deps", "__weakref__")

    def __init__(self, kind, conv=None, src=None, up2x=False):
        self.kind, self.conv, self.src, self.up2x = kind, conv, src, up2x
        self.residual, self.bn, self.relu, self.stage = None, None, False, 0
        # side: launch on the side stream (its output then lives in a BlockFeatures side buffer).  deps: the
        # "ready" events of ALL tensors the kernel reads, or None when one of them is unknown (the side stream
        # then waits for everything the main stream has queued so far)
        self.side, self.deps = False, None

    def clone(self):
        q = _Pending(self.kind, self.conv, self.src, self.up2x)
        q.residual, q.bn, q.relu, q.stage = self.residual, self.bn, self.relu, self.stage
        q.deps = None if self.deps is None else list(self.deps)
        return q

    def reads(self, t: torch.Tensor) -> bool:
        ptr = t.data_ptr()
        for u in (self.src, self.residual, self.conv["src"] if self.conv else None):
            if u is not None and u.data_ptr() == ptr:
                return True
        return False


class _SideStateT(threading.local):
    """Second CUDA stream for side branches of the wrapped CNN (active inside a side_stream_scope only).

    In a ladder decoder the skip bottlenecks (BN -> ReLU -> 1x1 conv on an encoder feature) depend on nothing
    but that encoder feature, yet the model's call order puts them between decoder kernels that are each too
    small to fill the GPU.  Inside a scope such a unit is launched on the side stream, waiting only for the
    "ready" event of the feature it reads, so that -- eagerly, and as parallel branches of a captured CUDA
    graph -- it runs next to the deeper encoder layers / the pyramid pooling / the decoder.  Rules that keep
    this race-free:
      * outputs of side kernels live in persistent BlockFeatures side buffers, never in freshly recycled
        allocator blocks (see BlockFeatures._next_side_buf);
      * everything a side kernel reads is kept alive until the scope's join;
      * a main-stream launch that reads a side output first waits for its event (`sync_main`);
      * in-place torch ops on blocks and the end of the scope join the side stream completely.
    All of it is PER PYTHON THREAD (threading.local): wrappers driven from different threads on different CUDA
    streams do not share scopes, events or side streams (re-entrant w.r.t. distinct streams).
    """

    def __init__(self):
        self.active = False
        self.streams: Dict[int, "torch.cuda.Stream"] = {}
        self.done: Dict[int, "torch.cuda.Event"] = {}   # data_ptr of a side output -> event recorded after its kernel
        self.keep: List[Any] = []                       # tensors read / written by side kernels that were not joined yet
        self.last: Optional["torch.cuda.Event"] = None
        self.splitk_ws: Optional[torch.Tensor] = None   # split-K scratch of side-stream convs (None: the per-stream one)
        self.launches = 0                               # kernels issued on the side stream so far (diagnostics / tests)

    def stream(self, device) -> "torch.cuda.Stream":
        st = self.streams.get(device.index)
        if st is None:
            st = self.streams[device.index] = torch.cuda.Stream(device=device, priority=-1)
        return st

    def sync_main(self, t: Optional[torch.Tensor]) -> None:
        """The current (main) stream is about to read `t`: wait for its side-stream producer, if any."""
        if t is not None and self.done:
            ev = self.done.pop(t.data_ptr(), None)
            if ev is not None:
                torch.cuda.current_stream().wait_event(ev)

    def run(self, fn: Callable, keep: Tuple = ()):
        """Run fn() -- kernels whose results the main stream needs only after the scope's join -- on the side
        stream, after everything the main stream has queued so far.  `keep`: tensors fn reads."""
        if not self.active:
            return fn()
        main = torch.cuda.current_stream()
        side = self.stream(main.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            out = fn()
            ev = torch.cuda.Event()
            ev.record(side)
        self.last = ev
        self.keep.append(tuple(keep) + (out,))
        return out

    def join(self) -> None:
        if self.last is not None:
            torch.cuda.current_stream().wait_event(self.last)
        self.last = None
        self.done.clear()
        self.keep.clear()


_SideState = _SideStateT()


class side_stream_scope:
    """`with side_stream_scope():` -- side branches issued inside run on the side stream; the exit joins it.
    Used by BlockCopyModel around a whole frame in CUDA-graph mode (the branches become parallel graph nodes)."""

    def __init__(self, splitk_ws: Optional[torch.Tensor] = None):
        self.ws = splitk_ws  # scratch for split-K convs on the side stream (must not be the main stream's)

    def __enter__(self):
        self.prev = (_SideState.active, _SideState.splitk_ws)
        _SideState.active, _SideState.splitk_ws = SIDE_STREAM, self.ws
        return self

    def __exit__(self, *exc):
        _SideState.join()
        _SideState.active, _SideState.splitk_ws = self.prev
        return False


def run_on_side_stream(fn: Callable, keep: Tuple = ()):
    """See _SideState.run (no-op wrapper outside a side_stream_scope)."""
    return _SideState.run(fn, keep)


def _is_dense4(t: torch.Tensor) -> bool:
    return t.dim() == 4 and (t.is_contiguous() or t.is_contiguous(memory_format=torch.channels_last))


_META_METHODS = {"dim", "size", "stride", "numel", "is_contiguous", "element_size", "storage_offset",
                 "is_floating_point", "is_complex", "ndimension", "nelement", "get_device", "__len__"}
_META_PROPS = {"shape", "dtype", "device", "ndim", "is_cuda", "requires_grad", "layout", "names", "is_leaf",
               "grad_fn", "is_sparse", "is_quantized", "is_meta"}
_BN_CACHE = _C.TensorCache(4096)


def _bn_params(running_mean, running_var, weight, bias, eps):
    """fp32 (mean, invstd, weight, shift) of an eval-mode batch norm, cached per parameter version."""
    src = (running_mean, running_var, weight, bias)
    hit = _BN_CACHE.get(src, (float(eps),))
    if hit is None:
        with torch.no_grad():
            hit = (_raw(running_mean).detach().float().contiguous(),
                   torch.rsqrt(_raw(running_var).detach().float() + eps).contiguous(),
                   None if weight is None else _raw(weight).detach().float().contiguous(),
                   None if bias is None else _raw(bias).detach().float().contiguous())
        _BN_CACHE.put(src, (float(eps),), hit)
    return hit


_GN_WS: Dict[Tuple[int, int], torch.Tensor] = {}
_F32_CACHE = _C.TensorCache(1024)


def _gn_workspace(device: torch.device) -> torch.Tensor:
    """bc_gn_stats scratch (zeroed once), private to a (device, CUDA stream) pair."""
    key = (device.index, int(torch.cuda.current_stream(device).cuda_stream))
    ws = _GN_WS.get(key)
    if ws is None:
        ws = _GN_WS[key] = torch.zeros(_C.GN_STATS_WORKSPACE, dtype=torch.uint8, device=device)
    return ws


def _f32_vector(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """fp32 contiguous copy of a per-channel parameter vector, cached per parameter version."""
    if t is None:
        return None
    hit = _F32_CACHE.get((t,))
    if hit is None:
        with torch.no_grad():
            hit = _F32_CACHE.put((t,), (), _raw(t).detach().float().contiguous())
    return hit


def _dense(t: torch.Tensor) -> torch.Tensor:
    """Plain, dense (NCHW or channels_last) view/copy of t for the kernels (launches a deferred
    producer first: Tensor.as_subclass is not routed through __torch_function__)."""
    if isinstance(t, TensorWrapper):
        if t._pending is not None:
            t._materialize()
        t._ensure_tiles()
    t = t.as_subclass(torch.Tensor)
    _SideState.sync_main(t)
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
    _pending: Optional[_Pending] = None  # deferred producer, see _Pending
    _ready = None  # CUDA event recorded after the kernel that wrote this tensor (inside a side_stream_scope)
    _in_plane = None  # the plane whose executed cells hold this tensor while its own tile batch is unwritten

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
            slot, prev = feats._next_full()
            if not inplace and prev is not None and tuple(prev.shape) == shape and _only_the_store_holds(prev):
                inplace = True  # the old tensor is unobservable: update it instead of clone + scatter
            head = self._pending if self._pending is not None and self._pending.kind == "head" else None
            if head is not None and (prev is None or (tuple(prev.shape) == shape and _is_dense4(prev))):
                # output head + combine in one kernel: tiles and the dense output from the same pass
                c = head.conv
                self._pending = None
                _LIVE_PENDING.pop(id(self), None)
                if prev is None:
                    assert E == grid_idx.numel(), "first frame must execute every block"
                    out, src_prev = torch.empty(shape, dtype=self.dtype, device=self.device), None
                elif inplace:
                    out, src_prev = prev, None          # cells that are not executed simply keep their content
                else:
                    out, src_prev = torch.empty_like(prev), prev
                _SideState.sync_main(c["src"])
                _C.head_1x1(c["src"], c["w"], c["bias"], c["bn"], c["relu"], tiles_out=_raw(self), dense_out=out,
                            dense_prev=src_prev, grid_idx=grid_idx, mapping_exec=mapping)
                tiles, done = None, True
            elif (self._pending is not None and inplace and prev is not None and tuple(prev.shape) == shape
                    and prev.is_contiguous(memory_format=torch.channels_last) and not prev.is_contiguous()):
                self._materialize(plane_out=prev)  # producer epilogue writes the plane: no scatter kernel
                tiles, out, done = None, prev, True
            else:
                tiles, done = _dense(self), False
            if done:
                pass
            elif prev is not None:
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
        N, _, GH, GW = feats._grid_idx.shape
        if self._pending is not None:
            E, C, BS, _ = self.shape
            plane = feats._next_plane(None, (N, C, GH * BS, GW * BS), self.dtype, self.device, True)
            if _C.layout_of(plane) == _C.BC_NHWC:
                self._materialize(plane_out=plane)
            else:
                self._materialize()
                _C.scatter(_match_layout(_dense(self), plane), plane, feats._mapping_exec, E)
        else:
            tiles = _dense(self)
            E, C, BS, _ = tiles.shape
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
        op = getattr(func, "__name__", "")
        # shape / dtype queries never need the data (and must not launch deferred producers)
        if op in _META_METHODS or (op == "__get__" and getattr(getattr(func, "__self__", None), "__name__", "")
                                   in _META_PROPS):
            return super().__torch_function__(func, types, args, kwargs)
        src = _first_wrapper(args)
        if src is None and kwargs:
            src = _first_wrapper(tuple(kwargs.values()))
        if src is None or not src._is_blocks:
            if LAZY_FUSION and src is not None and op in ("relu", "relu_", "batch_norm"):
                out = src._lazy_dispatch(op, args, kwargs)  # dense planes of a noblocks region: BN+ReLU in one kernel
                if out is not NotImplemented:
                    return out
            _materialize_all(args, kwargs)
            out = cls._dense_dispatch(func, args, kwargs) if src is not None else None
            if out is None:
                out = super().__torch_function__(func, types, args, kwargs)
            if src is not None and isinstance(out, TensorWrapper) and out is not src and out._features is None:
                out._inherit(src)
            return out

        if LAZY_FUSION:
            out = src._lazy_dispatch(op, args, kwargs)
            if out is not NotImplemented:
                return out
        if op not in OPS["PADDED"]:  # padded ops materialise their input themselves (dual write into the plane)
            _materialize_all(args, kwargs)
        if op.endswith("_") or op.startswith("__i") or kwargs.get("out", None) is not None:
            _flush_readers_of(args[0] if args else None)  # about to mutate: run deferred readers first

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
            out = src._try_fused_conv_transpose(args, kwargs) if op == "conv_transpose2d" else None
            if out is None:
                out = super().__torch_function__(func, types, args, kwargs)

        if isinstance(out, TensorWrapper) and out is not src and out._features is None:
            out._inherit(src)
        return out

    # ------------------------------------------------------------------ lazy epilogue fusion
    def _materialize(self, plane_out: Optional[torch.Tensor] = None) -> bool:
        """Launch the deferred producer of this tensor (if any).  With `plane_out` (the NHWC plane of
        the padded op about to consume it) the same kernel also scatters the result into the plane.
        Returns True if a kernel ran (and therefore wrote `plane_out`)."""
        p = self._pending
        if p is None:
            return False
        self._pending = None
        _LIVE_PENDING.pop(id(self), None)
        out = _raw(self)
        reads = (p.src, p.residual, p.conv["src"] if p.conv else None)
        if p.side and _SideState.active and plane_out is None:
            side, main = _SideState.stream(out.device), torch.cuda.current_stream()
            if p.deps is None:
                side.wait_stream(main)
            else:
                for ev in p.deps:
                    side.wait_event(ev)
            with torch.cuda.stream(side), _C.splitk_workspace_scope(_SideState.splitk_ws):
                ran = self._launch(p, out, None, True)
                ev = torch.cuda.Event()
                ev.record(side)
            _SideState.done[out.data_ptr()] = _SideState.last = self._ready = ev
            _SideState.launches += 1
            _SideState.keep.append((out,) + reads)
            return ran
        for u in reads:
            _SideState.sync_main(u)
        # Tile-less when the consumer is a padded op (it reads the plane).  For convs only without an absorbed
        # residual: a block output is also read as the next block's identity, and gathering it back from the plane
        # would cost more than the epilogue's second store stream; conv -> ReLU -> conv chains and the stem are not.
        tileless = TILELESS and plane_out is not None and \
            (p.kind == "ew" or (p.kind in ("conv", "stem") and p.residual is None))
        ran = self._launch(p, out, plane_out, write_tiles=not tileless)
        if tileless:
            self._in_plane = plane_out
        if _SideState.active:
            self._ready = torch.cuda.Event()
            self._ready.record()
        return ran

    def _ensure_tiles(self) -> None:
        """The packed tile batch is about to be read: if its producer wrote only the consumer's plane, gather it."""
        plane = self._in_plane
        if plane is None:
            return
        self._in_plane = None
        _C.gather(_raw(self), plane, self._features._mapping_exec, self.shape[0])
        if _SideState.active:  # side-stream readers must wait for the gather, not for the producer
            self._ready = torch.cuda.Event()
            self._ready.record()

    def _launch(self, p: _Pending, out: torch.Tensor, plane_out: Optional[torch.Tensor], write_tiles: bool = True) -> bool:
        feats = self._features
        if p.kind == "conv":
            c = p.conv
            _C.conv_igemm(out, c["src"], c["w"], c["bias"], p.residual, c["mapping"], c["E"], c["BS_in"],
                          c["stride"], c["pad"], relu=p.relu, plane_out=plane_out, out_mapping=feats._mapping_exec,
                          write_tiles=write_tiles)
        elif p.kind == "stem":
            c = p.conv
            _C.conv_stem(out, c["src"], c["w"], c["bias"], c["mapping"], c["E"], relu=p.relu, plane_out=plane_out,
                         write_tiles=write_tiles)
        elif p.kind == "head":
            c = p.conv
            _C.head_1x1(c["src"], c["w"], c["bias"], c["bn"], c["relu"], tiles_out=out)
            return False  # never writes a consumer's plane: the caller scatters
        elif p.kind == "pool":
            c = p.conv
            _C.maxpool_halo(out, c["src"], c["mapping"], c["E"], c["BS_in"], c["k"], c["stride"], c["pad"],
                            plane_out=plane_out)
        else:
            _C.ew_fused(out if write_tiles else None, p.src, p.residual, p.bn, p.relu, p.up2x, plane_out,
                        feats._mapping_exec if plane_out is not None else None)
        return True

    def _new_pending(self, shape, pending: _Pending, storage: Optional[torch.Tensor] = None) -> "TensorWrapper":
        t = storage if storage is not None else \
            torch.empty(shape, dtype=self.dtype, device=self.device, memory_format=torch.channels_last)
        t = t.as_subclass(TensorWrapper)._inherit(self)
        t._pending = pending
        _LIVE_PENDING[id(t)] = __import__("weakref").ref(t, lambda _r, k=id(t): _LIVE_PENDING.pop(k, None))
        return t

    def _tiles_nhwc(self) -> Optional[torch.Tensor]:
        """Materialised channels_last tiles of this block tensor (copy only if the layout differs)."""
        self._materialize()
        self._ensure_tiles()
        t = _raw(self)
        if t.dim() != 4:
            return None
        return t if t.is_contiguous(memory_format=torch.channels_last) else t.contiguous(memory_format=torch.channels_last)

    def _lazy_dispatch(self, op, args, kwargs):
        """Absorb ReLU / add / eval batch_norm / bilinear x2 on fp16 blocks into deferred descriptors.
        Returns NotImplemented for everything else (which then runs op by op on materialised tiles)."""
        x = args[0] if args else None
        if not isinstance(x, TensorWrapper) or not _C.lazy_supported(x):
            return NotImplemented
        if not x._is_blocks and op not in ("relu", "relu_", "batch_norm"):
            return NotImplemented
        if op in ("relu", "relu_"):
            inplace = op == "relu_" or bool(kwargs.get("inplace", args[1] if len(args) > 1 else False))
            return x._absorb("relu", None, inplace)
        if op in ("__iadd__", "add_", "__add__", "add"):
            other = args[1] if len(args) > 1 else kwargs.get("other")
            if kwargs.get("alpha", 1) != 1 or not isinstance(other, TensorWrapper) or not other._is_blocks \
                    or other.shape != x.shape or other.dtype != x.dtype or len(args) > 2:
                return NotImplemented
            return x._absorb("add", other, op in ("__iadd__", "add_"))
        if op == "batch_norm":
            training = kwargs.get("training", args[5] if len(args) > 5 else False)
            rm = args[1] if len(args) > 1 else kwargs.get("running_mean")
            rv = args[2] if len(args) > 2 else kwargs.get("running_var")
            if training or rm is None or rv is None:
                return NotImplemented
            w = kwargs.get("weight", args[3] if len(args) > 3 else None)
            b = kwargs.get("bias", args[4] if len(args) > 4 else None)
            eps = kwargs.get("eps", args[7] if len(args) > 7 else 1e-5)
            return x._absorb("bn", _bn_params(rm, rv, w, b, eps), False)
        if op == "interpolate":
            if kwargs.get("mode", "nearest") != "bilinear" or kwargs.get("align_corners", None) or \
                    kwargs.get("antialias", False):
                return NotImplemented
            E, C, h, w = x.shape
            size, sf = kwargs.get("size", args[1] if len(args) > 1 else None), kwargs.get("scale_factor", None)
            if size is not None:
                size = (size, size) if isinstance(size, int) else tuple(size)
                ok = size == (2 * h, 2 * w)
            else:
                sf = (sf, sf) if isinstance(sf, (int, float)) else tuple(sf or ())
                ok = sf == (2, 2) or sf == (2.0, 2.0)
            if not ok or h != w:
                return NotImplemented
            a = x._tiles_nhwc()
            return x._new_pending((E, C, 2 * h, 2 * w), _Pending("ew", src=a, up2x=True))
        return NotImplemented

    def _absorb(self, what: str, operand, inplace: bool):
        """Add one op to a deferred descriptor.  Canonical order inside a kernel: add -> bn -> relu."""
        stage = {"add": 1, "bn": 2, "relu": 3}[what]
        p = self._pending
        fits = p is not None and p.stage < stage and not (what == "bn" and p.kind in ("conv", "stem")) \
            and not (what == "add" and p.kind == "stem")
        if inplace:
            if not fits:
                return NotImplemented  # materialised (or chain out of order): plain torch in-place op
            target, q = self, p
        else:
            if fits:
                q = p.clone()  # `self` stays deferred and untouched; the new tensor re-derives from the same sources
            else:
                a = self._tiles_nhwc()
                if a is None:
                    return NotImplemented
                q = _Pending("ew", src=a)
                q.deps = _deps_of(self, a)
            target = self._new_pending(tuple(self.shape), q)
        if what == "add":
            r = operand._tiles_nhwc()
            if r is None:
                return NotImplemented
            q.residual = r
            more = _deps_of(operand, r)
            q.deps = None if (q.deps is None or more is None) else q.deps + more
        elif what == "bn":
            q.bn = operand
        else:
            q.relu = True
        q.stage = stage
        return target

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
        by block index (halo included); bias / residual / ReLU run in the epilogue.  Returns a
        DEFERRED tensor (see _Pending) or None if the conv is outside the kernel's envelope (then
        the generic gather-halo + torch path runs)."""
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
        if bias is not None and (bias.dtype != x.dtype or not bias.is_contiguous()):
            return None
        feats = self._features
        if LAZY_FUSION and FUSED_HEAD and _C.head_supported(x.dtype, weight, stride, padding, dilation, a["groups"]) \
                and _C.lazy_supported(x):
            # output head (few-channel 1x1 conv): one streaming kernel that also absorbs a preceding eval
            # batch_norm + ReLU and, when the consumer is combine(), writes the dense output directly
            q = x._pending
            if q is not None and q.kind == "ew" and not q.up2x and q.residual is None and q.src is not None:
                src, bn, relu = q.src, q.bn, q.relu  # x itself is only launched if someone else needs it
            else:
                src, bn, relu = x._tiles_nhwc(), None, False
            if src is not None:
                w2d = weight.detach().reshape(weight.shape[0], Cin)
                if not w2d.is_contiguous():
                    w2d = w2d.contiguous()
                pend = _Pending("head", conv=dict(src=src, w=w2d, bias=bias, E=E, bn=bn, relu=relu))
                return x._new_pending((E, weight.shape[0], BS, BS), pend)
        if _C.stem_supported(x.dtype, weight, BS, stride, padding, dilation, a["groups"]):
            # ResNet stem: 7x7/s2 on 3 channels as a 4x4/s1 implicit GEMM over a space-to-depth plane
            N, _, GH, GW = feats._grid_idx.shape
            if feats._plane_cursor < len(feats._planes) and feats._planes[feats._plane_cursor].shape[1] != 16:
                return None
            plane = feats._next_plane(None, (N, 16, GH * BS // 2, GW * BS // 2 + 2 * _C.STEM_XPAD), x.dtype, x.device,
                                      True, zero=True)  # zero pad columns = the conv's horizontal frame border
            _C.stem_pack(plane, _dense(x).contiguous(), feats._mapping_exec, E)
            pend = _Pending("stem", conv=dict(src=plane, w=_C.pack_stem_weight(weight), bias=bias,
                                              mapping=feats._mapping_exec, E=E))
            out = x._new_pending((E, weight.shape[0], BS // 2, BS // 2), pend)
            if not LAZY_FUSION:
                out._materialize()
            return out
        if not _C.conv_supported(x.dtype, weight, BS, stride, padding, dilation, a["groups"]):
            return None
        if padding > 0 and feats._plane_cursor < len(feats._planes) and \
                _C.layout_of(feats._planes[feats._plane_cursor]) != _C.BC_NHWC:
            return None  # this op's plane was created NCHW on the first frame: stay on the generic path
        w = weight.detach()
        if not w.is_contiguous(memory_format=torch.channels_last):
            w = w.contiguous(memory_format=torch.channels_last)
        Cout, BSo = weight.shape[0], BS // stride
        if padding > 0:
            N, _, GH, GW = feats._grid_idx.shape
            plane = feats._next_plane(None, (N, Cin, GH * BS, GW * BS), x.dtype, x.device, True)
            with timings.env("tensorwrapper/transfer", 10):
                if not x._materialize(plane_out=plane):  # producer epilogue wrote the plane, else scatter now
                    x._ensure_tiles()
                    _SideState.sync_main(_raw(x))
                    _C.scatter(_raw(x).contiguous(memory_format=torch.channels_last), plane, feats._mapping_exec, E)
            conv = dict(src=plane, w=w, bias=bias, mapping=feats._mapping_exec, E=E, BS_in=BS, stride=stride,
                        pad=padding)
            pend, storage = _Pending("conv", conv=conv), None
        else:
            q = x._pending
            # side branches: (a) a pre-activation unit on materialised tiles feeding a 1x1 conv (skip bottleneck);
            # (b) a 1x1 conv on tiles whose producer is known (residual downsample: runs next to the block's
            # first 3x3 conv).  If the consumer turns out to be a padded op, the conv still runs on the main stream.
            pre = q is not None and q.kind == "ew" and not q.up2x and q.residual is None
            side = _SideState.active and LAZY_FUSION and (pre or (SIDE_DOWNSAMPLE and q is None and x._ready is not None))
            if side and pre:
                # x is re-pointed at a persistent side buffer before its kernel is launched on the side stream
                with torch._C.DisableTorchFunctionSubclass():
                    x.set_(feats._next_side_buf(tuple(x.shape), x.dtype, x.device))
                q.side = True
            src = x._tiles_nhwc()
            conv = dict(src=src, w=w, bias=bias, mapping=None, E=E, BS_in=BS, stride=stride, pad=0)
            pend, storage = _Pending("conv", conv=conv), None
            if side:
                pend.side, pend.deps = True, _deps_of(x, src)
                storage = feats._next_side_buf((E, Cout, BSo, BSo), x.dtype, x.device)
        out = x._new_pending((E, Cout, BSo, BSo), pend, storage)
        if not LAZY_FUSION:
            out._materialize()
        return out

    def _try_fused_conv_transpose(self, args, kwargs):
        """conv_transpose2d on fp16 blocks (a pass-through op in the reference, tensorwrapper.py:519-520: every tile
        is its own sample, zeros beyond the tile edge): bc_conv_igemm over the tile batch viewed as E one-block
        frames, with s*s*Cout output channels (one group per output phase), then bc_depth_to_space.  None: outside
        the envelope (torch runs the op on the tile batch)."""
        names = ("input", "weight", "bias", "stride", "padding", "output_padding", "groups", "dilation")
        a = dict(bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1)
        a.update(zip(names, args))
        a.update(kwargs)
        one = lambda v: v if isinstance(v, int) else (v[0] if len(set(v)) == 1 else None)  # noqa: E731
        stride, padding, opad, dil = one(a["stride"]), one(a["padding"]), one(a["output_padding"]), one(a["dilation"])
        x, weight, bias = a["input"], a["weight"], a["bias"]
        if None in (stride, padding, opad, dil) or not isinstance(x, TensorWrapper) or not x._is_blocks \
                or isinstance(weight, TensorWrapper):
            return None
        if weight.dim() != 4 or x.dim() != 4 or x.shape[0] == 0 or weight.shape[0] != x.shape[1] or x.shape[2] != x.shape[3]:
            return None
        if not _C.deconv_supported(x.dtype, weight, x.shape[2], stride, padding, opad, dil, a["groups"]):
            return None
        if bias is not None and (bias.dtype != x.dtype or not bias.is_cuda or isinstance(bias, TensorWrapper)):
            return None
        tiles = x._tiles_nhwc()
        if tiles is None:
            return None
        E, Cin, h, _ = tiles.shape
        Cout = weight.shape[1]
        wp, bp = _C.pack_deconv_weight(weight, bias, stride)
        _SideState.sync_main(tiles)
        phases = torch.empty((E, stride * stride * Cout, h, h), dtype=x.dtype, device=x.device,
                             memory_format=torch.channels_last)
        _C.conv_igemm(phases, tiles, wp, bp, None, None, E, h, 1, wp.shape[2] // 2)
        out = torch.empty((E, Cout, stride * h, stride * h), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
        _C.depth_to_space(out, phases, stride)
        return out.as_subclass(TensorWrapper)._inherit(x)

    def _try_fused_pool(self, args, kwargs):
        """max_pool2d with padding on fp16 NHWC blocks through bc_maxpool_halo (reads the op's plane,
        halo included); deferred like the convs so that its result lands in the next op's plane."""
        names = ("input", "kernel_size", "stride", "padding", "dilation", "ceil_mode", "return_indices")
        a = dict(stride=None, padding=0, dilation=1, ceil_mode=False, return_indices=False)
        a.update(zip(names, args))
        a.update(kwargs)
        one = lambda v: v if isinstance(v, int) else (v[0] if len(set(v)) == 1 else None)  # noqa: E731
        k, padding, dilation = one(a["kernel_size"]), one(a["padding"]), one(a["dilation"])
        stride = k if a["stride"] is None or a["stride"] == [] else one(a["stride"])
        x = a["input"]
        if None in (k, padding, dilation, stride) or dilation != 1 or a["ceil_mode"] or a["return_indices"] \
                or padding <= 0 or not isinstance(x, TensorWrapper) or not x._is_blocks or not _C.lazy_supported(x):
            return None
        E, C, BS, _ = x.shape
        if E == 0 or BS % stride or (BS + 2 * padding - k) // stride + 1 != BS // stride:
            return None
        feats = self._features
        if feats._plane_cursor < len(feats._planes) and _C.layout_of(feats._planes[feats._plane_cursor]) != _C.BC_NHWC:
            return None
        N, _, GH, GW = feats._grid_idx.shape
        plane = feats._next_plane(None, (N, C, GH * BS, GW * BS), x.dtype, x.device, True)
        if not x._materialize(plane_out=plane):
            x._ensure_tiles()
            _SideState.sync_main(_raw(x))
            _C.scatter(_raw(x).contiguous(memory_format=torch.channels_last), plane, feats._mapping_exec, E)
        pend = _Pending("pool", conv=dict(src=plane, mapping=feats._mapping_exec, E=E, BS_in=BS, k=k, stride=stride,
                                          pad=padding))
        pend.stage = 3  # nothing is absorbed into a pooling kernel
        out = x._new_pending((E, C, BS // stride, BS // stride), pend)
        if not LAZY_FUSION:
            out._materialize()
        return out

    def _func_replace_padding(self, func, types, args, kwargs):
        """Padded op: take the padding from neighbouring blocks instead of zeros, then run the op
        itself with padding 0 (reference: _func_replace_paddding, tensorwrapper.py:529-575)."""
        if BLOCKPAD_WITH_ZEROES:
            _materialize_all(args, kwargs)
            return super().__torch_function__(func, types, args, kwargs)
        op = func.__name__
        if FUSED_CONV and op in ("conv2d", "max_pool2d"):
            fused = self._try_fused_conv(args, kwargs) if op == "conv2d" else self._try_fused_pool(args, kwargs)
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
        _materialize_all(args, kwargs)
        with timings.env("tensorwrapper/pad_func0", 11):
            return super().__torch_function__(func, types, tuple(args), kwargs)

    def _func_interpolate(self, func, types, args, kwargs):
        """Per-block interpolation: every tile is one batch element, so bilinear taps clamp at the
        BLOCK border exactly like the reference's trilinear rewrite (tensorwrapper.py:577-598)."""
        return super().__torch_function__(func, types, args, kwargs)

    def _try_fused_group_norm(self, args, kwargs):
        """group_norm on fp16 blocks: statistics over ALL executed blocks in one kernel (bc_gn_stats, the reference's
        fold of the tile batch into one sample), normalisation + affine as a deferred elementwise stage (bc_ew_fused)
        that absorbs a following ReLU and writes the consumer's plane.  None: outside the envelope."""
        names = ("input", "num_groups", "weight", "bias", "eps")
        a = dict(weight=None, bias=None, eps=1e-5)
        a.update(zip(names, args))
        a.update(kwargs)
        x, groups = a["input"], a["num_groups"]
        if not LAZY_FUSION or not isinstance(x, TensorWrapper) or not x._is_blocks or not isinstance(groups, int):
            return None
        if not _C.lazy_supported(x) or not _C.gn_supported(x, groups) or x.shape[0] == 0:
            return None
        w, b = a["weight"], a["bias"]
        for t in (w, b):
            if t is not None and (isinstance(t, TensorWrapper) or t.dim() != 1 or t.shape[0] != x.shape[1] or not t.is_cuda):
                return None
        tiles = x._tiles_nhwc()
        if tiles is None:
            return None
        C = x.shape[1]
        _SideState.sync_main(tiles)
        stats = torch.empty((2, C), dtype=torch.float32, device=tiles.device)
        _C.gn_stats(tiles, groups, float(a["eps"]), stats[0], stats[1], _gn_workspace(tiles.device))
        q = _Pending("ew", src=tiles)
        q.bn = (stats[0], stats[1], _f32_vector(w), _f32_vector(b))
        q.stage = 2
        return x._new_pending(tuple(x.shape), q)

    def _func_batched(self, func, types, args, kwargs):
        """Ops with per-sample statistics (group_norm): fold all executed blocks into ONE sample
        (reference tensorwrapper.py:600-633; batch size 1 only)."""
        if getattr(func, "__name__", "") == "group_norm":
            fused = self._try_fused_group_norm(args, kwargs)
            if fused is not None:
                return fused
            _materialize_all(args, kwargs)
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


class _LivePending(threading.local):
    """id -> weakref of deferred tensors that have not been launched yet; per Python thread."""

    def __init__(self):
        self.d: Dict[int, Any] = {}

    def pop(self, key, default=None):
        return self.d.pop(key, default)

    def __setitem__(self, key, value):
        self.d[key] = value

    def __bool__(self):
        return bool(self.d)

    def values(self):
        return self.d.values()


_LIVE_PENDING = _LivePending()


def _materialize_all(args, kwargs=None):
    """Launch the deferred producers of every TensorWrapper among the operands."""
    for a in args:
        if isinstance(a, TensorWrapper):
            if a._pending is not None:
                a._materialize()
            a._ensure_tiles()
            if _SideState.done:
                _SideState.sync_main(_raw(a))
        elif isinstance(a, (list, tuple)):
            _materialize_all(a)
    if kwargs:
        _materialize_all(tuple(kwargs.values()))


def _deps_of(t: "TensorWrapper", raw: torch.Tensor):
    """[ready event] of the materialised block tensor `t` whose tiles `raw` a descriptor is about to read, or
    None when there is none (materialised outside a side_stream_scope, or `raw` is a re-laid-out copy)."""
    ev = t._ready
    if ev is None or raw.data_ptr() != _raw(t).data_ptr():
        return None
    return [ev]


def _flush_readers_of(t):
    """An in-place op is about to change `t`: deferred tensors that still have to READ it go first."""
    if _SideState.last is not None:
        _SideState.join()  # side kernels may still be reading it
    if not isinstance(t, torch.Tensor) or not _LIVE_PENDING:
        return
    raw = _raw(t)
    for ref in list(_LIVE_PENDING.values()):
        h = ref()
        if h is not None and h is not t and h._pending is not None and h._pending.reads(raw):
            h._materialize()
