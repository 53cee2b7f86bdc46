"""BlockCopyModel -- per-clip stateful wrapper around a task CNN (reference core/blockcopy.py:7-122).

Per frame: the policy picks the blocks to execute; the input is split into those blocks; the
unmodified base model runs on the packed blocks (every torch call is intercepted by
TensorWrapper); the result is combined with the previous frame's output; the policy is
optimised online.  State machine, ``policy_meta`` keys and the ``num_exec == 0`` short cut are
the reference's.
"""
from __future__ import annotations

import functools

import torch
import torch.nn as nn

from ..utils.hints import get_num_exec_hint, set_num_exec_hint
from ..utils.profiler import timings
from .frame import U8Frame, as_tensor
from .tensorwrapper import TensorWrapper, run_on_side_stream, side_stream_scope, to_tensorwrapper


class BlockCopyModel(nn.Module):
    """Wraps ``base_model`` for block-sparse execution with temporal feature propagation.

    settings: the ``block_*`` dict produced by ``add_argparser_arguments`` (all keys are read).
    Optional extra keys (absent => reference behaviour):
      block_channels_last (bool, default True): store conv weights channels_last so that packed
          tiles and planes are NHWC, the layout the sm_100a kernels are written for.
      block_cuda_graphs (bool, default False): capture the block-sparse frame (index compaction, the gather of
          the executed input blocks, every layer, the combines, the copy into the output buffer) into one CUDA
          graph per executed-block count and replay it; the nodes that read the caller's grid and frame or write
          the output buffer are re-pointed before each replay (``_capture_whole_frame``; for ``U8Frame`` inputs and
          grids that need a conversion those steps run eagerly in front of the graph instead); side branches of
          the model (skip bottlenecks, residual downsamples, the frame_state scatter) are parallel branches of
          the graph.  The Python interception layer then runs only while
          capturing.  Differences to the eager mode: feature planes persist across ``reset_temporal`` (the
          first frame of a clip rewrites all of them), and the returned output tensor is one of two
          alternating buffers -- it stays valid until the next-but-one call (clone it to keep it longer).
      block_policy_fused (bool, default True): ``rl_*`` policies run the policy net's trunk on this repo's
          kernels on frames that are not followed by a policy update (policy/fused_net.py) and step the
          optimiser with one kernel (policy/fused_optim.py).  False: torch / cuDNN throughout.
      block_policy_shared (bool, default False): one policy for all ranks (gradient all-reduce).
      block_policy_device_sampling (bool, default True): ``rl_*`` policies draw the grid and round the executed-block
          count up on the device (bc_sample_grid).  False: the reference's host procedure (``random.sample``), which
          reproduces the reference's masks under ``random.seed``.
    """

    def __init__(self, base_model: nn.Module, settings: dict):
        super().__init__()
        from ..policy.policy import build_policy_from_settings

        self.is_blockcopy_manager = True  # marks the module that owns the temporal state
        self.base_model = base_model
        self.policy = build_policy_from_settings(settings)
        self.block_temporal_features = None
        self.reset_temporal()
        self.train_interval = settings["block_train_interval"]
        if settings.get("block_channels_last", True):
            self.base_model.to(memory_format=torch.channels_last)
        self._graphs = _GraphState() if settings.get("block_cuda_graphs", False) else None
        net = getattr(self.policy, "net", None)
        if self._graphs is not None and net is not None and hasattr(net, "use_cuda_graphs"):
            net.use_cuda_graphs = True  # policy trunk forward / backward as graph replays too
        if net is not None and hasattr(net, "channels_last") and settings.get("block_channels_last", True):
            net.channels_last = True  # fp32 policy trunk on cuDNN's NHWC kernels: forward 0.75 -> 0.57 ms, backward 1.6 -> 0.96 ms
            net.to(memory_format=torch.channels_last)

    def load_state_dict(self, state_dict, strict: bool = True):
        """Checkpoints are base-model checkpoints (reference core/blockcopy.py:30-32)."""
        return self.base_model.load_state_dict(state_dict, strict=strict)

    def reset_temporal(self):
        """Forget all temporal state; call at the start of every clip."""
        self.clip_length = 0
        if getattr(self, "_graphs", None) is not None and self.block_temporal_features is not None:
            # captured graphs hold the planes' addresses: keep them, only forget their content
            self.block_temporal_features.mark_reset()
        else:
            if self.block_temporal_features:
                self.block_temporal_features.clear()
            self.block_temporal_features = None
        self.policy_meta = {"inputs": None, "outputs": None, "outputs_prev": None}

    def forward(self, inputs, **kwargs):
        return self._forward_blockcopy(inputs, **kwargs)

    def _forward_blockcopy(self, inputs, **kwargs):
        self.clip_length += 1
        meta = self.policy_meta
        meta["inputs"] = inputs

        with timings.env("blockcopy/policy_forward", 3):
            # hint for trainable policies: will optim() below be a training step?  (same test as there)
            meta["policy_will_train"] = self.clip_length % self.train_interval == 0
            meta = self.policy(meta)  # sets grid, num_exec, num_total, perc_exec
            self.policy_meta = meta

        with timings.env("blockcopy/model", 3):
            if meta["num_exec"] == 0:
                # nothing to execute: the previous output object is returned, no state is touched
                meta = self.policy_meta = meta.copy()
                out = meta["outputs"]
            elif self._graphs is not None and not kwargs:
                meta["frame_state"], out = self._forward_graphed(inputs, meta["grid"], meta["num_exec"])
            else:
                x = to_tensorwrapper(as_tensor(inputs))
                self.block_temporal_features = x.process_temporal_features(self.block_temporal_features)
                blocks = x.to_blocks(meta["grid"])
                # frame state: for every block the most recently executed input pixels
                meta["frame_state"] = blocks.combine_().to_tensor()
                out = self.base_model(blocks, **kwargs)
                out = out.combine().to_tensor()
            meta["outputs_prev"] = meta["outputs"]
            meta["outputs"] = out

        with timings.env("blockcopy/policy_optim", 3):
            if self.policy is not None:
                train_policy = self.clip_length % self.train_interval == 0
                self.policy_meta = self.policy.optim(self.policy_meta, train=train_policy)
        return out


    # ------------------------------------------------------------------ CUDA-graph mode
    def _block_frame_inplace(self, inputs, grid, graph=None, patch=False):
        """One block-sparse frame with every combine IN PLACE into persistent planes (what a graph
        can replay): returns (frame_state plane, output plane, prefix), both persistent tensors.

        With `graph` (a torch.cuda.CUDAGraph) only the part AFTER the split is captured: index compaction and
        the gather of the executed input blocks run eagerly, straight from the caller's tensors, into buffers
        the captured part reads (`prefix`: what a replay has to refill).  A replay therefore needs no staging
        copy of the 12.6 MB frame."""
        from .. import _C

        gs = self._graphs
        if gs.splitk_ws is None or gs.splitk_ws.device != inputs.device:
            # graph replays of different models may overlap on different CUDA streams: the split-K scratch whose
            # address the graph bakes in belongs to this model, not to the (shared) capture stream
            gs.splitk_ws = torch.empty(_C.SPLITK_WS_BYTES, dtype=torch.uint8, device=inputs.device)
            gs.splitk_ws_side = torch.empty(_C.SPLITK_WS_BYTES // 2, dtype=torch.uint8, device=inputs.device)
        if graph is not None and patch:
            return self._capture_whole_frame(inputs, grid, graph)
        x = to_tensorwrapper(as_tensor(inputs))
        self.block_temporal_features = feats = x.process_temporal_features(self.block_temporal_features)
        feats.track_transfer_idx = False
        blocks = x.to_blocks(grid)  # bc_compact_mask + bc_gather
        tiles = blocks.as_subclass(torch.Tensor)
        prefix = (feats._grid_idx, feats._index_buf, tiles, tuple(inputs.shape), inputs.dtype, _C.layout_of(tiles))

        def body():
            # side branches of the model (skip bottlenecks) go to a second stream: parallel nodes of the captured graph
            with _C.splitk_workspace_scope(gs.splitk_ws), side_stream_scope(gs.splitk_ws_side):
                # frame_state is read by the next frame's policy only: its scatter runs beside the model
                frame_state = run_on_side_stream(lambda: blocks.combine_().to_tensor(), keep=(blocks,))
                out = self.base_model(blocks)
                return frame_state, out.combine_().to_tensor()

        if graph is None:
            return body() + (prefix,)
        if gs.capture_stream is None or gs.capture_stream.device != inputs.device:
            # kernel nodes inherit the capture stream's priority: the frame's latency-bound kernels get SMs before
            # lower-priority work that shares the GPU (uploads, driver-side kernels, other processes' streams)
            gs.capture_stream = torch.cuda.Stream(device=inputs.device, priority=-1)
        with torch.cuda.graph(graph, pool=gs.pool, stream=gs.capture_stream):
            res = body()
        return res + (prefix,)

    def _capture_whole_frame(self, inputs, grid, graph):
        """Graph-patch mode: index compaction, the input gather and the copy into the output buffer are captured too;
        before every replay their nodes are re-pointed at that frame's grid / frame / output buffer
        (``_C.graph_patch_next``), so that a steady frame is ONE graph launch and nothing else on the stream."""
        from .. import _C

        gs = self._graphs
        if gs.capture_stream is None or gs.capture_stream.device != inputs.device:
            gs.capture_stream = torch.cuda.Stream(device=inputs.device, priority=-1)
        image = inputs.as_subclass(torch.Tensor)
        with torch.cuda.graph(graph, pool=gs.pool, stream=gs.capture_stream):
            x = to_tensorwrapper(image)
            self.block_temporal_features = feats = x.process_temporal_features(self.block_temporal_features)
            feats.track_transfer_idx = False
            _C.graph_record(True)
            try:
                feats._process_grid(grid, x._features_prev)          # bc_compact_mask
                cm = _C.graph_last_node()
                blocks = x._split(image.shape[2] // grid.shape[2])   # bc_gather
                ga = _C.graph_last_node()
            finally:
                _C.graph_record(False)
            assert cm[0] != ga[0], "graph patching: the input gather created no node of its own"
            tiles = blocks.as_subclass(torch.Tensor)
            with _C.splitk_workspace_scope(gs.splitk_ws), side_stream_scope(gs.splitk_ws_side):
                frame_state = run_on_side_stream(lambda: blocks.combine_().to_tensor(), keep=(blocks,))
                dense = self.base_model(blocks).combine_().to_tensor()
            if gs.out_bufs is None or gs.out_bufs[0].shape != dense.shape:
                gs.out_bufs = [torch.empty_like(dense), torch.empty_like(dense)]
            cp = _C.graph_memcpy(gs.out_bufs[0], dense)
        graph.instantiate()
        prefix = (feats._grid_idx, feats._index_buf, tiles, tuple(inputs.shape), inputs.dtype, _C.layout_of(tiles))
        info = dict(exec=graph.raw_cuda_graph_exec(), cm=cm, ga=ga, cp=cp)
        return frame_state, dense, prefix, info

    def _patch_and_replay(self, entry, inputs, grid, num_exec):
        """Re-point the three per-frame nodes of a whole-frame graph, then launch it."""
        from .. import _C

        gs = self._graphs
        graph, frame_state, dense, launches, prefix, info = entry
        grid_idx, buf, tiles, shape, dtype, layout = prefix
        assert tuple(inputs.shape) == shape and inputs.dtype == dtype, \
            "input shape / dtype changed after CUDA graphs were captured"
        g = grid if (grid.dtype == torch.bool and grid.is_contiguous() and grid.device == tiles.device) else \
            grid.to(tiles.device, dtype=torch.bool).contiguous()
        G = g.numel()
        image = inputs.as_subclass(torch.Tensor)
        fmt = torch.channels_last if layout == _C.BC_NHWC else torch.contiguous_format
        if not image.is_contiguous(memory_format=fmt):
            image = image.contiguous(memory_format=fmt)
        try:
            _C.graph_patch_next(info["exec"], *info["cm"])
            _C.compact_mask(g.view(torch.uint8), grid_idx, buf[:G], buf[2 * G:])
            _C.graph_patch_next(info["exec"], *info["ga"])
            _C.gather(tiles, image, buf[:num_exec], num_exec)
        except _C.BlockCopyNativeError:
            # this frame's tensors would take another kernel variant than the captured one (other layout / alignment):
            # the caller drops the graph and runs the frame eagerly
            return None
        finally:
            _C.graph_record(False)  # disarm the hook should a call have failed before it reached its launch
        gs.flip ^= 1
        out = gs.out_bufs[gs.flip]
        _C.graph_patch_memcpy(info["exec"], info["cp"], out, dense)
        gs.keep = (g, image)  # what the re-pointed nodes read stays alive until the next frame's nodes are re-pointed
        graph.replay()
        _C.add_launches(launches)
        return frame_state, out

    @staticmethod
    def _refill_prefix(prefix, inputs, grid, num_exec):
        """What a replay runs eagerly before the graph: this frame's index tensors and executed input blocks."""
        from .. import _C

        grid_idx, buf, tiles, shape, dtype, layout = prefix
        assert tuple(inputs.shape) == shape and inputs.dtype == dtype, \
            "input shape / dtype changed after CUDA graphs were captured"
        g = grid if (grid.dtype == torch.bool and grid.is_contiguous() and grid.device == tiles.device) else \
            grid.to(tiles.device, dtype=torch.bool).contiguous()
        G = g.numel()
        _C.compact_mask(g.view(torch.uint8), grid_idx, buf[:G], buf[2 * G:])
        if isinstance(inputs, U8Frame) and layout == _C.BC_NCHW and tiles.shape[2] % 16 == 0:
            # decoded uint8 frame: normalisation fused into the gather, executed blocks only (bc_blocks_from_u8)
            _C.blocks_from_u8(tiles, inputs.u8, inputs.mean, inputs.std, buf[:num_exec], num_exec)
            return
        image = as_tensor(inputs).as_subclass(torch.Tensor)
        fmt = torch.channels_last if layout == _C.BC_NHWC else torch.contiguous_format
        if not image.is_contiguous(memory_format=fmt):
            image = image.contiguous(memory_format=fmt)
        _C.gather(tiles, image, buf[:num_exec], num_exec)

    def _forward_graphed(self, inputs, grid, num_exec):
        from .. import _C

        gs = self._graphs
        hint = get_num_exec_hint(grid)
        if hint is None:
            # no (or a stale) host-side count on this grid -- e.g. a custom policy edited it after the stats were
            # taken: count on the device like the reference (tensorwrapper.py:157), the graph is chosen by the truth
            hint = int(grid.sum())
            set_num_exec_hint(grid, hint)
        num_exec = hint
        if num_exec == 0:
            raise AssertionError("policy_meta['num_exec'] says blocks execute but the grid is empty")
        # whole-frame graphs with re-pointed nodes (graph-patch mode) for plain tensors whose grid needs no conversion;
        # a U8Frame keeps the eager prefix (its first gather is another kernel: graphs are keyed by the input kind)
        patch = gs.patch and not isinstance(inputs, U8Frame) and grid.dtype == torch.bool and grid.is_contiguous() \
            and grid.device == inputs.device and inputs.dim() == 4 \
            and (inputs.is_contiguous() or inputs.is_contiguous(memory_format=torch.channels_last))
        key = (num_exec, patch)
        entry = gs.graphs.get(key)
        if entry is None:
            seen = gs.seen.get(key, 0)
            gs.seen[key] = seen + 1
            if seen == 0 and not any(k[0] == num_exec for k in gs.graphs):
                # first time this block count shows up: run eagerly (allocates planes on the first
                # frame of the first clip, lets cuDNN pick its algorithms)
                frame_state, dense, _ = self._block_frame_inplace(inputs, grid)
            elif patch:
                graph = torch.cuda.CUDAGraph(keep_graph=True)
                n0 = _C.launch_count()
                frame_state, dense, prefix, info = self._block_frame_inplace(inputs, grid, graph=graph, patch=True)
                if gs.pool is None:
                    gs.pool = graph.pool()
                entry = gs.graphs[key] = (graph, frame_state, dense, _C.launch_count() - n0 - 2, prefix, info)
                # capturing does not execute: replay now (the nodes already point at this frame's grid and pixels; the
                # output node is re-pointed at the ping-pong buffer whose turn it is)
                self.block_temporal_features._was_reset = False
                return self._patch_and_replay(entry, inputs, grid, num_exec)
            else:
                graph = torch.cuda.CUDAGraph()
                n0 = _C.launch_count()
                frame_state, dense, prefix = self._block_frame_inplace(inputs, grid, graph=graph)
                if gs.pool is None:
                    gs.pool = graph.pool()
                # kernels of ours inside the graph (the eager prefix = 2 launches counts itself)
                entry = gs.graphs[key] = (graph, frame_state, dense, _C.launch_count() - n0 - 2, prefix, None)
                graph.replay()  # capturing does not execute; the prefix of THIS frame has just run eagerly
                _C.add_launches(0)
        else:
            graph, frame_state, dense, launches, prefix, info = entry
            if self.block_temporal_features._was_reset:
                # same contract as the eager path (tensorwrapper.py:164-165): the planes still hold the previous clip
                assert num_exec == grid.numel(), "No previous features known, first run should execute all blocks!"
            if info is not None:
                res = self._patch_and_replay(entry, inputs, grid, num_exec)
                if res is not None:
                    self.block_temporal_features._was_reset = False
                    return res
                gs.graphs.pop(key)  # captured for tensors of another layout: this frame eagerly, a new graph next time
                gs.seen[key] = 1
                frame_state, dense, _ = self._block_frame_inplace(inputs, grid)
                entry = None
            else:
                self._refill_prefix(prefix, inputs, grid, num_exec)
                graph.replay()
                _C.add_launches(launches)
        if entry is not None:
            frame_state, dense = entry[1], entry[2]
            # a replayed frame creates no new BlockFeatures: the kept one now holds this frame's history
            self.block_temporal_features._was_reset = False
        if gs.out_bufs is None or gs.out_bufs[0].shape != dense.shape:
            gs.out_bufs = [torch.empty_like(dense), torch.empty_like(dense)]
        gs.flip ^= 1
        out = gs.out_bufs[gs.flip]
        out.copy_(dense)
        return frame_state, out


class _GraphState:
    """Static buffers and captured graphs of one BlockCopyModel (block_cuda_graphs=True)."""

    def __init__(self):
        self.graphs = {}   # executed-block count -> (graph, frame_state plane, output plane, kernels of ours)
        self.seen = {}
        self.pool = None
        self.out_bufs = None
        self.flip = 0
        self.splitk_ws = None  # this model's split-K scratch (see _block_frame_inplace)
        self.splitk_ws_side = None  # ... and the one of convs issued on the side stream
        self.capture_stream = None  # high-priority stream the graphs are captured on
        # whole-frame graphs whose per-frame nodes are re-pointed before each replay (see _capture_whole_frame)
        self.patch = __import__("os").environ.get("BLOCKCOPY_GRAPH_PATCH", "1") != "0"
        self.keep = None


def _try_fused_dense(module, x):
    """Fused sm_100a implementation of a known dense module (today: SwiftNet-style pyramid pooling)."""
    import os

    if os.environ.get("BLOCKCOPY_FUSED_SPP", "1") == "0":
        return None
    from .fused_spp import try_fused_spp

    return try_fused_spp(module, x)


def blockcopy_noblocks(func):
    """Decorator for ``forward`` methods that cannot run on blocks (e.g. global pooling): the
    blocks are combined in place into the dense tensor, the method runs densely, and its result
    is split again with the same grid.  Costs a combine and a split."""

    @functools.wraps(func)
    def noblocks(self, x, *args):
        was_wrapper = isinstance(x, TensorWrapper)
        if was_wrapper:
            blocks = x
            # dense TensorWrapper (the reference hands over a plain tensor): numerically the same
            # object, but torch calls stay interceptable, see TensorWrapper._dense_dispatch
            x = x.combine_()
            fused = _try_fused_dense(self, x)
            if fused is not None:
                return to_tensorwrapper(fused).to_blocks_like(blocks)
        x = func(self, x)
        if was_wrapper:
            x = to_tensorwrapper(x.as_subclass(torch.Tensor)).to_blocks_like(blocks)
        return x

    return noblocks
