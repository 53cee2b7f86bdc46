"""Inference-only forward of PolicyNet's trunk on this repo's sm_100a kernels (forward half).

The policy net (reference policy/net.py:17-125, policy/resnet.py:60-123) is a ResNet-8 (width x2) plus three
stride-2 3x3 convs, run in TRAIN mode (batch statistics) on a (N,26,H/4,W/4) fp32 feature map every frame; only
every ``block_train_interval``-th frame is followed by a backward pass.  On the other frames nothing but the
sampled grid is used, so no autograd graph is needed and the trunk runs here as

    conv          bc_conv_igemm (tcgen05 implicit GEMM, fp16 operands, fp32 accumulate) over ALL cells of the dense
                  NHWC plane, result written straight into the next dense plane (``plane_out``)
    batch norm    bc_bn_stats (per-channel batch mean / invstd, one pass, reproducible) feeding bc_ew_fused's
                  (mean, invstd, weight, shift) pointers -- no host round trip
    relu / add    bc_ew_fused

Channel counts are padded to multiples of 64 (26 -> 64 inputs, 32 -> 64 in the first stage) with zero weights; the
final 128 -> 1 conv goes through bc_conv_fewout.  Weights are re-packed to fp16 from the live fp32 parameters (one
launch, ahead of the CUDA graph) whenever a parameter changed since the last pack -- in-place updates are seen through the
tensors' version counters, fused optimiser steps through fused_optim.PARAM_EPOCH, new storage through the pointers; code
that writes parameters through ``.data`` must bump PARAM_EPOCH itself (policy.py::sync_shared_policy does).

Numerics: fp16 operands / fp32 accumulation / fp16 activations against the reference's fp32 (TF32 on cuDNN)
trunk -- logits agree to ~1e-2 of their range (tests/test_gpu_policy.py states the bound); the grid is sampled
from them with torch's RNG exactly as on the torch path.  The train-mode side effect of the torch path -- the
BatchNorm running statistics and ``num_batches_tracked`` -- is reproduced by one table-driven launch at the end
(bc_bn_update_running), so the policy's ``state_dict`` after a clip matches the torch path's (to fp16-activation
accuracy).
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

from .. import _C


# BLOCKCOPY_BN_NORM=1: statistics + normalisation of a train-mode BatchNorm in ONE launch (bc_bn_norm, a grid-wide barrier
# between the phases).  Measured slower than the two launches (trunk forward 330 vs 308 us, rl_semseg 917 vs 942 frames/s:
# the second phase runs on the statistics kernel's 148 CTAs instead of bc_ew_fused's full grid, and every CTA repeats the
# final reduction), so it is an experiment switch only.
BN_NORM_ONE_LAUNCH = __import__("os").environ.get("BLOCKCOPY_BN_NORM", "0") == "1"


def _pad64(c: int) -> int:
    return (c + 63) // 64 * 64


def _block_edge(H: int, W: int, stride: int) -> Optional[int]:
    """Input block edge for a conv over a dense (H, W) plane: power of two, output edge in [4, 32]."""
    for bs in (32 * stride, 16 * stride, 8 * stride, 4 * stride):
        if H % bs == 0 and W % bs == 0:
            return bs
    return None


class _Conv:
    """One Conv2d of the trunk: live fp32 parameter -> padded fp16 channels_last copy."""

    def __init__(self, conv: nn.Conv2d):
        assert conv.bias is None and conv.groups == 1 and conv.dilation == (1, 1)
        self.conv = conv
        self.k, self.stride = conv.kernel_size[0], conv.stride[0]
        self.cout, self.cin = conv.out_channels, conv.in_channels
        self.w16: Optional[torch.Tensor] = None



class _BN:
    """Train-mode BatchNorm2d: batch statistics from bc_bn_stats, affine parameters padded with (1, 0)."""

    def __init__(self, bn: nn.BatchNorm2d):
        self.bn, self.c, self.cp = bn, bn.num_features, _pad64(bn.num_features)
        self.weight = self.shift = self.mean = self.invstd = None
        self.count = 0



class FusedPolicyTrunk:
    """no-grad forward of ``PolicyNet.backbone`` + ``PolicyNet.layers`` (see module docstring)."""

    def __init__(self, net: nn.Module):
        self.ok = False
        bb, layers = getattr(net, "backbone", None), getattr(net, "layers", None)
        try:
            self.stem = (_Conv(bb.conv1), _BN(bb.bn1))
            self.blocks = []
            for stage in (bb.layer1, bb.layer2, bb.layer3):
                for blk in stage:
                    ds = None if blk.downsample is None else (_Conv(blk.downsample[0]), _BN(blk.downsample[1]))
                    self.blocks.append((_Conv(blk.conv1), _BN(blk.bn1), _Conv(blk.conv2), _BN(blk.bn2), ds))
            self.head = []
            for seq in list(layers)[:-1]:
                conv, bn = seq[0], seq[1]
                assert isinstance(bn, nn.BatchNorm2d) and isinstance(seq[2], nn.ReLU)
                self.head.append((_Conv(conv), _BN(bn)))
            last = list(layers)[-1]
            assert len(last) == 1 and isinstance(last[0], nn.Conv2d)
            self.last = last[0]
            convs = [self.stem[0]] + [c for b in self.blocks for c in (b[0], b[2])] + \
                    [b[4][0] for b in self.blocks if b[4] is not None] + [h[0] for h in self.head]
            self.ok = all(c.k in (1, 3) and c.stride in (1, 2) and c.conv.padding == (c.k // 2, c.k // 2) for c in convs) \
                and all(_pad64(b.c) <= 128 for b in self._bns())
        except (AttributeError, AssertionError, TypeError, IndexError):
            self.ok = False
        self.in_channels = getattr(bb, "in_channels", None)
        self._ws = None      # bc_bn_stats workspace
        self._x16 = None     # padded fp16 NHWC input
        self._cells = {}
        self._graph = None   # (key, graph, static input, static output)
        self._pack_key = self._pack_table = None
        self._pack_total = 0
        self._pack_srcs = self._packed_token = None
        self._run_key = self._run_table = None

    def _bns(self) -> List[_BN]:
        out = [self.stem[1]]
        for b in self.blocks:
            out += [b[1], b[3]] + ([b[4][1]] if b[4] is not None else [])
        return out + [h[1] for h in self.head]

    # ------------------------------------------------------------------ shape envelope
    def supports(self, x: torch.Tensor) -> bool:
        return x.is_cuda and x.dim() == 4 and x.dtype == torch.float32 and self.supports_shape(tuple(x.shape), x.device)

    def supports_shape(self, shape, device) -> bool:
        """Can the trunk run on features of this (N, C, H, W) shape on `device`?"""
        if not (self.ok and len(shape) == 4 and shape[1] == self.in_channels and device.type == "cuda"):
            return False
        if any(b.bn.weight.dtype != torch.float32 or b.bn.weight.device != device for b in self._bns()):
            return False  # the affine vectors are handed to the kernels as they are
        H, W = shape[2], shape[3]
        for conv in [self.stem[0]] + [c for b in self.blocks for c in (b[0], b[2])] + [h[0] for h in self.head]:
            if _block_edge(H, W, conv.stride) is None:
                return False
            H, W = H // conv.stride, W // conv.stride
        return True

    # ------------------------------------------------------------------ pieces
    def _conv(self, x: torch.Tensor, c: _Conv) -> torch.Tensor:
        N, C, H, W = x.shape
        bs = _block_edge(H, W, c.stride)
        E = N * (H // bs) * (W // bs)
        cells = self._cells.get((E, x.device))
        if cells is None:
            cells = self._cells[(E, x.device)] = torch.arange(E, dtype=torch.int32, device=x.device)
        cl = dict(dtype=torch.float16, device=x.device, memory_format=torch.channels_last)
        cout = c.w16.shape[0]
        out = torch.empty((N, cout, H // c.stride, W // c.stride), **cl)
        # tile-less launch: only the dense output plane is written (`out` stands in for the unused tile batch)
        _C.conv_igemm(out, x, c.w16, None, None, cells, E, bs, c.stride, c.k // 2, plane_out=out, split_k=False,
                      write_tiles=False)
        return out

    def _bn(self, x: torch.Tensor, b: _BN, relu: bool, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        b.count = x.shape[0] * x.shape[2] * x.shape[3]
        out = torch.empty_like(x) if out is None else out
        if BN_NORM_ONE_LAUNCH:
            _C.bn_norm(out, x, b.mean, b.invstd, b.weight, b.shift, b.bn.eps, relu, self._ws)
        else:
            _C.bn_stats(x, b.mean, b.invstd, b.bn.eps, self._ws)
            _C.ew_fused(out, x, None, (b.mean, b.invstd, b.weight, b.shift), relu=relu)
        return out

    def _forward(self) -> torch.Tensor:
        """Everything after the input plane ``self._x16`` has been filled and the parameters packed."""
        h = self._bn(self._conv(self._x16, self.stem[0]), self.stem[1], relu=True)
        for c1, b1, c2, b2, ds in self.blocks:
            y = self._bn(self._conv(h, c1), b1, relu=True)
            y = self._bn(self._conv(y, c2), b2, relu=False)
            sc = h if ds is None else self._bn(self._conv(h, ds[0]), ds[1], relu=False)
            h = torch.empty_like(y)
            _C.ew_fused(h, y, sc, None, relu=True)   # relu(bn2(conv2) + shortcut)
        for c, b in self.head:
            h = self._bn(self._conv(h, c), b, relu=True)
        self._update_running_stats()
        last = self.last
        return _C.conv_fewout(h, last.weight.detach(), None if last.bias is None else last.bias.detach(),
                              last.stride[0], last.padding[0])

    def _update_running_stats(self):
        """running_mean / running_var / num_batches_tracked of every batch norm, one launch."""
        import struct

        bns = [b for b in self._bns() if b.bn.track_running_stats and b.bn.running_mean is not None]
        if not bns:
            return
        key = tuple((b.mean.data_ptr(), b.bn.running_mean.data_ptr(), b.bn.running_var.data_ptr(), b.count,
                     b.bn.momentum, b.bn.eps) for b in bns)
        if self._run_key != key:
            rows = []
            for b in bns:
                mom = -1.0 if b.bn.momentum is None else float(b.bn.momentum)
                bits = struct.unpack("<q", struct.pack("<ff", mom, float(b.bn.eps)))[0]
                nbt = b.bn.num_batches_tracked
                rows.append([b.mean.data_ptr(), b.invstd.data_ptr(), b.bn.running_mean.data_ptr(), b.bn.running_var.data_ptr(),
                             0 if nbt is None else nbt.data_ptr(), b.c, b.count, bits])
            self._run_table = torch.tensor(rows, dtype=torch.int64, device=bns[0].mean.device)
            self._run_key = key
        _C.bn_update_running(self._run_table)

    def _param_key(self):
        """Storage identity of every parameter the trunk reads (in-place optimiser steps keep it)."""
        srcs = [c.conv.weight for c in [self.stem[0]] + [c for b in self.blocks for c in (b[0], b[2])] +
                [b[4][0] for b in self.blocks if b[4] is not None] + [h[0] for h in self.head]]
        for b in self._bns():
            srcs += [b.bn.weight, b.bn.bias]
        return tuple((t.data_ptr(), t.stride()) for t in srcs)

    def _pack(self):
        """Live fp32 parameters -> fp16 channels_last conv weights / padded affine vectors, one launch
        (bc_pack_params) driven by a device table that is rebuilt only when a parameter's storage moves."""
        convs = [self.stem[0]] + [c for b in self.blocks for c in (b[0], b[2])] + \
            [b[4][0] for b in self.blocks if b[4] is not None] + [h[0] for h in self.head]
        key = self._param_key()
        if self._pack_key != key:
            dev = convs[0].conv.weight.device
            rows, off = [], 0
            for c in convs:
                w = c.conv.weight
                c.w16 = torch.zeros((_pad64(c.cout), _pad64(c.cin), c.k, c.k), dtype=torch.float16,
                                    device=dev).contiguous(memory_format=torch.channels_last)
                rows.append([w.data_ptr(), c.w16.data_ptr(), off, c.cout, c.cin, c.k, _pad64(c.cin), *w.stride(), 1])
                off += w.numel()
            for b in self._bns():
                b.mean = torch.empty(b.cp, dtype=torch.float32, device=dev)
                b.invstd = torch.empty(b.cp, dtype=torch.float32, device=dev)
                b.weight = torch.ones(b.cp, dtype=torch.float32, device=dev)
                b.shift = torch.zeros(b.cp, dtype=torch.float32, device=dev)
                for src, dst in ((b.bn.weight, b.weight), (b.bn.bias, b.shift)):
                    rows.append([src.data_ptr(), dst.data_ptr(), off, b.c, 1, 1, 1, src.stride(0), 0, 0, 0, 0])
                    off += b.c
            self._pack_table = torch.tensor(rows, dtype=torch.int64, device=dev)
            self._pack_total, self._pack_key = off, key
        _C.pack_params(self._pack_table, self._pack_total)

    def _pack_if_needed(self):
        """Re-pack the fp16 parameter copies only when a parameter changed since the last pack: an in-place update
        (`_version` counters), a fused optimiser step (fused_optim.PARAM_EPOCH: raw-pointer updates) or new storage
        (`_param_key`).  The policy trains every block_train_interval-th frame only, so most frames skip the launch."""
        from .fused_optim import PARAM_EPOCH

        srcs = self._pack_srcs
        if srcs is None:
            srcs = self._pack_srcs = [c.conv.weight for c in [self.stem[0]] + [c for b in self.blocks for c in (b[0], b[2])] +
                                      [b[4][0] for b in self.blocks if b[4] is not None] + [h[0] for h in self.head]] + \
                [t for b in self._bns() for t in (b.bn.weight, b.bn.bias)]
        token = (PARAM_EPOCH[0], sum(t._version for t in srcs))
        if self._pack_key != self._param_key() or self._packed_token != token:
            self._pack()
            self._packed_token = token

    def _prepare(self, shape, device):
        N, C, H, W = shape
        if self._ws is None or self._ws.device != device:
            self._ws = torch.zeros(_C.BN_STATS_WORKSPACE, dtype=torch.uint8, device=device)
        shape16 = (N, _pad64(C), H, W)
        if self._x16 is None or tuple(self._x16.shape) != shape16 or self._x16.device != device:
            self._x16 = torch.zeros(shape16, dtype=torch.float16, device=device).contiguous(memory_format=torch.channels_last)
            self._graph = None  # captured on the old plane

    # ------------------------------------------------------------------ entry points
    @torch.no_grad()
    def __call__(self, x: torch.Tensor, use_cuda_graph: bool = False) -> torch.Tensor:
        """x: (N, in_channels, H, W) fp32 CUDA features -> (N, 1, H/32, W/32) fp32 logits."""
        return self.run(lambda x16: x16[:, : x.shape[1]].copy_(x), tuple(x.shape), x.device, use_cuda_graph)

    @torch.no_grad()
    def run(self, fill, shape, device, use_cuda_graph: bool = False) -> torch.Tensor:
        """fill(x16) writes the features (fp16, channels 0..C-1 of the padded NHWC plane x16) -- eagerly, every call;
        everything behind it is one CUDA graph when `use_cuda_graph`."""
        self._prepare(shape, device)
        self._pack_if_needed()  # eagerly, outside the graph: skipped on frames whose parameters did not change
        fill(self._x16)
        if not use_cuda_graph:
            return self._forward()
        key = (tuple(shape), device, self._param_key())
        g = self._graph
        if g is None or g[0] != key:
            # this call runs eagerly (it also creates the persistent buffers and the tables); the graph captured right
            # after it serves the following calls.  No extra execution: the trunk has a side effect (running statistics)
            out = self._forward()
            torch.cuda.synchronize(device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = self._forward()
            self._graph = (key, graph, static_out)
            return out
        g[1].replay()
        return g[2]
