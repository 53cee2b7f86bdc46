"""Training frames of the policy net on this repo's sm_100a kernels: forward WITH saved activations and the backward
pass.

The reference back-propagates the REINFORCE loss through PolicyNet with torch autograd every
``block_train_interval`` frames (policy/policy.py:319-370 -> policy/net.py:78-125, policy/resnet.py:60-115; train-mode
BatchNorm).  Here the trunk is one ``torch.autograd.Function``: its forward is the fused forward of
``policy/fused_net.py`` that keeps, per unit  conv -> BN(batch statistics) -> [+ shortcut] -> [ReLU],  the input plane x,
the conv output z, the batch statistics and the activated output; its backward walks the units in reverse:

    g      = out > 0 ? dOut : 0                                          (folded into the two BN kernels)
    sums   = [sum g, sum g * xhat]                                       bc_bn_bwd_reduce
    dz     = gamma * invstd * (g - sums0 / P - xhat * sums1 / P)         bc_bn_bwd_apply  (+ zero-interleaved copy, stride 2)
    dW, dgamma, dbeta                                                    bc_conv_wgrad    (fp32, into the .grad tensors)
    dx     = conv(dz | dz_up, flipped / transposed weights)              bc_conv_igemm    (tcgen05)
    joins  = dx(conv1) + dx(downsample)  |  dx(conv1) + (out > 0 ? dOut : 0)              bc_bwd_mask_add

Activation gradients are fp16 (fp32 accumulation inside every kernel), scaled by a power of two chosen on the device
from the gradient that enters the trunk (no host round trip) and divided out where the fp32 parameter gradients are
written.  The final 128 -> 1 conv (0.6 MFLOP) is differentiated with three torch ops on its im2col matrix.  Parameter
gradients land in persistent fp32 buffers that are (re-)attached as ``.grad`` after every backward, so a CUDA graph of the
backward pass and the fused RMSprop step keep their addresses.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn.functional as F

from .. import _C
from .fused_net import FusedPolicyTrunk, _BN, _Conv, _block_edge, _pad64


class _Unit:
    """One conv -> BN -> [ReLU] unit of a training forward: what the backward needs."""

    __slots__ = ("conv", "bn", "x", "z", "out", "relu")

    def __init__(self, conv: _Conv, bn: _BN, x, z, out, relu: bool):
        self.conv, self.bn, self.x, self.z, self.out, self.relu = conv, bn, x, z, out, relu


class _TrunkFunction(torch.autograd.Function):
    """logits = trunk(features); the parameters are passed only so that autograd calls backward()."""

    @staticmethod
    def forward(ctx, trainer, fill, shape, device, *params):
        ctx.trainer = trainer
        return trainer._run_forward(fill, shape, device)

    @staticmethod
    def backward(ctx, grad_logits):
        ctx.trainer._run_backward(grad_logits)
        return (None,) * (4 + len(ctx.trainer._params))


class FusedPolicyTrainer(FusedPolicyTrunk):
    """Forward + backward of PolicyNet.backbone / PolicyNet.layers on training frames (see module docstring)."""

    def __init__(self, net):
        super().__init__(net)
        self._units: List[_Unit] = []
        self._blocks_rec = []
        self._head_rec: List[_Unit] = []
        self._stem_rec: Optional[_Unit] = None
        self._h_last = None
        self._params = [q for q in list(net.backbone.parameters()) + list(net.layers.parameters())] if self.ok else []
        self._grads = {}
        self._bwd_ws = self._wg_ws = None
        self._bufs = {}
        self._fwd_graph = self._bwd_graph = None
        self._static_grad = None
        self.use_cuda_graph = False
        self._eager, self._warm_key = True, None
        if self.ok:
            # fc / avgpool of the CIFAR trunk exist for state_dict compatibility only: they receive no gradient
            used = {id(c.conv.weight) for c in self._all_convs()} | {id(t) for b in self._bns() for t in (b.bn.weight, b.bn.bias)}
            used |= {id(self.last.weight)} | ({id(self.last.bias)} if self.last.bias is not None else set())
            self._params = [q for q in self._params if id(q) in used]

    def _all_convs(self) -> List[_Conv]:
        return [self.stem[0]] + [c for b in self.blocks for c in (b[0], b[2])] + \
            [b[4][0] for b in self.blocks if b[4] is not None] + [h[0] for h in self.head]

    # ------------------------------------------------------------------ parameter packing (forward + data-gradient weights)
    def _pack(self):
        """As FusedPolicyTrunk._pack, plus the data-gradient weights of every conv but the stem: W'[ci, co, kh, kw] =
        W[co, ci, k-1-kh, k-1-kw] (the same bc_pack_params launch, through negative source strides)."""
        convs = self._all_convs()
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
            for c in convs[1:]:
                w = c.conv.weight
                c.w16d = torch.zeros((_pad64(c.cin), _pad64(c.cout), c.k, c.k), dtype=torch.float16,
                                     device=dev).contiguous(memory_format=torch.channels_last)
                s = w.stride()
                first = w.data_ptr() + ((c.k - 1) * s[2] + (c.k - 1) * s[3]) * w.element_size()
                rows.append([first, c.w16d.data_ptr(), off, c.cin, c.cout, c.k, _pad64(c.cout), s[1], s[0], -s[2], -s[3], 1])
                off += w.numel()
            for b in self._bns():
                b.mean = torch.empty(b.cp, dtype=torch.float32, device=dev)
                b.invstd = torch.empty(b.cp, dtype=torch.float32, device=dev)
                b.weight = torch.ones(b.cp, dtype=torch.float32, device=dev)
                b.shift = torch.zeros(b.cp, dtype=torch.float32, device=dev)
                b.sums = torch.zeros(2 * b.cp, dtype=torch.float32, device=dev)
                for src, dst in ((b.bn.weight, b.weight), (b.bn.bias, b.shift)):
                    rows.append([src.data_ptr(), dst.data_ptr(), off, b.c, 1, 1, 1, src.stride(0), 0, 0, 0, 0])
                    off += b.c
            self._pack_table = torch.tensor(rows, dtype=torch.int64, device=dev)
            self._pack_total, self._pack_key = off, key
            self._fwd_graph = self._bwd_graph = None
        _C.pack_params(self._pack_table, self._pack_total)

    # ------------------------------------------------------------------ forward with saved activations
    def _unit(self, x, c: _Conv, b: _BN, relu: bool) -> _Unit:
        z = self._conv(x, c)
        out = self._bn(z, b, relu=relu)
        return _Unit(c, b, x, z, out, relu)

    def _forward(self) -> torch.Tensor:
        self._stem_rec = self._unit(self._x16, self.stem[0], self.stem[1], True)
        h = self._stem_rec.out
        self._blocks_rec = []
        for c1, b1, c2, b2, ds in self.blocks:
            u1 = self._unit(h, c1, b1, True)
            u2 = self._unit(u1.out, c2, b2, False)
            uds = None if ds is None else self._unit(h, ds[0], ds[1], False)
            hout = torch.empty_like(u2.out)
            _C.ew_fused(hout, u2.out, h if uds is None else uds.out, None, relu=True)
            self._blocks_rec.append((u1, u2, uds, hout))
            h = hout
        self._head_rec = []
        for c, b in self.head:
            u = self._unit(h, c, b, True)
            self._head_rec.append(u)
            h = u.out
        self._update_running_stats()
        self._h_last = h
        last = self.last
        return _C.conv_fewout(h, last.weight.detach(), None if last.bias is None else last.bias.detach(),
                              last.stride[0], last.padding[0])

    # ------------------------------------------------------------------ backward
    def _grad_buffer(self, q: torch.Tensor) -> torch.Tensor:
        g = self._grads.get(id(q))
        if g is None or g.shape != q.shape or g.device != q.device:
            g = self._grads[id(q)] = torch.zeros_like(q)  # the parameter's own strides (channels_last weights): FusedRMSprop's fast path
        return g

    def _buf(self, name, shape, zero=False):
        t = self._bufs.get(name)
        if t is None or tuple(t.shape) != tuple(shape):
            make = torch.zeros if zero else torch.empty
            t = self._bufs[name] = make(shape, dtype=torch.float16, device=self._x16.device).contiguous(memory_format=torch.channels_last)
        return t

    def _dgrad(self, name, u: _Unit, dz: torch.Tensor, dz_up: Optional[torch.Tensor]) -> torch.Tensor:
        """dx of the unit's conv: a stride-1 conv with the flipped / transposed weights over dz (stride 1) or over the
        zero-interleaved dz_up (stride 2)."""
        c = u.conv
        src = dz if c.stride == 1 else dz_up
        N, _, H, W = src.shape
        bs = _block_edge(H, W, 1)
        E = N * (H // bs) * (W // bs)
        cells = self._cells.get((E, src.device))
        if cells is None:
            cells = self._cells[(E, src.device)] = torch.arange(E, dtype=torch.int32, device=src.device)
        out = self._buf(name, (N, c.w16d.shape[0], H, W))
        _C.conv_igemm(out, src, c.w16d, None, None, cells, E, bs, 1, c.k // 2, plane_out=out, split_k=False, write_tiles=False)
        return out

    def _unit_backward(self, name: str, u: _Unit, d_out: torch.Tensor, mask: Optional[torch.Tensor], need_dx: bool):
        """Gradients of one unit: parameter gradients into the persistent .grad buffers, returns dx (or None)."""
        b, c = u.bn, u.conv
        _C.bn_bwd_reduce(b.sums, d_out, mask, u.z, b.mean, b.invstd, self._bwd_ws)
        up = need_dx and c.stride == 2
        dz = self._buf(name + ".dz", tuple(u.z.shape))
        N, C, Ho, Wo = u.z.shape
        dz_up = self._buf(name + ".dz_up", (N, C, 2 * Ho, 2 * Wo), zero=True) if up else None
        _C.bn_bwd_apply(dz, dz_up, d_out, mask, u.z, b.mean, b.invstd, b.weight, b.sums)
        _C.conv_wgrad(self._grad_buffer(c.conv.weight), dz, u.x, c.stride, self._inv_scale, self._wg_ws, bn_sums=b.sums,
                      dgamma=self._grad_buffer(b.bn.weight), dbeta=self._grad_buffer(b.bn.bias))
        return self._dgrad(name + ".dx", u, dz, dz_up) if need_dx else None

    def _last_conv_backward(self, grad_logits: torch.Tensor) -> torch.Tensor:
        """The 128 -> 1 output conv in fp32 torch ops on its im2col matrix; returns d(h_last) fp32 (N,C,H,W)."""
        last, h = self.last, self._h_last
        N, Cp, H, W = h.shape
        C = last.in_channels
        k, s, p = last.kernel_size[0], last.stride[0], last.padding[0]
        g = grad_logits.reshape(N, last.out_channels, -1).float()                     # (N, Co, L)
        cols = F.unfold(h[:, :C].float(), k, padding=p, stride=s)                     # (N, C*k*k, L)
        self._grad_buffer(last.weight).copy_(torch.einsum("nol,nkl->ok", g, cols).view_as(last.weight))
        if last.bias is not None:
            self._grad_buffer(last.bias).copy_(g.sum((0, 2)))
        w = last.weight.detach().reshape(last.out_channels, -1)                       # (Co, C*k*k)
        dcols = torch.einsum("ok,nol->nkl", w, g)
        return F.fold(dcols, (H, W), k, padding=p, stride=s)                          # (N, C, H, W)

    def _backward(self, grad_logits: torch.Tensor):
        dev = self._x16.device
        if self._bwd_ws is None or self._bwd_ws.device != dev:
            self._bwd_ws = torch.zeros(_C.BN_STATS_WORKSPACE, dtype=torch.uint8, device=dev)
            self._wg_ws = torch.empty(_C.WGRAD_WORKSPACE, dtype=torch.uint8, device=dev)
        dh = self._last_conv_backward(grad_logits)
        # loss scale: the largest power of two that brings max |dh| to <= 1024 (all on the device)
        amax = dh.abs().amax().clamp_min(1e-30)
        scale = torch.exp2(torch.floor(torch.log2(1024.0 / amax))).clamp(2.0 ** -20, 2.0 ** 40)
        self._inv_scale = (1.0 / scale).reshape(1).float().contiguous()
        h = self._h_last
        d = self._buf("d_last", tuple(h.shape), zero=True)
        d[:, : dh.shape[1]].copy_(dh * scale)
        for i in reversed(range(len(self._head_rec))):
            u = self._head_rec[i]
            d = self._unit_backward(f"head{i}", u, d, u.out, True)
        for i in reversed(range(len(self._blocks_rec))):
            u1, u2, uds, hout = self._blocks_rec[i]
            d_y1 = self._unit_backward(f"b{i}.2", u2, d, hout, True)
            d_a = self._unit_backward(f"b{i}.1", u1, d_y1, u1.out, True)
            nxt = self._buf(f"b{i}.din", tuple(d_a.shape))
            if uds is not None:
                d_b = self._unit_backward(f"b{i}.ds", uds, d, hout, True)
                _C.bwd_mask_add(nxt, d_a, None, d_b)
            else:
                _C.bwd_mask_add(nxt, d, hout, d_a)
            d = nxt
        self._unit_backward("stem", self._stem_rec, d, self._stem_rec.out, False)

    # ------------------------------------------------------------------ entry points
    def _run_forward(self, fill, shape, device) -> torch.Tensor:
        self._prepare(shape, device)
        self._pack_if_needed()
        fill(self._x16)
        self._eager = True
        if not self.use_cuda_graph:
            return self._forward()
        key = (tuple(shape), device, self._param_key())
        g = self._fwd_graph
        if g is not None and g[0] == key:
            g[1].replay()
            self._eager = False
            return g[2].clone()
        if self._warm_key != key:
            # first training frame of this shape: forward and backward run eagerly (they create the persistent buffers and
            # tables); the graphs are captured on the next one.  No extra execution: the forward has a side effect (BatchNorm
            # running statistics)
            self._warm_key, self._fwd_graph, self._bwd_graph = key, None, None
            return self._forward()
        torch.cuda.synchronize(device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_out = self._forward()
        self._fwd_graph, self._bwd_graph = (key, graph, static_out), None
        graph.replay()
        self._eager = False
        return static_out.clone()

    def _attach_grads(self):
        for q in self._params:
            q.grad = self._grad_buffer(q)

    def _run_backward(self, grad_logits: torch.Tensor):
        if self._eager:
            self._backward(grad_logits.detach().contiguous())
        elif self._bwd_graph is None:
            self._static_grad = grad_logits.detach().clone().contiguous()
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, pool=self._fwd_graph[1].pool()):
                self._backward(self._static_grad)
            self._bwd_graph = (graph,)
            graph.replay()
        else:
            self._static_grad.copy_(grad_logits)
            self._bwd_graph[0].replay()
        self._attach_grads()

    def run_train(self, fill, shape, device) -> torch.Tensor:
        """Differentiable trunk forward: fill(x16) writes the fp16 features; returns fp32 logits (N,1,H/32,W/32) whose
        backward() leaves the parameter gradients in ``.grad`` of every trunk parameter."""
        return _TrunkFunction.apply(self, fill, shape, device, *self._params)
