"""Fused execution of a pyramid-pooling module inside ``@blockcopy_noblocks``.

SwiftNet's SpatialPyramidPooling (reference swiftnet/util.py:85-138) is the only module of the model that
runs densely; as ~25 tiny torch launches it costs more than the whole encoder stage next to it.  When the
decorated module has that structure (duck-typed, the module code itself is NOT touched) and runs in fp16 on
CUDA, the same math runs as 7 launches: BN+ReLU (bc_ew_fused), 1x1 conv (bc_conv_igemm), all pools
(bc_spp_pool), per-level BN+ReLU+1x1 conv (bc_spp_levels), concat+upsample+BN+ReLU (bc_spp_prep), 1x1 conv
(bc_conv_igemm).  Anything that does not match falls through to the module's own forward.
"""
from __future__ import annotations

import weakref

import torch
import torch.nn as nn

from .. import _C

_PLANS = weakref.WeakKeyDictionary()  # module -> _Plan (dies with the module: no id() reuse)


def _unit(u):
    """(_BNReluConv-like Sequential) -> (bn, conv) or None."""
    norm, conv = getattr(u, "norm", None), getattr(u, "conv", None)
    if not isinstance(norm, nn.BatchNorm2d) or not isinstance(conv, nn.Conv2d) or norm.training:
        return None
    if conv.kernel_size != (1, 1) or conv.stride != (1, 1) or conv.groups != 1 or conv.bias is not None:
        return None
    kinds = []
    for m in u.children():
        if isinstance(m, (nn.Dropout, nn.Dropout2d)) and m.training:
            return None
        if not isinstance(m, (nn.BatchNorm2d, nn.ReLU, nn.Conv2d, nn.Dropout, nn.Dropout2d, nn.Identity)):
            return None
        if isinstance(m, (nn.BatchNorm2d, nn.ReLU, nn.Conv2d)):
            kinds.append(type(m))
    if kinds != [nn.BatchNorm2d, nn.ReLU, nn.Conv2d]:  # the fused kernels compute conv(relu(bn(x))), nothing else
        return None
    return norm, conv


def _bn32(bn: nn.BatchNorm2d, pad_to: int = 0) -> torch.Tensor:
    """fp32 [4][C(+pad)] = mean, invstd, weight, shift."""
    C = bn.num_features
    P = max(C, pad_to)
    out = torch.zeros(4, P, dtype=torch.float32, device=bn.running_mean.device)
    out[0, :C] = bn.running_mean.float()
    out[1, :C] = torch.rsqrt(bn.running_var.float() + bn.eps)
    out[2, :C] = bn.weight.float() if bn.weight is not None else 1.0
    out[3, :C] = bn.bias.float() if bn.bias is not None else 0.0
    return out.contiguous()


class _Plan:
    def __init__(self, module, units, grids):
        (bn0, conv0), levels, (bnf, convf) = units[0], units[1:-1], units[-1]
        dev, dt = conv0.weight.device, conv0.weight.dtype
        self.grids = grids
        self.C0, self.bt = conv0.in_channels, conv0.out_channels
        self.Lc = levels[0][1].out_channels
        self.L = len(levels)
        self.Cout = convf.out_channels
        self.Ccat = self.bt + self.L * self.Lc
        self.Cp = ((self.Ccat + 63) // 64) * 64
        with torch.no_grad():
            self.bn0 = tuple(t.contiguous() for t in _bn32(bn0))
            self.w0 = conv0.weight.detach().contiguous(memory_format=torch.channels_last)
            self.bn_lv = torch.stack([_bn32(b) for b, _ in levels]).contiguous()
            self.w_lv = torch.stack([c.weight.detach().reshape(self.Lc, self.bt) for _, c in levels]).contiguous()
            self.bnf = _bn32(bnf, self.Cp)
            wf = torch.zeros(self.Cout, self.Cp, 1, 1, dtype=dt, device=dev)
            wf[:, :self.Ccat] = convf.weight.detach()
            self.wf = wf.contiguous(memory_format=torch.channels_last)
        self.versions = _versions(units)
        self.ok = (dt == torch.float16 and self.C0 % 64 == 0 and self.bt % 64 == 0 and self.Cout % 64 == 0
                   and self.bt % 8 == 0 and convf.in_channels == self.Ccat
                   and all(c.in_channels == self.bt and c.out_channels == self.Lc for _, c in levels))
        self.cells = {}


def _versions(units):
    """(source tensors, their in-place version counters): the plan keeps the tensors, so a replaced
    parameter is a different object even if the allocator hands out the same address again."""
    src = [t for bn, conv in units for t in (bn.running_mean, bn.running_var, bn.weight, bn.bias, conv.weight)]
    return src, tuple(None if t is None else t._version for t in src)


def _same(a, b):
    return a[1] == b[1] and len(a[0]) == len(b[0]) and all(x is y for x, y in zip(a[0], b[0]))


def _upsampling_is_bilinear(module) -> bool:
    """The fused kernels hard-code F.interpolate(level, size, mode='bilinear') (align_corners False): probe the
    module's own `upsampling_method` once on a tiny CPU tensor instead of trusting the attribute names."""
    fn = getattr(module, "upsampling_method", None)
    if fn is None:
        return True  # the module calls F.interpolate itself (consumers/swiftnet_rn18.py)
    ok = getattr(fn, "_bc_is_bilinear", None)
    if ok is None:
        probe = torch.arange(24, dtype=torch.float32).reshape(1, 2, 3, 4).sin()
        try:
            got = fn(probe, (6, 12))
            ok = bool(torch.equal(got, torch.nn.functional.interpolate(probe, (6, 12), mode="bilinear")))
        except Exception:
            ok = False
        try:
            fn._bc_is_bilinear = ok
        except AttributeError:
            pass
    return ok


def _match(module):
    spp, grids = getattr(module, "spp", None), getattr(module, "grids", None)
    if not isinstance(spp, nn.Sequential) or grids is None or len(spp) < 3 or len(spp) - 2 > 4:
        return None
    if getattr(module, "square_grid", False) or getattr(module, "fixed_size", None) is not None or module.training:
        return None
    units = [_unit(u) for u in spp.children()]
    if any(u is None for u in units) or len(grids) < len(units) - 2:
        return None
    if not _upsampling_is_bilinear(module):
        return None
    return units


def try_fused_spp(module, x: torch.Tensor):
    """x: dense (N,C,H,W) fp16 CUDA tensor (channels_last).  Returns the module's output or None."""
    if not isinstance(x, torch.Tensor) or not x.is_cuda or x.dtype != torch.float16 or x.dim() != 4:
        return None
    units = _match(module)
    if units is None:
        return None
    plan = _PLANS.get(module)
    if plan is None or not _same(plan.versions, _versions(units)):
        plan = _PLANS[module] = _Plan(module, units, list(module.grids))
    if not plan.ok or x.shape[1] != plan.C0:
        return None
    x = x.as_subclass(torch.Tensor)
    N, C, H, W = x.shape
    if not x.is_contiguous(memory_format=torch.channels_last):
        x = x.contiguous(memory_format=torch.channels_last)
    ar = W / H
    gh = [plan.grids[i] for i in range(plan.L)]
    gw = [max(1, round(ar * g)) for g in gh]
    if any(H % a or W % b for a, b in zip(gh, gw)):
        return None
    BS = 32
    while BS > 4 and (H % BS or W % BS):
        BS //= 2
    if H % BS or W % BS:
        return None
    dev = x.device
    key = (N, H, W, BS)
    cells = plan.cells.get(key)
    if cells is None:
        cells = plan.cells[key] = torch.arange(N * (H // BS) * (W // BS), dtype=torch.int32, device=dev)
    E = cells.numel()
    cl = dict(dtype=torch.float16, device=dev, memory_format=torch.channels_last)
    a = torch.empty((N, C, H, W), **cl)
    _C.ew_fused(a, x, None, plan.bn0, relu=True)                         # relu(bn(x))
    x0 = torch.empty((N, plan.bt, H, W), **cl)
    # tile-less launches: only the dense planes are written
    _C.conv_igemm(x0, a, plan.w0, None, None, cells, E, BS, 1, 0, plane_out=x0, split_k=False, write_tiles=False)   # 1x1 conv -> x0
    ncell = sum(N * p * q for p, q in zip(gh, gw))
    pooled = torch.empty((ncell, plan.bt), dtype=torch.float16, device=dev)
    _C.spp_pool(pooled, x0, gh, gw)
    lev = torch.empty((ncell, plan.Lc), dtype=torch.float16, device=dev)
    _C.spp_levels(lev, pooled, plan.bn_lv, plan.w_lv, x0.shape, gh, gw)
    y = torch.empty((N, plan.Cp, H, W), **cl)
    _C.spp_prep(y, x0, lev, plan.bnf, gh, gw)
    out = torch.empty((N, plan.Cout, H, W), **cl)
    _C.conv_igemm(out, y, plan.wf, None, None, cells, E, BS, 1, 0, plane_out=out, split_k=False, write_tiles=False)
    return out
