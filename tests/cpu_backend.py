"""TEST-ONLY: runs the host logic of the blockcopy package on CPU tensors by rebinding the
``blockcopy._C`` entry points to the C oracle.  This exists so that the `-m "not gpu"` suite can
exercise the wrapper / state machine / policy code in a container without a GPU; the product has
no such path (``_C`` raises on non-CUDA tensors).  Use as a context manager."""
import contextlib

import torch

from oracle import cpu_oracle as O


def _nchw(t):
    return t.as_subclass(torch.Tensor).contiguous()


def _store(dst, src):
    dst.as_subclass(torch.Tensor).copy_(src)


@contextlib.contextmanager
def cpu_backend():
    import blockcopy
    from blockcopy import _C
    from blockcopy.core import tensorwrapper as tw
    from blockcopy.core import blockcopy as bcm

    saved = {k: getattr(_C, k) for k in ("compact_mask", "gather", "scatter", "copy_blocks", "transfer",
                                         "gather_halo_tiles", "gather_halo", "conv_igemm", "ew_fused", "maxpool_halo",
                                         "conv_supported", "lazy_supported", "stem_supported", "stem_pack", "conv_stem",
                                         "head_supported", "head_1x1")}
    saved_tw = tw.to_tensorwrapper

    def compact_mask(grid_u8, grid_idx, mapping_exec, counts, prev_grid_idx=None, transfer_idx=None):
        g = grid_u8.view(torch.bool) if grid_u8.dtype == torch.uint8 else grid_u8
        gi, me = O.grid_mappings(g)
        grid_idx.copy_(gi.view_as(grid_idx))
        mapping_exec[: me.numel()].copy_(me)
        counts[0], counts[1] = me.numel(), g.numel() - me.numel()
        if transfer_idx is not None:
            ti = O.transfer_idx(g, prev_grid_idx.contiguous())
            transfer_idx[: ti.numel()].copy_(ti)

    def gather(blocks, image, mapping_exec, E):
        if E:
            _store(blocks, O.split(_nchw(image), mapping_exec[:E].contiguous(), blocks.shape[-1]))
        return blocks

    def scatter(blocks, image, mapping_exec, E):
        if E:
            tmp = _nchw(image).clone()
            O.combine_(_nchw(blocks), tmp, mapping_exec[:E].contiguous())
            _store(image, tmp)
        return image

    def copy_blocks(out, prev, blocks, grid_idx):
        tmp = _nchw(prev).clone()
        gi = grid_idx.flatten()
        cells = torch.nonzero(gi >= 0).squeeze(1).to(torch.int32)
        if cells.numel():
            order = gi[cells.long()].long()
            O.combine_(_nchw(blocks)[order].contiguous(), tmp, cells.contiguous())
        _store(out, tmp)
        return out

    def transfer(out, prev_exec, prev_transfer, transfer_idx, G, padding):
        tmp = _nchw(out).clone()
        O.transfer(tmp, _nchw(prev_exec), _nchw(prev_transfer), transfer_idx.contiguous(), G, padding)
        _store(out, tmp)
        return out

    def gather_halo_tiles(out, exec_t, transfer_t, grid_idx, mapping_exec, E, pad):
        if E:
            _store(out, O.repad(_nchw(exec_t), _nchw(transfer_t), grid_idx.contiguous(),
                                mapping_exec[:E].contiguous(), pad))
        return out

    def gather_halo(out, plane, mapping_exec, E, BS, pad):
        if E:
            _store(out, O.plane_halo(_nchw(plane), mapping_exec[:E].contiguous(), BS, pad))
        return out

    # --- torch restatements of the two compute kernels, so that the lazy-fusion bookkeeping of the wrapper
    #     (deferred convs, absorbed ReLU / add / BN / bilinear, dual write into planes) runs on CPU as well
    import torch.nn.functional as F

    def _scatter_plane(plane_out, tiles, mapping):
        tmp = _nchw(plane_out).clone()
        O.combine_(_nchw(tiles), tmp, mapping.contiguous())
        _store(plane_out, tmp)

    def conv_igemm(out, plane, weight_cl, bias, residual, mapping_exec, E, BS_in, stride, padding, relu=False,
                   plane_out=None, out_mapping=None, split_k=True, write_tiles=True):
        dil = padding if weight_cl.shape[-1] == 3 else 1  # bc_conv_igemm: a 3x3 conv with padding p has dilation p
        if mapping_exec is None:
            y = F.conv2d(_nchw(plane), weight_cl, bias, stride, padding, dil)
        else:
            full = F.conv2d(_nchw(plane), weight_cl, bias, stride, padding, dil)
            y = O.split(full.contiguous(), mapping_exec[:E].contiguous(), BS_in // stride)
        if residual is not None:
            y = y + _nchw(residual)
        if relu:
            y = y.relu()
        if write_tiles:
            _store(out, y)
        else:
            out.fill_(float("nan"))  # tile-less launch: whoever reads the tiles without gathering them first gets NaN
        if plane_out is not None:
            _scatter_plane(plane_out, y, (out_mapping if out_mapping is not None else mapping_exec)[:E])
        return out

    def ew_fused(out, a, residual=None, bn=None, relu=False, up2x=False, plane_out=None, mapping_exec=None):
        y = _nchw(a)
        if up2x:
            y = F.interpolate(y, scale_factor=2, mode="bilinear")
        if residual is not None:
            y = y + _nchw(residual)
        if bn is not None:
            mean, invstd, w, s = bn
            v = lambda t: t.view(1, -1, 1, 1).to(y.dtype)  # noqa: E731
            y = (y - v(mean)) * v(invstd)
            if w is not None:
                y = y * v(w)
            if s is not None:
                y = y + v(s)
        if relu:
            y = y.relu()
        if out is not None:
            _store(out, y)
        if plane_out is not None:
            _scatter_plane(plane_out, y, mapping_exec[: y.shape[0]])
        return out

    def maxpool_halo(out, plane, mapping_exec, E, BS_in, k, stride, padding, plane_out=None):
        full = F.max_pool2d(F.pad(_nchw(plane), (padding,) * 4), k, stride, 0)  # zero (not -inf) halo at the frame edge
        y = O.split(full.contiguous(), mapping_exec[:E].contiguous(), BS_in // stride)
        _store(out, y)
        if plane_out is not None:
            _scatter_plane(plane_out, y, mapping_exec[:E])
        return out

    def stem_supported(dtype, weight, BS_in, stride, padding, dilation=1, groups=1):
        Cout, Cin, kh, kw = weight.shape
        return Cin == 3 and kh == kw == 7 and stride == 2 and padding == 3 and Cout % 64 == 0 and BS_in % 2 == 0

    def stem_pack(s2d_plane, tiles, mapping_exec, E):
        t = _nchw(tiles)                                   # (E,3,BS,BS)
        E_, _, BS, _ = t.shape
        s2d = torch.zeros(E_, 16, BS // 2, BS // 2, dtype=t.dtype)
        for dy in range(2):
            for dx in range(2):
                ch = (dy * 2 + dx) * 3
                s2d[:, ch:ch + 3] = t[:, :, dy::2, dx::2]
        _scatter_plane(s2d_plane[..., 2:-2], s2d, mapping_exec[:E])  # BC_STEM_XPAD columns stay zero
        return s2d_plane

    def conv_stem(out, s2d_plane, weight_packed, bias, mapping_exec, E, relu=False, plane_out=None, write_tiles=True):
        Cout = weight_packed.shape[0]
        w = weight_packed.view(Cout, 4, 4, 16).permute(0, 3, 1, 2).contiguous()
        full = F.conv2d(F.pad(_nchw(s2d_plane)[..., 2:-2], (2, 1, 2, 1)), w, bias)  # taps oy-2 .. oy+1
        y = O.split(full.contiguous(), mapping_exec[:E].contiguous(), out.shape[-1])
        if relu:
            y = y.relu()
        if write_tiles:
            _store(out, y)
        else:
            out.fill_(float("nan"))
        if plane_out is not None:
            _scatter_plane(plane_out, y, mapping_exec[:E])
        return out

    def conv_supported(dtype, weight, BS_in, stride, padding, dilation=1, groups=1):
        Cout, Cin, kh, kw = weight.shape
        if kh != kw or kh not in (1, 3) or stride not in (1, 2) or groups != 1:
            return False
        if (kh == 1 and (padding != 0 or dilation != 1)) or (kh == 3 and (padding != dilation or not 1 <= dilation <= 4
                                                                          or (dilation > 1 and stride != 1))):
            return False
        bo = BS_in // stride
        return Cin % 64 == 0 and Cout % 64 == 0 and BS_in % stride == 0 and bo >= 1

    def lazy_supported(x):
        return x.dim() == 4 and x.shape[1] % 8 == 0

    def head_supported(dtype, weight, stride, padding, dilation=1, groups=1):
        if weight.dim() != 4:
            return False
        Cout, Cin, kh, kw = weight.shape
        return kh == kw == 1 and stride == 1 and padding == 0 and dilation == 1 and groups == 1 and Cout <= 32 and Cin % 8 == 0

    def head_1x1(tiles_in, weight2d, bias, bn, relu_in, tiles_out=None, dense_out=None, dense_prev=None, grid_idx=None,
                 mapping_exec=None):
        y = _nchw(tiles_in)
        if bn is not None:
            mean, invstd, w, s = bn
            v = lambda t: t.view(1, -1, 1, 1).to(y.dtype)  # noqa: E731
            y = (y - v(mean)) * v(invstd)
            if w is not None:
                y = y * v(w)
            if s is not None:
                y = y + v(s)
        if relu_in:
            y = y.relu()
        y = F.conv2d(y, weight2d.view(*weight2d.shape, 1, 1), bias)
        if tiles_out is not None:
            _store(tiles_out, y)
        if dense_out is not None:
            tmp = _nchw(dense_prev if dense_prev is not None else dense_out).clone()
            O.combine_(y.contiguous(), tmp, mapping_exec[: y.shape[0]].contiguous())
            _store(dense_out, tmp)
        return dense_out if dense_out is not None else tiles_out

    for k, v in dict(conv_igemm=conv_igemm, ew_fused=ew_fused, conv_supported=conv_supported,
                     lazy_supported=lazy_supported, maxpool_halo=maxpool_halo, stem_supported=stem_supported,
                     stem_pack=stem_pack, conv_stem=conv_stem, head_supported=head_supported,
                     head_1x1=head_1x1).items():
        setattr(_C, k, v)
    for k, v in dict(compact_mask=compact_mask, gather=gather, scatter=scatter, copy_blocks=copy_blocks,
                     transfer=transfer, gather_halo_tiles=gather_halo_tiles, gather_halo=gather_halo).items():
        setattr(_C, k, v)
    cpu_tw = lambda x: x.as_subclass(tw.TensorWrapper)  # noqa: E731  (lifts the CUDA assert)
    tw.to_tensorwrapper = bcm.to_tensorwrapper = blockcopy.to_tensorwrapper = cpu_tw
    try:
        yield
    finally:
        for k, v in saved.items():
            setattr(_C, k, v)
        tw.to_tensorwrapper = bcm.to_tensorwrapper = blockcopy.to_tensorwrapper = saved_tw
