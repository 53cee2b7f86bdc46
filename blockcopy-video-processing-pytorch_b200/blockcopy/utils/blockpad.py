"""``pad`` / ``BlockPadFunction`` with the reference's signature (utils/blockpad.py:14-75): packed
tiles (E,C,BS,BS) + transferred tiles -> (E,C,BS+2p,BS+2p) with neighbour halos, zeros outside the
frame.  Bound to ``bc_gather_halo_tiles`` of libblockcopy_sm100.so."""
import warnings

import torch
from torch.autograd import Function

from .. import _C
from .profiler import timings


def pad(features, transfer, grid_idx, exec_map, pad=1):
    return BlockPadFunction.apply(features, transfer, grid_idx, exec_map, pad)


class BlockPadFunction(Function):
    @staticmethod
    def forward(ctx, data_exec, data_transfer, grid_idx, mapping_exec, pad):
        assert data_exec.shape[1:] == data_transfer.shape[1:], (data_exec.shape, data_transfer.shape)
        assert len(mapping_exec) <= data_exec.shape[0]
        assert grid_idx.numel() - len(mapping_exec) <= data_transfer.shape[0], \
            (grid_idx.numel(), len(mapping_exec), data_transfer.shape)
        assert pad > 0
        B, C, BS, _ = data_exec.shape
        assert BS > 0
        if BS <= 2:
            warnings.warn(f"Block size of 2 or smaller can be inefficient! Got size: {BS}")
        fmt = torch.channels_last if (B > 0 and _C.layout_of(data_exec) == _C.BC_NHWC) else torch.contiguous_format
        out = torch.empty((B, C, BS + 2 * pad, BS + 2 * pad), device=data_exec.device, dtype=data_exec.dtype,
                          memory_format=fmt)
        with timings.env("block/pad_kernel", 20):
            _C.gather_halo_tiles(out, data_exec, data_transfer, grid_idx, mapping_exec, len(mapping_exec), int(pad))
        return out

    @staticmethod
    def backward(ctx, grad_x):
        raise NotImplementedError("Backward not implemented for BlockPad")
