"""Kernel-level entry points with the reference's names and call signatures
(utils/block_funcs.py: SplitFunction :10, CombineFunction :85, TransferFunction :161), bound to
libblockcopy_sm100.so instead of CuPy-JIT-compiled CUDA C strings.  A maintainer of the reference
can swap this module in unchanged; shape checks mirror the reference's asserts.
Forward only -- ``backward`` raises, like the reference.
"""
import torch
from torch.autograd import Function

from .. import _C
from .profiler import timings

CUDA_NUM_THREADS = 512  # kept for API compatibility; launch geometry is chosen by the library
CUDA_NUM_BLOCKS = 1024


def _check4(t, name):
    assert t.is_cuda, f"{name} must be a CUDA tensor"
    assert t.dim() == 4, f"{name} must be 4-D"


class SplitFunction(Function):
    @staticmethod
    def forward(ctx, blocks, image, mapping_exec, grid_idx):
        """blocks[b] <- image block of cell mapping_exec[b]; returns ``blocks``."""
        _check4(blocks, "blocks"); _check4(image, "image")
        assert mapping_exec.dtype == torch.int32 and grid_idx.dtype == torch.int32
        assert blocks.shape[2] == blocks.shape[3]
        assert blocks.shape[1] == image.shape[1]
        BS = blocks.shape[2]
        assert grid_idx.shape[2] * BS == image.shape[2], (grid_idx.shape[2], BS, image.shape[2])
        assert grid_idx.shape[3] * BS == image.shape[3], (grid_idx.shape[3], BS, image.shape[3])
        _C.gather(blocks, image, mapping_exec, len(mapping_exec))
        return blocks

    @staticmethod
    def backward(ctx, grad_x):
        raise NotImplementedError()


class CombineFunction(Function):
    @staticmethod
    def forward(ctx, blocks, out, grid_idx, mapping_exec):
        """out[cell mapping_exec[b]] <- blocks[b] (in place); returns ``out``."""
        _check4(blocks, "blocks"); _check4(out, "out")
        N, C, H, W = out.shape
        BS = blocks.shape[2]
        assert BS >= 1 and BS == blocks.shape[3]
        assert grid_idx.size(1) == 1 and grid_idx.size(0) == N
        assert grid_idx.shape[2] * BS == H and grid_idx.shape[3] * BS == W
        with timings.env("block/combine_kernel", 20):
            _C.scatter(blocks, out, mapping_exec, len(mapping_exec))
        return out

    @staticmethod
    def backward(ctx, grad_x):
        raise NotImplementedError()


class TransferFunction(Function):
    @staticmethod
    def forward(ctx, data_transfer, prev_computed, prev_transfer, grid_idx_prev, transfer_map_prev, padding):
        """Ring (width ``padding``) of every transferred block from the previous frame's executed /
        transferred tiles; interiors of ``data_transfer`` are left as they are."""
        assert data_transfer.shape[1:] == prev_computed.shape[1:]
        assert data_transfer.shape[1:] == prev_transfer.shape[1:]
        assert transfer_map_prev.dtype == torch.int32
        with timings.env("block/transfer_kernel", 20):
            _C.transfer(data_transfer, prev_computed, prev_transfer, transfer_map_prev, grid_idx_prev.numel(),
                        int(padding))
        return data_transfer

    @staticmethod
    def backward(ctx, grad_data_transfer):
        raise NotImplementedError()
