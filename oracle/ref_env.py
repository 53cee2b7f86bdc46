"""oracle/ref_env.py -- TEST INFRASTRUCTURE: the reference importable for the golden-vector generators.

The environment itself (staging, cupy stub / NVRTC shim) lives in baseline/ref_env.py; this file adds the
one oracle-specific piece: in CPU mode the reference's four kernel call sites (tensorwrapper.py:10-11
imports) are rebound to the C oracle, and the CUDA assert of ``to_tensorwrapper`` is lifted.
"""
from __future__ import annotations

import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from baseline.ref_env import reference_roots, stage_reference  # noqa: E402,F401
from baseline import ref_env as _env  # noqa: E402


def import_reference(mode: str):
    """mode in {'cpu', 'gpu'}; returns the reference ``blockcopy`` module."""
    blockcopy = _env.import_reference(mode)
    if mode == "cpu":
        _bind_oracle_kernels(blockcopy)
    return blockcopy


def _bind_oracle_kernels(blockcopy):
    """Rebind the four kernel call sites of the reference wrapper to the C oracle (CPU tensors)."""
    import torch

    sys.path.insert(0, ROOT)
    from oracle import cpu_oracle as O

    tw = sys.modules["blockcopy.core.tensorwrapper"]

    def plain(t):
        return t.as_subclass(torch.Tensor).contiguous()

    class Split:
        @staticmethod
        def apply(blocks, image, mapping_exec, grid_idx):
            if len(mapping_exec):
                blocks.copy_(O.split(plain(image), mapping_exec.contiguous(), blocks.shape[2]))
            return blocks

    class Combine:
        @staticmethod
        def apply(blocks, out, grid_idx, mapping_exec):
            assert out.is_contiguous()
            if len(mapping_exec):
                O.combine_(plain(blocks), out.as_subclass(torch.Tensor), mapping_exec.contiguous())
            return out

    class Transfer:
        @staticmethod
        def apply(data_transfer, prev_computed, prev_transfer, grid_idx_prev, transfer_map_prev, padding):
            # poison: the interior is never written by the reference kernel and must never be read
            data_transfer.fill_(float("nan"))
            if len(transfer_map_prev):
                O.transfer(data_transfer.as_subclass(torch.Tensor), plain(prev_computed), plain(prev_transfer),
                           transfer_map_prev.contiguous(), grid_idx_prev.numel(), padding)
            return data_transfer

    def pad(features, transfer, grid_idx, exec_map, pad=1):
        return O.repad(plain(features), plain(transfer), grid_idx.contiguous(), exec_map.contiguous(), pad)

    tw.SplitFunction, tw.CombineFunction, tw.TransferFunction, tw.pad = Split, Combine, Transfer, pad
    tw.to_tensorwrapper = lambda x: x.as_subclass(tw.TensorWrapper)
    blockcopy.to_tensorwrapper = tw.to_tensorwrapper
    if not torch.cuda.is_available():
        torch.cuda.empty_cache = lambda: None
