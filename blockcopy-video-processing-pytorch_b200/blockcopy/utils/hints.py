"""Executed-block count carried on a grid tensor.

The blockcopy API exposes ``num_exec`` as a Python int (reference policy/policy.py:82-94 computes
``int(grid.sum())``, a device round trip).  Policies that already know the count on the host attach it to the
grid as ``grid._bc_num_exec = (count, grid._version)``; it is honoured only while the tensor's in-place version
counter still has that value, so a grid that was edited after the count was taken (a custom policy mutating it,
anything touching ``policy_meta['grid']``) falls back to counting on the device like the reference.
"""
from __future__ import annotations

from typing import Optional

import torch

_ATTR = "_bc_num_exec"


def set_num_exec_hint(grid: torch.Tensor, count: int) -> torch.Tensor:
    try:
        setattr(grid, _ATTR, (int(count), grid._version))
    except AttributeError:
        pass
    return grid


def get_num_exec_hint(grid: torch.Tensor) -> Optional[int]:
    """The attached count if it is still valid for the tensor's current content, else None."""
    hint = getattr(grid, _ATTR, None)
    if not isinstance(hint, tuple) or len(hint) != 2 or hint[1] != grid._version:
        return None
    return hint[0]
