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

from ..utils.profiler import timings
from .tensorwrapper import TensorWrapper, to_tensorwrapper


class BlockCopyModel(nn.Module):
    """Wraps ``base_model`` for block-sparse execution with temporal feature propagation.

    settings: the ``block_*`` dict produced by ``add_argparser_arguments`` (all keys are read).
    Optional extra keys (absent => reference behaviour):
      block_channels_last (bool, default True): store conv weights channels_last so that packed
          tiles and planes are NHWC, the layout the sm_100a kernels are written for.
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

    def load_state_dict(self, state_dict, strict: bool = True):
        """Checkpoints are base-model checkpoints (reference core/blockcopy.py:30-32)."""
        return self.base_model.load_state_dict(state_dict, strict=strict)

    def reset_temporal(self):
        """Forget all temporal state; call at the start of every clip."""
        self.clip_length = 0
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
            meta = self.policy(meta)  # sets grid, num_exec, num_total, perc_exec
            self.policy_meta = meta

        with timings.env("blockcopy/model", 3):
            x = to_tensorwrapper(inputs)
            if meta["num_exec"] == 0:
                # nothing to execute: the previous output object is returned, no state is touched
                meta = self.policy_meta = meta.copy()
                out = meta["outputs"]
            else:
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


def blockcopy_noblocks(func):
    """Decorator for ``forward`` methods that cannot run on blocks (e.g. global pooling): the
    blocks are combined in place into the dense tensor, the method runs densely, and its result
    is split again with the same grid.  Costs a combine and a split."""

    @functools.wraps(func)
    def noblocks(self, x, *args):
        was_wrapper = isinstance(x, TensorWrapper)
        if was_wrapper:
            blocks = x
            x = x.combine_().to_tensor()
        x = func(self, x)
        if was_wrapper:
            x = to_tensorwrapper(x).to_blocks_like(blocks)
        return x

    return noblocks
