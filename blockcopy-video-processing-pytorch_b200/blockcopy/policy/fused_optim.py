"""torch.optim.RMSprop whose step() is ONE kernel over all parameter tensors (bc_rmsprop_step).

The reference builds ``torch.optim.RMSprop(net.parameters(), lr, weight_decay, momentum)`` for the online policy
update (policy/policy.py:56-59) and steps it every ``block_train_interval`` frames (:361-362).  On a GPU that step
is ~10 multi-tensor launches over ~40 small tensors: 0.45 ms, mostly host time, for 0.6 M parameters.  This
subclass keeps the optimizer's interface, hyper-parameters and ``state_dict`` layout (``step``, ``square_avg``,
``momentum_buffer``) and falls back to the stock implementation whenever a case is outside the kernel's envelope
(closure, centered, maximize, differentiable, capturable, CPU / non-fp32 / non-dense tensors, gradients laid out
differently from their parameters).
"""
from __future__ import annotations

import torch

from .. import _C


# Bumped by every fused step: the kernel updates parameters through raw pointers, which does not touch the tensors'
# `_version` counters -- consumers that cache values derived from parameters (policy/fused_net.py) watch both.
PARAM_EPOCH = [0]


def _dense(t: torch.Tensor) -> bool:
    return t.is_contiguous() or (t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last))


class FusedRMSprop(torch.optim.RMSprop):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._plans = {}  # group index -> (parameter ids, static rows, step tensors, verified gradient pointers)

    # anything that replaces state tensors or parameters invalidates the cached pointer tables
    def load_state_dict(self, state_dict):
        self._plans.clear()
        return super().load_state_dict(state_dict)

    def add_param_group(self, param_group):
        if hasattr(self, "_plans"):
            self._plans.clear()
        return super().add_param_group(param_group)

    @staticmethod
    def _group_ok(group) -> bool:
        return not (group.get("centered") or group.get("maximize") or group.get("differentiable") or group.get("capturable"))

    @staticmethod
    def _pair_ok(p, g, dev) -> bool:
        return (p.is_cuda and p.device == dev and p.dtype == torch.float32 and _dense(p) and g.dtype == torch.float32
                and not g.is_sparse and g.device == dev and g.stride() == p.stride())

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None or not all(self._group_ok(g) for g in self.param_groups):
            return super().step(closure)
        work = []
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            dev = params[0].device
            plan = self._plans.get(gi)
            ids = tuple(id(p) for p in params)
            if plan is None or plan[0] != ids:
                # (re)build the static part: state tensors (created like torch does), offsets, pointers
                rows, steps, off = [], [], 0
                for p in params:
                    if not p.is_cuda:
                        return super().step()
                    st = self.state[p]
                    if len(st) == 0:
                        st["step"] = torch.zeros((), dtype=torch.float32)
                        st["square_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                        if group["momentum"] > 0:
                            st["momentum_buffer"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    buf = st["momentum_buffer"].data_ptr() if group["momentum"] > 0 else 0
                    rows.append([p.data_ptr(), 0, st["square_avg"].data_ptr(), buf, off, p.numel()])
                    steps.append(st["step"])
                    off += p.numel()
                plan = self._plans[gi] = (ids, rows, steps, [0] * len(params))
            _, rows, steps, seen = plan
            stale = False
            for i, p in enumerate(params):  # ~40 cheap host comparisons: state replaced behind our back (state.clear(),
                st = self.state[p]          # manual assignment) or a parameter re-allocated (.data = ...)
                if rows[i][0] != p.data_ptr() or st.get("square_avg") is None or \
                        rows[i][2] != st["square_avg"].data_ptr() or steps[i] is not st.get("step") or \
                        (rows[i][3] and rows[i][3] != st["momentum_buffer"].data_ptr()):
                    stale = True
                    break
            if stale:
                self._plans.pop(gi, None)
                return self.step()
            for i, p in enumerate(params):
                g = p.grad
                ptr = g.data_ptr()
                if ptr != seen[i]:  # a gradient tensor not seen before: check layout / dtype once
                    if rows[i][0] != p.data_ptr() or not self._pair_ok(p, g, dev):
                        self._plans.pop(gi, None)
                        return super().step()
                    seen[i] = ptr
                rows[i][1] = ptr
            work.append((group, rows, steps, dev))
        if work:
            PARAM_EPOCH[0] += 1
        for group, rows, steps, dev in work:
            torch._foreach_add_(steps, 1)
            _C.rmsprop_step(rows, dev, group["lr"], group["alpha"], group["eps"], group["weight_decay"], group["momentum"])
        return None
