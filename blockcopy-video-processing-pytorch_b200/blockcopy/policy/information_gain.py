"""Information gain = the reward signal of the online policy (reference policy/information_gain.py).

SemSeg: KL(prev || cur) of the soft-maxed logits at 1/4 of the output resolution, averaged over
classes (:22-41).  The object-detection variant (box rasterisation with Python loops on the host,
:43-160) belongs to the Pedestron consumer and is outside this repo's hot-path scope (SURVEY.md
section 8(f) rank 3); requesting it raises NotImplementedError.
"""
import abc
from typing import Dict

import torch
import torch.nn as nn
import torch.nn.functional as F


class InformationGain(nn.Module, metaclass=abc.ABCMeta):
    def __init__(self, num_classes):
        super().__init__()
        self.num_classes = num_classes

    def get_output_repr(self, policy_meta: Dict) -> torch.Tensor:
        raise NotImplementedError

    def forward(self, policy_meta: Dict) -> torch.Tensor:
        raise NotImplementedError


class InformationGainSemSeg(InformationGain):
    def __init__(self, num_classes):
        super().__init__(num_classes)
        self.scale_factor = 1 / 4

    def get_output_repr(self, policy_meta: Dict) -> torch.Tensor:
        out = policy_meta["outputs"]
        assert out.size(1) == self.num_classes
        return out

    def forward(self, policy_meta: Dict) -> torch.Tensor:
        cur, prev = policy_meta["outputs"], policy_meta["outputs_prev"]
        assert cur is not None and prev is not None
        if cur.is_cuda and cur.dtype == torch.float16 and prev.dtype == torch.float16 and cur.shape == prev.shape \
                and cur.dim() == 4 and cur.shape[1] <= 64 and cur.shape[2] % 4 == 0 and cur.shape[3] % 4 == 0 \
                and self.scale_factor == 1 / 4:
            from blockcopy import _C

            return _C.info_gain(cur, prev)  # one sm_100a kernel (bc_info_gain) instead of six
        cur = F.interpolate(cur, scale_factor=self.scale_factor, mode="bilinear")
        prev = F.interpolate(prev, scale_factor=self.scale_factor, mode="bilinear")
        # elementwise p_prev * (log p_prev - log p_cur), then mean over classes
        kl = F.kl_div(input=F.log_softmax(cur, dim=1), target=F.log_softmax(prev, dim=1),
                      reduction="none", log_target=True)
        return kl.mean(1, keepdim=True)


class InformationGainObjectDetection(InformationGain):
    def get_output_repr(self, policy_meta: Dict) -> torch.Tensor:
        raise NotImplementedError("rl_objectdetection (Pedestron) is out of scope of this build; see DESIGN.md")

    forward = get_output_repr
