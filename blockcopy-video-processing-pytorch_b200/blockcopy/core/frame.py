"""A decoded uint8 video frame as the model's input: normalised lazily, block by block (input side).

The reference's driver normalises every frame on the CPU and uploads it as fp16 / fp32
(semantic_segmentation/lib/ext_transforms.py:317-372, test_swiftnet.py:64-65,187).  With ``U8Frame`` the host uploads
the uint8 frame as decoded, and ``BlockCopyModel`` -- which only ever reads the EXECUTED blocks of a steady frame --
normalises exactly those inside the frame's first gather (bc_blocks_from_u8): the normalised full frame is never
written.  Whoever needs all of it (the first frame of a clip, a policy network that looks at the whole frame) calls
``materialize()``, which is bc_frame_from_u8, cached per frame.

    frame = U8Frame(u8_hw3_cuda)                 # (H,W,3) or (N,H,W,3) uint8 on the GPU
    out = model(frame)                           # same bits as model(frame.materialize())
"""
from __future__ import annotations

from typing import Optional

import torch

CITYSCAPES_MEAN = (73.1584 / 255, 82.9090 / 255, 72.3924 / 255)  # lib/datasets/cityscapes_vid.py:29-30
CITYSCAPES_STD = (44.9149 / 255, 46.1529 / 255, 45.3192 / 255)


class U8Frame:
    """Stands in for the (N,3,H,W) normalised input tensor: has its shape / dtype / device, holds the uint8 pixels."""

    def __init__(self, u8: torch.Tensor, mean=CITYSCAPES_MEAN, std=CITYSCAPES_STD, dtype=torch.float16,
                 out: Optional[torch.Tensor] = None):
        """out: optional caller-owned (N,3,H,W) buffer for materialize()."""
        if u8.dim() == 3:
            u8 = u8.unsqueeze(0)
        assert u8.is_cuda and u8.dtype == torch.uint8 and u8.dim() == 4 and u8.shape[3] == 3 and u8.is_contiguous(), \
            "U8Frame: a contiguous (N,H,W,3) uint8 CUDA tensor"
        self.u8, self.mean, self.std, self.dtype = u8, tuple(mean), tuple(std), dtype
        N, H, W, _ = u8.shape
        self.shape = torch.Size((N, 3, H, W))
        self.device, self.is_cuda = u8.device, True
        self._out, self._dense = out, None

    def dim(self) -> int:
        return 4

    def size(self, i: Optional[int] = None):
        return self.shape if i is None else self.shape[i]

    def materialize(self) -> torch.Tensor:
        """The whole normalised frame (bc_frame_from_u8), computed once per frame object."""
        if self._dense is None:
            from .. import _C

            self._dense = _C.frame_from_u8(self.u8, self.mean, self.std, self.dtype, self._out)
        return self._dense


def as_tensor(frame) -> torch.Tensor:
    """The dense input tensor of `frame` (a tensor, or a U8Frame)."""
    return frame.materialize() if isinstance(frame, U8Frame) else frame
