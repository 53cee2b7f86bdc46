"""Driver-side steps around BlockCopyModel, each one sm_100a kernel.

The reference's driver decodes a frame on the CPU (ExtToTensor + ExtNormalize,
semantic_segmentation/lib/ext_transforms.py:317-372), uploads it as fp32/fp16 and, for the label map,
upsamples the logits to the frame size and takes the class maximum (test_swiftnet.py:187,196-197).  Here the
host uploads the uint8 frame (3 bytes per pixel) and the device normalises it; the label map comes from one
kernel that never writes the upsampled logits.
"""
from __future__ import annotations

import torch

from blockcopy import _C

CITYSCAPES_MEAN = (73.1584 / 255, 82.9090 / 255, 72.3924 / 255)  # lib/datasets/cityscapes_vid.py:29-30
CITYSCAPES_STD = (44.9149 / 255, 46.1529 / 255, 45.3192 / 255)


class FrameNormalizer:
    """uint8 (N,H,W,3) / (H,W,3) CUDA frame -> normalised (N,3,H,W) network input; same bits as
    to_tensor + normalize + .to(dtype)."""

    def __init__(self, mean=CITYSCAPES_MEAN, std=CITYSCAPES_STD, dtype=torch.float16):
        self.mean, self.std, self.dtype = tuple(mean), tuple(std), dtype

    def __call__(self, frame_u8: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        return _C.frame_from_u8(frame_u8, self.mean, self.std, self.dtype, out)


def predict_labels(logits: torch.Tensor, size=None, label_dtype=torch.uint8, out: torch.Tensor = None) -> torch.Tensor:
    """Label map of dense logits (N,K,h,w): argmax over classes of the bilinear upsampling to `size`
    (default 4x, SwiftNet's logits stride).  Equals F.interpolate(logits, size, mode='bilinear').max(1)[1]."""
    N, K, h, w = logits.shape
    if size is None:
        scale = 4
    else:
        H, W = size
        if H % h or W % w or H // h != W // w or H // h not in (1, 2, 4):
            raise NotImplementedError(f"predict_labels: {h}x{w} -> {H}x{W} (integer scale 1, 2 or 4)")
        scale = H // h
    return _C.upsample_argmax(logits, scale, label_dtype, out)


class BlockLabelMap:
    """Stateful label map of one video stream: the driver-side counterpart of BlockCopy's own idea.  The dense logits
    of a frame differ from the previous frame's only inside the executed blocks, so only those blocks (plus a ring of
    one logit pixel, which the bilinear taps of the neighbours reach) need a new arg-max; the rest of the label map is
    still right.  ``update`` returns the full label map, equal to ``predict_labels(logits)`` bit for bit.

        labels = BlockLabelMap()
        for frame in clip:
            out = model(frame)
            lab = labels.update(out, model.policy_meta["grid"])      # (N, 4h, 4w) uint8, updated in place
        labels.reset()                                               # with model.reset_temporal()
    """

    def __init__(self, scale: int = 4, label_dtype=torch.uint8, out: torch.Tensor = None):
        """out: optional caller-owned (N, scale*h, scale*w) label buffer to keep updated in place."""
        self.scale, self.label_dtype, self.labels = scale, label_dtype, out
        self._valid = False  # does `labels` hold the previous frame's label map?

    def reset(self):
        self._valid = False

    def update(self, logits: torch.Tensor, grid: torch.Tensor = None) -> torch.Tensor:
        N, K, h, w = logits.shape
        shape = (N, h * self.scale, w * self.scale)
        fits = self.labels is not None and tuple(self.labels.shape) == shape and self.labels.device == logits.device
        sparse = fits and self._valid and grid is not None and grid.dim() == 4 and grid.shape[0] == N \
            and h % grid.shape[2] == 0 and w % grid.shape[3] == 0 and h // grid.shape[2] == w // grid.shape[3]
        if sparse:
            _C.upsample_argmax_blocks(self.labels, logits, grid, self.scale)
        else:
            self.labels = _C.upsample_argmax(logits, self.scale, self.label_dtype, self.labels if fits else None)
            self._valid = True
        return self.labels
