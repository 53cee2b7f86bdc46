"""Synthetic Cityscapes-shaped clips and the seeded fixed-fraction policy used by the benchmark
and the parity tests: identical masks for the reference and this
implementation, generated on the HOST, so no device round trip is needed to learn E.
"""
from __future__ import annotations

import torch

from blockcopy.policy.policy import Policy


def synthetic_clip(num_frames: int, height: int = 1024, width: int = 2048, seed: int = 0, batch: int = 1,
                   dtype=torch.float16, device="cpu", change_prob: float = 0.3, change_scale: float = 0.3):
    """f0 = randn; f_{t+1} = f_t + change_scale * randn * (rand > 1 - change_prob).  Returns a list of
    (batch,3,H,W) tensors.  Generated in fp32 on the CPU generator so that it is device independent."""
    g = torch.Generator().manual_seed(seed)
    frame = torch.randn(batch, 3, height, width, generator=g)
    frames = []
    for _ in range(num_frames):
        frames.append(frame.to(dtype=dtype).to(device))
        noise = torch.randn(frame.shape, generator=g)
        gate = torch.rand(frame.shape, generator=g) > (1 - change_prob)
        frame = frame + change_scale * noise * gate
    return frames


class PolicyFixedFraction(Policy):
    """Executes all blocks on the first frame of a clip, then exactly ``num_exec`` randomly chosen
    blocks per image (seeded, host RNG).  A ``Policy`` plug-in like any other
    (``model.policy = PolicyFixedFraction(...)``)."""

    def __init__(self, block_size: int, fraction: float = 0.3, quantize: int = 8, seed: int = 0):
        super().__init__(block_size, verbose=False, quantize_number_exec=0)
        self.fraction = fraction
        self.quantize = quantize
        self.seed = seed
        self._gen = torch.Generator().manual_seed(seed)

    def reseed(self, seed=None):
        self._gen = torch.Generator().manual_seed(self.seed if seed is None else seed)

    def num_exec_for(self, cells_per_image: int) -> int:
        e = max(1, round(self.fraction * cells_per_image))
        if self.quantize > 1:
            e = min(cells_per_image, self.quantize * (1 + (e - 1) // self.quantize))
        return e

    def forward(self, policy_meta: dict) -> dict:
        N, _, H, W = policy_meta["inputs"].shape
        assert H % self.block_size == 0 and W % self.block_size == 0
        GH, GW = H // self.block_size, W // self.block_size
        shape = (N, 1, GH, GW)
        dev = policy_meta["inputs"].device
        if policy_meta.get("outputs", None) is None:
            grid = torch.ones(shape, dtype=torch.bool)
        else:
            cells = GH * GW
            e = self.num_exec_for(cells)
            grid = torch.zeros(N, cells, dtype=torch.bool)
            for n in range(N):
                grid[n, torch.randperm(cells, generator=self._gen)[:e]] = True
            grid = grid.view(shape)
        count = int(grid.sum())
        grid = grid.to(dev, non_blocking=True)
        grid._bc_num_exec = (count, grid._version)  # blockcopy/utils/hints.py (inline: this file also runs against the reference package)
        policy_meta["grid"] = grid
        return self.stats.add_policy_meta(policy_meta)


class PolicyReplay(Policy):
    """Replays a recorded list of grids (one per frame of a clip); used to feed the reference and
    this implementation exactly the same masks."""

    def __init__(self, block_size: int, grids):
        super().__init__(block_size, verbose=False, quantize_number_exec=0)
        self.grids = [g.clone() for g in grids]
        self.t = 0

    def rewind(self):
        self.t = 0

    def forward(self, policy_meta: dict) -> dict:
        g = self.grids[self.t].to(torch.bool)
        self.t += 1
        count = int(g.sum())
        g = g.to(policy_meta["inputs"].device)
        g._bc_num_exec = (count, g._version)
        policy_meta["grid"] = g
        return self.stats.add_policy_meta(policy_meta)


def deterministic_init_(model: torch.nn.Module, seed: int = 0, gain: float = 1.0) -> torch.nn.Module:
    """Fill every parameter and buffer from a generator seeded by (seed, tensor NAME), so that two
    implementations of the same architecture (the reference's module tree and this repo's) get
    identical weights without shipping a checkpoint.  Conv weights ~ N(0, gain^2 * 2/fan_out), BatchNorm
    affine / running statistics are made non-trivial on purpose."""
    import zlib

    def gen(name):
        return torch.Generator().manual_seed((zlib.crc32(name.encode()) + 7919 * seed) % (2 ** 31))

    with torch.no_grad():
        for name, p in list(model.named_parameters()) + list(model.named_buffers()):
            g = gen(name)
            if name.endswith("num_batches_tracked"):
                continue
            if p.dim() == 4:
                fan_out = p.shape[0] * p.shape[2] * p.shape[3]
                v = torch.randn(p.shape, generator=g) * (gain * (2.0 / fan_out) ** 0.5)
            elif name.endswith("running_var") or (name.endswith("weight") and p.dim() == 1):
                v = 0.75 + 0.5 * torch.rand(p.shape, generator=g)
            else:
                v = 0.1 * torch.randn(p.shape, generator=g)
            p.copy_(v.to(p.dtype))
    return model
