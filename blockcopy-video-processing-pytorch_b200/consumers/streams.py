"""Stream sharding for multi-GPU runs: video streams are the independent unit of the path
(SURVEY.md 8(e)); one process per GPU, static assignment, no collective on the frame path.  The
only communication is the throughput aggregation of a benchmark (one MAX and one SUM of a scalar)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_streams(num_streams: int, world_size: int, rank: int):
    """Stream ids owned by `rank` under the static `stream_id % world_size` assignment."""
    assert 0 <= rank < world_size
    return [s for s in range(num_streams) if s % world_size == rank]


def aggregate_throughput(frames_local: float, elapsed_ms_local: float, device="cpu"):
    """Whole-job throughput: (sum of frames over ranks) / (max elapsed over ranks).  Returns
    (frames_total, elapsed_ms_max, frames_per_s); identical on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return frames_local, elapsed_ms_local, frames_local / (elapsed_ms_local * 1e-3)
    t = torch.tensor([elapsed_ms_local], dtype=torch.float64, device=device)
    f = torch.tensor([frames_local], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(f, op=dist.ReduceOp.SUM)
    return float(f), float(t), float(f) / (float(t) * 1e-3)
