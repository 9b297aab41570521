"""Frame-parallel sharding of a camera-path batch over ranks (BASELINE config 5).

Frames are independent units (the reference's BeginFrame resets everything, SoftRast/Renderer.cpp:201-204), so the
path is cut into contiguous per-rank arcs with no data-path collective.  The only cross-rank step is the reduction of
the per-rank device time to its maximum, which bench.py does through torch.distributed."""
from __future__ import annotations

import numpy as np


def frames_for_rank(step: int, frames_per_step: int, rank: int, world: int, path_frames: int) -> np.ndarray:
    """Camera indices rank `rank` renders in step `step`: every rank walks its own arc of the closed path."""
    base = (step * frames_per_step + rank * (path_frames // max(1, world))) % path_frames
    return (base + np.arange(frames_per_step)) % path_frames


def tiles_for_rank(num_tiles: int, rank: int, world: int) -> np.ndarray:
    """Screen-tile split of one frame (BASELINE config 4): interleaved ownership, tile t belongs to rank t % world."""
    return np.arange(rank, num_tiles, world)


def reduce_max_ms(ms: float, dist=None, device=None) -> float:
    """Max over ranks of a device time (the job's step time)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return ms
    import torch

    t = torch.tensor([ms], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
