"""Sample sharding of the BEV-projection path across the GPUs of one node.

The path has no cross-sample dependency (``batch_idx = pt_idx / num_points``,
``ops/voxel_pooling/src/voxel_pooling_forward_cuda.cu:19`` of the reference; the voxelizer runs per
sample, ``models/bev_depth.py:180-181``), so multi-GPU is plain partitioning: one process per GPU,
every rank owns a contiguous block of frames, and there is NO collective on the data path.  The only
communication is the reduction of the timing scalars for a throughput report.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [begin, end) of ``total`` frames owned by ``rank``; sizes differ by at most 1
    and earlier ranks take the remainder (strong scaling at a fixed global batch)."""
    if world <= 0 or not 0 <= rank < world or total < 0:
        raise ValueError(f'bad shard request total={total} rank={rank} world={world}')
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def aggregate_throughput(frames_local: int, elapsed_ms_local: float, device='cpu') -> Tuple[float, float, int]:
    """Whole-job throughput = frames of ALL ranks / the SLOWEST rank's device time.
    Returns (frames_per_s, max_ms, total_frames); identical on every rank.  Works without an
    initialised process group (single process)."""
    t = torch.tensor([float(elapsed_ms_local)], dtype=torch.float64, device=device)
    n = torch.tensor([int(frames_local)], dtype=torch.int64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
    max_ms, total = float(t.item()), int(n.item())
    return (total / (max_ms * 1e-3) if max_ms > 0 else float('inf')), max_ms, total
