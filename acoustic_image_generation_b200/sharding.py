"""Multi-GPU plumbing of the path: frames (or whole clips) shard across ranks with no data-path
collective; the only exchange is one all-reduce(sum) of the int64[K+1] success-count vector
(pos[0..K-1], num) at the end of an evaluation.

One process per GPU, torch.distributed for the plumbing: backend "nccl" on the GPUs (NVLink /
NVSwitch; the payload is <= 8 KiB, so the collective is latency-bound), "gloo" in the CPU tests.
The reference itself is single-GPU and re-runs the evaluation once per threshold
(scripts/iou.bash:47-53); sharding + one reduction replaces that.
"""
from __future__ import annotations

import numpy as np


def shard_range(n_items, rank, world):
    """Contiguous [start, stop) of `n_items` owned by `rank`: sizes differ by at most one, earlier ranks
    take the remainder, empty shards are allowed."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError('bad rank/world %r/%r' % (rank, world))
    base, extra = divmod(int(n_items), world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_clips(n_clips, frames_per_clip, rank, world):
    """Clip-major sharding (BASELINE configs[3]: 10 s clips of 120 or 300 frames): returns
    (first clip, stop clip, first frame, stop frame); a clip never straddles two ranks."""
    c0, c1 = shard_range(n_clips, rank, world)
    return c0, c1, c0 * frames_per_clip, c1 * frames_per_clip


def allreduce_counts(counts, group=None):
    """Sum the int64 count vector over all ranks in place and return it.

    `counts` is a torch int64 tensor (CUDA for nccl, CPU for gloo) or a NumPy int64 array (gloo /
    single process).  Without an initialised process group this is the identity."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return counts
    if isinstance(counts, np.ndarray):
        t = torch.from_numpy(counts)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return counts
    if counts.dtype != torch.int64:
        raise TypeError('counts must be int64')
    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


def merge_counts(shards):
    """Host-side equivalent of the all-reduce for results gathered some other way."""
    return np.sum(np.stack([np.asarray(s, dtype=np.int64) for s in shards], 0), axis=0)
