"""Sample-sharded data parallelism for the pose path (one process per GPU).

The reference is single-process; every sample (frame triple) is independent, so
the stream is cut into contiguous blocks by rank, each rank runs the forward on
its block, and the ``[n_local, 2, 6]`` poses are gathered once: on GPUs by the
library's own ``davo_allgather_poses`` (ncclAllGather over NVLink on the compute
stream, include/davo_b200.h) when the ``DAVO`` object carries a communicator,
else by ``torch.distributed.all_gather_into_tensor`` (gloo on CPU for tests).  Shards are padded to equal length by repeating the last sample
-- the rule the reference uses to fill its final batch (reference
``utils/common_utils.py:8-13``) -- and the padding is trimmed after the gather.
Trajectory composition stays sequential on the host.
"""
from __future__ import annotations

from typing import Tuple


def complete_batch_size(input_list, batch_size):
    """Pad a list to a multiple of batch_size by repeating its last item
    (reference utils/common_utils.py:8-13)."""
    left = len(input_list) % batch_size
    if left != 0:
        input_list.extend([input_list[-1]] * (batch_size - left))
    return input_list


def is_valid_sample(frames, tgt_idx, seq_length):
    """A target index is valid when both neighbours exist in the same drive
    (reference utils/common_utils.py:16-28)."""
    n = len(frames)
    half = int((seq_length - 1) / 2)
    lo, hi = tgt_idx - half, tgt_idx + half
    if lo < 0 or hi >= n:
        return False
    drive = frames[tgt_idx].split(' ')[0]
    return frames[lo].split(' ')[0] == drive and frames[hi].split(' ')[0] == drive


def shard_range(n_samples: int, rank: int, world: int) -> Tuple[int, int, int]:
    """(start, stop, n_local): contiguous block of rank; n_local = ceil(N / world) is
    the padded per-rank count, [start, stop) the real samples (may be empty)."""
    n_local = -(-n_samples // world)
    start = min(rank * n_local, n_samples)
    stop = min(start + n_local, n_samples)
    return start, stop, n_local


def padded_indices(n_samples: int, rank: int, world: int):
    """Sample indices rank processes, padded to n_local by repeating the stream's last sample."""
    start, stop, n_local = shard_range(n_samples, rank, world)
    idx = list(range(start, stop))
    idx.extend([n_samples - 1] * (n_local - len(idx)))
    return idx


def gather_poses(local_poses, n_samples: int, system=None):
    """All-gather ``[n_local, 2, 6]`` from every rank and trim to ``[n_samples, 2, 6]``.

    ``local_poses`` is a torch tensor (CUDA with NCCL, CPU with gloo).  Without an
    initialised process group (single GPU) it is returned trimmed.  ``system`` is the
    rank's ``DAVO`` object: after ``system.init_comm()`` the gather is the C library's.
    """
    import torch
    import torch.distributed as dist
    if system is not None and system.comm_world() > 1 and local_poses.is_cuda:
        return system.allgather_poses(local_poses)[:n_samples]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_poses[:n_samples]
    world = dist.get_world_size()
    n_local = local_poses.shape[0]
    out = torch.empty((world * n_local,) + tuple(local_poses.shape[1:]), dtype=local_poses.dtype,
                      device=local_poses.device)
    dist.all_gather_into_tensor(out, local_poses.contiguous())
    return out[:n_samples]
