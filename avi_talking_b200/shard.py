"""Clip sharding across the GPUs of one node: every clip is independent, so inference partitions the clip
batch with NO data-path collective - clip c goes to rank c mod G and results stay on the producing GPU. torch.distributed
(NCCL on the B200s, gloo in the CPU tests) is used for the rendezvous, the timing barrier and the max-over-ranks reduction
of the measured time only."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def clip_shard(n_clips: int, rank: int, world: int) -> list[int]:
    """Indices of the clips rank `rank` of `world` processes (round-robin: equal clip lengths => perfect balance)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, n_clips, world))


def env_rank_world() -> tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1-process defaults)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def max_over_ranks(values: list[float], device="cpu") -> list[float]:
    """Element-wise max over ranks of a small list of floats (the timed milliseconds); identity when not distributed."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(values)
    t = torch.tensor(values, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def gather_clip_counts(n_local: int, device="cpu") -> list[int]:
    """How many clips every rank processed (for the whole-job frames/s figure)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [n_local]
    t = torch.zeros(dist.get_world_size(), dtype=torch.int64, device=device)
    t[dist.get_rank()] = n_local
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.tolist()
