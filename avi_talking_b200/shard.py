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


def bind_to_gpu_numa(local_rank: int) -> dict:
    """Pin the calling process to the CPUs (and thereby, under Linux first-touch, the host memory) closest to GPU `local_rank`.

    One process per GPU on a two-socket host otherwise allocates every rank's pinned staging buffers wherever the launcher happened to
    run: measured on an 8 x B200 box, the end-to-end (host-buffer) throughput then stops scaling beyond 2 GPUs - per-GPU D2H drops from
    50 GB/s to 11 GB/s - while the device-resident number scales 8.06x. NVML knows the ideal CPU set of each GPU
    (nvmlDeviceSetCpuAffinity); binding before the first pinned allocation keeps each rank's PCIe traffic on its own socket.
    Best effort: returns {"bound": False, "why": ...} when NVML or the affinity call is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            index = local_rank
            if visible:
                ids = [v.strip() for v in visible.split(",") if v.strip()]
                if local_rank < len(ids) and ids[local_rank].isdigit():
                    index = int(ids[local_rank])
            handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            pynvml.nvmlDeviceSetCpuAffinity(handle)
            cpus = sorted(os.sched_getaffinity(0))
            return {"bound": True, "cpus": f"{cpus[0]}-{cpus[-1]} ({len(cpus)})" if cpus else ""}
        finally:
            pynvml.nvmlShutdown()
    except Exception as e:  # noqa: BLE001 - NVML missing / container without the capability: run unbound
        return {"bound": False, "why": f"{type(e).__name__}: {e}"[:120]}
