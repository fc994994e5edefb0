"""Sharding of the env population over GPUs (one process per GPU) -- the only multi-GPU plumbing the path needs.

Every env owns its path, its state and its Philox counters; nothing in ``HedgingEnv.step``
(src/env/hedging_env_v2.py:175-294) or in the simulator's outer step (src/sim/rbergomi_sim.py:454-464) couples two
envs.  So rank ``g`` of ``G`` takes the contiguous block ``[g * N / G, (g + 1) * N / G)`` of the GLOBAL env index,
passes its ``env_offset`` to the kernels (which put the global index into the counters), and the only exchange is
``EpisodeStats.all_reduce`` -- a sum of a few KB over NCCL / NVLink per reporting interval.
"""
from __future__ import annotations

import os
from typing import Tuple


def rank_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1-process defaults)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def shard(total_envs: int, rank: int, world: int) -> Tuple[int, int]:
    """(env_offset, num_envs) of ``rank``: contiguous blocks, the first ``total_envs % world`` ranks get one more."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, extra = divmod(int(total_envs), int(world))
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, count


def init_process_group(backend: str = "nccl", device=None):
    """``torch.distributed`` rendezvous from the torchrun environment; no-op for a single process."""
    import torch.distributed as dist
    rank, local_rank, world = rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kw = {}
        if backend == "nccl" and device is not None:
            kw["device_id"] = device
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world
