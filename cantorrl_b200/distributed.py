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


def _parse_cpulist(text: str) -> set:
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_local_cpus(pci_bus_id: str, sysfs: str = "/sys/bus/pci/devices") -> set:
    """CPUs on the NUMA node whose PCIe root the GPU ``pci_bus_id`` ("0000:1b:00.0") hangs off (sysfs ``local_cpulist``);
    empty when sysfs does not say (a VM without NUMA information, node −1)."""
    try:
        with open(os.path.join(sysfs, pci_bus_id.lower(), "local_cpulist")) as f:
            return _parse_cpulist(f.read())
    except (OSError, ValueError):
        return set()


def bind_to_gpu_numa_node(device_index: int, sysfs: str = "/sys/bus/pci/devices"):
    """Restrict this process to the CPUs local to CUDA device ``device_index`` BEFORE it page-locks its staging buffers.

    The host-buffer entry points (``cantor_vecenv_step_host``) move 65 B per env-step over PCIe; Linux places page-locked
    memory on the NUMA node of the allocating thread, so an unbound rank of an 8-GPU job stages half of its traffic
    through the other socket.  One process per GPU + this call keeps every rank's DMA on its own root complex.
    Never fails: returns the CPU set it bound to, or None when the topology is unknown, the local CPUs are outside the
    cgroup's cpuset, or ``CANTOR_NO_NUMA_BIND`` is set.
    """
    if os.environ.get("CANTOR_NO_NUMA_BIND") or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        allowed = os.sched_getaffinity(0)
        cpus = gpu_local_cpus(bus, sysfs) & allowed
        if not cpus or cpus == allowed:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:           # topology probing must never take a job down
        return None
