"""Multi-GPU plumbing: one process per GPU, `torch.distributed` (NCCL over NVLink on the box, gloo in
the CPU tests).  The path shards by independent streams: every rank renders the partial
mix of its contiguous slice of the batch with no data-path exchange, then ONE all-reduce(sum, f32) of
the `frames * channels` mix (192 KB per second of mono 48 kHz audio) produces the MixerSource output
on every rank.  Inputs never move between GPUs.
"""
from __future__ import annotations

import os
from typing import Tuple


def shard_range(n_streams: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous index range [lo, hi) of rank `rank` — keeps the mixer's insertion order inside a shard
    (src/mixer.rs:185-198 sums in insertion order)."""
    if world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError("bad rank / world_size")
    return n_streams * rank // world_size, n_streams * (rank + 1) // world_size


class DeviceArray:
    """Zero-copy view of a device pointer for torch (`torch.as_tensor(DeviceArray(...), device='cuda')`)."""

    def __init__(self, ptr: int, n: int, typestr: str = "<f4"):
        self.__cuda_array_interface__ = {
            "shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2, "strides": None,
        }


def env_rank() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1-process defaults)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init_process_group(backend: str = "nccl"):
    """Initialise torch.distributed from MASTER_ADDR/MASTER_PORT/RANK/WORLD_SIZE (no-op for world size 1)."""
    import torch.distributed as dist
    rank, local_rank, world = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kw = {}
        if backend == "nccl":
            import torch
            # NCCL prints its log (version banner included) on STDOUT; the bench contract is one JSON line there, so the
            # log goes to stderr.  The level is the caller's (NCCL_DEBUG=INFO shows the communicator's ranks and NVLS).
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def all_reduce_mix(mix_tensor):
    """Sum the per-rank partial mixes in place (the only collective on the path)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(mix_tensor, op=dist.ReduceOp.SUM)
    return mix_tensor


def max_over_ranks(value: float, device=None) -> float:
    """Device-timed durations are reported as the max over ranks."""
    import torch
    import torch.distributed as dist
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
