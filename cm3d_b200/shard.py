"""Frame sharding across GPUs: one process per GPU, frames split by sample index, no collective
on the data path - only a host-side gather of the per-frame labels at the end.

The reference is single-process (src/nuscenes/2d_to_3d.py keeps everything in RAM and writes once
at :929-930); its only cross-frame state is list concatenation (:408-410,662-663), so any
partition of the sample indices gives the same labels once the partial results are merged in
sample order.
"""
from __future__ import annotations

import os
from typing import Callable, Dict, List, Sequence


def world_from_env():
    """(rank, world_size, local_rank) from the torchrun environment (1 process when absent)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init_distributed():
    """(rank, world, local_rank); under torchrun (WORLD_SIZE > 1) also makes sure a process group
    exists for the host-side label gather (gloo: the gather moves pickled label dicts, not tensors)."""
    rank, world, local_rank = world_from_env()
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group("gloo", rank=rank, world_size=world)
    return rank, world, local_rank


def stage_device(cfg_device: str, world: int, local_rank: int) -> str:
    """One process per GPU: under torchrun rank r lifts on cuda:<LOCAL_RANK>, else on the script's DEVICE.
    With fewer visible GPUs than local ranks the ranks share them round-robin (correct, just not
    faster): that is how the sharded path is exercised on a one-GPU box."""
    if world <= 1:
        return cfg_device
    import torch
    n = torch.cuda.device_count()
    return f"cuda:{local_rank % n}" if n else f"cuda:{local_rank}"


def shard_indices(n: int, rank: int, world: int, mode: str = "interleaved") -> List[int]:
    """Sample indices owned by `rank`.  "interleaved": i -> rank i mod G (even load when frame
    cost varies slowly with time); "blocked": contiguous blocks (keeps a scene's frames together)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    if mode == "interleaved":
        return list(range(rank, n, world))
    if mode == "blocked":
        per, rem = divmod(n, world)
        start = rank * per + min(rank, rem)
        return list(range(start, start + per + (1 if rank < rem else 0)))
    raise ValueError(f"unknown shard mode {mode!r}")


def merge_shards(parts: Sequence[Dict[int, object]], n: int) -> List[object]:
    """Union of per-rank {sample index: labels} dicts, returned in sample order; every index in
    [0, n) must be present exactly once."""
    merged: Dict[int, object] = {}
    for p in parts:
        for k, v in p.items():
            if k in merged:
                raise ValueError(f"sample {k} lifted by two ranks")
            merged[k] = v
    missing = [i for i in range(n) if i not in merged]
    if missing:
        raise ValueError(f"samples never lifted: {missing[:8]}{'...' if len(missing) > 8 else ''}")
    return [merged[i] for i in range(n)]


def gather_labels(local: Dict[int, object], n: int, dst: int = 0):
    """Host-side gather of the label dicts to rank `dst` (torch.distributed object gather: works
    on gloo and on nccl process groups).  Returns the merged list on `dst`, None elsewhere."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return merge_shards([local], n)
    world, rank = dist.get_world_size(), dist.get_rank()
    out = [None] * world if rank == dst else None
    dist.gather_object(local, out, dst=dst)
    return merge_shards(out, n) if rank == dst else None


def lift_sharded(n: int, load_frame: Callable[[int], object], lift_batch: Callable[[list], list],
                 batch: int = 64, mode: str = "interleaved", rank: int = None, world: int = None):
    """Run `lift_batch` over this rank's share of samples [0, n) in batches; returns
    {sample index: result}.  `load_frame(i)` builds the FrameSpec of sample i."""
    if rank is None or world is None:
        rank, world, _ = world_from_env()
    mine = shard_indices(n, rank, world, mode)
    out: Dict[int, object] = {}
    for b in range(0, len(mine), batch):
        idx = mine[b:b + batch]
        res = lift_batch([load_frame(i) for i in idx])
        out.update(zip(idx, res))
    return out
