"""Batch-axis sharding of stereo sequences over the GPUs of one box.

The hot path has no cross-sample reduction (the only batch statistic, the splat metric's mean of the
previous disparity, is taken over the LOCAL batch by the reference as well —
projects/TemporalStereo/TemporalStereo.py:364, 380, 418), so sequences are independent units: each rank
owns a contiguous slice of the batch and its own recurrent `prev_info`; there is no data-path collective.
`torch.distributed` is used for the rendezvous, barriers and the max-over-ranks reduction of the timing only.
The engine is inference-only, so the training collective of SURVEY.md §8e (one flat gradient all-reduce) has no
producer here and is not shipped.
The time axis is a sequential recurrence and is never sharded.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple

import torch


def env_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1 process when absent)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def bind_to_gpu_numa(local_rank: int) -> Optional[List[int]]:
    """Pin this process to the CPU cores NVML reports as local to GPU `local_rank` (its NUMA node), so that pinned host
    staging buffers allocated afterwards are first-touched next to the GPU's PCIe root port.  With one process per GPU and
    eight GPUs on two sockets, un-pinned ranks stage half of their uploads across the inter-socket link.  Returns the core
    list, or None when NVML / the affinity call is unavailable (the bench reports which)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = int(visible.split(",")[local_rank]) if visible and visible.split(",")[0].isdigit() else local_rank
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [w * 64 + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:  # noqa: BLE001  (no NVML, container without the call, ...)
        return None


def shard_range(total: int, world: int, rank: int) -> Tuple[int, int]:
    """[start, stop) of the `total` sequences owned by `rank`: contiguous, sizes differ by at most one,
    the first `total % world` ranks take the extra one (an empty range when total < world)."""
    if world < 1 or not 0 <= rank < world or total < 0:
        raise ValueError(f"bad shard request total={total} world={world} rank={rank}")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_batch(tensors: Sequence[torch.Tensor], world: int, rank: int) -> List[torch.Tensor]:
    """This rank's slice (dim 0) of every tensor of a batch; all tensors must share the batch size."""
    sizes = {int(t.shape[0]) for t in tensors}
    if len(sizes) != 1:
        raise ValueError(f"tensors disagree on the batch size: {sorted(sizes)}")
    a, b = shard_range(sizes.pop(), world, rank)
    return [t[a:b] for t in tensors]


def barrier(dist=None) -> None:
    """Process-group barrier followed by a device synchronize (both sides of every timed region)."""
    if dist is not None and dist.is_initialized():
        dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def max_over_ranks(value: float, dist=None, device="cpu") -> float:
    """Slowest rank's value: a multi-GPU step is as slow as its slowest replica."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, dist=None, device="cpu") -> float:
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_throughput(units_this_rank: int, ms_this_rank: float, dist=None, device="cpu") -> Tuple[float, float, int]:
    """Whole-job throughput: (units of all ranks) / (max time over ranks).  Returns (units/s, ms, units)."""
    ms = max_over_ranks(ms_this_rank, dist, device)
    units = int(round(sum_over_ranks(units_this_rank, dist, device)))
    return units / (ms * 1e-3), ms, units
