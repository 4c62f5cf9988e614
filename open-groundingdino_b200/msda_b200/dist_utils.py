"""Multi-GPU plumbing for the op.  Multi-scale deformable attention has no cross-sample dependency
(every output row (b, q, :) reads image b only; ms_deform_im2col_cuda.cuh:263,269), so it shards by batch
with NO data-path collective: each rank runs the kernels on its own images ("replicas only").  The
helpers below are the whole host side of that: a deterministic batch partition, and the max/sum
reductions used to time a multi-rank run on the device (max over ranks) and to aggregate throughput.
They work on any torch.distributed backend (nccl on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world_size: int, rank: int) -> tuple[int, int]:
    """[lo, hi) of the batch elements rank `rank` owns: contiguous, sizes differ by at most one, and the
    union over ranks is exactly range(n_items)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _reduce_scalar(x: float, op, device) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=op)
    return float(t.item())


def max_over_ranks(x: float, device="cpu") -> float:
    return _reduce_scalar(x, dist.ReduceOp.MAX, device)


def sum_over_ranks(x: float, device="cpu") -> float:
    return _reduce_scalar(x, dist.ReduceOp.SUM, device)


def aggregate_throughput(units_this_rank: float, seconds_this_rank: float, device="cpu") -> float:
    """Whole-job throughput: all ranks' units divided by the slowest rank's time."""
    return sum_over_ranks(units_this_rank, device) / max_over_ranks(seconds_this_rank, device)
