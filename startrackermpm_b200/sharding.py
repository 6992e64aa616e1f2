"""Multi-GPU sharding of the trial loop: one process per GPU, no data-path collective.

Trial i depends only on (seed, global trial index i) and the replicated catalog, so ranks take contiguous index ranges
(or whole segments of a sweep) and the ONLY communication is one all-reduce of the packed statistics buffer at the end:
every entry is a sum except slot 5 (x_max), which is max-reduced.  Works with the `nccl` backend on CUDA tensors and
with `gloo` on CPU tensors (tests)."""
from __future__ import annotations

X_MAX_SLOT = 5


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [begin, end) of n items for `rank` of `world`; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(int(n), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_segments(n_segments: int, rank: int, world: int) -> range:
    """Round-robin whole segments (grid points / orbit steps): per-segment statistics stay rank-local.  Use this when the
    cost per segment varies systematically along the index (orbit steps); for uniform sweeps prefer a contiguous block
    `range(*shard_range(n_segments, rank, world))`, which `studies.run_segments` batches into few launches."""
    return range(rank, int(n_segments), int(world))


def allreduce_stats(stats, group=None):
    """In-place merge of packed statistics [..., stats_len] across ranks: one sum all-reduce (+ a max for x_max)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return stats
    xmax = stats[..., X_MAX_SLOT].clone()
    dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(xmax, op=dist.ReduceOp.MAX, group=group)
    stats[..., X_MAX_SLOT] = xmax
    return stats
