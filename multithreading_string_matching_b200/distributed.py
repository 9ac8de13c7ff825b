"""Multi-GPU plumbing: one process per GPU (torch.distributed), packets split over ranks exactly as
the reference's MPI variant splits them (mpi_dumping.c:149-157), per-pattern count vectors summed with
one all-reduce (the MPI_Reduce(SUM) of mpi_dumping.c:202 -- NCCL over NVLink on GPUs, gloo in CPU tests).
There is no data-path collective: every rank reads (or generates) only its own contiguous slice.
Device-resident callers can skip the collective altogether: Matcher.count_device_into adds a rank's counts
to every rank's vector (symmetric memory) from inside the match kernel -- see bench.py."""
import torch
import torch.distributed as dist

from .matcher import shard_range


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def rank_slice(total_packets, rank=None, world_size=None):
    """(first, count) of this rank's contiguous packet slice: N/P each, rank 0 also takes N%P."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    return shard_range(total_packets, world_size, rank)


def reduce_counts(counts):
    """Sum a per-pattern count vector (torch int64 tensor, on the device the backend needs) over ranks,
    in place; every rank gets the total (all-reduce rather than reduce-to-0: 8 bytes per pattern)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts


def sharded_count(count_slice, total_packets, n_patterns, device="cpu"):
    """count_slice(first, count) -> sequence of n_patterns per-pattern counts for that packet slice.
    Returns the global counts as a list on every rank."""
    first, count = rank_slice(total_packets)
    local = torch.as_tensor([int(c) for c in count_slice(first, count)], dtype=torch.int64, device=device)
    assert local.numel() == n_patterns
    return reduce_counts(local).cpu().tolist()
