"""Data-parallel plumbing for the decode path: image batches shard
trivially across ranks, the data path needs NO collective; torch.distributed is used only to
(a) agree on the shards, (b) optionally gather the per-rank [b_i, K, D] detections on rank 0
(what the reference's DataParallel `gather` does, models/data_parallel.py:70-84) and
(c) reduce timings (max over ranks).  Works with any backend (nccl on GPUs, gloo in CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(n_images, rank, world, chunk_sizes=None):
    """Contiguous slice [lo, hi) of the batch owned by `rank`.  `chunk_sizes` reproduces the
    reference's uneven split (`opt.chunk_sizes`, opts.py:260-269); default is an even split with
    the remainder spread over the first ranks."""
    if chunk_sizes is not None:
        if len(chunk_sizes) != world or sum(chunk_sizes) != n_images:
            raise ValueError("chunk_sizes must have one entry per rank and sum to the batch size")
        lo = sum(chunk_sizes[:rank])
        return lo, lo + chunk_sizes[rank]
    base, rem = divmod(n_images, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_detections(local_dets, n_images, chunk_sizes=None):
    """All ranks pass their [b_i, K, D] detections; returns the [n_images, K, D] batch on every rank
    (shards may be uneven, so they are padded to the largest shard for the collective)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_dets
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(n_images, r, world, chunk_sizes) for r in range(world)]
    biggest = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((biggest,) + tuple(local_dets.shape[1:]), dtype=local_dets.dtype, device=local_dets.device)
    pad[: local_dets.shape[0]] = local_dets
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)


def max_over_ranks(value, device="cpu"):
    """Max of a python float over all ranks (timing reduction of bench.py)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
