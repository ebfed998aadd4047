"""Multi-GPU sharding of the path: receipts are independent units, so ranks take disjoint slices and nothing is
exchanged on the data path (reference: one image at a time, ocr/pipeline.py:65-87).  torch.distributed is
used only for the start/stop barrier, the max-over-ranks timing and gathering result lists to rank 0."""
import torch.distributed as dist


def shard_indices(n_items, rank, world):
    """Round-robin assignment: item i belongs to rank i % world.  Returns this rank's item indices in order."""
    return list(range(rank, n_items, world))


def gather_in_order(local_results, n_items, rank, world):
    """All ranks contribute [(item index, result), ...]; rank 0 returns the results ordered by item index."""
    if world == 1 or not dist.is_initialized():
        parts = [local_results]
    else:
        parts = [None] * world
        dist.all_gather_object(parts, local_results)
    if rank != 0:
        return None
    out = [None] * n_items
    for part in parts:
        for i, r in part:
            out[i] = r
    return out


def max_over_ranks(value, device=None):
    """Max of a python float over all ranks (timing is the slowest rank's)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
