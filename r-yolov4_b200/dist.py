"""Data-parallel plumbing: shard a batch across ranks and average the flat gradient
buffer with ONE all-reduce.  torch.distributed (NCCL over NVLink on the GPUs, gloo in CPU tests) carries it."""
import torch
import torch.distributed as dist


def shard_batch(imgs, targets, rank, world):
    """Contiguous image shard of a global batch + its target rows with the image index re-based
    (targets[:, 0] is the image index inside the batch, datasets/base_dataset.py:161-167)."""
    B = imgs.shape[0]
    assert B % world == 0, "global batch must divide evenly across ranks"
    per = B // world
    lo, hi = rank * per, (rank + 1) * per
    sel = (targets[:, 0] >= lo) & (targets[:, 0] < hi)
    t = targets[sel].clone()
    t[:, 0] -= lo
    return imgs[lo:hi], t


def allreduce_mean(flat, group=None):
    """In-place mean over ranks of a flat buffer.  NCCL has a native AVG; gloo gets SUM then divide."""
    if not (dist.is_available() and dist.is_initialized()):
        return flat
    world = dist.get_world_size(group)
    if world == 1:
        return flat
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
    return flat
