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


def plan_buckets(spans, final, n_buckets):
    """Cut a flat gradient buffer into <= n_buckets contiguous ranges that become final as EARLY as possible.

    spans[i] = (lo, hi) of parameter i in the flat buffer (ascending, contiguous); final[i] = position in the backward
    pass (larger = later) at which its gradient is complete (-1: never touched).  The backward pass finalises the tail of
    the buffer first, so buckets are grown from the end with ~equal byte counts.  Returns [(lo, hi, when, members)] in
    flush order, `when` = max(final) over the members."""
    n = len(spans)
    assert n == len(final) and n > 0
    total = spans[-1][1] - spans[0][0]
    target = max(1, total // max(1, n_buckets))
    out, members, size = [], [], 0
    for i in range(n - 1, -1, -1):
        members.append(i)
        size += spans[i][1] - spans[i][0]
        if (size >= target and len(out) < n_buckets - 1) or i == 0:
            out.append((spans[members[-1]][0], spans[members[0]][1], max(final[j] for j in members), members[::-1]))
            members, size = [], 0
    # a bucket can never be flushed before one that lies later in the backward pass was due: keep `when` monotone
    when = -1
    fixed = []
    for lo, hi, w, m in out:
        when = max(when, w)
        fixed.append((lo, hi, when, m))
    return fixed
