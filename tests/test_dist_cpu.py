"""world_size-2 gloo tests (CPU) of the data-parallel host logic: batch sharding and gradient averaging."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.util import ROOT, make_targets


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ryolo_b200.dist import allreduce_mean, shard_batch
    imgs = torch.arange(8 * 3 * 4 * 4, dtype=torch.float32).view(8, 3, 4, 4)
    tg = make_targets(0, 8, 5, 2, False)
    si, st = shard_batch(imgs, tg, rank, world)
    assert si.shape[0] == 4 and torch.equal(si, imgs[rank * 4:(rank + 1) * 4])
    assert st.shape[0] == 20 and st[:, 0].min() == 0 and st[:, 0].max() == 3
    assert torch.equal(st[:, 1:], tg[(tg[:, 0] >= rank * 4) & (tg[:, 0] < rank * 4 + 4)][:, 1:])
    flat = torch.full((1000,), float(rank + 1))
    allreduce_mean(flat)
    assert torch.allclose(flat, torch.full((1000,), 1.5))
    # rank-local "gradients" averaged == gradient of the mean of per-shard losses
    w = torch.ones(3, requires_grad=True)
    loss = (w * (rank + 1.0)).pow(2).sum()
    loss.backward()
    g = w.grad.clone()
    allreduce_mean(g)
    assert torch.allclose(g, torch.full((3,), (2.0 * 1 + 2.0 * 4) / 2))
    dist.destroy_process_group()
    if rank == 0:
        open(out, "w").write("ok")


def test_gloo_world2(tmp_path):
    out = str(tmp_path / "ok")
    mp.spawn(_worker, args=(2, 29533, out), nprocs=2, join=True)
    assert open(out).read() == "ok"


def test_plan_buckets_covers_the_buffer_and_orders_by_readiness():
    """Host logic of the bucketed all-reduce (train_step.py): contiguous cover of the flat buffer, ~equal bytes, each
    bucket due when its LAST member is final, flush order = backward order."""
    import random
    from ryolo_b200.dist import plan_buckets
    rnd = random.Random(0)
    for n, nb in ((1, 4), (5, 3), (327, 4), (327, 1), (40, 8)):
        sizes = [rnd.choice((8, 64, 4096, 1 << 16)) for _ in range(n)]
        spans, off = [], 0
        for s in sizes:
            spans.append((off, off + s))
            off += s
        final = [n - 1 - i + rnd.choice((-2, 0, 3)) for i in range(n)]          # roughly reverse, locally permuted
        b = plan_buckets(spans, final, nb)
        assert 1 <= len(b) <= nb
        assert b[0][1] == off and b[-1][0] == 0
        for (lo, hi, when, mem), nxt in zip(b, b[1:] + [None]):
            assert lo == spans[mem[0]][0] and hi == spans[mem[-1]][1]
            assert when >= max(final[j] for j in mem)
            if nxt is not None:
                assert nxt[1] == lo and nxt[2] >= when
        assert sorted(j for x in b for j in x[3]) == list(range(n))
