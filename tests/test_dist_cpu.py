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
