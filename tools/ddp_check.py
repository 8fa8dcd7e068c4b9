"""2+ GPU check of the bucketed, backward-overlapped gradient all-reduce (TrainStep, world > 1):
   torchrun --nproc-per-node 2 tools/ddp_check.py
Every rank runs forward + backward on ITS OWN shard from the same weights twice: with n_buckets=4 (buckets folded and
all-reduced while the backward pass is still running) and with n_buckets=1 (one all-reduce of the flat buffer after
backward), on the well-conditioned fixture of tests/test_gpu_backward_model.py (at the reference's init two backward
runs over the SAME activations already differ by percents: order-dependent fp32 atomics amplified by the chaotic
stack).  Checks: the averaged gradients agree to the run-to-run noise of the backward pass; every rank holds the SAME
averaged gradient bit for bit (a bucket that missed its all-reduce would differ, the shards are different); then 3
optimizer steps leave identical replicas."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import ryolo_b200 as R
from tests.util import CFG, HYP, det_init, make_targets
from tests.test_gpu_backward_model import _calm

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
res = {}
for ver, nc in (("yolov4", 2), ("yolov7", 16)):
    img = torch.rand(4, 3, 256, 256, generator=torch.Generator().manual_seed(10 + rank)).to(dev)
    tg = make_targets(20 + rank, 4, 10, nc, True).to(dev)
    grads, flats = [], []
    for nb in (4, 1, 1):
        m = _calm(det_init(R.Yolo(nc, CFG, "csl", ver))).to(dev).train()
        crit = R.ComputeCSLLoss(m, HYP)
        crit.sync_items = False
        step = R.TrainStep(m, crit, lr=0.01, n_buckets=nb)
        step.zero_grad()
        step.forward_backward(img, tg)
        if not step._reduced:
            from ryolo_b200.dist import allreduce_mean
            allreduce_mean(step.grad, step.pg)
        torch.cuda.synchronize()
        grads.append(step.grad.clone())
        step._reduced = True
        step.step()
        for _ in range(2):
            step(img, tg)
        torch.cuda.synchronize()
        flats.append(step.flat.clone())
        if nb == 4:
            nbk = sum(len(v) for v in step._plan.values())
            sizes = sorted((b["hi"] - b["lo"]) * 4 / 1e6 for v in step._plan.values() for b in v)
            when = sorted(step._plan.keys())
    rel = float((grads[0] - grads[1]).norm() / grads[1].norm())
    noise = float((grads[2] - grads[1]).norm() / grads[1].norm())      # two un-bucketed runs: the backward's own noise
    g = [torch.empty_like(grads[0]) for _ in range(dist.get_world_size())]
    dist.all_gather(g, grads[0])
    same_grad = all(torch.equal(g[0], x) for x in g)
    dist.all_gather(g, flats[0])
    same = all(torch.equal(g[0], x) for x in g)
    if rank == 0:
        print(f"{ver}: buckets {nbk} (MB {['%.1f' % s for s in sizes]}) flushed after reverse-tape entries {when}; "
              f"averaged gradient, bucketed vs single all-reduce: rel-L2 {rel:.3e} (two single runs: {noise:.3e}); "
              f"identical on all ranks: {same_grad}; replicas identical after 3 steps: {same}")
        assert rel < max(5 * noise, 1e-3) and same_grad and same
dist.destroy_process_group()
if rank == 0:
    print("ddp_check ok")
