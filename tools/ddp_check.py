"""2+ GPU check of the bucketed, backward-overlapped gradient all-reduce (TrainStep, world > 1):
   torchrun --nproc-per-node 2 tools/ddp_check.py
Every rank trains the same initial weights on ITS OWN shard for 3 steps twice: with n_buckets=4 (overlapped buckets)
and with n_buckets=1 (one all-reduce of the flat buffer after backward).  The updated weights must agree (NCCL may sum
slices in another order than the whole buffer: 1e-5 relative), and all ranks must hold identical weights."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import ryolo_b200 as R
from tests.util import CFG, HYP, det_init, make_targets

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
res = {}
for ver, nc in (("yolov4", 2), ("yolov7", 16)):
    img = torch.rand(4, 3, 256, 256, generator=torch.Generator().manual_seed(10 + rank)).to(dev)
    tg = make_targets(20 + rank, 4, 10, nc, True).to(dev)
    flats = []
    for nb in (4, 1):
        m = det_init(R.Yolo(nc, CFG, "csl", ver)).to(dev).train()
        crit = R.ComputeCSLLoss(m, HYP)
        crit.sync_items = False
        step = R.TrainStep(m, crit, lr=0.01, n_buckets=nb)
        for _ in range(3):
            step(img, tg)
        torch.cuda.synchronize()
        flats.append(step.flat.clone())
        if nb == 4:
            nbk = sum(len(v) for v in step._plan.values())
            sizes = sorted((b["hi"] - b["lo"]) * 4 / 1e6 for v in step._plan.values() for b in v)
            when = sorted(step._plan.keys())
    a, b = flats
    rel = float((a - b).norm() / b.norm())
    mx = float((a - b).abs().max())
    # replicas identical across ranks
    g = [torch.empty_like(a) for _ in range(dist.get_world_size())]
    dist.all_gather(g, a)
    same = all(torch.equal(g[0], x) for x in g)
    if rank == 0:
        print(f"{ver}: buckets {nbk} (MB {['%.1f' % s for s in sizes]}) flushed after reverse-tape entries {when}; "
              f"bucketed vs single all-reduce: rel-L2 {rel:.3e}, max abs {mx:.3e}; replicas identical: {same}")
        assert rel < 1e-4 and same
dist.destroy_process_group()
if rank == 0:
    print("ddp_check ok")
