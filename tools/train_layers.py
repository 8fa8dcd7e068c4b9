"""Per-launch timing of the tensor-core kernels (conv fwd, dgrad, wgrad) inside one yolov4 800x800 bs=32 train step."""
import json, os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
import ryolo_b200 as R
from ryolo_b200 import ops
from bench import CFG, HYP, S, make_targets, weights_init_normal

bs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ver = sys.argv[2] if len(sys.argv) > 2 else "yolov4"
nc = 2 if ver == "yolov4" else 16
torch.manual_seed(42)
m = R.Yolo(nc, CFG, "csl", ver)
m.apply(weights_init_normal)
m = m.cuda().train()
crit = R.ComputeCSLLoss(m, HYP)
crit.sync_items = False
step = R.TrainStep(m, crit)
img = torch.rand(bs, 3, S, S, device="cuda")
tg = make_targets(0, bs, nc).cuda()
for _ in range(2):
    step(img, tg)
torch.cuda.synchronize()
ops.PROFILE = []
step(img, tg)
torch.cuda.synchronize()
prof, ops.PROFILE = ops.PROFILE, None
tot = {}
for tag, a, b in prof:
    ms = a.elapsed_time(b)
    kind, M, N, K = tag
    tot[kind] = tot.get(kind, 0) + ms
    print(f"{kind:6s} M={M:9d} N={N:5d} K={K:5d} {ms:7.3f} ms {2.0*M*N*K/ms/1e9:7.1f} TF/s")
print(tot)
