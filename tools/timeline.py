"""Absolute start/end times (CUDA events, one base event) of the conv / dgrad / wgrad / BatchNorm-backward launches of
one yolov4 800x800 train step, across the high-priority backward stream and the wgrad side stream: shows whether the
weight-gradient GEMMs run under the HBM-bound passes or time-slice with them.  Event pairs around every launch break
programmatic dependent launch, so absolute numbers are a little slower than the un-instrumented step.

    python tools/timeline.py [batch] [first_row] [rows]"""
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import ryolo_b200 as R  # noqa: E402
from ryolo_b200 import ops  # noqa: E402
from bench import CFG, HYP, S, make_targets, weights_init_normal  # noqa: E402

bs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
nrows = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 6
torch.manual_seed(42)
m = R.Yolo(2, CFG, "csl", "yolov4")
m.apply(weights_init_normal)
m = m.cuda().train()
crit = R.ComputeCSLLoss(m, HYP)
crit.sync_items = False
step = R.TrainStep(m, crit)
img = torch.rand(bs, 3, S, S, device="cuda")
tg = make_targets(0, bs, 2).cuda()
for _ in range(2):
    step(img, tg)
torch.cuda.synchronize()
base = torch.cuda.Event(enable_timing=True)
ops.PROFILE = []
base.record()
step(img, tg)
torch.cuda.synchronize()
prof, ops.PROFILE = ops.PROFILE, None
rows = sorted(((base.elapsed_time(a), base.elapsed_time(b), tag) for tag, a, b in prof), key=lambda r: r[0])
bw = [r for r in rows if r[2][0] != "conv"]
t0 = bw[0][0]
busy = {}
for s, e, tag in bw:
    busy[tag[0]] = busy.get(tag[0], 0.0) + (e - s)
span = max(e for s, e, _ in bw) - t0
print(f"backward span {span:.2f} ms; summed durations {busy}")
# overlap accounting: time covered by >= 1 wgrad interval AND >= 1 bn_bwd interval
ev = []
for s, e, tag in bw:
    if tag[0] in ("wgrad", "bn_bwd"):
        ev += [(s, tag[0], 1), (e, tag[0], -1)]
ev.sort()
cnt = {"wgrad": 0, "bn_bwd": 0}
both = 0.0
last = None
for tt, k, d in ev:
    if last is not None and cnt["wgrad"] > 0 and cnt["bn_bwd"] > 0:
        both += tt - last
    cnt[k] += d
    last = tt
print(f"time with a wgrad AND a bn_bwd interval open: {both:.2f} ms")
for s, e, tag in bw[first:first + nrows]:
    print(f"{s - t0:8.3f} {e - t0:8.3f} {e - s:7.3f}  {tag[0]:6s} M={tag[1]:9d} N={tag[2]:5d} K={tag[3]:5d}")
