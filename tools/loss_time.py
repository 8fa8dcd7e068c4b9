"""Times ComputeCSLLoss / ComputeKFIoULoss forward + gradient at 800x800, 32 images, 100 targets per image (CUDA events,
L2 flushed between iterations).  Prints ms per call."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
import ryolo_b200 as R
from tests.util import CFG, HYP, make_targets
from oracle import hotpath as hp      # anchor tables only
from tools.microbench import _M as MM


gen = torch.Generator(device="cuda").manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for mode, nc in (("csl", 2), ("kfiou", 2), ("csl", 16)):
    csl = mode == "csl"
    na, ch = (3, nc + 185) if csl else (18, nc + 6)
    levels = [torch.randn(32, na, 800 // s, 800 // s, ch, device="cuda", generator=gen).requires_grad_(True) for s in (8, 16, 32)]
    targets = make_targets(1, 32, 100, nc, csl).cuda()
    anchors = hp.make_anchors(CFG["anchors"]) if csl else hp.make_rotated_anchors(CFG["anchors"], CFG["angles"])
    fn = (R.ComputeCSLLoss if csl else R.ComputeKFIoULoss)(MM(anchors, nc), HYP)
    fn.sync_items = False
    ts = []
    for it in range(13):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn(levels, targets)
        b.record()
        torch.cuda.synchronize()
        if it >= 3:
            ts.append(a.elapsed_time(b))
    print(f"loss fwd+grad {mode} nc={nc}: {sorted(ts)[len(ts) // 2]:.3f} ms")
