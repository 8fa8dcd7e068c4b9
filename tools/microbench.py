"""Micro-benchmarks of the non-conv kernels (decode / loss / post_process) on one B200.
Writes gpurun_out/microbench.json.  Timing: CUDA events on the current stream, L2 flushed between
iterations (256 MiB memset)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import ryolo_b200 as R  # noqa: E402
from tests.util import CFG, HYP, make_targets  # noqa: E402
from oracle import hotpath as hp  # noqa: E402  (anchor tables only)

PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


class _M:
    def __init__(self, anchors, nc):
        self.anchors, self.nc = anchors, nc
        self._p = torch.nn.Parameter(torch.zeros(1, device="cuda"))

    def parameters(self):
        return iter([self._p])


def main():
    out = {}
    gen = torch.Generator(device="cuda").manual_seed(0)
    # ---- post_process, BASELINE config 5: 64 x 100k rows
    for nc, conf, iou in ((2, 0.001, 0.65), (2, 0.7, 0.2), (16, 0.001, 0.65)):
        B, Rr = 64, 100000
        centres = torch.rand(B, 200, 2, device="cuda", generator=gen) * 800
        pick = torch.randint(0, 200, (B, Rr), device="cuda", generator=gen)
        xy = torch.gather(centres, 1, pick[..., None].expand(B, Rr, 2)) + torch.randn(B, Rr, 2, device="cuda", generator=gen) * 6
        w = torch.rand(B, Rr, 1, device="cuda", generator=gen) * 116 + 4
        h = w * (1 + 3 * torch.rand(B, Rr, 1, device="cuda", generator=gen))
        th = (torch.rand(B, Rr, 1, device="cuda", generator=gen) - 0.5) * np.pi * 0.9999
        oc = torch.rand(B, Rr, 1 + nc, device="cuda", generator=gen)
        pred = torch.cat((xy, w, h, th, oc), 2).contiguous()
        ms = timeit(lambda: R.post_process_device(pred, conf, iou, mutate=False), iters=5, warm=2)
        _, _, n = R.post_process_device(pred, conf, iou, mutate=False)
        out[f"post_process_nc{nc}_conf{conf}_iou{iou}"] = dict(
            ms=ms, boxes_per_s=B * Rr / ms * 1e3, front_end_bytes=B * Rr * (6 + nc) * 4,
            mean_survivors=float(n.float().mean()))
        del pred
    # ---- loss fwd+bwd at 800^2 bs=32
    for mode, nc in (("csl", 2), ("kfiou", 2), ("csl", 16)):
        csl = mode == "csl"
        na, ch = (3, nc + 185) if csl else (18, nc + 6)
        levels = [torch.randn(32, na, 800 // s, 800 // s, ch, device="cuda", generator=gen).requires_grad_(True)
                  for s in (8, 16, 32)]
        targets = make_targets(1, 32, 100, nc, csl).cuda()
        anchors = hp.make_anchors(CFG["anchors"]) if csl else hp.make_rotated_anchors(CFG["anchors"], CFG["angles"])
        fn = (R.ComputeCSLLoss if csl else R.ComputeKFIoULoss)(_M(anchors, nc), HYP)
        fn.sync_items = False

        def step():
            loss, _ = fn(levels, targets)
            return loss
        ms = timeit(step, iters=10)
        nbytes = sum(l.numel() for l in levels) * 4
        out[f"loss_fwd_bwd_{mode}_nc{nc}_800_bs32"] = dict(ms=ms, head_bytes=nbytes,
                                                          gbs_head_write=nbytes / ms / 1e6)
        with torch.no_grad():
            lv = [l.detach() for l in levels]
            ms = timeit(lambda: fn(lv, targets), iters=10)
        out[f"loss_fwd_{mode}_nc{nc}_800_bs32"] = dict(ms=ms)
        # ---- decode
        layer = (R.YoloCSLLayer(nc, anchors, [8, 16, 32]) if csl else R.YoloKFIoULayer(nc, anchors, [8, 16, 32]))
        ms = timeit(lambda: layer(list(lv), training=False), iters=10)
        outb = 32 * sum(na * (800 // s) ** 2 for s in (8, 16, 32)) * (nc + 6) * 4
        out[f"decode_{mode}_nc{nc}_800_bs32"] = dict(ms=ms, gbs=(nbytes + outb) / ms / 1e6,
                                                    frac_hbm=(nbytes + outb) / ms / 1e6 / PEAKS["hbm_gbs"])
        del levels, lv
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "microbench.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
