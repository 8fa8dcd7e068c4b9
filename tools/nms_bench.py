"""BASELINE config 5: post_process on 64 images x 100k candidate rows (clustered boxes), boxes/s."""
import json, os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import ryolo_b200 as R

def make(B, Rr, nc, gen):
    centres = torch.rand(B, 200, 2, device="cuda", generator=gen) * 800
    pick = torch.randint(0, 200, (B, Rr), device="cuda", generator=gen)
    xy = torch.gather(centres, 1, pick[..., None].expand(B, Rr, 2)) + torch.randn(B, Rr, 2, device="cuda", generator=gen) * 6
    w = torch.rand(B, Rr, 1, device="cuda", generator=gen) * 116 + 4
    h = w * (1 + 3 * torch.rand(B, Rr, 1, device="cuda", generator=gen))
    th = (torch.rand(B, Rr, 1, device="cuda", generator=gen) - 0.5) * np.pi * 0.9999
    oc = torch.rand(B, Rr, 1 + nc, device="cuda", generator=gen)
    return torch.cat((xy, w, h, th, oc), 2).contiguous()

if __name__ == "__main__":
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    gen = torch.Generator(device="cuda").manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = {}
    for nc, conf, iou in ((2, 0.001, 0.65), (2, 0.7, 0.2), (16, 0.001, 0.65)):
        pred = make(64, 100000, nc, gen)
        for _ in range(2):
            R.post_process_device(pred, conf, iou, mutate=False)
        ts = []
        for _ in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); a.record()
            d, r, n = R.post_process_device(pred, conf, iou, mutate=False)
            b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = float(np.median(ts))
        out[f"nc{nc}_conf{conf}_iou{iou}"] = dict(ms=ms, boxes_per_s=64 * 100000 / ms * 1e3, survivors=float(n.float().mean()))
        del pred
    # ---- BASELINE config 4: KFIoU loss on 50k pairs/image x 256 images (563 MB of pairs: larger than L2)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    N = 50000 * 256
    pr = torch.cat((torch.rand(N, 2, device="cuda") * 2 - 0.5, torch.rand(N, 2, device="cuda") * 8 + 0.5,
                    (torch.rand(N, 1, device="cuda") - 0.5) * 3.14), 1).contiguous()
    tg = torch.cat((torch.rand(N, 2, device="cuda"), torch.rand(N, 2, device="cuda") * 8 + 0.5,
                    (torch.rand(N, 1, device="cuda") - 0.5) * 3.14), 1).contiguous()
    kf = R.KFLoss()
    for grad, bpp in ((False, 44), (True, 64)):
        p = pr.clone().requires_grad_(grad)
        ts = []
        for i in range(iters + 2):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); a.record()
            for _ in range(10):                      # back to back: the 563 MB of pairs never fit the 126 MB L2
                loss, k = kf(p, tg)
            b.record(); torch.cuda.synchronize()
            if i >= 2:
                ts.append(a.elapsed_time(b) / 10)
        ms = float(np.median(ts))
        out["kfloss_fwd_bwd" if grad else "kfloss_fwd"] = dict(ms=ms, pairs=N, gbs=N * bpp / ms / 1e6,
                                                                frac_hbm=N * bpp / ms / 1e6 / peaks["hbm_gbs"])
    print(json.dumps(out, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "nms_bench.json"), "w"), indent=1)
