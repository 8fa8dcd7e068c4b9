"""One launch of every non-conv hot-path kernel at its BASELINE size, for `ncu --set full` (tools/gpu_job_r2d.sh):
decode (csl / kfiou heads of 800x800 bs=32), fused CSL / KFIoU loss value+gradient, standalone KFLoss on 12.8 M pairs
(forward, forward+backward), post_process on 64 x 100 000 rows, pairwise skew IoU.  A warm-up call of each precedes the
profiled one (ncu is pointed at the LAST launches by -s/-c, see the job script)."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import ryolo_b200 as R
from bench import CFG, HYP, make_pred, make_targets

dev = torch.device("cuda")
torch.manual_seed(0)
gen = torch.Generator(device=dev).manual_seed(0)


class M:
    def __init__(self, anchors, nc):
        self.anchors, self.nc = anchors, nc
        self._p = torch.nn.Parameter(torch.zeros(1, device=dev))

    def parameters(self):
        return iter([self._p])


def run(reps):
    for _ in range(reps):
        for mode, nc in (("csl", 2), ("kfiou", 2)):
            m = R.Yolo(nc, CFG, mode, "yolov4")
            na, ch = m.na, m.ch
            levels = [torch.randn(32, na, 800 // s, 800 // s, ch, device=dev) for s in (8, 16, 32)]
            with torch.no_grad():
                m.yolo([l.clone() for l in levels], False)                       # decode_csl / decode_kfiou
            crit = (R.ComputeCSLLoss if mode == "csl" else R.ComputeKFIoULoss)(M(m.anchors, nc), HYP)
            crit.sync_items = False
            tg = make_targets(0, 32, nc).to(dev)
            crit.value_and_grad(levels, tg if mode == "csl" else tg[:, :7].contiguous())
            del levels
        N = 50000 * 256
        pr = torch.cat((torch.rand(N, 2, device=dev) * 2 - 0.5, torch.rand(N, 2, device=dev) * 8 + 0.5,
                        (torch.rand(N, 1, device=dev) - 0.5) * 3.14), 1).contiguous()
        tgp = torch.cat((torch.rand(N, 2, device=dev), torch.rand(N, 2, device=dev) * 8 + 0.5,
                         (torch.rand(N, 1, device=dev) - 0.5) * 3.14), 1).contiguous()
        kf = R.KFLoss()
        kf(pr, tgp)
        kf(pr.clone().requires_grad_(True), tgp)
        del pr, tgp
        pred = make_pred(64, 100000, 2, gen, dev)
        R.post_process_device(pred, 0.001, 0.65, mutate=False)
        a = pred[0, :1500, :5].clone()
        a[:, 4] *= 180 / np.pi
        R.pairwise_iou_rotated(a, a[:200].contiguous())
        del pred
        torch.cuda.synchronize()


run(int(sys.argv[1]) if len(sys.argv) > 1 else 2)
print("nonconv_profile done")
