"""Diagnostic: product Yolo (GPU) vs oracle conv stack (CPU fp32) at a better-conditioned size."""
import sys, os, json
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
import ryolo_b200 as R
from oracle import model_cpu
from tests.util import CFG, det_init

def cmp(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten()
    return float(torch.dot(a, b) / (a.norm() * b.norm())), float((a - b).norm() / b.norm())

for ver, mode, nc in (("yolov4", "csl", 2), ("yolov7", "csl", 16)):
    for S, bs in ((64, 2), (256, 4)):
        m = det_init(R.Yolo(nc, CFG, mode, ver))
        sd = {k: v.clone() for k, v in m.state_dict().items()}
        img = torch.rand(bs, 3, S, S, generator=torch.Generator().manual_seed(1))
        ref, _, _ = model_cpu.forward(sd, img, ver, mode, nc, train=True)
        m = m.cuda().train()
        out = m(img.cuda(), training=True)
        print(ver, S, bs, "train", [cmp(a, b) for a, b in zip(out, ref)], flush=True)
