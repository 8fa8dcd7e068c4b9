"""Diagnose a GPU-vs-oracle NMS disagreement on the adversarial sets: first differing survivor, the pair that
decided it, and what each implementation says about that pair."""
import ctypes, os, subprocess, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import ryolo_b200 as R
from oracle import rotated as orot
from tests.test_rotated_iou_host import THRESHOLDS, adversarial_pairs, run_pairs

so = "/tmp/libhr_dbg.so"
subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-I",
                os.path.join(ROOT, "r-yolov4_b200", "csrc"), os.path.join(ROOT, "tests", "host", "rotated_iou_host.cpp"),
                "-o", so], check=True)
hr = ctypes.CDLL(so)
np.set_printoptions(precision=9, suppress=False, linewidth=200)
for kind in ("tiny", "any", "edge", "corner", "normal"):
    rng = np.random.default_rng(len(kind))
    A, B = adversarial_pairs(rng, 1200, kind)
    boxes = torch.from_numpy(np.concatenate([A, B], 0))
    perm = torch.from_numpy(rng.permutation(boxes.shape[0]))
    boxes = boxes[perm].contiguous()
    scores = torch.from_numpy(rng.random(boxes.shape[0]).astype(np.float32))
    scores[::11] = scores[0]
    n = boxes.shape[0]
    gpu_m = R.pairwise_iou_rotated(boxes.cuda(), boxes.cuda()).cpu()
    ref_m = orot.pairwise_iou_rotated(boxes, boxes)
    neq = (gpu_m.view(torch.int32) != ref_m.view(torch.int32))
    print(f"== {kind}: pairwise matrix mismatching entries {int(neq.sum())} of {n * n}")
    if neq.any():
        ij = neq.nonzero()[:5]
        for i, j in ij.tolist():
            print("   pair", i, j, boxes[i].numpy(), boxes[j].numpy(), "gpu", float(gpu_m[i, j]), "oracle", float(ref_m[i, j]))
    for thr in THRESHOLDS:
        ref = orot.nms_rotated(boxes, scores, thr)
        out = R.nms_rotated(boxes.cuda(), scores.cuda(), thr).cpu()
        if torch.equal(out, ref):
            print(f"   thr {thr}: equal ({ref.numel()} kept)")
            continue
        k = 0
        while k < min(out.numel(), ref.numel()) and out[k] == ref[k]:
            k += 1
        # the box at position k differs: one side kept a box the other suppressed
        cand = int(ref[k]) if k < ref.numel() else -1
        cand2 = int(out[k]) if k < out.numel() else -1
        print(f"   thr {thr}: differ at position {k}: oracle keeps {cand}, gpu keeps {cand2} (kept {ref.numel()} vs {out.numel()})")
        order = torch.sort(scores, descending=True, stable=True).indices
        rank = {int(b): r for r, b in enumerate(order.tolist())}
        j = cand if (cand2 < 0 or rank[cand] < rank[cand2]) else cand2       # the earlier one in score order was dropped by someone
        who = "gpu dropped it" if j == cand else "oracle dropped it"
        kept_before = [int(x) for x in ref[:k].tolist()]
        for i in kept_before:
            v = float(ref_m[i, j])
            a, b = boxes[i:i + 1].numpy(), boxes[j:j + 1].numpy()
            fast, full, gate, dec = run_pairs(hr, a, b, thr)
            if dec[0] or v > thr or full[0] > thr:
                print(f"      {who}: box {j} vs kept {i}: oracle IoU {v!r} gpu-pairwise {float(gpu_m[i, j])!r} host fast {fast[0]!r} "
                      f"full {full[0]!r} gate {gate[0]} host-decide {dec[0]}")
                print("         A", a[0], "B", b[0])
        break
