"""Geometry + post-process, drop-in for the reference's lib/general.py hot-path functions.

  post_process(predictions, conf_thres, iou_thres)   lib/general.py:136-183
  norm_angle(theta)                                  lib/general.py:7-20
  nms_rotated / pairwise_iou_rotated                 detectron2 ops the reference calls at
                                                     lib/general.py:177 and test.py:135
All device work happens in libryolo_b200.so; one host read (detections per image) ends post_process
because the reference API returns a ragged python list.
"""
import math

import torch

from .. import _lib as L

MAX_WH, MAX_NMS, MAX_DET = 4096.0, 5000, 1500      # lib/general.py:147-149


def norm_angle(theta):
    """lib/general.py:7-20 (single wrap + range assert); tiny elementwise helper kept in torch."""
    half = math.pi / 2
    theta = torch.where(theta >= half, theta - math.pi, theta)
    theta = torch.where(theta < -half, theta + math.pi, theta)
    assert torch.logical_and(-half <= theta, theta < half).all(), \
        "Theta of oriented bounding boxes are not within the boundary [-pi / 2, pi / 2)"
    return theta


def pairwise_iou_rotated(boxes1, boxes2):
    """[N,5] x [M,5] (cx, cy, w, h, angle degrees) -> [N,M] fp32 skew IoU."""
    L.require_cuda(boxes1, "boxes1")
    L.require_cuda(boxes2, "boxes2")
    a = boxes1.detach().contiguous().float()
    b = boxes2.detach().contiguous().float()
    n, m = a.shape[0], b.shape[0]
    out = torch.empty((n, m), dtype=torch.float32, device=a.device)
    if n == 0 or m == 0:
        return out
    lib = L.lib()
    nb = lib.ryolo_pairwise_iou_rotated_workspace(n, m)
    ws = L.workspace(nb, a.device, "iou")
    L.check(lib.ryolo_pairwise_iou_rotated(L.ptr(a), n, L.ptr(b), m, L.ptr(out), L.ptr(ws), nb, L.stream()))
    return out


def nms_rotated(boxes, scores, iou_threshold):
    """detectron2.layers.nms.nms_rotated: int64 keep indices, score-descending."""
    L.require_cuda(boxes, "boxes")
    b = boxes.detach().contiguous().float()
    s = scores.detach().contiguous().float()
    n = b.shape[0]
    if n == 0:
        return torch.zeros(0, dtype=torch.int64, device=b.device)
    lib = L.lib()
    keep = torch.empty(n, dtype=torch.int64, device=b.device)
    nk = torch.zeros(1, dtype=torch.int32, device=b.device)
    nb = lib.ryolo_nms_rotated_workspace(n)
    ws = L.workspace(nb, b.device, "nms")
    L.check(lib.ryolo_nms_rotated(L.ptr(b), L.ptr(s), n, float(iou_threshold), L.ptr(keep), L.ptr(nk), L.ptr(ws), nb,
                                  L.stream()))
    return keep[: int(nk.item())]


def post_process_device(predictions, conf_thres=0.5, iou_thres=0.4, mutate=True):
    """Batched post_process without any host synchronisation.

    Returns (dets [B,1500,7], rows [B,1500] int64 source rows, n [B] int32), all on the device.
    """
    L.require_cuda(predictions, "predictions")
    assert predictions.dim() == 3 and predictions.dtype == torch.float32 and predictions.is_contiguous(), \
        "predictions must be a contiguous fp32 [B, R, 6+nc] tensor"
    B, R, C = predictions.shape
    nc = C - 6
    dev = predictions.device
    dets = torch.empty((B, MAX_DET, 7), dtype=torch.float32, device=dev)
    rows = torch.empty((B, MAX_DET), dtype=torch.int64, device=dev)
    n = torch.zeros((B,), dtype=torch.int32, device=dev)
    if B == 0:
        return dets, rows, n
    lib = L.lib()
    nb = lib.ryolo_post_process_workspace(B, R, nc, MAX_NMS)
    ws = L.workspace(nb, dev, "pp")
    L.check(lib.ryolo_post_process(L.ptr(predictions), B, R, nc, float(conf_thres), float(iou_thres), MAX_NMS, MAX_DET,
                                   MAX_WH, 1 if mutate else 0, L.ptr(dets), L.ptr(rows), L.ptr(n), L.ptr(ws), nb,
                                   L.stream()))
    L.count(3 + 3 * (((MAX_NMS + 63) // 64 + 7) // 8))    # score, select/sort, gather + 3 launches per NMS band (8 tiles)
    return dets, rows, n


def post_process(predictions, conf_thres=0.5, iou_thres=0.4, return_indices=False):
    """Drop-in for lib/general.py:136-183: python list (len B) of [n_i, 7] fp32 tensors
    (x, y, w, h, theta, score, class).  Like the reference it multiplies the class columns of
    `predictions` by objectness in place."""
    dets, rows, n = post_process_device(predictions, conf_thres, iou_thres, mutate=True)
    counts = n.tolist()                                   # the only host read
    outs = [dets[i, :c] for i, c in enumerate(counts)]
    if return_indices:
        return outs, [rows[i, :c] for i, c in enumerate(counts)]
    return outs


non_max_suppression = post_process      # north_star alias


def encode_labels(polys, csl=True):
    """Polygon targets -> loss-ready label rows on the device (datasets/base_dataset.py:137-154 in one launch).

    polys: CUDA fp32 [T, 10] rows (image index, class, x1, y1, ..., x4, y4), vertices clockwise, i.e. the reference's
    `targets` after `collate_fn` stamped the sample index (base_dataset.py:161-167).  Returns [T, 187] rows
    (image, class, x, y, w, h, theta, csl[180]) when `csl`, else [T, 7]; T == 0 gives the reference's empty
    zeros((0, 187)) / zeros((0, 7))."""
    L.require_cuda(polys, "polys")
    assert polys.dim() == 2 and polys.shape[1] == 10, "polys must be [T, 10]"
    p = polys.contiguous().float()
    out = torch.empty((p.shape[0], 187 if csl else 7), dtype=torch.float32, device=p.device)
    L.check(L.lib().ryolo_encode_labels(L.ptr(p), p.shape[0], 1 if csl else 0, L.ptr(out), L.stream()))
    L.count(1)
    return out


def xyxyxyxy2xywha(boxes):
    """Drop-in for lib/general.py:70-104 on the device: [N, 8] clockwise vertices -> [N, 5] (x, y, w, h, theta)."""
    L.require_cuda(boxes, "boxes")
    polys = torch.cat((boxes.new_zeros(boxes.shape[0], 2), boxes.float()), 1)
    return encode_labels(polys, csl=False)[:, 2:]
