"""Evaluation matching + AP, drop-in for the reference's test.py helpers:

  get_batch_statistics(outputs, targets, iouv, niou)   test.py:100-149
  ap_per_class(tp, conf, pred_cls, target_cls)         test.py:14-67
  compute_ap(recall, precision)                        test.py:70-97

The skew-IoU matrices come from the device kernel behind `pairwise_iou_rotated` (csrc/nms.cu); the greedy
"each target is claimed once, in detection order" bookkeeping is inherently sequential and stays on the host,
but works on one device->host copy per (image, class) instead of one `.item()` per detection.
"""
import numpy as np
import torch

from .general import pairwise_iou_rotated as _device_iou

_trapz = getattr(np, "trapezoid", None) or np.trapz


def get_batch_statistics(outputs, targets, iouv, niou, iou_fn=None):
    """Per image: (true_positives [n_pred, niou] bool, scores, labels, target classes) like test.py:100-149.
    outputs: list of [n,7] (x,y,w,h,theta rad,score,cls); targets [T, >=7] (img, cls, x,y,w,h,theta rad).
    Like the reference it converts the prediction angles to degrees IN PLACE."""
    iou_fn = iou_fn or _device_iou
    batch_stats = []
    for sample_i, pred in enumerate(outputs):
        tar = targets[targets[:, 0] == sample_i, 1:]
        nl = len(tar)
        tcls = tar[:, 0].tolist() if nl else []
        if len(pred) == 0:
            if nl:
                batch_stats.append((np.zeros((0, niou), dtype=bool), np.empty(0), np.empty(0), tcls))
            continue
        pred_boxes, pred_scores, pred_labels = pred[:, :5], pred[:, 5], pred[:, 6]
        true_positives = torch.zeros(pred.shape[0], niou, dtype=torch.bool, device=targets.device)
        if nl:
            n_detected = 0
            target_labels = tar[:, 0]
            target_boxes = tar[:, 1:6]
            pred_boxes[:, 4] = pred_boxes[:, 4] / np.pi * 180          # test.py:124 (in place on the caller's tensor)
            target_boxes[:, 4] = target_boxes[:, 4] / np.pi * 180
            for cls in torch.unique(target_labels):
                ti = (cls == target_labels).nonzero(as_tuple=False).view(-1)
                pi = (cls == pred_labels).nonzero(as_tuple=False).view(-1)
                if not pi.shape[0]:
                    continue
                ious, best = iou_fn(pred_boxes[pi], target_boxes[ti]).max(1)
                hit = ious > iouv[0]
                cand = hit.nonzero(as_tuple=False).view(-1)
                if not cand.numel():
                    continue
                # one host copy per class: candidate detections (in detection order) and the target each one wants
                cand_h = cand.tolist()
                want_h = ti[best[cand]].tolist()
                claimed, rows = set(), []
                for j, d in zip(cand_h, want_h):
                    if d not in claimed:
                        claimed.add(d)
                        n_detected += 1
                        rows.append(j)
                        if n_detected == nl:                             # test.py:143 (breaks this class only)
                            break
                if rows:
                    r = torch.tensor(rows, dtype=torch.long, device=pred.device)
                    true_positives[pi[r]] = ious[r][:, None] > iouv[None, :]
        batch_stats.append((true_positives.cpu(), pred_scores.cpu(), pred_labels.cpu(), tcls))
    return batch_stats


def compute_ap(recall, precision):
    """test.py:70-97 — 101-point interpolated AP with the precision envelope."""
    mrec = np.concatenate(([0.], recall, [recall[-1] + 0.01]))
    mpre = np.concatenate(([1.], precision, [0.]))
    mpre = np.flip(np.maximum.accumulate(np.flip(mpre)))
    x = np.linspace(0, 1, 101)
    ap = _trapz(np.interp(x, mrec, mpre), x)
    return ap, mpre, mrec


def ap_per_class(tp, conf, pred_cls, target_cls):
    """test.py:14-67 — returns (p, r, ap, f1, unique_classes) at the max-mean-F1 confidence."""
    order = np.argsort(-conf)
    tp, conf, pred_cls = tp[order], conf[order], pred_cls[order]
    unique_classes = np.unique(target_cls)
    nc = unique_classes.shape[0]
    px = np.linspace(0, 1, 1000)
    ap, p, r = np.zeros((nc, tp.shape[1])), np.zeros((nc, 1000)), np.zeros((nc, 1000))
    for ci, c in enumerate(unique_classes):
        sel = pred_cls == c
        n_l, n_p = (target_cls == c).sum(), sel.sum()
        if n_p == 0 or n_l == 0:
            continue
        fpc = (1 - tp[sel]).cumsum(0)
        tpc = tp[sel].cumsum(0)
        recall = tpc / (n_l + 1e-16)
        r[ci] = np.interp(-px, -conf[sel], recall[:, 0], left=0)
        precision = tpc / (tpc + fpc)
        p[ci] = np.interp(-px, -conf[sel], precision[:, 0], left=1)
        for j in range(tp.shape[1]):
            ap[ci, j], _, _ = compute_ap(recall[:, j], precision[:, j])
    f1 = 2 * p * r / (p + r + 1e-16)
    i = f1.mean(0).argmax()
    return p[:, i], r[:, i], ap, f1[:, i], unique_classes.astype('int32')


def calculate_eval_stats(stats, num_classes):
    """Per-class and mean P / R / AP@0.5 / AP@0.5:0.95 from the concatenated batch statistics (test.py:152-165)."""
    p, r, f1, mp, mr, map50, map_ = 0., 0., 0., 0., 0., 0., 0.
    ap50, ap, ap_class = [], [], []
    if len(stats) and stats[0].any():
        p, r, ap, f1, ap_class = ap_per_class(*stats)
        ap50, ap = ap[:, 0], ap.mean(1)
        mp, mr, map50, map_ = p.mean(), r.mean(), ap50.mean(), ap.mean()
        nt = np.bincount(stats[3].astype(np.int64), minlength=num_classes)
    else:
        nt = torch.zeros(1)
    return nt, p, r, ap50, ap, f1, ap_class, mp, mr, map50, map_


def fitness(x):
    """Model fitness = 0.1 mAP@0.5 + 0.9 mAP@0.5:0.95 over rows (P, R, mAP@0.5, mAP@0.5:0.95) (train.py:41-44)."""
    w = [0.0, 0.0, 0.1, 0.9]
    return (x * w).sum(0)
