"""Evaluation matching + AP for the reference's validation loop (test.py:14-165, train.py:41-44).

The expensive half — `get_batch_statistics`' per-image, per-class, per-detection Python loops with one `.item()` each
(test.py:100-149) — is ONE kernel launch for the whole batch here (`ryolo_eval_match`, csrc/nms.cu: exact skew IoU of
every detection against the targets of its class + the sequential claim bookkeeping, one CTA per image) followed by
one device->host copy.  It takes `post_process_device`'s padded outputs directly (`match_batch`), or the reference's
list-of-tensors form (`get_batch_statistics`).

The precision/recall/AP reduction (`ap_per_class`) follows the definitions test.py implements — per class, detections in
descending confidence; recall = TP / n_labels, precision = TP / (TP + FP); AP = area under the monotone precision
envelope sampled at 101 recall points; P / R / F1 reported at the confidence that maximises the mean F1 — computed
for all IoU thresholds at once on a class-sorted layout instead of class-by-class boolean masks.  Pinned to the
reference's outputs by tests/golden/metrics.pt and eval_stats.pt.
"""
import numpy as np
import torch

from .. import _lib as L

_trapz = getattr(np, "trapezoid", None) or np.trapz
_RECALL_GRID = np.linspace(0, 1, 101)          # test.py:93
_CONF_GRID = np.linspace(0, 1, 1000)           # test.py:33


def match_batch(dets, n_det, targets, iouv):
    """Device form of test.py:100-149.  dets [B, max_det, 7] (x, y, w, h, theta rad, score, class), n_det int32 [B]
    (e.g. from post_process_device), targets [T, >=7] (image, class, x, y, w, h, theta rad), iouv [niou].
    Returns tp uint8 [B, max_det, niou] on the device (rows >= n_det[b] are undefined)."""
    L.require_cuda(dets, "dets")
    L.require_cuda(targets, "targets")
    assert dets.dim() == 3 and dets.shape[2] == 7 and dets.dtype == torch.float32 and dets.is_contiguous()
    B, max_det, _ = dets.shape
    dev = dets.device
    tg = targets.detach().contiguous().float()
    iv = iouv.detach().to(dev).contiguous().float()
    niou = iv.numel()
    tp = torch.empty((B, max_det, niou), dtype=torch.uint8, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    if B == 0:
        return tp
    lib = L.lib()
    nb = lib.ryolo_eval_match_workspace(B, max_det)
    ws = L.workspace(nb, dev, "match")
    L.check(lib.ryolo_eval_match(L.ptr(dets), L.ptr(n_det.to(torch.int32).contiguous()), B, max_det, L.ptr(tg),
                                 tg.shape[0], tg.shape[1] if tg.dim() == 2 and tg.shape[0] else 7, L.ptr(iv), niou,
                                 L.ptr(tp), L.ptr(status), L.ptr(ws), nb, L.stream()))
    L.count(1)
    match_batch.last_status = status
    return tp


def get_batch_statistics(outputs, targets, iouv, niou):
    """test.py:100-149 with the reference's call shape: `outputs` is post_process' list of [n_i, 7] tensors.
    Returns the reference's list of (true_positives [n_i, niou] bool, scores, labels, target classes) per image (images
    without predictions AND without labels are skipped, like the reference).  Like the reference it rewrites the
    prediction angles of images that have labels to degrees IN PLACE (test.py:124)."""
    B = len(outputs)
    dev = targets.device
    counts = [int(o.shape[0]) for o in outputs]
    max_det = max(1, max(counts) if counts else 1)
    dets = torch.zeros((B, max_det, 7), dtype=torch.float32, device=dev)
    for i, o in enumerate(outputs):
        if counts[i]:
            dets[i, :counts[i]] = o
    n_det = torch.tensor(counts, dtype=torch.int32, device=dev)
    tp = match_batch(dets, n_det, targets, iouv[:niou]).cpu().numpy().astype(bool)     # the one device->host copy
    if int(match_batch.last_status.item()):
        raise L.RyoloError("get_batch_statistics: an image has more than 1024 targets")
    timg = targets[:, 0].long().cpu().numpy() if targets.shape[0] else np.zeros(0, np.int64)
    tcls_all = targets[:, 1].cpu().numpy() if targets.shape[0] else np.zeros(0, np.float32)
    stats = []
    for i, o in enumerate(outputs):
        tcls = tcls_all[timg == i].tolist()
        if counts[i] == 0:
            if tcls:
                stats.append((np.zeros((0, niou), dtype=bool), np.empty(0), np.empty(0), tcls))
            continue
        if tcls:
            o[:, 4] = o[:, 4] / np.pi * 180
        stats.append((torch.from_numpy(tp[i, :counts[i]].copy()), o[:, 5].cpu(), o[:, 6].cpu(), tcls))
    return stats


def compute_ap(recall, precision):
    """Area under the precision envelope, 101-point interpolation (test.py:70-97).  recall / precision: [n] curves of
    one class at one IoU threshold.  Returns (ap, envelope precision, padded recall)."""
    r = np.concatenate(([0.0], recall, [recall[-1] + 0.01]))
    p = np.concatenate(([1.0], precision, [0.0]))
    env = np.maximum.accumulate(p[::-1])[::-1]
    return _trapz(np.interp(_RECALL_GRID, r, env), _RECALL_GRID), env, r


def ap_per_class(tp, conf, pred_cls, target_cls):
    """P, R, AP[class, iou], F1 and the class ids (test.py:14-67).  tp [n, niou] bool, conf [n], pred_cls [n],
    target_cls [n_labels]."""
    rank = np.argsort(-conf)
    tp, conf, pred_cls = np.asarray(tp)[rank], np.asarray(conf)[rank], np.asarray(pred_cls)[rank]
    classes, n_labels = np.unique(target_cls, return_counts=True)
    niou = tp.shape[1]
    ap = np.zeros((len(classes), niou))
    p_at = np.zeros((len(classes), _CONF_GRID.size))
    r_at = np.zeros((len(classes), _CONF_GRID.size))
    # class-sorted layout (stable: keeps the confidence order inside a class), one cumulative sum for everything
    by_cls = np.argsort(pred_cls, kind="stable")
    hits = np.cumsum(tp[by_cls], axis=0, dtype=np.float64)
    miss = np.cumsum(~tp[by_cls].astype(bool), axis=0, dtype=np.float64)
    cls_sorted, conf_sorted = pred_cls[by_cls], conf[by_cls]
    lo = np.searchsorted(cls_sorted, classes, side="left")
    hi = np.searchsorted(cls_sorted, classes, side="right")
    for k, (a, b, nl) in enumerate(zip(lo, hi, n_labels)):
        if a == b or nl == 0:
            continue
        base_h = hits[a - 1] if a else 0.0
        base_m = miss[a - 1] if a else 0.0
        tpc, fpc = hits[a:b] - base_h, miss[a:b] - base_m
        recall = tpc / (nl + 1e-16)
        precision = tpc / (tpc + fpc)
        r_at[k] = np.interp(-_CONF_GRID, -conf_sorted[a:b], recall[:, 0], left=0)
        p_at[k] = np.interp(-_CONF_GRID, -conf_sorted[a:b], precision[:, 0], left=1)
        for j in range(niou):
            ap[k, j] = compute_ap(recall[:, j], precision[:, j])[0]
    f1 = 2 * p_at * r_at / (p_at + r_at + 1e-16)
    best = f1.mean(0).argmax()
    return p_at[:, best], r_at[:, best], ap, f1[:, best], classes.astype("int32")


def calculate_eval_stats(stats, num_classes):
    """(labels per class, P, R, AP50, AP, F1, class ids, mean P, mean R, mAP50, mAP) from the concatenated batch
    statistics (test.py:152-165); zeros when there is nothing to score."""
    out = dict(p=0.0, r=0.0, f1=0.0, mp=0.0, mr=0.0, map50=0.0, map=0.0, ap50=[], ap=[], cls=[], nt=torch.zeros(1))
    if len(stats) and stats[0].any():
        p, r, ap, f1, cls = ap_per_class(*stats)
        out.update(p=p, r=r, f1=f1, cls=cls, ap50=ap[:, 0], ap=ap.mean(1))
        out.update(mp=p.mean(), mr=r.mean(), map50=out["ap50"].mean(), map=out["ap"].mean(),
                   nt=np.bincount(stats[3].astype(np.int64), minlength=num_classes))
    return (out["nt"], out["p"], out["r"], out["ap50"], out["ap"], out["f1"], out["cls"], out["mp"], out["mr"],
            out["map50"], out["map"])


def fitness(x):
    """0.1 * mAP@0.5 + 0.9 * mAP@0.5:0.95 over rows (P, R, mAP@0.5, mAP@0.5:0.95) (train.py:41-44)."""
    return (np.asarray(x) * np.array([0.0, 0.0, 0.1, 0.9])).sum(0)
