"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by r-yolov4_b200/).

CPU (torch fp32, differentiable) restatement of the reference's decode / target assignment /
loss / post-process glue.  Each function cites the reference file:line it follows.  PINNED: every
function here is checked against outputs of the reference's own Python (tests/golden/*.pt, made by
tests/golden/make_golden.py from /root/reference) in tests/test_oracle_golden.py.

Deliberate restatement choices (all documented in DESIGN.md):
  * KFLoss is O(N): the reference's accidental [N,1]+[N] -> [N,N] broadcast (lib/loss.py:114,148)
    equals mean(xy_loss) + mean(kf_loss) whenever no element is clamped; we reproduce the exact
    value, including the clamp, with an O(N) separable form (see kf_loss()).
  * duplicate objectness cells: last writer in reference order wins.
  * post_process sort is stable descending (lower row first on ties) — Appendix C #6.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import rotated as _rot

HALF_PI = np.pi / 2


# ----------------------------------------------------------------------------- geometry helpers
def norm_angle(theta):
    """lib/general.py:7-20 — single wrap into [-pi/2, pi/2) + range assert."""
    theta = torch.where(theta >= HALF_PI, theta - np.pi, theta)
    theta = torch.where(theta < -HALF_PI, theta + np.pi, theta)
    assert bool(((theta >= -HALF_PI) & (theta < HALF_PI)).all()), \
        "Theta of oriented bounding boxes are not within the boundary [-pi / 2, pi / 2)"
    return theta


# ----------------------------------------------------------------------------- anchors
def make_anchors(anchors, strides=(8, 16, 32)):
    """model/yolo.py:53-61 — 3 (w,h) anchors per level in grid units."""
    return [[[lvl[i] / s, lvl[i + 1] / s] for i in range(0, len(lvl), 2)] for s, lvl in zip(strides, anchors)]


def make_rotated_anchors(anchors, angles_deg, strides=(8, 16, 32)):
    """model/yolo.py:63-72 — 3 sizes x 6 angles (angle fastest) per level, (w,h,rad)."""
    ang = [a * np.pi / 180 for a in angles_deg]
    return [[[lvl[i] / s, lvl[i + 1] / s, t] for i in range(0, len(lvl), 2) for t in ang]
            for s, lvl in zip(strides, anchors)]


# ----------------------------------------------------------------------------- decode
def head_to_grid(x, na, ch):
    """model/yololayer.py:25,76 — NCHW head -> [B,na,gs,gs,ch] contiguous."""
    b, _, gh, gw = x.shape
    return x.view(b, na, ch, gh, gw).permute(0, 1, 3, 4, 2).contiguous()


def _grid(gs, device):
    ys, xs = torch.meshgrid(torch.arange(gs, device=device), torch.arange(gs, device=device), indexing="ij")
    return torch.stack((xs, ys), -1).view(1, 1, gs, gs, 2)


def decode_csl(levels, anchors, nc, strides=(8, 16, 32)):
    """model/yololayer.py:28-56 — levels: 3 x [B,3,gs,gs,nc+185] -> [B,R,nc+6]."""
    outs = []
    for p, anc, s in zip(levels, anchors, strides):
        b, na, gs = p.shape[0], p.shape[1], p.shape[2]
        a = torch.tensor(anc, device=p.device)[:, :2].view(1, na, 1, 1, 2)
        y = p.sigmoid()
        xy = (y[..., 0:2] * 2 - 0.5 + _grid(gs, p.device)) * s
        wh = (y[..., 2:4] * 2) ** 2 * a * s
        bins = y[..., 5 + nc:].argmax(-1, keepdim=True)  # first max wins (sigmoid space)
        th = (bins - 90) / 180 * np.pi
        outs.append(torch.cat((xy, wh, th, y[..., 4:5], y[..., 5:5 + nc]), -1).view(b, -1, nc + 6))
    return torch.cat(outs, 1)


def decode_kfiou(levels, anchors, nc, strides=(8, 16, 32)):
    """model/yololayer.py:79-105 — levels: 3 x [B,18,gs,gs,nc+6] -> [B,R,nc+6]."""
    outs = []
    for p, anc, s in zip(levels, anchors, strides):
        b, na, gs = p.shape[0], p.shape[1], p.shape[2]
        at = torch.tensor(anc, device=p.device)
        awh = at[:, :2].view(1, na, 1, 1, 2)
        aa = at[:, 2].view(1, na, 1, 1, 1)
        y = p.sigmoid()
        xy = (y[..., 0:2] * 2 - 0.5 + _grid(gs, p.device)) * s
        wh = (y[..., 2:4] * 2) ** 2 * awh * s
        ang = (y[..., 4:5] - 0.5) * 0.5236 + aa
        outs.append(torch.cat((xy, wh, ang, y[..., 5:6], y[..., 6:]), -1).view(b, -1, nc + 6))
    return torch.cat(outs, 1)


# ----------------------------------------------------------------------------- target assignment
_OFFS = ((0.0, 0.0), (0.5, 0.0), (0.0, 0.5), (-0.5, 0.0), (0.0, -0.5))


def assign_targets(targets, grids, anchors, rotated):
    """lib/loss.py:270-331 (csl) / :427-492 (kfiou).

    targets [T, 7+] = (img, cls, x, y, w, h, theta, ...); grids: list of (gh, gw); anchors: list of
    [na, 2|3].  Returns per level a dict with b, a, gj, gi (int64), tbox [n,4|5], tcls, anch, row
    (index of the source target row) in the reference's emission order: offset-major
    (centre, x-left, y-up, x-right, y-down) -> anchor -> target row.
    """
    out = []
    T = targets.shape[0]
    dev = targets.device
    for (gh, gw), anc in zip(grids, anchors):
        A = torch.as_tensor(anc, dtype=torch.float32, device=dev)
        na = A.shape[0]
        if T == 0:
            z = torch.zeros(0, dtype=torch.long, device=dev)
            out.append(dict(b=z, a=z, gj=z, gi=z, tbox=torch.zeros(0, 5 if rotated else 4, device=dev),
                            tcls=z, anch=A[z], row=z, ta=torch.zeros(0, 1, device=dev)))
            continue
        gwf, ghf = torch.tensor(float(gw)), torch.tensor(float(gh))
        tx, ty = targets[:, 2] * gwf, targets[:, 3] * ghf      # :292  t = targets * gain (fp32)
        tw, th = targets[:, 4] * gwf, targets[:, 5] * ghf
        rw, rh = tw[None] / A[:, 0:1], th[None] / A[:, 1:2]    # :297
        worst = torch.maximum(torch.maximum(rw, 1.0 / rw), torch.maximum(rh, 1.0 / rh))
        match = worst < 4.0                                    # :298  [na,T]
        if rotated:
            d = torch.abs(torch.cos(targets[:, 6][None] - A[:, 2:3]))  # :458
            match = match & (d > 0.866)
        ix, iy = gwf - tx, ghf - ty                            # :304
        fj = (torch.remainder(tx, 1.0) < 0.5) & (tx > 1.0)     # :305
        fk = (torch.remainder(ty, 1.0) < 0.5) & (ty > 1.0)
        fl = (torch.remainder(ix, 1.0) < 0.5) & (ix > 1.0)     # :306
        fm = (torch.remainder(iy, 1.0) < 0.5) & (iy > 1.0)
        sel = torch.stack((torch.ones_like(fj), fj, fk, fl, fm))  # [5,T]
        incl = sel[:, None, :] & match[None, :, :]             # [5,na,T]
        o, a, r = incl.nonzero(as_tuple=True)                  # row-major == reference order
        offs = torch.tensor(_OFFS, dtype=torch.float32, device=dev)
        gx, gy = tx[r], ty[r]
        gi = (gx - offs[o, 0]).long().clamp(0, gw - 1)         # :319,324 trunc then clamp
        gj = (gy - offs[o, 1]).long().clamp(0, gh - 1)
        cols = [gx - gi, gy - gj, tw[r], th[r]]                # :325 (uses the clamped gij)
        if rotated:
            cols.append(targets[r, 6])                         # :488
        out.append(dict(b=targets[r, 0].long(), a=a, gj=gj, gi=gi, tbox=torch.stack(cols, 1),
                        tcls=targets[r, 1].long(), anch=A[a], row=r,
                        ta=targets[r, 6:7] * 180 / np.pi))      # :326
    return out


# ----------------------------------------------------------------------------- box losses
def bbox_ciou(p, t):
    """lib/loss.py:36-78 — CIoU of (x,y,w,h) pairs, alpha detached, clamp to [-1,1]."""
    assert p.shape == t.shape
    x1, y1, w1, h1 = p.unbind(-1)
    x2, y2, w2, h2 = t.unbind(-1)
    l1, r1, t1, b1 = x1 - w1 / 2, x1 + w1 / 2, y1 - h1 / 2, y1 + h1 / 2
    l2, r2, t2, b2 = x2 - w2 / 2, x2 + w2 / 2, y2 - h2 / 2, y2 + h2 / 2
    iw = (torch.min(r1, r2) - torch.max(l1, l2)).clamp(min=0)
    ih = (torch.min(b1, b2) - torch.max(t1, t2)).clamp(min=0)
    inter = iw * ih
    centre = (x2 - x1) ** 2 + (y2 - y1) ** 2
    ow = (torch.max(r1, r2) - torch.min(l1, l2)).clamp(min=0)
    oh = (torch.max(b1, b2) - torch.min(t1, t2)).clamp(min=0)
    diag = ow ** 2 + oh ** 2
    union = w1 * h1 + w2 * h2 - inter
    u = centre / (diag + 1e-15)
    iou = inter / (union + 1e-15)
    v = (4 / (np.pi ** 2)) * torch.pow(torch.atan(w2 / h2) - torch.atan(w1 / h1), 2)
    with torch.no_grad():
        alpha = v / ((1 - iou) + v)
    return torch.clamp(iou - (u + alpha * v), min=-1.0, max=1.0)


def kf_terms(pred, target):
    """Per-pair terms of KFLoss (lib/loss.py:100-146 + lib/general.py:107-133), O(N).

    Returns xy_loss[N], kf_loss[N], KFIoU[N].
    """
    wp, hp = pred[:, 2].clamp(1e-4, 1e4), pred[:, 3].clamp(1e-4, 1e4)      # general.py:121
    wt, ht = target[:, 2].clamp(1e-4, 1e4), target[:, 3].clamp(1e-4, 1e4)
    rp, rt = pred[:, 4], target[:, 4]
    c, s = torch.cos(rt), torch.sin(rt)
    a, b = (0.5 * wt) ** 2, (0.5 * ht) ** 2                                 # general.py:129
    # Sigma_t = R diag(a,b) R^T   (general.py:127-131)
    s00, s01, s11 = c * c * a + s * s * b, c * s * (a - b), s * s * a + c * c * b
    det = s00 * s11 - s01 * s01
    dx, dy = pred[:, 0] - target[:, 0], pred[:, 1] - target[:, 1]
    quad = (dx * dx * s11 - 2 * dx * dy * s01 + dy * dy * s00) / det        # loss.py:114  d^T Sigma^-1 d
    xy_loss = torch.log(quad + 1)
    wp2, hp2, wt2, ht2 = wp ** 2, hp ** 2, wt ** 2, ht ** 2                 # loss.py:132-134
    c2, s2 = torch.cos(rp - rt) ** 2, torch.sin(rp - rt) ** 2
    A = torch.sqrt(1 + (wp2 * hp2) / (wt2 * ht2) + (wp2 / wt2 + hp2 / ht2) * c2 + (wp2 / ht2 + hp2 / wt2) * s2)
    B = torch.sqrt(1 + (wt2 * ht2) / (wp2 * hp2) + (wt2 / wp2 + ht2 / hp2) * c2 + (wt2 / hp2 + ht2 / wp2) * s2)
    kfiou = 1.0 / (A + B - 3.0)                                             # loss.py:139 (alpha = 3)
    kf_loss = torch.exp(1 - kfiou) - 1                                      # loss.py:144
    return xy_loss, kf_loss, kfiou


def kf_loss(pred, target):
    """KFLoss.forward (lib/loss.py:100-150) value in O(N).

    The reference forms loss[i,j] = clamp(xy_loss[i] + kf_loss[j], 0) over an N x N grid and takes
    its mean (F6).  xy_loss >= 0 (log of >= 1) and kf_loss = exp(1-KFIoU)-1 >= 0 because
    KFIoU = 1/(A+B-3) <= 1 (A,B >= sqrt(1+1+2)=2 at best), so the clamp never fires and the mean
    separates exactly into mean(xy_loss) + mean(kf_loss).  We still apply clamp(0) per term for
    robustness against rounding (it is the identity on valid boxes).
    """
    xy, kf, kfiou = kf_terms(pred, target)
    return xy.clamp(0).mean() + kf.clamp(0).mean(), kfiou


def bce_logits(x, t, pos_weight=1.0):
    """nn.BCEWithLogitsLoss(reduction='none') formula (ATen binary_cross_entropy_with_logits)."""
    lw = 1 + (pos_weight - 1) * t
    return (1 - t) * x + lw * (torch.log1p(torch.exp(-x.abs())) + (-x).clamp(min=0))


def focal(x, t, pos_weight, gamma, alpha=0.25):
    """lib/loss.py:10-33 FocalLoss wrapping BCEWithLogits ('mean' reduction applied by caller)."""
    loss = bce_logits(x, t, pos_weight)
    if gamma > 0:
        p = torch.sigmoid(x)
        p_t = t * p + (1 - t) * (1 - p)
        loss = loss * (t * alpha + (1 - t) * (1 - alpha)) * (1.0 - p_t) ** gamma
    return loss


def _scatter_last(tconf, b, a, gj, gi, val):
    """tconf[b,a,gj,gi] = val with the LAST duplicate (reference order) winning — Appendix C #8."""
    shape = tconf.shape
    lin = ((b * shape[1] + a) * shape[2] + gj) * shape[3] + gi
    lin_np = lin.cpu().numpy()
    _, first_in_rev = np.unique(lin_np[::-1], return_index=True)
    last = torch.from_numpy((len(lin_np) - 1 - first_in_rev).astype(np.int64))
    flat = tconf.view(-1)
    flat[lin[last]] = val[last]
    return tconf


def csl_loss(levels, targets, anchors, nc, hyp):
    """ComputeCSLLoss.__call__ (lib/loss.py:191-268). levels: 3 x [B,3,gs,gs,nc+185] (may require grad).

    Returns (loss[1], dict of python floats in the reference key order).
    """
    dev = targets.device
    g = hyp.get("fl_gamma", 0.0)
    reg, conf, cls, theta = (torch.zeros(1, device=dev) for _ in range(4))
    asg = assign_targets(targets, [(p.shape[2], p.shape[3]) for p in levels], anchors, rotated=False)
    for p, s in zip(levels, asg):
        tconf = torch.zeros_like(p[..., 0])
        if s["b"].numel() > 0:
            ps = p[s["b"], s["a"], s["gj"], s["gi"]]
            pxy = ps[:, 0:2].sigmoid() * 2 - 0.5
            pwh = (ps[:, 2:4].sigmoid() * 2) ** 2 * s["anch"][:, :2]
            ciou = bbox_ciou(torch.cat((pxy, pwh), -1), s["tbox"])
            reg = reg + (1.0 - ciou).mean()
            _scatter_last(tconf, s["b"], s["a"], s["gj"], s["gi"], ciou.detach().clamp(0))
            if nc > 1:
                t = torch.zeros_like(ps[:, 5:5 + nc])
                t[torch.arange(t.shape[0]), s["tcls"]] = 1
                cls = cls + focal(ps[:, 5:5 + nc], t, hyp["cls_pw"], g).mean()
            tg = targets[s["row"], 7:7 + 180]
            theta = theta + focal(ps[:, 5 + nc:], tg, 1.0, g).mean()
        conf = conf + focal(p[..., 4], tconf, hyp["obj_pw"], g).mean()
    reg, theta, conf, cls = hyp["box"] * reg, 0.5 * theta, hyp["obj"] * conf, hyp["cls"] * cls
    loss = reg + conf + cls + theta
    items = {"reg_loss": reg.item(), "theta_loss": theta.item(), "conf_loss": conf.item(),
             "cls_loss": cls.item(), "total_loss": loss.item()}
    return loss, items


def kfiou_loss(levels, targets, anchors, nc, hyp):
    """ComputeKFIoULoss.__call__ (lib/loss.py:368-425). levels: 3 x [B,18,gs,gs,nc+6]."""
    dev = targets.device
    g = hyp.get("fl_gamma", 0.0)
    reg, conf, cls = (torch.zeros(1, device=dev) for _ in range(3))
    asg = assign_targets(targets, [(p.shape[2], p.shape[3]) for p in levels], anchors, rotated=True)
    for p, s in zip(levels, asg):
        tconf = torch.zeros_like(p[..., 0])
        if s["b"].numel() > 0:
            ps = p[s["b"], s["a"], s["gj"], s["gi"]]
            pxy = ps[:, 0:2].sigmoid() * 2 - 0.5
            pwh = (ps[:, 2:4].sigmoid() * 2) ** 2 * s["anch"][:, :2]
            pa = norm_angle((ps[:, 4:5].sigmoid() - 0.5) * 1.1 + s["anch"][:, 2:])
            kl, kfiou = kf_loss(torch.cat((pxy, pwh, pa), -1), s["tbox"])
            reg = reg + kl
            _scatter_last(tconf, s["b"], s["a"], s["gj"], s["gi"], kfiou.detach().clamp(0))
            if nc > 1:
                t = torch.zeros_like(ps[:, 6:])
                t[torch.arange(t.shape[0]), s["tcls"]] = 1
                cls = cls + focal(ps[:, 6:], t, hyp["cls_pw"], g).mean()
        conf = conf + focal(p[..., 5], tconf, hyp["obj_pw"], g).mean()
    reg, conf, cls = hyp["box"] * reg, hyp["obj"] * conf, hyp["cls"] * cls
    loss = reg + conf + cls
    items = {"reg_loss": reg.item(), "conf_loss": conf.item(), "cls_loss": cls.item(),
             "total_loss": loss.item()}
    return loss, items


# ----------------------------------------------------------------------------- post-process
MAX_WH, MAX_NMS, MAX_DET = 4096, 5000, 1500


def post_process(predictions, conf_thres=0.5, iou_thres=0.4, return_indices=False):
    """lib/general.py:136-183 with the restated detectron2 nms_rotated plugged in.

    predictions [B,R,nc+6] is mutated in place (cls *= obj) like the reference (:155).
    Sort is STABLE descending; NMS suppresses on IoU > thr (frozen spec).
    """
    outs, idxs = [], []
    for img in predictions:
        img[:, 6:] *= img[:, 5:6]
        score, cid = img[:, 6:].max(1)
        rows = torch.nonzero(score > conf_thres).view(-1)
        if rows.numel() == 0:
            outs.append(torch.zeros((0, 7)))
            idxs.append(rows)
            continue
        order = torch.sort(score[rows], descending=True, stable=True).indices[:MAX_NMS]
        rows = rows[order]
        dets = torch.cat((img[rows, :5], score[rows, None], cid[rows, None].float()), 1)
        rb = dets[:, :5].clone()
        rb[:, :2] = rb[:, :2] + dets[:, 6:7] * MAX_WH
        rb[:, 4] = rb[:, 4] / np.pi * 180
        keep = _rot.nms_rotated(rb, dets[:, 5], iou_thres)[:MAX_DET]
        outs.append(dets[keep])
        idxs.append(rows[keep])
    return (outs, idxs) if return_indices else outs


def get_batch_statistics(outputs, targets, iouv, niou):
    """ORACLE restatement of test.py:100-149 (per image, per target class, per detection; the C++ restatement of
    detectron2's pairwise_iou_rotated behind it).  Pinned by tests/golden/metrics.pt (generated by running the
    reference's own function).  Mutates the prediction angles to degrees in place like the reference (:124)."""
    stats = []
    for si, pred in enumerate(outputs):
        tar = targets[targets[:, 0] == si, 1:]
        nl = len(tar)
        tcls = tar[:, 0].tolist() if nl else []
        if len(pred) == 0:
            if nl:
                stats.append((np.zeros((0, niou), dtype=bool), np.empty(0), np.empty(0), tcls))
            continue
        tp = torch.zeros(pred.shape[0], niou, dtype=torch.bool)
        if nl:
            found = 0
            pred[:, 4] = pred[:, 4] / np.pi * 180
            tb = tar[:, 1:6].clone()
            tb[:, 4] = tb[:, 4] / np.pi * 180
            for c in torch.unique(tar[:, 0]):
                ti = (tar[:, 0] == c).nonzero().view(-1)
                pi = (pred[:, 6] == c).nonzero().view(-1)
                if not pi.numel():
                    continue
                ious, arg = _rot.pairwise_iou_rotated(pred[pi, :5], tb[ti]).max(1)
                taken = set()
                for j in (ious > iouv[0]).nonzero().view(-1).tolist():
                    d = int(ti[arg[j]])
                    if d in taken:
                        continue
                    taken.add(d)
                    found += 1
                    tp[pi[j]] = ious[j] > iouv
                    if found == nl:
                        break
        stats.append((tp, pred[:, 5].clone(), pred[:, 6].clone(), tcls))
    return stats
