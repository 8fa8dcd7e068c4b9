"""Loss + anchor assignment, drop-in for the reference's lib/loss.py.

  ComputeCSLLoss(model, hyp)(outputs, targets)     lib/loss.py:153-268, build_targets :270-331
  ComputeKFIoULoss(model, hyp)(outputs, targets)   lib/loss.py:334-425, build_targets :427-492
Value AND gradient are produced by one fused pass of CUDA kernels (csrc/loss.cu); autograd only sees
a single Function whose backward hands out the precomputed gradients.
"""
import ctypes

import numpy as np
import torch

from .. import _lib as L


def _grid_hw(levels):
    arr = (ctypes.c_int32 * 6)()
    for i, p in enumerate(levels):
        arr[2 * i], arr[2 * i + 1] = p.shape[2], p.shape[3]
    return arr


def _run_loss(cfg, targets, levels, need_grad):
    """One fused pass: returns (items fp32[8] on device, list of d loss / d level or None)."""
    mode, na, nc, anchors, hyp = cfg
    lib = L.lib()
    dev = levels[0].device
    for p in levels:
        L.require_cuda(p, "outputs")
        if p.dtype != torch.float32:
            raise L.RyoloError(f"outputs must be float32 head levels (got {p.dtype}): the loss kernels read fp32")
    T = targets.shape[0]
    if T and targets.device != dev:
        raise L.RyoloError(f"targets are on {targets.device} but the head levels are on {dev} (train.py:187 moves "
                           "them to the model's device)")
    lv = [p.detach().contiguous() for p in levels]
    B = lv[0].shape[0]
    tg = targets.detach().contiguous().float() if T else torch.zeros((0, 187 if mode == 0 else 7), device=dev)
    ghw = _grid_hw(lv)
    nb = lib.ryolo_loss_workspace(B, na, ghw, T)
    ws = L.workspace(nb, dev, "loss")
    grads = [torch.empty_like(p) for p in lv] if need_grad else None
    items = torch.empty(8, dtype=torch.float32, device=dev)
    lp = (ctypes.c_void_p * 3)(*[p.data_ptr() for p in lv])
    gp = (ctypes.c_void_p * 3)(*[g.data_ptr() for g in grads]) if need_grad else None
    L.check(lib.ryolo_loss(mode, lp, gp, B, na, nc, ghw, L.ptr(tg), T, tg.shape[1], L.ptr(anchors), hyp,
                           L.ptr(items), L.ptr(ws), nb, L.stream()))
    L.count(10 if T else 4)
    return items, grads


class _KFLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target):
        L.require_cuda(pred, "pred")
        p, t = pred.detach().contiguous().float(), target.detach().contiguous().float()
        N = p.shape[0]
        dev = p.device
        kfiou = torch.empty(N, dtype=torch.float32, device=dev)
        grad = torch.empty_like(p) if ctx.needs_input_grad[0] else None
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        ws = L.workspace(16, dev, "kfloss")
        L.check(L.lib().ryolo_kfloss(L.ptr(p), L.ptr(t), N, L.ptr(kfiou), L.ptr(grad), L.ptr(loss), L.ptr(ws), 16,
                                     L.stream()))
        L.count(2)
        ctx.grad = grad
        ctx.mark_non_differentiable(kfiou)
        return loss.reshape(()), kfiou

    @staticmethod
    def backward(ctx, gloss, _gk):
        g = ctx.grad
        ctx.grad = None
        return (g * gloss if g is not None else None), None


class KFLoss(torch.nn.Module):
    """Drop-in for lib/loss.py:81-150 (fun='exp', alpha=3): forward(pred[N,5], target[N,5]) -> (loss, KFIoU[N])."""

    def __init__(self, fun='exp', alpha=3.0):
        super().__init__()
        assert fun == 'exp' and alpha == 3.0, "only the reference's live configuration (exp, alpha=3) is implemented"

    def forward(self, pred, target):
        return _KFLossFn.apply(pred, target)


class _FusedLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg, targets, *levels):
        items, grads = _run_loss(cfg, targets, levels, any(ctx.needs_input_grad[2:]))
        ctx.grads = grads
        ctx.mark_non_differentiable(items)
        return items[4:5].clone(), items

    @staticmethod
    def backward(ctx, gloss, _gitems):
        grads = ctx.grads
        ctx.grads = None
        if grads is None:
            return (None, None, None, None, None)
        return (None, None) + tuple(g.mul_(gloss) for g in grads)


class _ComputeLoss:
    MODE = 0
    KEYS = ()

    def __init__(self, model, hyp):
        device = next(model.parameters()).device            # lib/loss.py:155
        L.require_cuda(torch.empty(0, device=device), "model parameters")
        self.device = device
        self.hyp = dict(hyp)
        self.lambda_coord, self.lambda_conf_scale, self.lambda_cls_scale = hyp['box'], hyp['obj'], hyp['cls']
        self.lambda_theta, self.gr = 0.5, 1.0               # lib/loss.py:160-161
        an = np.zeros((3, len(model.anchors[0]), 3), dtype=np.float32)
        for i, lvl in enumerate(model.anchors):
            for j, a in enumerate(lvl):
                an[i, j, :len(a)] = a
        self.anchors = torch.tensor(model.anchors, device=device)
        self._anchors3 = torch.from_numpy(an).to(device)
        self.na, self.nl, self.nc = an.shape[1], 3, model.nc
        self._hyp = (ctypes.c_float * 7)(hyp['box'], hyp['obj'], hyp['cls'], hyp['obj_pw'], hyp['cls_pw'],
                                         hyp['fl_gamma'], self.lambda_theta)
        self.loss_items = {k: 0 for k in self.KEYS}          # must exist before the first call (train.py:178)
        self.sync_items = True

    def __call__(self, outputs, target):
        for p in outputs:
            L.require_cuda(p, "outputs")
        cfg = (self.MODE, self.na, self.nc, self._anchors3, self._hyp)
        loss, items = _FusedLoss.apply(cfg, target, *outputs)
        self.last_items_device = items
        if self.sync_items:
            v = items.tolist()                               # ONE device->host read (the reference does 4-5)
            vals = dict(reg_loss=v[0], theta_loss=v[1], conf_loss=v[2], cls_loss=v[3], total_loss=v[4])
            self.loss_items.update({k: vals[k] for k in self.KEYS})
        return loss, self.loss_items

    def value_and_grad(self, outputs, target):
        """Native path (no autograd): (items fp32[8] on device = reg, theta, conf, cls, total, n_pos x3;
        [d loss / d level] x 3)."""
        cfg = (self.MODE, self.na, self.nc, self._anchors3, self._hyp)
        items, grads = _run_loss(cfg, target, outputs, True)
        self.last_items_device = items
        return items, grads

    def build_targets(self, p, targets):
        """Reference-format assignment tuples (bit-exact indices, reference emission order)."""
        lib = L.lib()
        L.require_cuda(targets, "targets")
        dev = targets.device
        T = targets.shape[0]
        B = p[0].shape[0]
        ghw = _grid_hw(p)
        cap = 5 * self.na * T
        rec = torch.zeros((3, max(cap, 1), 12), dtype=torch.int32, device=dev)
        counts = torch.zeros(3, dtype=torch.int32, device=dev)
        tg = targets.detach().contiguous().float()
        if T:
            nb = lib.ryolo_loss_workspace(B, self.na, ghw, T)
            ws = L.workspace(nb, dev, "loss")
            L.check(lib.ryolo_build_targets(L.ptr(tg), T, tg.shape[1], 1 if self.MODE == 1 else 0,
                                            L.ptr(self._anchors3), self.na, B, ghw, L.ptr(rec), L.ptr(counts),
                                            L.ptr(ws), nb, L.stream()))
        n = counts.tolist()
        out = []
        for i in range(3):
            r = rec[i, :n[i]]
            f = r.view(torch.float32)
            idx = tuple(r[:, k].long() for k in range(4))
            out.append(dict(indices=idx, tbox=f[:, 4:9] if self.MODE == 1 else f[:, 4:8], angle=f[:, 8:9],
                            tcls=r[:, 9].long(), row=r[:, 10].long(), anch=self.anchors[i][idx[1]]))
        return self._pack(out, tg)


class ComputeCSLLoss(_ComputeLoss):
    MODE = 0
    KEYS = ("reg_loss", "theta_loss", "conf_loss", "cls_loss", "total_loss")

    def _pack(self, out, tg):
        tcls = [o["tcls"] for o in out]
        tbox = [o["tbox"] for o in out]
        ta = [o["angle"] * 180 / np.pi for o in out]
        tgl = [tg[o["row"], 7:187] for o in out]
        return tcls, tbox, ta, tgl, [o["indices"] for o in out], [o["anch"] for o in out]


class ComputeKFIoULoss(_ComputeLoss):
    MODE = 1
    KEYS = ("reg_loss", "conf_loss", "cls_loss", "total_loss")

    def _pack(self, out, tg):
        return [o["tcls"] for o in out], [o["tbox"] for o in out], [o["indices"] for o in out], \
            [o["anch"] for o in out]
