"""`Yolo(n_classes, model_config, mode, ver)` — drop-in for model/yolo.py:9-72 running on the
B200-native conv stack.  forward(imgs[B,3,S,S] fp32, training) returns the three
[B, na, gs, gs, ch] fp32 levels (and the decoded [B, R, nc+6] tensor when training is False)."""
import numpy as np
import torch
import torch.nn as nn

from .. import _lib as L
from .backbone import Backbonev4, Backbonev5, Backbonev7
from .blocks import Ctx
from .neck import Neckv4, Neckv5, Neckv7
from .yololayer import YoloCSLLayer, YoloKFIoULayer


class _StackFn(torch.autograd.Function):
    """Makes the natively executed conv stack one node of torch's autograd graph, so the reference's
    `loss.backward(); optimizer.step()` (train.py:198-202) works unchanged."""

    @staticmethod
    def forward(ctx, model, img, *params):
        ctx.model = model
        heads = model._run(img, True)
        ctx.fwd = model.last_ctx
        return tuple(heads)

    @staticmethod
    def backward(ctx, *dlevels):
        from .backward import run_backward
        model = ctx.model
        params = list(model.parameters())
        grads = {id(p): torch.zeros_like(p, dtype=torch.float32) for p in params}
        run_backward(model, ctx.fwd, [d.contiguous() for d in dlevels], grads)
        return (None, None) + tuple(grads[id(p)] for p in params)


class Yolo(nn.Module):
    def __init__(self, n_classes, model_config, mode, ver):
        super().__init__()
        anchors = model_config["anchors"]
        angles = [a * np.pi / 180 for a in model_config["angles"]]
        strides = [8, 16, 32]                                               # model/yolo.py:21
        if mode == "csl":
            self.na, self.ch = 3, 4 + 180 + 1 + n_classes                   # model/yolo.py:24
            an = self._make_anchors(strides, anchors)
            layer = YoloCSLLayer(n_classes, an, strides)
        elif mode == "kfiou":
            self.na, self.ch = 18, 5 + 1 + n_classes                        # model/yolo.py:28
            an = self._make_rotated_anchors(strides, anchors, angles)
            layer = YoloKFIoULayer(n_classes, an, strides)
        else:
            raise NotImplementedError("Loss mode : {} not found.".format(mode))
        self.anchors, self.nc, self.mode, self.ver = an, n_classes, mode, ver
        families = {'yolov4': (Backbonev4, Neckv4), 'yolov5': (Backbonev5, Neckv5), 'yolov7': (Backbonev7, Neckv7)}
        self.backbone = families[ver][0]()
        self.neck = families[ver][1](self.na * self.ch)
        self.yolo = layer
        self.autograd = True          # False: the caller drives Yolo.backward() itself (TrainStep)
        self._flat = self._flat_grad = self._grad_views = None
        self._bn_channels = sum(m.num_features for m in self.modules() if isinstance(m, nn.BatchNorm2d))
        self._bn_layers = sum(1 for m in self.modules() if isinstance(m, nn.BatchNorm2d))
        self.last_ctx = None

    def forward(self, i, training):
        L.require_cuda(i, "imgs")
        train = bool(training) and self.training
        if train and torch.is_grad_enabled() and self.autograd:
            heads = _StackFn.apply(self, i, *self.parameters())       # loss.backward() reaches run_backward
        else:
            heads = self._run(i, train)
        return self.yolo(list(heads), training)

    def _run(self, i, train):
        ctx = Ctx(self, train, i.device)
        with torch.no_grad():
            d3, d4, d5 = self.backbone(ctx, i)
            heads = self.neck(ctx, d5, d4, d3, self.na, self.ch)
        self.last_ctx = ctx
        return heads

    # ---- backward (native; torch autograd only sees _StackFn) -----------------------------------------------
    def backward(self, dlevels, param_grads=None):
        """d loss / d parameters from d loss / d levels (3 fp32 [B,na,gs,gs,ch] tensors), for the last
        training-mode forward.  Gradients are ADDED into `param_grads` (default: the flat gradient buffer
        behind every parameter's .grad, see flatten_parameters)."""
        from .backward import run_backward
        if param_grads is None:
            if self._flat_grad is None:
                self.flatten_parameters()
            param_grads = self._grad_views
        return run_backward(self, self.last_ctx, list(dlevels), param_grads)

    def flatten_parameters(self):
        """Re-homes all parameters into one flat fp32 buffer (and .grad into a second one) so that the optimizer
        step and the gradient all-reduce are single-buffer operations.  state_dict keys/values are unchanged."""
        params = list(self.parameters())
        dev = params[0].device
        sizes = [(p.numel() + 3) // 4 * 4 for p in params]
        total = sum(sizes)
        flat = torch.empty(total, dtype=torch.float32, device=dev)
        grad = torch.zeros(total, dtype=torch.float32, device=dev)
        off, views = 0, {}
        for p, n in zip(params, sizes):
            flat[off:off + p.numel()].copy_(p.data.reshape(-1))
            p.data = flat[off:off + p.numel()].view(p.shape)
            g = grad[off:off + p.numel()].view(p.shape)
            p.grad = g
            views[id(p)] = g
            off += n
        self._flat, self._flat_grad, self._grad_views = flat, grad, views
        return flat, grad

    def _implicit_head_grads(self, conv, gl, y, mul, param_grads):
        """yolov7 head y = im * (conv(x + ia) + b): gradients of ImplicitM / ImplicitA (model/utils.py:163-186)."""
        neck = self.neck
        i = {id(getattr(neck, f"conv{4 + j}")): j for j in (1, 2, 3)}[id(conv)]
        ia, im = getattr(neck, f"ia{i}"), getattr(neck, f"im{i}")
        B, na, H, W, ch = gl.shape
        d_im = (gl * y).sum((0, 2, 3)).reshape(-1) / mul                     # d/d im_c = sum dY * pre_c
        param_grads[id(im.implicit)].add_(d_im.view_as(im.implicit))
        dpre_sum = (gl.sum((0, 2, 3)).reshape(-1) * mul)                      # sum over pixels of d pre
        w = conv.conv[0].weight.data.float().flatten(1)                       # [Cout, Cin]
        param_grads[id(ia.implicit)].add_((w.t() @ dpre_sum).view_as(ia.implicit))

    def _repconv_backward(self, mod, x, rd, r1, dout, affs, G, sums, param_grads):
        """out = silu(bn_d(conv3x3(x)) + bn_1(conv1x1(x)))   (model/utils.py:209-215)."""
        from .. import ops as O
        C = mod.c2
        ds = O.Act.empty(rd.N, rd.H, rd.W, C, rd.buf.device)
        O.act_bwd2(dout, rd, affs[0][0], affs[0][1], r1, affs[1][0], affs[1][1], "swish", ds)
        first = True
        for raw, aff, seq, k, packed_t in ((rd, affs[0], mod.rbr_dense, mod.k, mod._pdt), (r1, affs[1], mod.rbr_1x1, 1, mod._p1t)):
            bn = seq[1]
            O.bn_act_bwd(ds, raw, aff[0], aff[1], aff[2], aff[3], "linear", sums[:2 * C] if first else sums[2 * C:],
                         raw, param_grads[id(bn.weight)], param_grads[id(bn.bias)])
            O.conv2d_wgrad(x, raw, C, k, mod.s, param_grads[id(seq[0].weight)])
            gx, acc = G.writable(x)
            O.conv2d_dgrad(raw, packed_t.get(seq[0].weight, transpose=True), x.C, k, mod.s, gx, acc)
            G.mark(x)
            first = False

    @staticmethod
    def _make_anchors(strides, anchors):
        return [[[a[i] / s, a[i + 1] / s] for i in range(0, len(a), 2)] for s, a in zip(strides, anchors)]

    @staticmethod
    def _make_rotated_anchors(strides, anchors, angles):
        return [[[a[i] / s, a[i + 1] / s, t] for i in range(0, len(a), 2) for t in angles]
                for s, a in zip(strides, anchors)]
