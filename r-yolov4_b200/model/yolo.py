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
        self._pack_table = self._pack_params = self._pack_sig = None
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

    def _pack_signature(self):
        """Cheap staleness key of the pinned bf16 operand copies: any torch-visible write to a conv weight
        (load_state_dict, .apply(weights_init_normal), an EMA swap-in, a re-homed .data) changes it.  Writers that go
        through raw pointers (TrainStep's optimizer kernels) repack themselves and bump WEIGHT_EPOCH."""
        from .blocks import WEIGHT_EPOCH
        return (sum(p._version for p in self._pack_params), WEIGHT_EPOCH[0],
                tuple(p.data_ptr() for p in self._pack_params[:4]))

    def _run(self, i, train):
        if getattr(self, "_pack_params", None) is not None and self._pack_signature() != self._pack_sig:
            self.repack_weights()
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
        off, views, self._flat_offsets = 0, {}, {}
        for p, n in zip(params, sizes):
            self._flat_offsets[id(p)] = (off, p.numel())
            flat[off:off + p.numel()].copy_(p.data.reshape(-1))
            p.data = flat[off:off + p.numel()].view(p.shape)
            g = grad[off:off + p.numel()].view(p.shape)
            p.grad = g
            views[id(p)] = g
            off += n
        self._flat, self._flat_grad, self._grad_views = flat, grad, views
        return flat, grad

    def enable_fused_pack(self):
        """Pre-allocates the bf16 operand copies of every conv weight and refreshes ALL of them with one kernel
        launch (repack_weights) instead of one launch per layer and layout."""
        import numpy as np_
        from .. import ops as ops_
        from .blocks import Conv, RepConv
        dev = next(self.parameters()).device
        ents, first = [], 0
        for m in self.modules():
            todo = []
            if isinstance(m, Conv):
                todo.append((m.conv[0].weight, m._packed, None if (m.stem or not m.has_bn) else m._packed_t, m.stem))
            elif isinstance(m, RepConv):
                todo.append((m.rbr_dense[0].weight, m._pd, m._pdt, False))
                todo.append((m.rbr_1x1[0].weight, m._p1, m._p1t, False))
            for w, pk, pkt, stem in todo:
                Cout, Cin, k, _ = w.shape
                pk.w = torch.zeros((Cout, ops_.stem_kpad(k) if stem else k * k * Cin), dtype=torch.bfloat16, device=dev)
                pk.pinned = True
                dt = 0
                if pkt is not None:
                    pkt.w = torch.empty((Cin, k * k * Cout), dtype=torch.bfloat16, device=dev)
                    pkt.pinned = True
                    dt = pkt.w.data_ptr()
                assert w.numel() % 8 == 0, "ryolo_pack_weights_multi packs 8 elements per thread"
                ents.append((w, pk.w.data_ptr(), dt, first, Cout, Cin, k, ops_.stem_kpad(k) if stem else 0))
                first += w.numel()
        self._pack_params = [e[0] for e in ents]
        self._pack_meta = [e[1:] for e in ents]
        self._pack_total = first
        self._pack_table = None
        self.repack_weights()

    def wgrad_scratch(self):
        """fp32 K-major scratch for every conv weight gradient (one flat buffer) + the table that lets ONE launch
        add all of them into the flat OIHW gradient buffer.  Needs flatten_parameters() + enable_fused_pack()."""
        import numpy as np_
        if getattr(self, "_wg_flat", None) is None or self._wg_grad_ptr != self._flat_grad.data_ptr():
            dev = self._flat_grad.device
            sizes = []
            for p, (_, _, _, Cout, Cin, k, stem) in zip(self._pack_params, self._pack_meta):
                sizes.append((Cout * (stem if stem else k * k * Cin) + 3) // 4 * 4)
            self._wg_flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
            rows = np_.zeros((len(sizes), 6), dtype=np_.int64)
            self._wg_views, off = {}, 0
            for i, (p, (_, _, first, Cout, Cin, k, stem), n) in enumerate(zip(self._pack_params, self._pack_meta, sizes)):
                v = self._wg_flat[off:off + Cout * (stem if stem else k * k * Cin)]
                self._wg_views[id(p)] = v
                rows[i, 0], rows[i, 1], rows[i, 3] = v.data_ptr(), self._grad_views[id(p)].data_ptr(), first
                rows[i, 4], rows[i, 5] = Cout | (Cin << 32), k | (stem << 32)
                off += n
            self._wg_table = torch.from_numpy(rows).to(dev)
            self._wg_grad_ptr = self._flat_grad.data_ptr()
        return self._wg_views

    def unpack_wgrads(self):
        from .. import _lib as L_
        L_.check(L_.lib().ryolo_unpack_wgrad_multi(L_.ptr(self._wg_table), len(self._pack_meta), self._pack_total,
                                                   L_.stream()))
        import os as _os
        self._wg_clean = L_.lib().ryolo_knob(12) != 0 and _os.environ.get("RYOLO_WG_CLEAR", "1") != "0"   # knob "ssa": the tiled fold clears what it folds
        L_.count(1)

    def wgrad_subtable(self, param_ids):
        """Table for ryolo_unpack_wgrad_multi restricted to the conv weights in `param_ids` (element offsets re-based):
        lets TrainStep fold one gradient bucket at a time.  Returns (device table, rows, elements) or None."""
        self.wgrad_scratch()
        rows = self._wg_table.cpu().numpy()
        keep = [i for i, p in enumerate(self._pack_params) if id(p) in param_ids]
        if not keep:
            return None
        sub = rows[keep].copy()
        first = 0
        for j, i in enumerate(keep):
            sub[j, 3] = first
            first += self._pack_params[i].numel()
        return torch.from_numpy(sub).to(self._flat_grad.device), len(keep), first

    def unpack_wgrads_sub(self, sub):
        from .. import _lib as L_
        table, n, total = sub
        L_.check(L_.lib().ryolo_unpack_wgrad_multi(L_.ptr(table), n, total, L_.stream()))
        L_.count(1)

    def repack_weights(self):
        import numpy as np_
        from .. import _lib as L_
        if self._pack_table is None or any(p.data_ptr() != a for p, a in zip(self._pack_params, self._pack_addrs)):
            rows = np_.zeros((len(self._pack_meta), 6), dtype=np_.int64)          # 48-byte ryolo_pack_entry
            for i, (p, (dst, dst_t, first, Cout, Cin, k, stem)) in enumerate(zip(self._pack_params, self._pack_meta)):
                rows[i, 0], rows[i, 1], rows[i, 2], rows[i, 3] = p.data_ptr(), dst, dst_t, first
                rows[i, 4] = Cout | (Cin << 32)
                rows[i, 5] = k | (stem << 32)
            self._pack_table = torch.from_numpy(rows).to(next(self.parameters()).device)
            self._pack_addrs = [p.data_ptr() for p in self._pack_params]
        L_.check(L_.lib().ryolo_pack_weights_multi(L_.ptr(self._pack_table), len(self._pack_meta), self._pack_total,
                                                   L_.stream()))
        L_.count(1)
        self._pack_sig = self._pack_signature()

    @staticmethod
    def _make_anchors(strides, anchors):
        return [[[a[i] / s, a[i + 1] / s] for i in range(0, len(a), 2)] for s, a in zip(strides, anchors)]

    @staticmethod
    def _make_rotated_anchors(strides, anchors, angles):
        return [[[a[i] / s, a[i + 1] / s, t] for i in range(0, len(a), 2) for t in angles]
                for s, a in zip(strides, anchors)]
