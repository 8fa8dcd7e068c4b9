"""Rotated-box decode layers, drop-in for the reference's model/yololayer.py.

  YoloCSLLayer.forward     model/yololayer.py:15-56
  YoloKFIoULayer.forward   model/yololayer.py:66-105
forward(out, training) takes the three NCHW head tensors, rewrites `out[i]` in place to the
[B, na, gs, gs, ch] layout (like the reference, :25/:76) and, when training is False, also returns
the fused [B, R, nc+6] prediction tensor produced by the decode kernels (csrc/decode.cu).
"""
import ctypes

import numpy as np
import torch
import torch.nn as nn

from .. import _lib as L


def _to_grid(x, na, ch):
    if x.dim() == 5:                      # already [B, na, gs, gs, ch] (written directly by the head conv)
        return x
    b, _, gh, gw = x.shape
    return x.view(b, na, ch, gh, gw).permute(0, 1, 3, 4, 2).contiguous()


class _YoloLayer(nn.Module):
    def __init__(self, num_classes, anchors, stride):
        super().__init__()
        self.num_classes, self.anchors, self.stride = num_classes, anchors, stride
        self.na = len(anchors[0])
        self._dev_anchors = {}

    def _rows(self, out):
        return sum(self.na * p.shape[2] * p.shape[3] for p in out)


class YoloCSLLayer(_YoloLayer):
    def forward(self, out, training):
        nc = self.num_classes
        for i in range(len(out)):
            out[i] = _to_grid(out[i], self.na, nc + 185)
        if training:
            return out
        L.require_cuda(out[0], "head outputs")
        lib = L.lib()
        B, R = out[0].shape[0], self._rows(out)
        infer = torch.empty((B, R, nc + 6), dtype=torch.float32, device=out[0].device)
        row0 = 0
        for p, anc, s in zip(out, self.anchors, self.stride):
            gs = p.shape[2]
            awh = (ctypes.c_float * 6)(*[float(v) for a in anc for v in a[:2]])
            L.check(lib.ryolo_decode_csl(L.ptr(p.detach()), B, gs, nc, float(s), awh, L.ptr(infer), row0, R,
                                         L.stream()))
            L.count(1)
            row0 += self.na * gs * gs
        return out, infer


class YoloKFIoULayer(_YoloLayer):
    def _anchors_on(self, device):
        key = str(device)
        if key not in self._dev_anchors:
            self._dev_anchors[key] = [torch.tensor(np.asarray(a, dtype=np.float32), device=device).contiguous()
                                      for a in self.anchors]
        return self._dev_anchors[key]

    def forward(self, out, training):
        nc = self.num_classes
        for i in range(len(out)):
            out[i] = _to_grid(out[i], self.na, nc + 6)
        if training:
            return out
        L.require_cuda(out[0], "head outputs")
        lib = L.lib()
        B, R = out[0].shape[0], self._rows(out)
        infer = torch.empty((B, R, nc + 6), dtype=torch.float32, device=out[0].device)
        row0 = 0
        for p, anc, s in zip(out, self._anchors_on(out[0].device), self.stride):
            gs = p.shape[2]
            L.check(lib.ryolo_decode_kfiou(L.ptr(p.detach()), B, self.na, gs, nc, float(s), L.ptr(anc), L.ptr(infer),
                                           row0, R, L.stream()))
            L.count(1)
            row0 += self.na * gs * gs
        return out, infer
