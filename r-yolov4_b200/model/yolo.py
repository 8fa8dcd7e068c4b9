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
        self._bn_channels = sum(m.num_features for m in self.modules() if isinstance(m, nn.BatchNorm2d))
        self._bn_layers = sum(1 for m in self.modules() if isinstance(m, nn.BatchNorm2d))
        self.last_ctx = None

    def forward(self, i, training):
        L.require_cuda(i, "imgs")
        ctx = Ctx(self, bool(training) and self.training, i.device)
        with torch.no_grad():
            d3, d4, d5 = self.backbone(ctx, i)
            heads = self.neck(ctx, d5, d4, d3, self.na, self.ch)
        self.last_ctx = ctx
        return self.yolo(list(heads), training)

    @staticmethod
    def _make_anchors(strides, anchors):
        return [[[a[i] / s, a[i + 1] / s] for i in range(0, len(a), 2)] for s, a in zip(strides, anchors)]

    @staticmethod
    def _make_rotated_anchors(strides, anchors, angles):
        return [[[a[i] / s, a[i + 1] / s, t] for i in range(0, len(a), 2) for t in angles]
                for s, a in zip(strides, anchors)]
