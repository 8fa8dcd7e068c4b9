"""Backbones on the tcgen05 conv stack: CSPDarknet53+SPP (model/backbone.py:4-36) and the yolov7
ELAN backbone (model/backbone.py:69-101).  Attribute names follow the reference so state_dict keys match."""
import torch.nn as nn

from .. import ops
from .blocks import C3, CSP, ELAN1, SPP, SPPCSPC, SPPF, Conv, MaxConv


class Backbonev4(nn.Module):
    def __init__(self):
        super().__init__()
        self.cbm0 = Conv(3, 32, 3, 1, "mish")
        chans = [(32, 64, 1), (64, 128, 2), (128, 256, 8), (256, 512, 8), (512, 1024, 4)]
        for i, (c1, c2, n) in enumerate(chans, start=1):
            setattr(self, f"cbm{i}", Conv(c1, c2, 3, 2, "mish"))
            setattr(self, f"csp{i}", CSP(c2, c2, n))
        self.spp = SPP(1024, 512)

    def forward(self, ctx, img):
        x = self.cbm0(ctx, ops.stem_im2col(img, 3, 1))
        feats = []
        for i in range(1, 6):
            x = getattr(self, f"csp{i}")(ctx, getattr(self, f"cbm{i}")(ctx, x))
            feats.append(x)
        return feats[2], feats[3], self.spp(ctx, feats[4])


class Backbonev7(nn.Module):
    def __init__(self):
        super().__init__()
        self.cbs0 = Conv(3, 32, 3, 1, "swish")
        self.cbs1 = Conv(32, 64, 3, 2, "swish")
        self.cbs2 = Conv(64, 64, 3, 1, "swish")
        self.cbs3 = Conv(64, 128, 3, 2, "swish")
        self.elan1 = ELAN1(128, 256)
        self.mc1 = MaxConv(256)
        self.elan2 = ELAN1(256, 512)
        self.mc2 = MaxConv(512)
        self.elan3 = ELAN1(512, 1024)
        self.mc3 = MaxConv(1024)
        self.elan4 = ELAN1(1024, 1024, e1=0.25, e2=0.25)
        self.spp = SPPCSPC(1024, 512)

    def forward(self, ctx, img):
        x = self.cbs0(ctx, ops.stem_im2col(img, 3, 1))
        x = self.cbs3(ctx, self.cbs2(ctx, self.cbs1(ctx, x)))
        x = self.elan1(ctx, x)
        d3 = self.elan2(ctx, self.mc1(ctx, x))
        d4 = self.elan3(ctx, self.mc2(ctx, d3))
        d5 = self.elan4(ctx, self.mc3(ctx, d4))
        return d3, d4, self.spp(ctx, d5)


class Backbonev5(nn.Module):
    """yolov5 backbone (model/backbone.py:39-66): 6x6/s2 stem, C3 stages, SPPF."""

    def __init__(self):
        super().__init__()
        self.cbs0 = Conv(3, 64, 6, 2, "swish")
        self.cbs1 = Conv(64, 128, 3, 2, "swish")
        self.csp1 = C3(128, 128, 3)
        self.cbs2 = Conv(128, 256, 3, 2, "swish")
        self.csp2 = C3(256, 256, 6)
        self.cbs3 = Conv(256, 512, 3, 2, "swish")
        self.csp3 = C3(512, 512, 9)
        self.cbs4 = Conv(512, 1024, 3, 2, "swish")
        self.csp4 = C3(1024, 1024, 3)
        self.spp = SPPF(1024, 1024)

    def forward(self, ctx, img):
        x = self.cbs0(ctx, ops.stem_im2col(img, 6, 2))
        x = self.csp1(ctx, self.cbs1(ctx, x))
        d3 = self.csp2(ctx, self.cbs2(ctx, x))
        d4 = self.csp3(ctx, self.cbs3(ctx, d3))
        d5 = self.csp4(ctx, self.cbs4(ctx, d4))
        return d3, d4, self.spp(ctx, d5)
