"""Necks + heads: PANet (model/neck.py:4-81) and the yolov7 ELAN/RepConv/Implicit neck (:150-217).
The three head convs write fp32 straight into the [B, na, gs, gs, ch] layout the loss/decode kernels
read (the view/permute/contiguous of model/yololayer.py:25,76 never runs)."""
import torch
import torch.nn as nn

from .. import ops
from .blocks import C3, C5, ELAN2, Conv, ImplicitA, ImplicitM, MaxConv, RepConv


def _copy_into(ctx, src, dst, factor=1):
    ops.resize_copy(src, factor, out=dst)
    if ctx.tape is not None:
        ctx.tape.append(("resize", src, dst, factor))


class Neckv4(nn.Module):
    def __init__(self, output_ch):
        super().__init__()
        self.conv7 = Conv(512, 256, 1, 1, 'leaky')
        self.up1 = nn.Upsample(scale_factor=2)
        self.conv8 = Conv(512, 256, 1, 1, 'leaky')
        self.conv9 = C5(512, 256)
        self.conv14 = Conv(256, 128, 1, 1, 'leaky')
        self.up2 = nn.Upsample(scale_factor=2)
        self.conv15 = Conv(256, 128, 1, 1, 'leaky')
        self.conv16 = C5(256, 128)
        self.conv21 = Conv(128, 256, 3, 1, 'leaky')
        self.conv22 = Conv(256, output_ch, 1, 1, 'linear', bn=False, bias=True)
        self.conv23 = Conv(128, 256, 3, 2, 'leaky')
        self.conv24 = C5(512, 256)
        self.conv29 = Conv(256, 512, 3, 1, 'leaky')
        self.conv30 = Conv(512, output_ch, 1, 1, 'linear', bn=False, bias=True)
        self.conv31 = Conv(256, 512, 3, 2, 'leaky')
        self.conv32 = C5(1024, 512)
        self.conv37 = Conv(512, 1024, 3, 1, 'leaky')
        self.conv38 = Conv(1024, output_ch, 1, 1, 'linear', bn=False, bias=True)

    def _head(self, ctx, pre, head, x, na, ch):
        return head(ctx, pre(ctx, x), head=(na, ch), head_shift=head.conv[0].bias.data)

    def forward(self, ctx, x1, x2, x3, na, ch):
        """x1 = d5 (stride 32), x2 = d4, x3 = d3.  Returns the heads for strides 8, 16, 32."""
        N = x1.N
        cat16 = ctx.new(N, x2.H, x2.W, 512)                       # [conv8(d4) | up(conv7(d5))]
        self.conv8(ctx, x2, out=cat16.slice(0, 256))
        _copy_into(ctx, self.conv7(ctx, x1), cat16.slice(256, 256), 2)
        pan16 = ctx.new(N, x2.H, x2.W, 512)                       # [conv23(p8) | p16]  (filled later / now)
        p16 = self.conv9(ctx, cat16, out=pan16.slice(256, 256))
        cat8 = ctx.new(N, x3.H, x3.W, 256)                        # [conv15(d3) | up(conv14(p16))]
        self.conv15(ctx, x3, out=cat8.slice(0, 128))
        _copy_into(ctx, self.conv14(ctx, p16), cat8.slice(128, 128), 2)
        p8 = self.conv16(ctx, cat8)
        h8 = self._head(ctx, self.conv21, self.conv22, p8, na, ch)
        self.conv23(ctx, p8, out=pan16.slice(0, 256))
        pan32 = ctx.new(N, x1.H, x1.W, 1024)                      # [conv31(n16) | d5]
        n16 = self.conv24(ctx, pan16)
        h16 = self._head(ctx, self.conv29, self.conv30, n16, na, ch)
        self.conv31(ctx, n16, out=pan32.slice(0, 512))
        _copy_into(ctx, x1, pan32.slice(512, 512))
        n32 = self.conv32(ctx, pan32)
        h32 = self._head(ctx, self.conv37, self.conv38, n32, na, ch)
        return h8, h16, h32


class Neckv7(nn.Module):
    def __init__(self, output_ch):
        super().__init__()
        self.conv1 = Conv(512, 256, 1, 1, 'swish')
        self.up1 = nn.Upsample(scale_factor=2, mode='nearest')
        self.elan1 = ELAN2(512, 256)
        self.conv2 = Conv(256, 128, 1, 1, 'swish')
        self.up2 = nn.Upsample(scale_factor=2, mode='nearest')
        self.elan2 = ELAN2(256, 128)
        self.conv3 = Conv(1024, 256, 1, 1, 'swish')
        self.conv4 = Conv(512, 128, 1, 1, 'swish')
        self.mc1 = MaxConv(128, e=1.0)
        self.elan3 = ELAN2(512, 256)
        self.mc2 = MaxConv(256, e=1.0)
        self.elan4 = ELAN2(1024, 512)
        for i, (c1, c2) in enumerate(((128, 256), (256, 512), (512, 1024)), start=1):
            setattr(self, f"repVgg{i}", RepConv(c1, c2))
            setattr(self, f"ia{i}", ImplicitA(c2))
            setattr(self, f"conv{4 + i}", Conv(c2, output_ch, 1, 1, 'linear', bn=False, bias=True))
            setattr(self, f"im{i}", ImplicitM(output_ch))

    def _head(self, ctx, i, x, na, ch):
        """im * (conv(x + ia) + b)  ==  conv(x) * im + (W.ia + b) * im   (model/neck.py:201,208,215)."""
        rep, ia, conv, im = (getattr(self, f"{n}{j}") for n, j in (("repVgg", i), ("ia", i), ("conv", 4 + i), ("im", i)))
        y = rep(ctx, x)
        w = conv.conv[0].weight.data.float().flatten(1)
        m = im.implicit.data.float().flatten()
        shift = (w @ ia.implicit.data.float().flatten() + conv.conv[0].bias.data.float()) * m
        return conv(ctx, y, head=(na, ch), head_scale=m.contiguous(), head_shift=shift.contiguous())

    def forward(self, ctx, x1, x2, x3, na, ch):
        N = x1.N
        cat16 = ctx.new(N, x2.H, x2.W, 512)                       # [conv3(d4) | up(conv1(d5))]
        self.conv3(ctx, x2, out=cat16.slice(0, 256))
        _copy_into(ctx, self.conv1(ctx, x1), cat16.slice(256, 256), 2)
        pan16 = ctx.new(N, x2.H, x2.W, 512)                       # [p16 | mc1(p8)]
        p16 = self.elan1(ctx, cat16, out=pan16.slice(0, 256))
        cat8 = ctx.new(N, x3.H, x3.W, 256)                        # [conv4(d3) | up(conv2(p16))]
        self.conv4(ctx, x3, out=cat8.slice(0, 128))
        _copy_into(ctx, self.conv2(ctx, p16), cat8.slice(128, 128), 2)
        p8 = self.elan2(ctx, cat8)
        h8 = self._head(ctx, 1, p8, na, ch)
        self.mc1(ctx, p8, out=pan16.slice(256, 256))
        n16 = self.elan3(ctx, pan16)
        h16 = self._head(ctx, 2, n16, na, ch)
        pan32 = ctx.new(N, x1.H, x1.W, 1024)                      # [d5 | mc2(n16)]
        _copy_into(ctx, x1, pan32.slice(0, 512))
        self.mc2(ctx, n16, out=pan32.slice(512, 512))
        n32 = self.elan4(ctx, pan32)
        h32 = self._head(ctx, 3, n32, na, ch)
        return h8, h16, h32


class Neckv5(nn.Module):
    """yolov5 PAN neck (model/neck.py:84-147)."""

    def __init__(self, output_ch):
        super().__init__()
        self.conv7 = Conv(1024, 512, 1, 1, 'swish')
        self.up1 = nn.Upsample(scale_factor=2, mode='nearest')
        self.csp1 = C3(1024, 512, 3, shortcut=False)
        self.conv14 = Conv(512, 256, 1, 1, 'swish')
        self.up2 = nn.Upsample(scale_factor=2, mode='nearest')
        self.csp2 = C3(512, 256, 3, shortcut=False)
        self.conv15 = Conv(256, output_ch, 1, 1, 'linear', bn=False, bias=True)
        self.conv16 = Conv(256, 256, 3, 2, 'swish')
        self.csp3 = C3(512, 512, 3, shortcut=False)
        self.conv17 = Conv(512, output_ch, 1, 1, 'linear', bn=False, bias=True)
        self.conv18 = Conv(512, 512, 3, 2, 'swish')
        self.csp4 = C3(1024, 1024, 3, shortcut=False)
        self.conv19 = Conv(1024, output_ch, 1, 1, 'linear', bn=False, bias=True)

    @staticmethod
    def _head(ctx, head, x, na, ch):
        return head(ctx, x, head=(na, ch), head_shift=head.conv[0].bias.data)

    def forward(self, ctx, x1, x2, x3, na, ch):
        N = x1.N
        pan32 = ctx.new(N, x1.H, x1.W, 1024)                      # [conv7(d5) | conv18(n16)]
        t32 = self.conv7(ctx, x1, out=pan32.slice(0, 512))
        cat16 = ctx.new(N, x2.H, x2.W, 1024)                      # [d4 | up(t32)]
        _copy_into(ctx, x2, cat16.slice(0, 512))
        _copy_into(ctx, t32, cat16.slice(512, 512), 2)
        pan16 = ctx.new(N, x2.H, x2.W, 512)                       # [conv14(csp1) | conv16(p8)]
        t16 = self.conv14(ctx, self.csp1(ctx, cat16), out=pan16.slice(0, 256))
        cat8 = ctx.new(N, x3.H, x3.W, 512)                        # [d3 | up(t16)]
        _copy_into(ctx, x3, cat8.slice(0, 256))
        _copy_into(ctx, t16, cat8.slice(256, 256), 2)
        p8 = self.csp2(ctx, cat8)
        h8 = self._head(ctx, self.conv15, p8, na, ch)
        self.conv16(ctx, p8, out=pan16.slice(256, 256))
        n16 = self.csp3(ctx, pan16)
        h16 = self._head(ctx, self.conv17, n16, na, ch)
        self.conv18(ctx, n16, out=pan32.slice(512, 512))
        n32 = self.csp4(ctx, pan32)
        h32 = self._head(ctx, self.conv19, n32, na, ch)
        return h8, h16, h32
