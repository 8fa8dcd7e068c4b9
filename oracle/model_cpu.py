"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by r-yolov4_b200/).

Functional torch-CPU fp32 restatement of the reference conv stack (model/utils.py, backbone.py,
neck.py, yolo.py), driven directly by a reference-format ``state_dict`` (keys such as
``backbone.cbm0.conv.0.weight``).  It is the parity checker for the tcgen05 conv stack and the
CPU arm (`--impl reference`, kind "port") of bench.py.  PINNED against tests/golden/model_*.pt.
"""
import torch
import torch.nn.functional as F

from . import hotpath as hp


class _RoundBoth(torch.autograd.Function):
    """bf16 storage point of the B200 stack in BOTH directions: the activation is stored in bf16 on the way forward
    and its gradient (the dgrad / BN-backward output that lands in the mirrored gradient buffer) on the way back."""

    @staticmethod
    def forward(ctx, t):
        return t.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


class Net:
    """Evaluates the reference graph from a flat state_dict. train=True uses batch statistics
    (model/utils.py:16-17 BatchNorm2d in train mode) and records updated running stats."""

    def __init__(self, sd, train, eps=1e-5, momentum=0.1, emulate_bf16=False):
        """emulate_bf16: round conv weights, raw conv outputs and block outputs to bf16 (straight-through
        gradient), i.e. the storage points of the B200 stack, so that discontinuous derivatives (LeakyReLU's
        kink, max-pool arg-max) are decided on the same numbers.  Arithmetic stays fp32 like the product's.
        emulate_bf16="grad" also rounds the GRADIENTS arriving at those storage points (the product keeps d out and
        d raw in bf16 buffers): the yardstick for sums that cancel to ~0 in exact arithmetic (BatchNorm biases in front
        of another conv + BatchNorm), where that rounding noise is the whole signal."""
        self.sd, self.train, self.eps, self.mom = sd, train, eps, momentum
        if emulate_bf16 == "grad":
            self.q = _RoundBoth.apply
        else:
            self.q = (lambda t: t + (t.bfloat16().float() - t).detach()) if emulate_bf16 else (lambda t: t)
        self.new_stats = {}
        self.trace = None      # optional list of (module path, input, output) per Conv / RepConv

    # -- primitives -------------------------------------------------------------------------
    def bn(self, x, pre):
        """nn.BatchNorm2d forward (model/utils.py:16-17) through the same ATen op the reference runs."""
        g, b = self.sd[pre + ".weight"], self.sd[pre + ".bias"]
        rm, rv = self.sd[pre + ".running_mean"].clone(), self.sd[pre + ".running_var"].clone()
        y = F.batch_norm(x, rm, rv, g, b, self.train, self.mom, self.eps)
        if self.train:
            self.new_stats[pre + ".running_mean"], self.new_stats[pre + ".running_var"] = rm, rv
        return y

    def conv(self, x, pre, act, s=1):
        """model/utils.py:6-32 Conv = conv(+BN)(+act). `pre` is the module path of the Conv."""
        w = self.sd[pre + ".conv.0.weight"]
        bias = self.sd.get(pre + ".conv.0.bias")
        y = self.q(F.conv2d(x, self.q(w), bias, stride=s, padding=(w.shape[2] - 1) // 2))
        if (pre + ".conv.1.weight") in self.sd:
            y = self.bn(y, pre + ".conv.1")
        y = self.act(y, act)
        if getattr(self, "_defer_round", False) is False:
            y = self.q(y)
        if self.trace is not None:
            self.trace.append((pre, x, y))
        return y

    @staticmethod
    def act(x, a):
        if a == "mish":
            return F.mish(x)
        if a == "leaky":
            return F.leaky_relu(x, 0.1)
        if a == "swish":
            return F.silu(x)
        return x

    # -- composites (model/utils.py) -----------------------------------------------------------
    def bottleneck(self, x, pre, act, add):
        y = self.conv(x, pre + ".cv1", act)
        if not add:
            return self.conv(y, pre + ".cv2", act)
        self._defer_round = True               # the product rounds once, after the residual add
        y = self.conv(y, pre + ".cv2", act)
        self._defer_round = False
        return self.q(x + y)

    def csp(self, x, pre, n):  # utils.py:49-64
        y = self.conv(x, pre + ".cv1", "mish")
        for i in range(n):
            y = self.bottleneck(y, f"{pre}.m.{i}", "mish", True)
        y1 = self.conv(y, pre + ".cv3", "mish")
        y2 = self.conv(x, pre + ".cv2", "mish")
        return self.conv(torch.cat((y1, y2), 1), pre + ".cv4", "mish")

    def c5(self, x, pre):  # utils.py:67-80
        for i in range(1, 6):
            x = self.conv(x, f"{pre}.cv{i}", "leaky")
        return x

    def c3(self, x, pre, n, shortcut):  # utils.py:83-95
        y = self.conv(x, pre + ".cv1", "swish")
        for i in range(n):
            y = self.bottleneck(y, f"{pre}.m.{i}", "swish", shortcut)
        return self.conv(torch.cat((y, self.conv(x, pre + ".cv2", "swish")), 1), pre + ".cv3", "swish")

    def spp(self, x, pre):  # utils.py:218-244
        x = self.conv(self.conv(self.conv(x, pre + ".cv1", "leaky"), pre + ".cv2", "leaky"), pre + ".cv3", "leaky")
        x = torch.cat([F.max_pool2d(x, 13, 1, 6), F.max_pool2d(x, 9, 1, 4), F.max_pool2d(x, 5, 1, 2), x], 1)
        return self.conv(self.conv(self.conv(x, pre + ".cv4", "leaky"), pre + ".cv5", "leaky"), pre + ".cv6", "leaky")

    def sppf(self, x, pre):  # utils.py:247-261
        x = self.conv(x, pre + ".cv1", "swish")
        y1 = F.max_pool2d(x, 5, 1, 2)
        y2 = F.max_pool2d(y1, 5, 1, 2)
        return self.conv(torch.cat([x, y1, y2, F.max_pool2d(y2, 5, 1, 2)], 1), pre + ".cv2", "swish")

    def sppcspc(self, x, pre):  # utils.py:264-282
        x1 = self.conv(self.conv(self.conv(x, pre + ".cv1", "swish"), pre + ".cv3", "swish"), pre + ".cv4", "swish")
        cat = torch.cat([x1] + [F.max_pool2d(x1, k, 1, k // 2) for k in (5, 9, 13)], 1)
        y1 = self.conv(self.conv(cat, pre + ".cv5", "swish"), pre + ".cv6", "swish")
        y2 = self.conv(x, pre + ".cv2", "swish")
        return self.conv(torch.cat((y1, y2), 1), pre + ".cv7", "swish")

    def elan1(self, x, pre):  # utils.py:98-118
        x1 = self.conv(x, pre + ".cv1", "swish")
        x2 = self.conv(x, pre + ".cv2", "swish")
        x3 = self.conv(self.conv(x2, pre + ".cv3", "swish"), pre + ".cv4", "swish")
        x4 = self.conv(self.conv(x3, pre + ".cv5", "swish"), pre + ".cv6", "swish")
        return self.conv(torch.cat((x1, x2, x3, x4), 1), pre + ".cv7", "swish")

    def elan2(self, x, pre):  # utils.py:121-143
        x1 = self.conv(x, pre + ".cv1", "swish")
        x2 = self.conv(x, pre + ".cv2", "swish")
        x3 = self.conv(x2, pre + ".cv3", "swish")
        x4 = self.conv(x3, pre + ".cv4", "swish")
        x5 = self.conv(x4, pre + ".cv5", "swish")
        x6 = self.conv(x5, pre + ".cv6", "swish")
        return self.conv(torch.cat((x1, x2, x3, x4, x5, x6), 1), pre + ".cv7", "swish")

    def maxconv(self, x, pre):  # utils.py:146-160
        x1 = self.conv(F.max_pool2d(x, 2, 2), pre + ".cv1", "swish")
        x2 = self.conv(self.conv(x, pre + ".cv2", "swish"), pre + ".cv3", "swish", s=2)
        return torch.cat((x1, x2), 1)

    def repconv(self, x, pre):  # utils.py:189-215
        d = self.bn(self.q(F.conv2d(x, self.q(self.sd[pre + ".rbr_dense.0.weight"]), None, 1, 1)), pre + ".rbr_dense.1")
        o = self.bn(self.q(F.conv2d(x, self.q(self.sd[pre + ".rbr_1x1.0.weight"]), None, 1, 0)), pre + ".rbr_1x1.1")
        y = d + o
        if (pre + ".rbr_identity.weight") in self.sd:
            y = y + self.bn(x, pre + ".rbr_identity")
        y = self.q(F.silu(y))
        if self.trace is not None:
            self.trace.append((pre, x, y))
        return y

    # -- backbones / necks (model/backbone.py, model/neck.py) -------------------------------------
    def backbone_v4(self, x):  # backbone.py:4-36
        p = "backbone."
        x = self.conv(x, p + "cbm0", "mish")
        x = self.csp(self.conv(x, p + "cbm1", "mish", 2), p + "csp1", 1)
        x = self.csp(self.conv(x, p + "cbm2", "mish", 2), p + "csp2", 2)
        d3 = self.csp(self.conv(x, p + "cbm3", "mish", 2), p + "csp3", 8)
        d4 = self.csp(self.conv(d3, p + "cbm4", "mish", 2), p + "csp4", 8)
        d5 = self.csp(self.conv(d4, p + "cbm5", "mish", 2), p + "csp5", 4)
        return d3, d4, self.spp(d5, p + "spp")

    def neck_v4(self, x1, x2, x3):  # neck.py:47-81
        p = "neck."
        up = lambda t: F.interpolate(t, scale_factor=2.0, mode="nearest")
        x2 = torch.cat([self.conv(x2, p + "conv8", "leaky"), up(self.conv(x1, p + "conv7", "leaky"))], 1)
        x2 = self.c5(x2, p + "conv9")
        x3 = torch.cat([self.conv(x3, p + "conv15", "leaky"), up(self.conv(x2, p + "conv14", "leaky"))], 1)
        x3 = self.c5(x3, p + "conv16")
        x6 = self.conv(self.conv(x3, p + "conv21", "leaky"), p + "conv22", "linear")
        x2 = self.c5(torch.cat([self.conv(x3, p + "conv23", "leaky", 2), x2], 1), p + "conv24")
        x5 = self.conv(self.conv(x2, p + "conv29", "leaky"), p + "conv30", "linear")
        x1 = self.c5(torch.cat([self.conv(x2, p + "conv31", "leaky", 2), x1], 1), p + "conv32")
        x4 = self.conv(self.conv(x1, p + "conv37", "leaky"), p + "conv38", "linear")
        return x6, x5, x4

    def backbone_v5(self, x):  # backbone.py:39-66
        p = "backbone."
        x = self.conv(x, p + "cbs0", "swish", 2)
        x = self.c3(self.conv(x, p + "cbs1", "swish", 2), p + "csp1", 3, True)
        d3 = self.c3(self.conv(x, p + "cbs2", "swish", 2), p + "csp2", 6, True)
        d4 = self.c3(self.conv(d3, p + "cbs3", "swish", 2), p + "csp3", 9, True)
        d5 = self.c3(self.conv(d4, p + "cbs4", "swish", 2), p + "csp4", 3, True)
        return d3, d4, self.sppf(d5, p + "spp")

    def neck_v5(self, x1, x2, x3):  # neck.py:110-147
        p = "neck."
        up = lambda t: F.interpolate(t, scale_factor=2.0, mode="nearest")
        x1 = self.conv(x1, p + "conv7", "swish")
        x2 = self.c3(torch.cat([x2, up(x1)], 1), p + "csp1", 3, False)
        x2 = self.conv(x2, p + "conv14", "swish")
        x3 = self.c3(torch.cat([x3, up(x2)], 1), p + "csp2", 3, False)
        x6 = self.conv(x3, p + "conv15", "linear")
        x2 = self.c3(torch.cat([x2, self.conv(x3, p + "conv16", "swish", 2)], 1), p + "csp3", 3, False)
        x5 = self.conv(x2, p + "conv17", "linear")
        x1 = self.c3(torch.cat([x1, self.conv(x2, p + "conv18", "swish", 2)], 1), p + "csp4", 3, False)
        x4 = self.conv(x1, p + "conv19", "linear")
        return x6, x5, x4

    def backbone_v7(self, x):  # backbone.py:69-101
        p = "backbone."
        x = self.conv(self.conv(self.conv(x, p + "cbs0", "swish"), p + "cbs1", "swish", 2), p + "cbs2", "swish")
        x = self.elan1(self.conv(x, p + "cbs3", "swish", 2), p + "elan1")
        d3 = self.elan1(self.maxconv(x, p + "mc1"), p + "elan2")
        d4 = self.elan1(self.maxconv(d3, p + "mc2"), p + "elan3")
        d5 = self.elan1(self.maxconv(d4, p + "mc3"), p + "elan4")
        return d3, d4, self.sppcspc(d5, p + "spp")

    def head_v7(self, x, i):  # neck.py:201,208,215
        p = "neck."
        y = self.repconv(x, f"{p}repVgg{i}") + self.sd[f"{p}ia{i}.implicit"]
        return self.conv(y, f"{p}conv{4 + i}", "linear") * self.sd[f"{p}im{i}.implicit"]

    def neck_v7(self, x1, x2, x3):  # neck.py:188-217
        p = "neck."
        up = lambda t: F.interpolate(t, scale_factor=2.0, mode="nearest")
        x2 = torch.cat([self.conv(x2, p + "conv3", "swish"), up(self.conv(x1, p + "conv1", "swish"))], 1)
        x2 = self.elan2(x2, p + "elan1")
        x3 = torch.cat([self.conv(x3, p + "conv4", "swish"), up(self.conv(x2, p + "conv2", "swish"))], 1)
        x3 = self.elan2(x3, p + "elan2")
        x6 = self.head_v7(x3, 1)
        x2 = self.elan2(torch.cat([x2, self.maxconv(x3, p + "mc1")], 1), p + "elan3")
        x5 = self.head_v7(x2, 2)
        x1 = self.elan2(torch.cat([x1, self.maxconv(x2, p + "mc2")], 1), p + "elan4")
        x4 = self.head_v7(x1, 3)
        return x6, x5, x4


def forward(sd, img, ver, mode, nc, train, anchors_cfg=None, angles=None, decode=True, trace=None):
    """Yolo.forward (model/yolo.py:46-51).  Returns (levels, infer|None, new_running_stats)."""
    net = Net(sd, train)
    net.trace = trace
    d3, d4, d5 = getattr(net, "backbone_" + ver[-2:])(img)
    heads = getattr(net, "neck_" + ver[-2:])(d5, d4, d3)
    na, ch = (3, nc + 185) if mode == "csl" else (18, nc + 6)
    levels = [hp.head_to_grid(h, na, ch) for h in heads]
    infer = None
    if (not train) and decode:
        if mode == "csl":
            infer = hp.decode_csl(levels, hp.make_anchors(anchors_cfg), nc)
        else:
            infer = hp.decode_kfiou(levels, hp.make_rotated_anchors(anchors_cfg, angles), nc)
    return levels, infer, net.new_stats
