"""Conv-stack building blocks on NHWC bf16 views, executed by the tcgen05 conv kernel and its
HBM-bound companions.  The nn.Module tree only *holds* parameters under the reference's names
(model/utils.py) so that state_dict keys/order, .apply(weights_init_normal), .train()/.eval() and
optimizers keep working; the arithmetic never goes through torch.nn.

Block semantics follow model/utils.py: Conv :6-32, Bottleneck :35-46, CSP :49-64, C5 :67-80,
ELAN1 :98-118, ELAN2 :121-143, MaxConv :146-160, ImplicitA/M :163-186, RepConv :189-215,
SPP :218-244, SPPCSPC :264-282.  torch.cat is replaced by writing producers into channel slices of
one buffer.
"""
import os

import torch
import torch.nn as nn

from .. import ops
from ..ops import Act

_ACTS = {"mish": nn.Mish, "leaky": lambda: nn.LeakyReLU(0.1, inplace=True), "swish": nn.SiLU}


class Ctx:
    """Per-forward execution context: mode, BN statistic scratch, and (in training) the tape."""

    def __init__(self, model, training, device):
        self.training = training
        self.device = device
        self.model = model
        self.tape = [] if training else None
        # one zeroed scratch for the cross-CTA BN accumulators (every launch leaves it zeroed again; layers run back
        # to back on one stream) + a zeroed counter each
        self._partial = torch.zeros(4 * 2048, dtype=torch.float32, device=device) if training else None
        self._counters = torch.zeros(model._bn_layers, dtype=torch.int32, device=device) if training else None
        self._ctr_off = 0
        # deferred finalize (plain Conv blocks): one zeroed arena, a private 4*C-float accumulator per layer
        self._acc = torch.zeros(4 * model._bn_channels, dtype=torch.float32, device=device) \
            if training and DEFER_BN_FINALIZE else None
        self._acc_off = 0
        if training:
            BN_EPOCH[0] += 1      # the fused conv epilogue rewrites running statistics through raw pointers

    def new(self, N, H, W, C):
        return Act.empty(N, H, W, C, self.device)

    def acc_slot(self, C):
        a = self._acc[self._acc_off:self._acc_off + 4 * C]
        self._acc_off += 4 * C
        return a

    def stat_slot(self, C):
        ctr = self._counters[self._ctr_off:self._ctr_off + 1]
        self._ctr_off += 1
        return self._partial, ctr


# The conv kernel only accumulates its BatchNorm totals and scale_shift_act_bn finalizes them (no fence / arrival counter /
# last-CTA pass in every conv launch's tail).  RYOLO_BN_DEFER=0 restores the in-conv finalize (A/B switch).
DEFER_BN_FINALIZE = os.environ.get("RYOLO_BN_DEFER", "1") != "0"

WEIGHT_EPOCH = [0]     # bumped by whoever rewrites parameters through raw pointers (TrainStep's SGD kernel)
BN_EPOCH = [0]         # bumped by every training-mode forward: running statistics change without a torch version bump


class _Packed:
    """bf16 [Cout][kh][kw][Cin] copy of an fp32 OIHW parameter, refreshed when the parameter changes."""

    def __init__(self):
        self.key, self.w, self.pinned = None, None, False

    def get(self, p, stem=False, transpose=False):
        if self.pinned:                       # refreshed in bulk by Yolo.repack_weights()
            return self.w
        key = (p.data_ptr(), p._version, p.device, WEIGHT_EPOCH[0])
        if key != self.key:
            self.w = ops.pack_weights(p.data, stem=stem, transpose=transpose)
            self.key = key
        return self.w


def _bn_eval_affine(bn, cache):
    key = (bn.weight._version, bn.bias._version, bn.running_mean._version, bn.running_var._version,
           bn.weight.data_ptr(), WEIGHT_EPOCH[0], BN_EPOCH[0], bn.num_batches_tracked.data_ptr())
    if cache.get("key") != key:
        scale = bn.weight.data.float() * torch.rsqrt(bn.running_var.float() + bn.eps)
        cache["scale"], cache["shift"] = scale.contiguous(), (bn.bias.data.float() - bn.running_mean.float() * scale)
        cache["key"] = key
    return cache["scale"], cache["shift"]


class Conv(nn.Module):
    """Conv2d(bias=False) + BatchNorm2d + activation, or the biased linear head conv (bn=False)."""

    def __init__(self, c1, c2, k, s, activation, bn=True, bias=False):
        super().__init__()
        if activation not in ("mish", "leaky", "swish", "linear"):
            raise NotImplementedError("Acativation function not found.")
        layers = [nn.Conv2d(c1, c2, k, s, (k - 1) // 2, bias=bias)]
        if bn:
            layers.append(nn.BatchNorm2d(c2))
        if activation != "linear":
            layers.append(_ACTS[activation]())
        self.conv = nn.ModuleList(layers)
        self.c1, self.c2, self.k, self.s, self.act, self.has_bn = c1, c2, k, s, activation, bn
        self.stem = (c1 == 3)
        self._packed, self._packed_t, self._affine = _Packed(), _Packed(), {}

    def weight(self):
        return self._packed.get(self.conv[0].weight, stem=self.stem)

    def weight_t(self):
        """[Cin][kh][kw][Cout] bf16 copy for dgrad."""
        return self._packed_t.get(self.conv[0].weight, transpose=True)

    def forward(self, ctx, x, out=None, residual=None, head=None, head_scale=None, head_shift=None):
        w = self.weight()
        k, st = (1, 1) if self.stem else (self.k, self.s)     # the stem's k x k / stride lives in its im2col
        if head is not None:                       # biased linear 1x1 -> fp32 [B,na,gs,gs,ch]
            y = ops.conv2d(x, w, self.c2, k, st, scale=head_scale, shift=head_shift, act="linear", head=head)
            if ctx.tape is not None:
                ctx.tape.append(("head", self, x, y, head_scale))
            return y
        bn = self.conv[1]
        if not ctx.training:
            scale, shift = _bn_eval_affine(bn, self._affine)
            return ops.conv2d(x, w, self.c2, k, st, out=out, scale=scale, shift=shift, act=self.act,
                              residual=residual)
        aff = torch.empty(4 * self.c2, dtype=torch.float32, device=ctx.device)
        scale, shift, mean, invstd = aff[:self.c2], aff[self.c2:2 * self.c2], aff[2 * self.c2:3 * self.c2], \
            aff[3 * self.c2:]
        if ctx._acc is not None:
            bnf = ops.bn_fuse(ctx.acc_slot(self.c2), None, bn, scale, shift, mean, invstd)
            raw = ops.conv2d(x, w, self.c2, k, st, bn=bnf)
            if out is None:
                out = ctx.new(raw.N, raw.H, raw.W, self.c2)
            ops.scale_shift_act_bn(raw, bnf, self.act, out, residual=residual)
        else:
            part, ctr = ctx.stat_slot(self.c2)
            raw = ops.conv2d(x, w, self.c2, k, st, bn=ops.bn_fuse(part, ctr, bn, scale, shift, mean, invstd))
            if out is None:
                out = ctx.new(raw.N, raw.H, raw.W, self.c2)
            ops.scale_shift_act(raw, scale, shift, self.act, out, residual=residual)
        ctx.tape.append(("conv", self, x, raw, out, residual, scale, shift, mean, invstd))
        return out


class Bottleneck(nn.Module):
    def __init__(self, c1, c2, shortcut=True, e=0.5, act=None):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, 1, 1, act)
        self.cv2 = Conv(c_, c2, 3, 1, act)
        self.add = shortcut and c1 == c2

    def forward(self, ctx, x, out=None):
        return self.cv2(ctx, self.cv1(ctx, x), out=out, residual=x if self.add else None)


class CSP(nn.Module):
    def __init__(self, c1, c2, n=1, shortcut=True, e=0.5):
        super().__init__()
        c_ = int(c1 * e)
        self.cv1 = Conv(c1, c_, 1, 1, "mish")
        self.cv2 = Conv(c1, c_, 1, 1, "mish")
        self.cv3 = Conv(c_, c_, 1, 1, "mish")
        self.cv4 = Conv(2 * c_, c2, 1, 1, "mish")
        self.m = nn.Sequential(*(Bottleneck(c_, c_, shortcut, e=1.0, act="mish") for _ in range(n)))
        self.c_ = c_

    def forward(self, ctx, x, out=None):
        cat = ctx.new(x.N, x.H, x.W, 2 * self.c_)
        y = self.cv1(ctx, x)
        for b in self.m:
            y = b(ctx, y)
        self.cv3(ctx, y, out=cat.slice(0, self.c_))
        self.cv2(ctx, x, out=cat.slice(self.c_, self.c_))
        return self.cv4(ctx, cat, out=out)


class C5(nn.Module):
    def __init__(self, c1, c2, e=0.5):
        super().__init__()
        c_ = int(c1 * e)
        self.cv1 = Conv(c1, c_, 1, 1, "leaky")
        self.cv2 = Conv(c_, c1, 3, 1, "leaky")
        self.cv3 = Conv(c1, c_, 1, 1, "leaky")
        self.cv4 = Conv(c_, c1, 3, 1, "leaky")
        self.cv5 = Conv(c1, c2, 1, 1, "leaky")

    def forward(self, ctx, x, out=None):
        for cv in (self.cv1, self.cv2, self.cv3, self.cv4):
            x = cv(ctx, x)
        return self.cv5(ctx, x, out=out)


def _spp_pools(ctx, src, dsts):
    """5/9/13 stride-1 max-pools of `src`.  Inference cascades 5x5 pools (pool9 = pool5 o pool5, pool13 =
    pool5 o pool9: identical values, 3.7x fewer loads).  Training pools directly so that the backward pass routes
    gradients to the same arg-max positions as the reference's three independent pools even when values tie."""
    if ctx.tape is None:
        cur = src
        for dst, _ in dsts:
            ops.maxpool(cur, 5, 1, 2, out=dst)
            cur = dst
        return
    for dst, k in dsts:
        ops.maxpool(src, k, 1, k // 2, out=dst)
    if src.H * src.W <= 1024 and len(dsts) == 3:      # one fused backward launch for the three pools (ryolo_spp_bwd)
        ctx.tape.append(("spp3", src, [d for d, _ in dsts], [k for _, k in dsts]))
    else:
        for dst, k in dsts:
            ctx.tape.append(("maxpool", src, dst, k, 1, k // 2))


class SPP(nn.Module):
    def __init__(self, c1, c2):
        super().__init__()
        c_ = c1 // 2
        self.cv1 = Conv(c1, c_, 1, 1, 'leaky')
        self.cv2 = Conv(c_, c1, 3, 1, 'leaky')
        self.cv3 = Conv(c1, c_, 1, 1, 'leaky')
        self.m1 = nn.MaxPool2d(kernel_size=5, stride=1, padding=5 // 2)
        self.m2 = nn.MaxPool2d(kernel_size=9, stride=1, padding=9 // 2)
        self.m3 = nn.MaxPool2d(kernel_size=13, stride=1, padding=13 // 2)
        self.cv4 = Conv(c_ * 4, c_, 1, 1, 'leaky')
        self.cv5 = Conv(c_, c1, 3, 1, 'leaky')
        self.cv6 = Conv(c1, c2, 1, 1, 'leaky')
        self.c_ = c_

    def forward(self, ctx, x, out=None):
        c_ = self.c_
        cat = ctx.new(x.N, x.H, x.W, 4 * c_)                      # order m3, m2, m1, x  (model/utils.py:241)
        src = cat.slice(3 * c_, c_)
        self.cv3(ctx, self.cv2(ctx, self.cv1(ctx, x)), out=src)
        _spp_pools(ctx, src, [(cat.slice(2 * c_, c_), 5), (cat.slice(c_, c_), 9), (cat.slice(0, c_), 13)])
        return self.cv6(ctx, self.cv5(ctx, self.cv4(ctx, cat)), out=out)


class C3(nn.Module):
    """yolov5 CSP bottleneck with 3 convs (model/utils.py:83-95)."""

    def __init__(self, c1, c2, n=1, shortcut=True, e=0.5):
        super().__init__()
        c_ = int(c1 * e)
        self.cv1 = Conv(c1, c_, 1, 1, "swish")
        self.cv2 = Conv(c1, c_, 1, 1, "swish")
        self.cv3 = Conv(2 * c_, c2, 1, 1, "swish")
        self.m = nn.Sequential(*(Bottleneck(c_, c_, shortcut, e=1.0, act="swish") for _ in range(n)))
        self.c_ = c_

    def forward(self, ctx, x, out=None):
        cat = ctx.new(x.N, x.H, x.W, 2 * self.c_)                 # [m(cv1(x)) | cv2(x)]
        y = self.cv1(ctx, x)
        for i, b in enumerate(self.m):
            y = b(ctx, y, out=cat.slice(0, self.c_) if i == len(self.m) - 1 else None)
        self.cv2(ctx, x, out=cat.slice(self.c_, self.c_))
        return self.cv3(ctx, cat, out=out)


class SPPF(nn.Module):
    """yolov5 SPP-Fast (model/utils.py:247-261): three chained 5x5 pools."""

    def __init__(self, c1, c2, k=5):
        super().__init__()
        c_ = c1 // 2
        self.cv1 = Conv(c1, c_, 1, 1, "swish")
        self.cv2 = Conv(c_ * 4, c2, 1, 1, "swish")
        self.m = nn.MaxPool2d(kernel_size=k, stride=1, padding=k // 2)
        self.c_, self.k = c_, k

    def forward(self, ctx, x, out=None):
        c_, k = self.c_, self.k
        cat = ctx.new(x.N, x.H, x.W, 4 * c_)                      # [x, y1, y2, y3]
        src = self.cv1(ctx, x, out=cat.slice(0, c_))
        for i in range(1, 4):
            dst = cat.slice(i * c_, c_)
            ops.maxpool(src, k, 1, k // 2, out=dst)
            if ctx.tape is not None:
                ctx.tape.append(("maxpool", src, dst, k, 1, k // 2))
            src = dst
        return self.cv2(ctx, cat, out=out)


class SPPCSPC(nn.Module):
    def __init__(self, c1, c2, e=0.5, k=(5, 9, 13)):
        super().__init__()
        c_ = int(2 * c2 * e)
        self.cv1 = Conv(c1, c_, 1, 1, "swish")
        self.cv2 = Conv(c1, c_, 1, 1, "swish")
        self.cv3 = Conv(c_, c_, 3, 1, "swish")
        self.cv4 = Conv(c_, c_, 1, 1, "swish")
        self.m = nn.ModuleList([nn.MaxPool2d(kernel_size=x, stride=1, padding=x // 2) for x in k])
        self.cv5 = Conv(4 * c_, c_, 1, 1, "swish")
        self.cv6 = Conv(c_, c_, 3, 1, "swish")
        self.cv7 = Conv(2 * c_, c2, 1, 1, "swish")
        self.c_, self.ks = c_, tuple(k)

    def forward(self, ctx, x, out=None):
        c_ = self.c_
        cat4 = ctx.new(x.N, x.H, x.W, 4 * c_)                     # order x1, m5, m9, m13 (model/utils.py:279)
        src = cat4.slice(0, c_)
        self.cv4(ctx, self.cv3(ctx, self.cv1(ctx, x)), out=src)
        assert self.ks == (5, 9, 13)
        _spp_pools(ctx, src, [(cat4.slice((i + 1) * c_, c_), k) for i, k in enumerate(self.ks)])
        cat2 = ctx.new(x.N, x.H, x.W, 2 * c_)
        self.cv6(ctx, self.cv5(ctx, cat4), out=cat2.slice(0, c_))
        self.cv2(ctx, x, out=cat2.slice(c_, c_))
        return self.cv7(ctx, cat2, out=out)


class ELAN1(nn.Module):
    def __init__(self, c1, c2, e1=0.5, e2=0.5):
        super().__init__()
        h1, h2 = int(c1 * e1), int(c1 * e2)
        self.cv1 = Conv(c1, h1, 1, 1, "swish")
        self.cv2 = Conv(c1, h1, 1, 1, "swish")
        self.cv3 = Conv(h1, h2, 3, 1, "swish")
        self.cv4 = Conv(h1, h2, 3, 1, "swish")
        self.cv5 = Conv(h2, h2, 3, 1, "swish")
        self.cv6 = Conv(h2, h2, 3, 1, "swish")
        self.cv7 = Conv((h1 + h2) * 2, c2, 1, 1, "swish")
        self.h1, self.h2 = h1, h2

    def forward(self, ctx, x, out=None):
        h1, h2 = self.h1, self.h2
        cat = ctx.new(x.N, x.H, x.W, 2 * (h1 + h2))               # x1, x2, x3, x4
        self.cv1(ctx, x, out=cat.slice(0, h1))
        x2 = self.cv2(ctx, x, out=cat.slice(h1, h1))
        x3 = self.cv4(ctx, self.cv3(ctx, x2), out=cat.slice(2 * h1, h2))
        self.cv6(ctx, self.cv5(ctx, x3), out=cat.slice(2 * h1 + h2, h2))
        return self.cv7(ctx, cat, out=out)


class ELAN2(nn.Module):
    def __init__(self, c1, c2, e1=0.5, e2=0.25):
        super().__init__()
        h1, h2 = int(c1 * e1), int(c1 * e2)
        self.cv1 = Conv(c1, h1, 1, 1, "swish")
        self.cv2 = Conv(c1, h1, 1, 1, "swish")
        self.cv3 = Conv(h1, h2, 3, 1, "swish")
        self.cv4 = Conv(h2, h2, 3, 1, "swish")
        self.cv5 = Conv(h2, h2, 3, 1, "swish")
        self.cv6 = Conv(h2, h2, 3, 1, "swish")
        self.cv7 = Conv(h1 * 2 + h2 * 4, c2, 1, 1, "swish")
        self.h1, self.h2 = h1, h2

    def forward(self, ctx, x, out=None):
        h1, h2 = self.h1, self.h2
        cat = ctx.new(x.N, x.H, x.W, 2 * h1 + 4 * h2)             # x1 .. x6
        self.cv1(ctx, x, out=cat.slice(0, h1))
        y = self.cv2(ctx, x, out=cat.slice(h1, h1))
        for i, cv in enumerate((self.cv3, self.cv4, self.cv5, self.cv6)):
            y = cv(ctx, y, out=cat.slice(2 * h1 + i * h2, h2))
        return self.cv7(ctx, cat, out=out)


class MaxConv(nn.Module):
    def __init__(self, c1, e=0.5):
        super().__init__()
        c_ = int(c1 * e)
        self.m = nn.MaxPool2d(kernel_size=2, stride=2)
        self.cv1 = Conv(c1, c_, 1, 1, "swish")
        self.cv2 = Conv(c1, c_, 1, 1, "swish")
        self.cv3 = Conv(c_, c_, 3, 2, "swish")
        self.c_ = c_

    def forward(self, ctx, x, out=None):
        """out: a 2*c_-channel view (x1 | x2), typically a slice of the consumer's concat buffer."""
        c_ = self.c_
        if out is None:
            out = ctx.new(x.N, x.H // 2, x.W // 2, 2 * c_)
        pooled = ops.maxpool(x, 2, 2, 0)
        if ctx.tape is not None:
            ctx.tape.append(("maxpool", x, pooled, 2, 2, 0))
        self.cv1(ctx, pooled, out=out.slice(0, c_))
        self.cv3(ctx, self.cv2(ctx, x), out=out.slice(c_, c_))
        return out


class ImplicitA(nn.Module):
    def __init__(self, channel, mean=0., std=.02):
        super().__init__()
        self.implicit = nn.Parameter(torch.zeros(1, channel, 1, 1))
        nn.init.normal_(self.implicit, mean=mean, std=std)


class ImplicitM(nn.Module):
    def __init__(self, channel, mean=1., std=.02):
        super().__init__()
        self.implicit = nn.Parameter(torch.ones(1, channel, 1, 1))
        nn.init.normal_(self.implicit, mean=mean, std=std)


class RepConv(nn.Module):
    """SiLU(BN(conv3x3(x)) + BN(conv1x1(x)) [+ BN(x)])  (model/utils.py:189-215)."""

    def __init__(self, c1, c2, k=3, s=1, p=1):
        super().__init__()
        self.silu = nn.SiLU()
        self.rbr_identity = nn.BatchNorm2d(num_features=c1) if c2 == c1 and s == 1 else None
        self.rbr_dense = nn.Sequential(nn.Conv2d(c1, c2, k, s, p, bias=False), nn.BatchNorm2d(num_features=c2))
        self.rbr_1x1 = nn.Sequential(nn.Conv2d(c1, c2, 1, s, 0, bias=False), nn.BatchNorm2d(num_features=c2))
        self.c1, self.c2, self.k, self.s = c1, c2, k, s
        self._pd, self._p1, self._ad, self._a1 = _Packed(), _Packed(), {}, {}
        self._pdt, self._p1t = _Packed(), _Packed()

    def reparam(self):
        """(bf16 packed [c2, k*k*c1] weight, fp32 bias[c2]) of the single k x k conv this block equals in eval mode.
        Cached until a parameter or a running statistic changes."""
        sd, bd = _bn_eval_affine(self.rbr_dense[1], self._ad)
        s1, b1 = _bn_eval_affine(self.rbr_1x1[1], self._a1)
        wd_, w1_ = self.rbr_dense[0].weight, self.rbr_1x1[0].weight
        key = (self._ad["key"], self._a1["key"], wd_._version, w1_._version, wd_.data_ptr(), w1_.data_ptr())
        if getattr(self, "_rep_key", None) != key:
            w = wd_.data.float() * sd.view(-1, 1, 1, 1)
            c = self.k // 2
            w[:, :, c, c] += w1_.data.float()[:, :, 0, 0] * s1.view(-1, 1)
            self._rep = (ops.pack_weights(w), (bd + b1).contiguous())
            self._rep_key = key
        return self._rep

    def forward(self, ctx, x, out=None):
        if self.rbr_identity is not None:
            raise NotImplementedError("RepConv identity branch is not used by any reference network")
        wd, w1 = self._pd.get(self.rbr_dense[0].weight), self._p1.get(self.rbr_1x1[0].weight)
        if out is None:
            out = ctx.new(x.N, x.H, x.W, self.c2)
        if not ctx.training:
            # Inference-time re-parameterisation (RepVGG, the paper model/utils.py:189-191 cites): with BatchNorm folded,
            #   SiLU(s_d * conv3x3(x) + b_d + s_1 * conv1x1(x) + b_1) = SiLU(conv3x3(x; s_d W_d + pad(s_1 W_1)) + b_d + b_1)
            # ONE tensor-core launch with the bias + SiLU epilogue instead of two convs and an elementwise pass.
            wf, bf = self.reparam()
            return ops.conv2d(x, wf, self.c2, self.k, self.s, out=out, shift=bf, act="swish")
        affs, raws = [], []
        for wgt, kk, bn in ((wd, self.k, self.rbr_dense[1]), (w1, 1, self.rbr_1x1[1])):
            part, ctr = ctx.stat_slot(self.c2)
            aff = torch.empty(4, self.c2, dtype=torch.float32, device=ctx.device)
            raws.append(ops.conv2d(x, wgt, self.c2, kk, self.s,
                                   bn=ops.bn_fuse(part, ctr, bn, aff[0], aff[1], aff[2], aff[3])))
            affs.append(aff)
        rd, r1 = raws
        ops.scale_shift_act(rd, affs[0][0], affs[0][1], "swish", out, x2=r1, scale2=affs[1][0], shift2=affs[1][1])
        ctx.tape.append(("repconv", self, x, rd, r1, out, affs))
        return out
