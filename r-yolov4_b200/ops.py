"""Thin Python wrappers over the conv-stack entry points of libryolo_b200.so.

`Act` is an NHWC bf16 *view*: a channel slice [coff, coff+C) of a [N,H,W,Cbuf] buffer, so producers
write straight into their slot of a concat buffer (torch.cat of the reference never materialises).
"""
import ctypes

import torch

from . import _lib as L

ACT = {"linear": 0, "leaky": 1, "mish": 2, "swish": 3}

# Optional per-launch timing of the conv kernel (bench.py): list of (tag, start_event, end_event).
PROFILE = None
_pending = []


def _prof_begin():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    _pending.append(e)


def _prof_end(tag):
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    PROFILE.append((tag, _pending.pop(), e))


class Act:
    __slots__ = ("buf", "N", "H", "W", "C", "coff")

    def __init__(self, buf, C=None, coff=0):
        assert buf.dtype == torch.bfloat16 and buf.dim() == 4 and buf.is_contiguous()
        self.buf = buf
        self.N, self.H, self.W = buf.shape[0], buf.shape[1], buf.shape[2]
        self.C = buf.shape[3] - coff if C is None else C
        self.coff = coff
        assert coff % 8 == 0 and self.C % 8 == 0 and coff + self.C <= buf.shape[3]

    @staticmethod
    def empty(N, H, W, C, device):
        return Act(torch.empty((N, H, W, C), dtype=torch.bfloat16, device=device))

    @property
    def ptr(self):
        return self.buf.data_ptr() + 2 * self.coff

    @property
    def pitch(self):
        return self.buf.shape[3]

    @property
    def P(self):
        return self.N * self.H * self.W

    def slice(self, coff, C):
        return Act(self.buf, C, self.coff + coff)

    def torch(self):
        """[N,H,W,C] torch view (for tests)."""
        return self.buf[..., self.coff:self.coff + self.C]


def _vp(x):
    return ctypes.c_void_p(x)


def _tp(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def out_hw(H, W, k, s):
    p = (k - 1) // 2
    return (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1


BN_PARTIAL_ROWS = 160   # RYOLO_BN_PARTIAL_ROWS


def bn_fuse(partial, counter, bn, scale, shift, save_mean=None, save_invstd=None):
    """ryolo_bn_fuse for an nn.BatchNorm2d `bn` (fused statistics + finalize in the conv epilogue).
    partial: ZEROED fp32 scratch with >= 4*C elements (left zeroed by the kernel); counter: zeroed int32[1]."""
    f = L.BnFuse()
    assert partial.numel() >= 4 * bn.num_features and partial.data_ptr() % 8 == 0
    # counter None: deferred finalize (the conv only accumulates; scale_shift_act_bn finalizes; `partial` is then a
    # per-layer accumulator nobody clears)
    f.partial, f.sum, f.sumsq = partial.data_ptr(), None, None
    f.counter = counter.data_ptr() if counter is not None else None
    f.gamma, f.beta = bn.weight.data_ptr(), bn.bias.data_ptr()
    f.running_mean, f.running_var = bn.running_mean.data_ptr(), bn.running_var.data_ptr()
    f.num_batches = bn.num_batches_tracked.data_ptr()
    f.eps, f.momentum = bn.eps, bn.momentum
    f.scale, f.shift = scale.data_ptr(), shift.data_ptr()
    f.save_mean = save_mean.data_ptr() if save_mean is not None else None
    f.save_invstd = save_invstd.data_ptr() if save_invstd is not None else None
    return f


def conv2d(x, w, Cout, k, stride, out=None, scale=None, shift=None, act="linear", residual=None, head=None,
           reference=False, bn=None):
    """x: Act; w: packed bf16 [Cout, k*k*Cin]; out: Act (bf16 NHWC) or, with head=(na, ch), an fp32
    [N, na, Ho, Wo, ch] tensor.  Returns out."""
    Ho, Wo = out_hw(x.H, x.W, k, stride)
    dev = x.buf.device
    d = L.ConvDesc()
    d.x, d.N, d.H, d.W, d.Cin, d.x_cpitch = x.ptr, x.N, x.H, x.W, x.C, x.pitch
    d.w, d.Cout, d.ksize, d.stride = w.data_ptr(), Cout, k, stride
    assert w.dtype == torch.bfloat16 and w.numel() == Cout * k * k * x.C, (w.shape, Cout, k, x.C)
    if head is None:
        if out is None:
            out = Act.empty(x.N, Ho, Wo, Cout, dev)
        assert (out.N, out.H, out.W, out.C) == (x.N, Ho, Wo, Cout)
        d.out, d.out_mode, d.out_cpitch = out.ptr, 0, out.pitch
    else:
        na, ch = head
        if out is None:
            out = torch.empty((x.N, na, Ho, Wo, ch), dtype=torch.float32, device=dev)
        assert out.dtype == torch.float32 and tuple(out.shape) == (x.N, na, Ho, Wo, ch) and out.is_contiguous()
        d.out, d.out_mode, d.out_cpitch, d.head_na, d.head_ch = out.data_ptr(), 1, 0, na, ch
    d.scale = scale.data_ptr() if scale is not None else None
    d.shift = shift.data_ptr() if shift is not None else None
    d.act = ACT[act]
    if residual is not None:
        d.residual, d.res_cpitch = residual.ptr, residual.pitch
    if bn is not None:
        d.bn = ctypes.pointer(bn)
    fn = L.lib().ryolo_conv2d_reference if reference else L.lib().ryolo_conv2d_forward
    if PROFILE is not None:
        _prof_begin()
    L.check(fn(ctypes.byref(d), L.stream()))
    L.count(1)
    if PROFILE is not None:
        _prof_end(('conv', x.N * Ho * Wo, Cout, k * k * x.C))
    return out


def stem_kpad(k):
    """K of the 1x1 conv the 3-channel k x k stem turns into (3*k*k values padded to a multiple of 32)."""
    return (3 * k * k + 31) // 32 * 32


def pack_weights(w_oihw, stem=False, transpose=False):
    """fp32 OIHW (state-dict layout) -> bf16 [Cout, kh*kw*Cin] ([Cout, stem_kpad(k)] for the stem;
    [Cin, kh*kw*Cout] with transpose=True, the dgrad operand)."""
    w = w_oihw.detach().contiguous().float()
    Cout, Cin, k, _ = w.shape
    kp = stem_kpad(k) if stem else 0
    shape = (Cin, k * k * Cout) if transpose else (Cout, kp if stem else k * k * Cin)
    out = torch.empty(shape, dtype=torch.bfloat16, device=w.device)
    L.check(L.lib().ryolo_pack_weights(_tp(w), Cout, Cin, k, 2 if transpose else kp, _tp(out), L.stream()))
    L.count(1)
    return out


def stem_im2col(img, k=3, stride=1):
    img = img.contiguous().float()
    N, C, H, W = img.shape
    assert C == 3
    Ho, Wo = out_hw(H, W, k, stride)
    kp = stem_kpad(k)
    out = Act.empty(N, Ho, Wo, kp, img.device)
    L.check(L.lib().ryolo_stem_im2col(_tp(img), N, H, W, k, stride, kp, _vp(out.ptr), L.stream()))
    L.count(1)
    return out


def bn_stats(x, sum_, sumsq):
    L.check(L.lib().ryolo_bn_stats(_vp(x.ptr), x.pitch, x.P, x.C, _tp(sum_), _tp(sumsq), L.stream()))
    L.count(1)


def bn_finalize(sum_, sumsq, count, gamma, beta, eps, momentum, running_mean, running_var, num_batches, scale, shift,
                save_mean=None, save_invstd=None):
    L.check(L.lib().ryolo_bn_finalize(_tp(sum_), _tp(sumsq), float(count), gamma.numel(), _tp(gamma), _tp(beta),
                                      float(eps), float(momentum), _tp(running_mean), _tp(running_var),
                                      _tp(num_batches), _tp(scale), _tp(shift), _tp(save_mean), _tp(save_invstd),
                                      L.stream()))
    L.count(1)


def scale_shift_act(x, scale, shift, act, out, residual=None, x2=None, scale2=None, shift2=None):
    L.check(L.lib().ryolo_scale_shift_act(
        _vp(x.ptr), x.pitch, _tp(scale), _tp(shift), _vp(x2.ptr if x2 is not None else 0),
        x2.pitch if x2 is not None else 0, _tp(scale2), _tp(shift2), ACT[act],
        _vp(residual.ptr if residual is not None else 0), residual.pitch if residual is not None else 0,
        _vp(out.ptr), out.pitch, x.P, x.C, L.stream()))
    L.count(1)
    return out


def scale_shift_act_bn(x, bnf, act, out, residual=None):
    """act(BN_train(x)) [+ residual] with the finalize of the producing conv's deferred statistics folded in."""
    L.check(L.lib().ryolo_scale_shift_act_bn(
        _vp(x.ptr), x.pitch, ctypes.byref(bnf), float(x.P), ACT[act],
        _vp(residual.ptr if residual is not None else 0), residual.pitch if residual is not None else 0,
        _vp(out.ptr), out.pitch, x.P, x.C, L.stream()))
    L.count(1)
    return out


def maxpool(x, k, stride, pad, out=None):
    Ho, Wo = (x.H + 2 * pad - k) // stride + 1, (x.W + 2 * pad - k) // stride + 1
    if out is None:
        out = Act.empty(x.N, Ho, Wo, x.C, x.buf.device)
    L.check(L.lib().ryolo_maxpool(_vp(x.ptr), x.pitch, x.N, x.H, x.W, x.C, k, stride, pad, _vp(out.ptr), out.pitch,
                                  L.stream()))
    L.count(1)
    return out


def resize_copy(x, factor, out=None):
    if out is None:
        out = Act.empty(x.N, x.H * factor, x.W * factor, x.C, x.buf.device)
    L.check(L.lib().ryolo_resize_copy(_vp(x.ptr), x.pitch, x.N, x.H, x.W, x.C, factor, _vp(out.ptr), out.pitch,
                                      L.stream()))
    L.count(1)
    return out


# ------------------------------------------------------------------------------------ backward wrappers
def conv2d_dgrad(dy, wt, Cin, k, stride, dx, accumulate):
    """dx (Act [N,H,W,Cin]) (+)= conv^T(dy, w);  wt = pack_weights(w, transpose=True) with Cout == dy.C."""
    assert wt.dtype == torch.bfloat16 and wt.numel() == Cin * k * k * dy.C
    if PROFILE is not None:
        _prof_begin()
    L.check(L.lib().ryolo_conv2d_dgrad(_vp(dy.ptr), dy.pitch, dx.N, dx.H, dx.W, Cin, dy.C, k, stride, _tp(wt),
                                       _vp(dx.ptr), dx.pitch, 1 if accumulate else 0, L.stream()))
    L.count(stride * stride)
    if PROFILE is not None:
        _prof_end(('dgrad', dx.P, Cin, k * k * dy.C))


def conv2d_wgrad(x, dy, Cout, k, stride, dwk):
    """dwk (fp32, K-major [Cout, k*k*Cin]; the stem passes its im2col input and [Cout, 64]) +=
    conv_backward_weight(x, dy)."""
    assert dwk.dtype == torch.float32 and dwk.is_contiguous() and dwk.numel() == Cout * k * k * x.C
    if PROFILE is not None:
        _prof_begin()
    L.check(L.lib().ryolo_conv2d_wgrad(_vp(x.ptr), x.pitch, x.N, x.H, x.W, x.C, _vp(dy.ptr), dy.pitch, dy.C, Cout, k,
                                       stride, _tp(dwk), L.stream()))
    L.count(1)
    if PROFILE is not None:
        _prof_end(('wgrad', dy.P, Cout, k * k * x.C))


def wgrad_to_oihw(dwk, Cout, Cin, k, stem=False):
    """K-major wgrad scratch -> OIHW tensor (torch glue for the non-fused paths / tests)."""
    if stem:
        return dwk.view(Cout, stem_kpad(k))[:, :3 * k * k].reshape(Cout, k, k, 3).permute(0, 3, 1, 2).contiguous()
    return dwk.view(Cout, k * k, Cin).permute(0, 2, 1).reshape(Cout, Cin, k, k).contiguous()


def bn_act_bwd(dout, raw, scale, shift, mean, invstd, act, sums, draw, dgamma, dbeta, dres=None):
    """dres: optional Act that receives a copy of dout (gradient of a residual operand, first contribution)."""
    if PROFILE is not None:
        _prof_begin()
    L.check(L.lib().ryolo_bn_act_bwd(_vp(dout.ptr), dout.pitch, _vp(raw.ptr), raw.pitch, _tp(scale), _tp(shift),
                                     _tp(mean), _tp(invstd), ACT[act], raw.P, raw.C, _tp(sums), _vp(draw.ptr),
                                     draw.pitch, _tp(dgamma), _tp(dbeta), _vp(dres.ptr if dres is not None else 0),
                                     dres.pitch if dres is not None else 0, L.stream()))
    L.count(2)
    if PROFILE is not None:
        _prof_end(('bn_bwd', raw.P, raw.C, 0))


def act_bwd2(dout, x1, s1, b1, x2, s2, b2, act, ds):
    L.check(L.lib().ryolo_act_bwd2(_vp(dout.ptr), dout.pitch, _vp(x1.ptr), x1.pitch, _tp(s1), _tp(b1), _vp(x2.ptr),
                                   x2.pitch, _tp(s2), _tp(b2), ACT[act], _vp(ds.ptr), ds.pitch, x1.P, x1.C,
                                   L.stream()))
    L.count(1)


def add_into(dst, src, accumulate):
    L.check(L.lib().ryolo_add_into(_vp(dst.ptr), dst.pitch, _vp(src.ptr), src.pitch, src.P, src.C,
                                   1 if accumulate else 0, L.stream()))
    L.count(1)


def maxpool_bwd(x, dy, k, stride, pad, dx, accumulate):
    scratch = torch.empty(x.P * x.C, dtype=torch.float32, device=x.buf.device)
    L.check(L.lib().ryolo_maxpool_bwd(_vp(x.ptr), x.pitch, _vp(dy.ptr), dy.pitch, x.N, x.H, x.W, x.C, k, stride, pad,
                                      _vp(dx.ptr), dx.pitch, 1 if accumulate else 0, _tp(scratch), L.stream()))
    L.count(2)


def spp_bwd(x, dys, ks, dx, accumulate):
    """dx (+)= gradients of the three stride-1 'same' pools `ks` of x (SPP, model/utils.py:231-241) in one launch."""
    assert len(dys) == 3 and dys[0].pitch == dys[1].pitch == dys[2].pitch
    L.check(L.lib().ryolo_spp_bwd(_vp(x.ptr), x.pitch, _vp(dys[0].ptr), _vp(dys[1].ptr), _vp(dys[2].ptr), dys[0].pitch,
                                  x.N, x.H, x.W, x.C, ks[0], ks[1], ks[2], _vp(dx.ptr), dx.pitch, 1 if accumulate else 0,
                                  L.stream()))
    L.count(1)


def upsample2x_bwd(dy, dx, accumulate):
    L.check(L.lib().ryolo_upsample2x_bwd(_vp(dy.ptr), dy.pitch, dx.N, dx.H, dx.W, dx.C, _vp(dx.ptr), dx.pitch,
                                         1 if accumulate else 0, L.stream()))
    L.count(1)


def head_grad_pack(glev, Cpad, mul, dbias, yhead=None, dsum=None, dmul=None):
    """fp32 head gradient -> bf16 NHWC d pre (+ bias gradient); with `yhead` (yolov7's implicit head) the same pass also
    returns sum d pre per channel in `dsum` and accumulates the ImplicitM gradient into `dmul`."""
    B, na, H, W, ch = glev.shape
    out = Act.empty(B, H, W, Cpad, glev.device)
    if yhead is not None:
        assert yhead.shape == glev.shape and yhead.dtype == torch.float32 and yhead.is_contiguous()
    L.check(L.lib().ryolo_head_grad_pack(_tp(glev.contiguous()), B, na, H, W, ch, Cpad, _tp(mul), _vp(out.ptr),
                                         _tp(dbias), _tp(yhead), _tp(dsum), _tp(dmul), L.stream()))
    L.count(1)
    return out


def sgd_step(param, grad, buf, lr, momentum, weight_decay, nesterov, first):
    L.check(L.lib().ryolo_sgd_step(_tp(param), _tp(grad), _tp(buf), param.numel(), float(lr), float(momentum),
                                   float(weight_decay), 1 if nesterov else 0, 1 if first else 0, L.stream()))
    L.count(1)


def sgd_step_lrdev(param, grad, buf, lr_dev, momentum, weight_decay, nesterov):
    """SGD update with the learning rate in a device float (CUDA-graph friendly, see TrainStep.capture)."""
    L.check(L.lib().ryolo_sgd_step_lrdev(_tp(param), _tp(grad), _tp(buf), param.numel(), _tp(lr_dev), float(momentum),
                                         float(weight_decay), 1 if nesterov else 0, L.stream()))
    L.count(1)


def adam_step(param, grad, exp_avg, exp_avg_sq, lr, step, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
    L.check(L.lib().ryolo_adam_step(_tp(param), _tp(grad), _tp(exp_avg), _tp(exp_avg_sq), param.numel(), float(lr),
                                    float(betas[0]), float(betas[1]), float(eps), float(weight_decay), int(step),
                                    L.stream()))
    L.count(1)
