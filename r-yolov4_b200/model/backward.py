"""Backward pass of the conv stack: walks the tape recorded by the training forward (model/blocks.py)
in reverse and drives the dgrad / wgrad / BN-activation / pooling / upsample backward kernels.
This is what autograd does for the reference at train.py:198; here nothing goes through torch autograd.

Gradients of activations live in bf16 NHWC buffers that mirror the forward buffers (so a slice of a
concat buffer has its gradient in the same slice of the mirrored buffer); weight gradients are fp32 and
are ACCUMULATED into the tensors handed in by `param_grads` (zeroed by the caller).
"""
import os

import torch

from .. import ops
from ..ops import Act


class GradStore:
    """bf16 gradient buffers mirroring forward activation buffers, with per-channel-range init tracking."""

    def __init__(self):
        self.bufs = {}       # buf.data_ptr() -> (grad tensor, list of initialised (lo, hi))

    def _entry(self, act):
        key = act.buf.data_ptr()
        e = self.bufs.get(key)
        if e is None:
            e = (torch.empty_like(act.buf), [])
            self.bufs[key] = e
        return e

    def view(self, act):
        g, _ = self._entry(act)
        return Act(g, act.C, act.coff)

    def is_init(self, act):
        _, rng = self._entry(act)
        lo, hi = act.coff, act.coff + act.C
        covered = lo
        for a, b in sorted(rng):
            if a <= covered < b:
                covered = b
        return covered >= hi

    def mark(self, act):
        self._entry(act)[1].append((act.coff, act.coff + act.C))

    def writable(self, act):
        """(grad view, accumulate flag) for an op that can either overwrite or add."""
        gv = self.view(act)
        acc = self.is_init(act)
        if not acc:
            _, rng = self._entry(act)
            lo, hi = act.coff, act.coff + act.C
            if any(a < hi and lo < b for a, b in rng):       # partial overlap: fall back to zero + accumulate
                self.zero(act)
                acc = True
        return gv, acc

    def zero(self, act):
        gv = self.view(act)
        gv.torch().zero_()
        self.mark(act)
        return gv

    def add_only(self, act):
        """grad view that holds valid numbers (zeros if nothing was written yet): for atomically-added gradients."""
        if not self.is_init(act):
            return self.zero(act)
        return self.view(act)


def _accumulate(G, act, src):
    gv, acc = G.writable(act)
    ops.add_into(gv, src, acc)
    G.mark(act)


def _implicit_head_grads(model, conv, dpre_sum, param_grads):
    """yolov7 head y = im * (conv(x + ia) + b) (model/utils.py:163-186): what is left of the ImplicitA gradient once
    ryolo_head_grad_pack has produced dpre_sum[c] = sum over pixels of d pre_c (and d ImplicitM) in its one pass over the
    head gradient: a [Cin] matrix-vector product and a rank-1 update on Cout x Cin numbers."""
    neck = model.neck
    i = {id(getattr(neck, f"conv{4 + j}")): j for j in (1, 2, 3)}[id(conv)]
    ia = getattr(neck, f"ia{i}")
    w = conv.conv[0].weight.data.flatten(1)                               # [Cout, Cin] fp32
    param_grads[id(ia.implicit)].add_((dpre_sum @ w).view_as(ia.implicit))
    # the conv's input is x + ia: the wgrad GEMM sees x only, the constant part contributes (sum d pre) (x) ia
    param_grads[id(conv.conv[0].weight)].view(w.shape).addr_(dpre_sum, ia.implicit.data.flatten())


def _repconv_backward(mod, x, rd, r1, dout, affs, G, sums, param_grads, sink):
    """out = silu(bn_d(conv3x3(x)) + bn_1(conv1x1(x)))   (model/utils.py:209-215)."""
    C = mod.c2
    ds = ops.Act.empty(rd.N, rd.H, rd.W, C, rd.buf.device)
    ops.act_bwd2(dout, rd, affs[0][0], affs[0][1], r1, affs[1][0], affs[1][1], "swish", ds)
    first = True
    for raw, aff, seq, k, packed_t in ((rd, affs[0], mod.rbr_dense, mod.k, mod._pdt), (r1, affs[1], mod.rbr_1x1, 1, mod._p1t)):
        bn = seq[1]
        ops.bn_act_bwd(ds, raw, aff[0], aff[1], aff[2], aff[3], "linear", sums[:2 * C] if first else sums[2 * C:],
                     raw, param_grads[id(bn.weight)], param_grads[id(bn.bias)])
        sink.wgrad(x, raw, C, k, mod.s, seq[0].weight)
        gx, acc = G.writable(x)
        ops.conv2d_dgrad(raw, packed_t.get(seq[0].weight, transpose=True), x.C, k, mod.s, gx, acc)
        G.mark(x)
        first = False



_SIDE = {}
_MAIN = {}
WGRAD_SIDE_STREAM = os.environ.get("RYOLO_WGRAD_SIDE", "1") != "0"   # False: weight-gradient GEMMs on the main stream
# Experiment switch (RYOLO_BWD_PRIO=1, off by default: measured neutral on the whole step): the critical
# path of the backward pass (BatchNorm backward -> dgrad -> BatchNorm backward ...) runs on a HIGH-priority stream, the
# weight-gradient GEMMs on a default-priority one.  When a layer's BatchNorm backward finishes, its dgrad and
# its wgrad become runnable together and both want every SM (one CTA with > 200 KB of shared memory each): with equal
# priorities the wgrad, enqueued first, wins and the chain behind the dgrad waits.  With priorities the dgrad goes first,
# and the wgrad then runs UNDER the next layer's HBM-bound BatchNorm backward, whose blocks fit on the same SMs (two
# 96-register blocks + 16 KB next to the 56-register wgrad CTA with 13 x 16 KB boxes): tensor-core work hidden behind
# memory-bound work instead of time-sliced with it.
BACKWARD_HIGH_PRIORITY = os.environ.get("RYOLO_BWD_PRIO", "0") != "0"
RESIDUAL_COPY_IN_BN_BWD = os.environ.get("RYOLO_RES_FUSE", "1") != "0"     # A/B switch, see run_backward / "conv" entries


def _side_stream(dev):
    """One extra stream per device for the weight-gradient GEMMs (see run_backward)."""
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(dev, priority=0)
    return _SIDE[key]


def _main_stream(dev):
    """High-priority stream carrying the backward pass's critical path."""
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _MAIN:
        _MAIN[key] = torch.cuda.Stream(dev, priority=-1)
    return _MAIN[key]


class _WgradSink:
    """Where conv_wgrad writes.  Fused path (TrainStep): the model's flat K-major scratch, folded into the flat OIHW
    gradients by one launch at the end.  Generic path (autograd node, block tests): per-layer scratch + torch glue."""

    def __init__(self, model, param_grads):
        self.param_grads = param_grads
        self.fused = getattr(model, "_pack_table", None) is not None and \
            param_grads is getattr(model, "_grad_views", None)
        self.model = model
        self.pending = []
        self.keep = []
        self.side = _side_stream(next(iter(param_grads.values())).device)
        if self.fused:
            self.views = model.wgrad_scratch()
            if not getattr(model, "_wg_clean", False):      # the tiled fold leaves the scratch zeroed (see unpack_wgrads)
                model._wg_flat.zero_()
            model._wg_clean = False                          # dirty until every weight gradient has been folded again

    def wgrad(self, x, dy, Cout, k, stride, weight, stem=False, keep=()):
        """Launch conv_wgrad on the side stream: it only reads x and dy (= d raw, final once the BN backward of this
        layer has run), so it overlaps the dgrad of this layer and the HBM-bound BN backward of the next one instead
        of serialising with them (its CTAs leave most of an SM's bandwidth idle)."""
        buf = self.buffer(weight, stem)
        if not WGRAD_SIDE_STREAM:
            ops.conv2d_wgrad(x, dy, Cout, k, stride, buf)
            return
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(main)
        with torch.cuda.stream(self.side):
            self.side.wait_event(ev)
            ops.conv2d_wgrad(x, dy, Cout, k, stride, buf)
        self.keep.extend(keep)               # temporaries the side stream still reads: freed after the join

    def buffer(self, weight, stem=False):
        if self.fused:
            return self.views[id(weight)]
        Cout, Cin, k, _ = weight.shape
        buf = torch.zeros(Cout * (ops.stem_kpad(k) if stem else k * k * Cin), dtype=torch.float32,
                          device=weight.device)
        self.pending.append((weight, buf, stem))
        return buf

    def finish(self):
        torch.cuda.current_stream().wait_stream(self.side)      # join before anything the GEMMs read is released
        self.keep.clear()
        if self.fused:
            if not getattr(self.model, "_bucketed_unpack", False):     # TrainStep unpacks bucket by bucket itself
                self.model.unpack_wgrads()
            return
        for weight, buf, stem in self.pending:
            Cout, Cin, k, _ = weight.shape
            self.param_grads[id(weight)].add_(ops.wgrad_to_oihw(buf, Cout, Cin, k, stem))


def run_backward(model, ctx, dlevels, param_grads, seed=(), on_entry=None):
    """Runs the backward pass (see _run_backward) on the high-priority stream, fenced against the caller's stream on
    both sides, so callers keep ordinary single-stream semantics."""
    if not (BACKWARD_HIGH_PRIORITY and WGRAD_SIDE_STREAM):
        return _run_backward(model, ctx, dlevels, param_grads, seed, on_entry)
    cur = torch.cuda.current_stream()
    hi = _main_stream(ctx.device)
    hi.wait_stream(cur)
    with torch.cuda.stream(hi):
        G = _run_backward(model, ctx, dlevels, param_grads, seed, on_entry)
    cur.wait_stream(hi)
    return G


def _run_backward(model, ctx, dlevels, param_grads, seed=(), on_entry=None):
    """dlevels: 3 fp32 tensors [B, na, gs, gs, ch] (d loss / d level, strides 8, 16, 32).
    param_grads: dict id(parameter) -> fp32 tensor (same shape) that receives += d loss / d parameter.
    seed: optional [(activation, gradient Act)] pairs that pre-load output gradients (block-level tests).
    on_entry: optional callback(i, sink) after the i-th tape entry (reverse order) has launched all of its kernels —
    TrainStep uses it to hand finished gradient buckets to the all-reduce while the backward pass keeps going."""
    from .blocks import Conv
    tape = ctx.tape
    assert tape is not None, "backward needs a training-mode forward"
    dev = ctx.device
    G = GradStore()
    sink = _WgradSink(model, param_grads)
    for act, g in seed:
        ops.add_into(G.view(act), g, False)
        G.mark(act)
    sums = torch.zeros(2 * model._bn_channels + 16, dtype=torch.float32, device=dev)
    soff = 0
    state = {"head_i": 2, "soff": 0}
    na, ch = model.na, model.ch

    def pg(p):
        return param_grads[id(p)]

    def handle(e):
        soff, head_i = state["soff"], state["head_i"]
        kind = e[0]
        if kind == "head":
            _, mod, x, y, mul = e
            gl = dlevels[head_i]
            state["head_i"] = head_i - 1
            conv = mod.conv[0]
            Cout = na * ch
            Cpad = (Cout + 7) // 8 * 8
            if mul is not None:        # yolov7: y = im * (conv(x + ia) + b)   (model/neck.py:201,208,215)
                j = {id(getattr(model.neck, f"conv{4 + q}")): q for q in (1, 2, 3)}[id(mod)]
                im = getattr(model.neck, f"im{j}")
                dsum = torch.zeros(Cout, dtype=torch.float32, device=dev)
                dpre = ops.head_grad_pack(gl, Cpad, mul, pg(conv.bias), yhead=y.detach().contiguous(), dsum=dsum,
                                          dmul=pg(im.implicit).view(-1))
                _implicit_head_grads(model, mod, dsum, param_grads)
            else:
                dpre = ops.head_grad_pack(gl, Cpad, mul, pg(conv.bias))
            sink.wgrad(x, dpre, Cout, 1, 1, conv.weight, keep=(dpre,))
            w = conv.weight.data
            if Cpad != Cout:
                w = torch.cat((w, w.new_zeros(Cpad - Cout, *w.shape[1:])), 0)
            wt = ops.pack_weights(w, transpose=True)
            gx, acc = G.writable(x)
            ops.conv2d_dgrad(dpre, wt, x.C, 1, 1, gx, acc)
            G.mark(x)
        elif kind == "conv":
            _, mod, x, raw, out, residual, scale, shift, mean, invstd = e
            if not G.is_init(out):
                return                                     # output never used downstream (cannot happen in these nets)
            dout = G.view(out)
            dres = None
            if residual is not None:                       # out = residual + act(bn(raw)): identity path
                gv, acc = G.writable(residual)
                if acc or not RESIDUAL_COPY_IN_BN_BWD:     # something already flowed into it: read-add-write pass
                    ops.add_into(gv, dout, acc)
                else:                                      # first contribution: a copy, made by the BN-backward pass
                    dres = gv
                G.mark(residual)
            bn = mod.conv[1]
            C = mod.c2
            ops.bn_act_bwd(dout, raw, scale, shift, mean, invstd, mod.act, sums[soff:soff + 2 * C], raw,
                           pg(bn.weight), pg(bn.bias), dres=dres)     # d raw overwrites raw in place
            state["soff"] = soff + 2 * C
            k, st = (1, 1) if mod.stem else (mod.k, mod.s)
            sink.wgrad(x, raw, C, k, st, mod.conv[0].weight, mod.stem)
            if not mod.stem:
                gx, acc = G.writable(x)
                ops.conv2d_dgrad(raw, mod.weight_t(), x.C, mod.k, mod.s, gx, acc)
                G.mark(x)
        elif kind == "repconv":
            _, mod, x, rd, r1, out, affs = e
            dout = G.view(out)
            _repconv_backward(mod, x, rd, r1, dout, affs, G, sums[soff:soff + 4 * mod.c2], param_grads, sink)
            state["soff"] = soff + 4 * mod.c2
        elif kind == "maxpool":
            _, src, dst, k, s, p = e
            if not G.is_init(dst):
                return
            gs, acc = G.writable(src)
            ops.maxpool_bwd(src, G.view(dst), k, s, p, gs, acc)
            G.mark(src)
        elif kind == "spp3":
            _, src, dsts, ks = e
            if not all(G.is_init(d) for d in dsts):       # a pool output nobody consumed: route the others one by one
                for d, k in zip(dsts, ks):
                    if G.is_init(d):
                        gs, acc = G.writable(src)
                        ops.maxpool_bwd(src, G.view(d), k, 1, k // 2, gs, acc)
                        G.mark(src)
                return
            gs, acc = G.writable(src)
            ops.spp_bwd(src, [G.view(d) for d in dsts], ks, gs, acc)
            G.mark(src)
        elif kind == "resize":
            _, src, dst, factor = e
            if not G.is_init(dst):
                return
            gs, acc = G.writable(src)
            if factor == 2:
                ops.upsample2x_bwd(G.view(dst), gs, acc)
            else:
                ops.add_into(gs, G.view(dst), acc)
            G.mark(src)
        else:
            raise RuntimeError(f"unknown tape entry {kind}")

    for i, e in enumerate(reversed(tape)):
        handle(e)
        if on_entry is not None:
            on_entry(i, sink)
    sink.finish()
    ctx.tape = None            # the tape's raw buffers now hold gradients: a second backward would be wrong
    return G
