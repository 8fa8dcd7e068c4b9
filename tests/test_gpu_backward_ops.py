"""GPU parity of the backward kernels (dgrad on the conv kernel, tcgen05 wgrad, BN/act/pool/upsample backward)
against torch CPU autograd on the same bf16-rounded operands."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

SHAPES = [  # N, H, W, Cin, Cout, k, stride
    (2, 25, 25, 64, 128, 3, 1),
    (2, 16, 16, 64, 64, 1, 1),
    (2, 50, 50, 128, 256, 3, 2),
    (2, 13, 13, 256, 512, 3, 1),
    (1, 52, 52, 32, 64, 3, 2),
    (1, 40, 40, 32, 32, 3, 1),
    (2, 9, 7, 192, 96, 3, 1),
    (1, 25, 25, 1024, 512, 1, 1),
    (1, 25, 25, 512, 1024, 3, 1),
    (1, 27, 27, 64, 128, 3, 2),      # odd size, stride 2
]


def _case(shape, seed=0):
    N, H, W, Cin, Cout, k, s = shape
    gen = torch.Generator().manual_seed(seed + sum(shape))
    x = torch.randn(N, H, W, Cin, generator=gen).bfloat16()
    w = (torch.randn(Cout, Cin, k, k, generator=gen) * (2.0 / (Cin * k * k)) ** 0.5)
    Ho, Wo = (H + 2 * ((k - 1) // 2) - k) // s + 1, (W + 2 * ((k - 1) // 2) - k) // s + 1
    dy = torch.randn(N, Ho, Wo, Cout, generator=gen).bfloat16()
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = w.bfloat16().float().requires_grad_(True)
    y = F.conv2d(xr, wr, None, s, (k - 1) // 2)
    y.backward(dy.float().permute(0, 3, 1, 2))
    return x, w, dy, xr.grad.permute(0, 2, 3, 1).contiguous(), wr.grad


@pytest.mark.parametrize("shape", SHAPES)
def test_dgrad(shape, epi):
    from ryolo_b200 import ops
    N, H, W, Cin, Cout, k, s = shape
    x, w, dy, dx_ref, _ = _case(shape)
    wt = ops.pack_weights(w.cuda(), transpose=True)
    dx = ops.Act(torch.full((N, H, W, Cin), 3.0).bfloat16().cuda())
    ops.conv2d_dgrad(ops.Act(dy.cuda()), wt, Cin, k, s, dx, accumulate=False)
    got = dx.torch().float().cpu()
    tol = 2e-2 * dx_ref.abs().max()
    assert (got - dx_ref).abs().max() < tol
    ops.conv2d_dgrad(ops.Act(dy.cuda()), wt, Cin, k, s, dx, accumulate=True)      # dx += same thing
    assert (dx.torch().float().cpu() - 2 * dx_ref).abs().max() < 2 * tol
    # gradient into a channel slice of a wider (concat) buffer: the neighbours stay untouched
    big = torch.full((N, H, W, Cin + 48), 5.0).bfloat16().cuda()
    ops.conv2d_dgrad(ops.Act(dy.cuda()), wt, Cin, k, s, ops.Act(big, Cin, 16), accumulate=False)
    b = big.float().cpu()
    assert (b[..., 16:16 + Cin] - dx_ref).abs().max() < tol
    assert (b[..., :16] == 5).all() and (b[..., 16 + Cin:] == 5).all()


@pytest.mark.parametrize("shape", SHAPES)
def test_wgrad(shape):
    from ryolo_b200 import ops
    N, H, W, Cin, Cout, k, s = shape
    x, w, dy, _, dw_ref = _case(shape, seed=1)
    dwk = torch.zeros(Cout * k * k * Cin, device="cuda")                          # K-major [Cout][tap][Cin]
    ops.conv2d_wgrad(ops.Act(x.cuda()), ops.Act(dy.cuda()), Cout, k, s, dwk)
    got = ops.wgrad_to_oihw(dwk, Cout, Cin, k).cpu()
    assert (got - dw_ref).abs().max() < 1e-2 * dw_ref.abs().max(), (got - dw_ref).abs().max() / dw_ref.abs().max()
    ops.conv2d_wgrad(ops.Act(x.cuda()), ops.Act(dy.cuda()), Cout, k, s, dwk)     # accumulates
    assert (ops.wgrad_to_oihw(dwk, Cout, Cin, k).cpu() - 2 * dw_ref).abs().max() < 2e-2 * dw_ref.abs().max()


def test_wgrad_channel_slices_and_padded_dy():
    """x and dy are slices of wider buffers; dy carries padded channels beyond Cout (head case)."""
    from ryolo_b200 import ops
    shape = (2, 20, 20, 64, 40, 1, 1)
    x, w, dy, _, dw_ref = _case(shape, seed=2)
    xb = torch.randn(2, 20, 20, 160).bfloat16()
    xb[..., 32:96] = x
    dyb = torch.zeros(2, 20, 20, 48).bfloat16()
    dyb[..., :40] = dy
    dyb[..., 40:] = 5.0          # garbage in the padding must not leak into rows < Cout
    dwk = torch.zeros(40 * 64, device="cuda")
    ops.conv2d_wgrad(ops.Act(xb.cuda(), 64, 32), ops.Act(dyb.cuda()), 40, 1, 1, dwk)
    assert (ops.wgrad_to_oihw(dwk, 40, 64, 1).cpu() - dw_ref).abs().max() < 1e-2 * dw_ref.abs().max()


def test_wgrad_stem():
    from ryolo_b200 import ops
    gen = torch.Generator().manual_seed(3)
    img = torch.rand(2, 3, 40, 36, generator=gen)
    dy = torch.randn(2, 40, 36, 32, generator=gen).bfloat16()
    w = torch.zeros(32, 3, 3, 3, requires_grad=True)
    y = F.conv2d(img.bfloat16().float(), w, None, 1, 1)
    y.backward(dy.float().permute(0, 3, 1, 2))
    dwk = torch.zeros(32 * ops.stem_kpad(3), device="cuda")
    ops.conv2d_wgrad(ops.stem_im2col(img.cuda()), ops.Act(dy.cuda()), 32, 1, 1, dwk)
    assert (ops.wgrad_to_oihw(dwk, 32, 3, 3, stem=True).cpu() - w.grad).abs().max() < 1e-2 * w.grad.abs().max()
    # 6x6 / stride-2 stem of yolov5
    w6 = torch.zeros(64, 3, 6, 6, requires_grad=True)
    dy6 = torch.randn(2, 20, 18, 64, generator=gen).bfloat16()
    F.conv2d(img.bfloat16().float(), w6, None, 2, 2).backward(dy6.float().permute(0, 3, 1, 2))
    dwk6 = torch.zeros(64 * ops.stem_kpad(6), device="cuda")
    ops.conv2d_wgrad(ops.stem_im2col(img.cuda(), 6, 2), ops.Act(dy6.cuda()), 64, 1, 1, dwk6)
    assert (ops.wgrad_to_oihw(dwk6, 64, 3, 6, stem=True).cpu() - w6.grad).abs().max() < 1e-2 * w6.grad.abs().max()


def test_spp_backward_three_pools_one_launch():
    """ryolo_spp_bwd (5 / 9 / 13 stride-1 pools of one map, gradients in channel slices of one concat buffer, ties in the
    map) against torch autograd, overwriting and accumulating into a slice of a wider gradient buffer."""
    from ryolo_b200 import ops
    gen = torch.Generator().manual_seed(11)
    N, H, W, C = 3, 25, 25, 64
    x = (torch.randn(N, H, W, C, generator=gen) * 2).round().div(2).bfloat16()       # quantised: many ties per window
    xn = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    cat = torch.randn(N, H, W, 4 * C, generator=gen).bfloat16()                      # d cat: slices 13 | 9 | 5 | x
    ks = [5, 9, 13]
    sl = {5: 2 * C, 9: C, 13: 0}
    tot = sum(F.max_pool2d(xn, k, 1, k // 2) * cat[..., sl[k]:sl[k] + C].float().permute(0, 3, 1, 2) for k in ks)
    tot.sum().backward()
    ref = xn.grad.permute(0, 2, 3, 1)
    g = ops.Act(cat.clone().cuda())
    dys = [g.slice(sl[k], C) for k in ks]
    dx = ops.Act(torch.full((N, H, W, C), 7.0).bfloat16().cuda())
    ops.spp_bwd(ops.Act(x.cuda()), dys, ks, dx, False)
    assert (dx.torch().float().cpu() - ref).abs().max() <= 1e-2 * ref.abs().max()
    # accumulate into the x slice of the same gradient buffer (what the SPP block's backward does)
    ops.spp_bwd(ops.Act(x.cuda()), dys, ks, g.slice(3 * C, C), True)
    got = g.buf.float().cpu()
    want = cat[..., 3 * C:].float() + ref
    assert (got[..., 3 * C:] - want).abs().max() <= 1e-2 * want.abs().max()
    assert torch.equal(got[..., :3 * C], cat[..., :3 * C].float())                  # the dy slices are untouched


@pytest.mark.parametrize("act", ["linear", "leaky", "mish", "swish"])
def test_bn_act_backward(act):
    from ryolo_b200 import ops
    gen = torch.Generator().manual_seed(4)
    N, H, W, C = 3, 20, 20, 96
    raw = (torch.randn(N, H, W, C, generator=gen) * 1.5 + 0.3).bfloat16()
    dout = torch.randn(N, H, W, C, generator=gen).bfloat16()
    gamma = (torch.rand(C, generator=gen) + 0.5).requires_grad_(True)
    beta = torch.randn(C, generator=gen).requires_grad_(True)
    xr = raw.float().permute(0, 3, 1, 2).requires_grad_(True)
    y = F.batch_norm(xr, None, None, gamma, beta, True, 0.1, 1e-5)
    z = {"linear": y, "leaky": F.leaky_relu(y, 0.1), "mish": F.mish(y), "swish": F.silu(y)}[act]
    z.backward(dout.float().permute(0, 3, 1, 2))
    mean = raw.float().mean((0, 1, 2))
    var = raw.float().var((0, 1, 2), unbiased=False)
    invstd = torch.rsqrt(var + 1e-5)
    scale = gamma.detach() * invstd
    shift = beta.detach() - mean * scale
    sums = torch.zeros(2 * C, device="cuda")
    draw = ops.Act.empty(N, H, W, C, "cuda")
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")     # gradients accumulate
    ops.bn_act_bwd(ops.Act(dout.cuda()), ops.Act(raw.cuda()), scale.cuda(), shift.cuda(), mean.cuda(), invstd.cuda(), act,
                   sums, draw, dg, db)
    ref = xr.grad.permute(0, 2, 3, 1)
    assert (draw.torch().float().cpu() - ref).abs().max() < 2e-2 * ref.abs().max()
    assert (dg.cpu() - gamma.grad).abs().max() < 5e-3 * gamma.grad.abs().max()
    assert (db.cpu() - beta.grad).abs().max() < 5e-3 * beta.grad.abs().max()


def test_pool_upsample_add_headpack_backward():
    from ryolo_b200 import ops
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(2, 25, 25, 64, generator=gen).bfloat16()
    xn = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    dy = torch.randn(2, 25, 25, 64, generator=gen).bfloat16()
    F.max_pool2d(xn, 5, 1, 2).backward(dy.float().permute(0, 3, 1, 2))
    dx = ops.Act(torch.full((2, 25, 25, 64), 9.0).bfloat16().cuda())
    ops.maxpool_bwd(ops.Act(x.cuda()), ops.Act(dy.cuda()), 5, 1, 2, dx, False)
    ref = xn.grad.permute(0, 2, 3, 1)
    assert (dx.torch().float().cpu() - ref).abs().max() < 1e-2 * ref.abs().max()
    xn.grad = None
    F.max_pool2d(xn, 13, 1, 6).backward(dy.float().permute(0, 3, 1, 2))          # up to 169 routes per element
    ops.maxpool_bwd(ops.Act(x.cuda()), ops.Act(dy.cuda()), 13, 1, 6, dx, True)
    ref13 = ref + xn.grad.permute(0, 2, 3, 1)
    assert (dx.torch().float().cpu() - ref13).abs().max() < 1e-2 * ref13.abs().max()
    # 2x2 / stride 2
    x2 = x[:, :24, :24].contiguous()
    xn = x2.float().permute(0, 3, 1, 2).requires_grad_(True)
    dy2 = torch.randn(2, 12, 12, 64, generator=gen).bfloat16()
    F.max_pool2d(xn, 2, 2).backward(dy2.float().permute(0, 3, 1, 2))
    dx = ops.Act(torch.zeros(2, 24, 24, 64).bfloat16().cuda())
    ops.maxpool_bwd(ops.Act(x2.cuda()), ops.Act(dy2.cuda()), 2, 2, 0, dx, False)
    assert torch.equal(dx.torch().float().cpu(), xn.grad.permute(0, 2, 3, 1))
    ops.maxpool_bwd(ops.Act(x2.cuda()), ops.Act(dy2.cuda()), 2, 2, 0, dx, True)          # accumulates
    assert (dx.torch().float().cpu() - 2 * xn.grad.permute(0, 2, 3, 1)).abs().max() <= 2e-2 * xn.grad.abs().max()
    # upsample x2
    g = torch.randn(2, 20, 20, 32, generator=gen).bfloat16()
    ref = g.float().view(2, 10, 2, 10, 2, 32).sum((2, 4))
    dx = ops.Act(torch.ones(2, 10, 10, 32).bfloat16().cuda())
    ops.upsample2x_bwd(ops.Act(g.cuda()), dx, accumulate=True)
    assert (dx.torch().float().cpu() - (ref + 1)).abs().max() < 2e-2 * ref.abs().max()
    # add_into on slices
    a = torch.randn(2, 6, 6, 64, generator=gen).bfloat16()
    b = torch.randn(2, 6, 6, 32, generator=gen).bfloat16()
    A = ops.Act(a.clone().cuda(), 32, 16)
    ops.add_into(A, ops.Act(b.cuda()), accumulate=True)
    out = A.buf.float().cpu()
    assert (out[..., 16:48] - (a[..., 16:48].float() + b.float())).abs().max() < 2e-2
    assert torch.equal(out[..., :16], a[..., :16].float()) and torch.equal(out[..., 48:], a[..., 48:].float())
    # head gradient pack
    gl = torch.randn(2, 3, 7, 7, 187, generator=gen)
    mul = torch.rand(561, generator=gen) + 0.5
    dbias = torch.zeros(561, device="cuda")
    packed = ops.head_grad_pack(gl.cuda(), 568, mul.cuda(), dbias)
    ref = (gl.permute(0, 2, 3, 1, 4).reshape(2, 7, 7, 561) * mul)
    got = packed.torch().float().cpu()
    assert (got[..., :561] - ref).abs().max() < 1e-2 * ref.abs().max() and (got[..., 561:] == 0).all()
    assert (dbias.cpu() - ref.sum((0, 1, 2))).abs().max() < 1e-3 * ref.sum((0, 1, 2)).abs().max()


def test_sgd_step():
    from ryolo_b200 import ops
    gen = torch.Generator().manual_seed(6)
    p0 = torch.randn(10007, generator=gen)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.SGD([ref], lr=0.01, momentum=0.937, nesterov=True, weight_decay=5e-4)
    p, buf = p0.clone().cuda(), torch.zeros(10007, device="cuda")
    for it in range(3):
        g = torch.randn(10007, generator=gen)
        ref.grad = g.clone()
        opt.step()
        ops.sgd_step(p, g.cuda(), buf, 0.01, 0.937, 5e-4, True, it == 0)
    assert torch.allclose(p.cpu(), ref.detach(), rtol=1e-5, atol=1e-6)


def test_adam_step():
    from ryolo_b200 import ops
    gen = torch.Generator().manual_seed(7)
    p0 = torch.randn(5003, generator=gen)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-3)
    p, m, v = p0.clone().cuda(), torch.zeros(5003, device="cuda"), torch.zeros(5003, device="cuda")
    for it in range(1, 4):
        g = torch.randn(5003, generator=gen)
        ref.grad = g.clone()
        opt.step()
        ops.adam_step(p, g.cuda(), m, v, 1e-3, it)
    assert torch.allclose(p.cpu(), ref.detach(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("act", ["linear", "leaky", "mish", "swish"])
@pytest.mark.parametrize("C,coff,Cbuf", [(32, 0, 32), (96, 8, 128), (256, 64, 384)])
def test_pipelined_elementwise_kernels_many_rows_per_thread(act, C, coff, Cbuf):
    """The register double-buffered kernels (knobs bn_bwd = 3, ssa = 1: resident-capacity grid, several trips per thread,
    odd row counts so that the pair / quad loops, their last trip and the left-over rows all run) on channel slices of
    wider buffers, against torch autograd, and against the original kernels (bn_bwd = 0, ssa = 0) at rounding level."""
    import ryolo_b200._lib as L
    from ryolo_b200 import ops
    gen = torch.Generator().manual_seed(C + len(act))
    N, H, W = 1, 367, 373                                   # 136 891 pixels: 3-8 trips per thread at 296 blocks
    raw = (torch.randn(N, H, W, Cbuf, generator=gen) * 1.5 + 0.3).bfloat16()
    dout = torch.randn(N, H, W, Cbuf, generator=gen).bfloat16()
    resid = torch.randn(N, H, W, Cbuf, generator=gen).bfloat16()
    gamma = (torch.rand(C, generator=gen) + 0.5).requires_grad_(True)
    beta = torch.randn(C, generator=gen).requires_grad_(True)
    xr = raw[..., coff:coff + C].float().permute(0, 3, 1, 2).requires_grad_(True)
    y = F.batch_norm(xr, None, None, gamma, beta, True, 0.1, 1e-5)
    z = {"linear": y, "leaky": F.leaky_relu(y, 0.1), "mish": F.mish(y), "swish": F.silu(y)}[act]
    z.backward(dout[..., coff:coff + C].float().permute(0, 3, 1, 2))
    mean = raw[..., coff:coff + C].float().mean((0, 1, 2))
    invstd = torch.rsqrt(raw[..., coff:coff + C].float().var((0, 1, 2), unbiased=False) + 1e-5)
    scale = gamma.detach() * invstd
    shift = beta.detach() - mean * scale
    dev = [t.cuda() for t in (scale, shift, mean, invstd)]
    rawc, doutc, resc = raw.cuda(), dout.cuda(), resid.cuda()
    got = {}
    try:
        for variant in (3, 0, 13):                          # 13 = variant 3 through the one-launch fused kernel (bn_fuse)
            L.tune(bn_bwd=variant % 10, ssa=1 if variant else 0, bn_fuse=1 if variant == 13 else 0)
            # forward: act(raw * scale + shift) (+ residual) into a slice of a wider buffer
            for use_res in (False, True):
                out = torch.full((N, H, W, Cbuf), 7.0, dtype=torch.bfloat16, device="cuda")
                ops.scale_shift_act(ops.Act(rawc, C, coff), dev[0], dev[1], act, ops.Act(out, C, coff),
                                    residual=ops.Act(resc, C, coff) if use_res else None)
                assert float((out[..., :coff].float() - 7).abs().max() if coff else 0.0) == 0.0      # neighbours intact
                assert float((out[..., coff + C:].float() - 7).abs().max() if coff + C < Cbuf else 0.0) == 0.0
                got[("fwd", use_res, variant)] = out[..., coff:coff + C].float().cpu()
            sums = torch.zeros(2 * C, device="cuda")
            draw = torch.full((N, H, W, Cbuf), 7.0, dtype=torch.bfloat16, device="cuda")
            dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
            ops.bn_act_bwd(ops.Act(doutc.clone(), C, coff), ops.Act(rawc, C, coff), *dev, act, sums, ops.Act(draw, C, coff),
                           dg, db)
            assert float((draw[..., coff + C:].float() - 7).abs().max() if coff + C < Cbuf else 0.0) == 0.0
            got[("bwd", variant)] = (draw[..., coff:coff + C].float().cpu(), dg.cpu(), db.cpu())
    finally:
        L.tune(bn_bwd=3, ssa=1, bn_fuse=0)
    zr = z.detach().permute(0, 2, 3, 1)
    for use_res in (False, True):
        ref = zr + (resid[..., coff:coff + C].float() if use_res else 0.0)
        a, b = got[("fwd", use_res, 3)], got[("fwd", use_res, 0)]
        assert (a - ref).abs().max() < 1e-2 * ref.abs().max()
        assert (a - b).abs().max() <= 2 ** -7 * ref.abs().max()              # both round the same value to bf16
        assert float(((a - b).abs() > 0).float().mean()) < 0.02             # ... and almost always to the same one
    ref = xr.grad.permute(0, 2, 3, 1)
    d3, g3, b3 = got[("bwd", 3)]
    d0, g0, b0 = got[("bwd", 0)]
    df, gf, bf = got[("bwd", 13)]                           # same kernels' bodies, one launch: same rounding
    assert float((df - d3).norm() / ref.norm()) < 2e-3 and (gf - g3).abs().max() < 2e-3 * g3.abs().max()
    # LeakyReLU's kink: the kernels form z = raw * scale + shift, torch (x - mean) * invstd * gamma + beta; where z is
    # within rounding of 0 the derivative flips between 1 and 0.1, so isolated elements differ by 0.9 |d out|
    for other in (ref, d0):
        bad = (d3 - other).abs() > 2e-2 * ref.abs().max()
        assert float(bad.float().mean()) < (1e-4 if act == "leaky" else 0.0) + 1e-12, float(bad.float().mean())
        assert float((d3 - other).norm() / ref.norm()) < 1e-2
    tol = 1.5e-2 if act == "leaky" else 5e-3                 # a flipped element moves its channel's sums by ~|d out|
    assert (g3 - gamma.grad).abs().max() < tol * gamma.grad.abs().max()
    assert (b3 - beta.grad).abs().max() < tol * beta.grad.abs().max()
