"""GPU parity of the tcgen05 implicit-GEMM conv (csrc/conv.cu) and its HBM-bound companions
(csrc/elementwise.cu) against torch CPU fp32 on the same bf16-rounded operands, and against the
library's own CUDA-core checker kernel.  Tolerance: bf16 output rounding (rel 2^-8) + fp32 accumulation."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _mk(gen, N, H, W, Cin, Cout, k):
    x = torch.randn(N, H, W, Cin, generator=gen).bfloat16()
    w = (torch.randn(Cout, Cin, k, k, generator=gen) * (2.0 / (Cin * k * k)) ** 0.5)
    return x, w


def _ref_conv(x_nhwc_bf16, w_oihw, stride):
    k = w_oihw.shape[2]
    y = F.conv2d(x_nhwc_bf16.float().permute(0, 3, 1, 2), w_oihw.bfloat16().float(), None, stride, (k - 1) // 2)
    return y.permute(0, 2, 3, 1).contiguous()


SHAPES = [  # N, H, W, Cin, Cout, k, stride
    (2, 25, 25, 64, 128, 3, 1),
    (1, 16, 16, 64, 64, 1, 1),
    (2, 50, 50, 128, 256, 3, 2),
    (3, 13, 13, 256, 512, 3, 1),
    (2, 26, 26, 128, 64, 1, 1),
    (1, 52, 52, 32, 64, 3, 2),     # Cin < 64: OOB-filled K
    (1, 40, 40, 32, 32, 3, 1),
    (1, 100, 100, 64, 32, 1, 1),
    (1, 7, 9, 192, 96, 3, 1),      # ragged everything
    (2, 25, 25, 1024, 512, 1, 1),
    (1, 25, 25, 512, 1024, 3, 1),
]


@pytest.mark.parametrize("shape", SHAPES)
def test_conv_raw(shape, epi):
    from ryolo_b200 import ops
    N, H, W, Cin, Cout, k, s = shape
    gen = torch.Generator().manual_seed(sum(shape))
    x, w = _mk(gen, N, H, W, Cin, Cout, k)
    ref = _ref_conv(x, w, s)
    xa = ops.Act(x.cuda())
    wp = ops.pack_weights(w.cuda())
    out = ops.conv2d(xa, wp, Cout, k, s)
    chk = ops.conv2d(xa, wp, Cout, k, s, reference=True)
    torch.cuda.synchronize()
    got, got_chk = out.torch().float().cpu(), chk.torch().float().cpu()
    tol = 2e-2 * ref.abs().max()
    assert (got_chk - ref).abs().max() < tol, "CUDA-core checker disagrees with torch CPU"
    assert (got - ref).abs().max() < tol, f"tcgen05 conv max err {(got - ref).abs().max()} vs tol {tol}"
    assert (got - got_chk).abs().max() <= 1.5e-2 * ref.abs().max()


@pytest.mark.parametrize("shape", SHAPES + [(5, 25, 25, 256, 512, 3, 1), (3, 20, 24, 128, 96, 1, 1)])
def test_conv_operand_ring_variants(shape, issue):
    """Cluster pairs (odd tile counts leave rank 1 a ghost tile), grouped K blocks and the plain ring against the
    CUDA-core checker: raw output, affine + activation + residual epilogue, and an accumulating dgrad."""
    from ryolo_b200 import ops
    N, H, W, Cin, Cout, k, s = shape
    gen = torch.Generator().manual_seed(sum(shape) + 1)
    x, w = _mk(gen, N, H, W, Cin, Cout, k)
    xa, wp = ops.Act(x.cuda()), ops.pack_weights(w.cuda())
    got = ops.conv2d(xa, wp, Cout, k, s).torch().float().cpu()
    chk = ops.conv2d(xa, wp, Cout, k, s, reference=True).torch().float().cpu()
    ref = _ref_conv(x, w, s)
    assert (chk - ref).abs().max() < 2e-2 * ref.abs().max()
    assert (got - chk).abs().max() <= 1.5e-2 * ref.abs().max()
    Ho, Wo = got.shape[1], got.shape[2]
    scale, shift = (torch.rand(Cout, generator=gen) + 0.5).cuda(), torch.randn(Cout, generator=gen).cuda()
    res = ops.Act(torch.randn(N, Ho, Wo, Cout, generator=gen).bfloat16().cuda())
    kw = dict(scale=scale, shift=shift, act="mish", residual=res)
    got2 = ops.conv2d(xa, wp, Cout, k, s, **kw).torch().float().cpu()
    chk2 = ops.conv2d(xa, wp, Cout, k, s, reference=True, **kw).torch().float().cpu()
    assert (got2 - chk2).abs().max() <= 2e-2 * chk2.abs().max()
    # dgrad (the same kernel with mirrored taps), accumulating into a pre-filled gradient
    dy = torch.randn(N, Ho, Wo, Cout, generator=gen).bfloat16()
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    F.conv2d(xr, w.bfloat16().float(), None, s, (k - 1) // 2).backward(dy.float().permute(0, 3, 1, 2))
    dx_ref = xr.grad.permute(0, 2, 3, 1)
    dx = ops.Act(torch.full((N, H, W, Cin), 1.0).bfloat16().cuda())
    ops.conv2d_dgrad(ops.Act(dy.cuda()), ops.pack_weights(w.cuda(), transpose=True), Cin, k, s, dx, accumulate=True)
    assert (dx.torch().float().cpu() - 1.0 - dx_ref).abs().max() < 2e-2 * dx_ref.abs().max() + 2e-2


@pytest.mark.parametrize("act", ["linear", "leaky", "mish", "swish"])
def test_conv_epilogue_scale_shift_act_residual_concat_slice(act, epi):
    from ryolo_b200 import ops
    gen = torch.Generator().manual_seed(5)
    N, H, W, Cin, Cout = 2, 20, 20, 64, 64
    x, w = _mk(gen, N, H, W, Cin, Cout, 3)
    scale, shift = torch.rand(Cout, generator=gen) + 0.5, torch.randn(Cout, generator=gen)
    res = torch.randn(N, H, W, Cout, generator=gen).bfloat16()
    y = _ref_conv(x, w, 1) * scale + shift
    y = {"linear": y, "leaky": F.leaky_relu(y, 0.1), "mish": F.mish(y), "swish": F.silu(y)}[act] + res.float()
    big = torch.full((N, H, W, 160), 7.0).bfloat16().cuda()
    # input is itself a slice of a wider buffer; output goes to channels [32, 96) of `big`
    xin = torch.zeros(N, H, W, 96).bfloat16()
    xin[..., 16:80] = x
    out = ops.conv2d(ops.Act(xin.cuda(), 64, 16), ops.pack_weights(w.cuda()), Cout, 3, 1, out=ops.Act(big, 64, 32),
                     scale=scale.cuda(), shift=shift.cuda(), act=act, residual=ops.Act(res.cuda()))
    got = out.torch().float().cpu()
    assert (got - y).abs().max() < 2e-2 * y.abs().max()
    b = big.float().cpu()
    assert (b[..., :32] == 7).all() and (b[..., 96:] == 7).all()      # neighbours of the slice untouched


@pytest.mark.parametrize("na,ch,gs,Cin", [(3, 187, 13, 256), (18, 8, 26, 128), (3, 201, 7, 64)])
def test_conv_head_layout(na, ch, gs, Cin):
    from ryolo_b200 import ops
    gen = torch.Generator().manual_seed(na)
    Cout = na * ch
    x, w = _mk(gen, 2, gs, gs, Cin, Cout, 1)
    bias = torch.randn(Cout, generator=gen)
    ref = _ref_conv(x, w, 1) + bias                                           # [N,gs,gs,na*ch]
    ref = ref.view(2, gs, gs, na, ch).permute(0, 3, 1, 2, 4).contiguous()     # model/yololayer.py:25
    out = ops.conv2d(ops.Act(x.cuda()), ops.pack_weights(w.cuda()), Cout, 1, 1, shift=bias.cuda(), head=(na, ch))
    assert out.shape == ref.shape and out.dtype == torch.float32
    assert (out.cpu() - ref).abs().max() < 5e-3 * ref.abs().max()


@pytest.mark.parametrize("shape", [(2, 40, 40, 32, 32, 3, 1), (3, 25, 25, 128, 96, 1, 1), (2, 26, 30, 64, 160, 3, 2),
                                   (1, 50, 50, 64, 320, 1, 1)])
def test_conv_fused_bn_statistics(shape, epi):
    """Train-mode BatchNorm statistics + finalize fused into the conv epilogue (EPI_RAW): scale/shift/mean/invstd and
    the running statistics equal nn.BatchNorm2d's on the stored (bf16) conv output."""
    from ryolo_b200 import ops
    N, H, W, Cin, Cout, k, s = shape
    gen = torch.Generator().manual_seed(11 + sum(shape))
    x, w = _mk(gen, N, H, W, Cin, Cout, k)
    bn = torch.nn.BatchNorm2d(Cout)
    bn.weight.data = torch.rand(Cout, generator=gen) + 0.5
    bn.bias.data = torch.randn(Cout, generator=gen)
    bn = bn.cuda().train()
    part = torch.zeros(4 * Cout, device="cuda")
    ctr = torch.zeros(1, dtype=torch.int32, device="cuda")
    scale, shift, mean, invstd = (torch.empty(Cout, device="cuda") for _ in range(4))
    raw = ops.conv2d(ops.Act(x.cuda()), ops.pack_weights(w.cuda()), Cout, k, s,
                     bn=ops.bn_fuse(part, ctr, bn, scale, shift, mean, invstd))
    torch.cuda.synchronize()
    assert float(part.abs().max()) == 0.0                          # the scratch is handed back zeroed
    y = raw.torch().float().cpu()                                  # [N,Ho,Wo,Cout] as stored
    ref = _ref_conv(x, w, s)
    assert (y - ref).abs().max() < 2e-2 * ref.abs().max()
    m = y.reshape(-1, Cout).mean(0)
    v = y.reshape(-1, Cout).var(0, unbiased=False)
    assert torch.allclose(mean.cpu(), m, rtol=1e-4, atol=1e-5)
    assert torch.allclose(invstd.cpu(), (v + bn.eps).rsqrt(), rtol=1e-4, atol=1e-5)
    sc = bn.weight.data.cpu() * (v + bn.eps).rsqrt()
    assert torch.allclose(scale.cpu(), sc, rtol=1e-4, atol=1e-5)
    assert torch.allclose(shift.cpu(), bn.bias.data.cpu() - m * sc, rtol=1e-3, atol=1e-4)
    cnt = y.numel() // Cout
    assert torch.allclose(bn.running_mean.cpu(), 0.1 * m, rtol=1e-4, atol=1e-5)
    assert torch.allclose(bn.running_var.cpu(), 0.9 + 0.1 * v * cnt / (cnt - 1), rtol=1e-4, atol=1e-5)


def test_stem_im2col_conv():
    from ryolo_b200 import ops
    gen = torch.Generator().manual_seed(2)
    img = torch.rand(2, 3, 48, 40, generator=gen)
    w = torch.randn(32, 3, 3, 3, generator=gen) * 0.2
    ref = F.conv2d(img.bfloat16().float(), w.bfloat16().float(), None, 1, 1).permute(0, 2, 3, 1)
    cols = ops.stem_im2col(img.cuda())
    out = ops.conv2d(cols, ops.pack_weights(w.cuda(), stem=True), 32, 1, 1)
    assert (out.torch().float().cpu() - ref).abs().max() < 2e-2 * ref.abs().max()
    # yolov5 stem: 6x6 / stride 2 / pad 2 (model/backbone.py:42)
    w6 = torch.randn(64, 3, 6, 6, generator=gen) * 0.1
    ref6 = F.conv2d(img.bfloat16().float(), w6.bfloat16().float(), None, 2, 2).permute(0, 2, 3, 1)
    cols6 = ops.stem_im2col(img.cuda(), 6, 2)
    assert (cols6.H, cols6.W, cols6.C) == (24, 20, 128)
    out6 = ops.conv2d(cols6, ops.pack_weights(w6.cuda(), stem=True), 64, 1, 1)
    assert (out6.torch().float().cpu() - ref6).abs().max() < 2e-2 * ref6.abs().max()


def test_bn_train_stats_apply():
    from ryolo_b200 import ops
    gen = torch.Generator().manual_seed(4)
    N, H, W, C = 4, 30, 30, 96
    x = (torch.randn(N, H, W, C, generator=gen) * 2 + 0.5).bfloat16()
    gamma, beta = torch.rand(C, generator=gen) + 0.5, torch.randn(C, generator=gen)
    rm, rv = torch.zeros(C), torch.ones(C)
    bn = torch.nn.BatchNorm2d(C)
    bn.weight.data, bn.bias.data = gamma.clone(), beta.clone()
    bn.train()
    ref = F.mish(bn(x.float().permute(0, 3, 1, 2))).permute(0, 2, 3, 1)
    xa = ops.Act(x.cuda())
    s, q = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    sc, sh = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    rmc, rvc, nb = rm.cuda(), rv.cuda(), torch.zeros(1, dtype=torch.int64, device="cuda")
    ops.bn_stats(xa, s, q)
    ops.bn_finalize(s, q, N * H * W, gamma.cuda(), beta.cuda(), 1e-5, 0.1, rmc, rvc, nb, sc, sh)
    out = ops.scale_shift_act(xa, sc, sh, "mish", ops.Act.empty(N, H, W, C, "cuda"))
    assert (out.torch().float().cpu() - ref).abs().max() < 2e-2 * ref.abs().max()
    assert torch.allclose(rmc.cpu(), bn.running_mean, rtol=1e-4, atol=1e-5)
    assert torch.allclose(rvc.cpu(), bn.running_var, rtol=1e-4, atol=1e-5)
    assert int(nb) == 1


def test_maxpool_and_upsample():
    from ryolo_b200 import ops
    gen = torch.Generator().manual_seed(6)
    x = torch.randn(2, 25, 25, 64, generator=gen).bfloat16()
    xa = ops.Act(x.cuda())
    xn = x.float().permute(0, 3, 1, 2)
    for k in (5, 9, 13):
        ref = F.max_pool2d(xn, k, 1, k // 2).permute(0, 2, 3, 1)
        assert torch.equal(ops.maxpool(xa, k, 1, k // 2).torch().float().cpu(), ref)
    ref = F.max_pool2d(xn[..., :24, :24], 2, 2).permute(0, 2, 3, 1)
    x24 = ops.Act(x[:, :24, :24].contiguous().cuda())
    assert torch.equal(ops.maxpool(x24, 2, 2, 0).torch().float().cpu(), ref)
    up = ops.resize_copy(xa, 2).torch().float().cpu()
    assert torch.equal(up, F.interpolate(xn, scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1))
