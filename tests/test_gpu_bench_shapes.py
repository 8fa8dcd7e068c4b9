"""Conv forward / dgrad / wgrad AT THE BENCHMARKED SIZES (N = 32; 800x800, 400x400, 200x200, 100x100 maps — the tile
index, pitch and TMA-coordinate ranges of bench.py's workload, 20.5 M-row GEMMs) under a checker:
  forward   vs the library's CUDA-core direct convolution (ryolo_conv2d_reference, itself checked against torch CPU on
            the small shapes of tests/test_gpu_conv.py), every output element;
  dgrad     stride 1: the same checker on the mirrored / transposed weights; stride 2: torch CPU conv_transpose2d on the
            first and the last image (dgrad is per-image independent; the last image has the largest coordinates);
  wgrad     vs fp32 torch.matmul per tap over all 32 images (TF32 off; operands are bf16-exact so products are exact).
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

N = 32
FWD = [  # H, W, Cin, Cout, k, stride          (yolov4 @800^2)
    (800, 800, 32, 32, 1, 1),     # the stem as the kernel sees it: K = 32 im2col columns
    (800, 800, 32, 64, 3, 2),
    (400, 400, 64, 32, 1, 1),
    (400, 400, 32, 32, 3, 1),
    (400, 400, 64, 128, 3, 2),
    (200, 200, 64, 64, 3, 1),
    (200, 200, 128, 64, 1, 1),
    (100, 100, 128, 128, 3, 1),
    (50, 50, 256, 512, 3, 1),     # N = 256 tiles, 36 K blocks: runs as cluster pairs with weight-tile multicast
    (25, 25, 512, 1024, 3, 1),    # 160 m-tiles x 4 n-tiles, 72 K blocks (pairs; 5 tiles per image)
]


def _mk(gen, H, W, Cin, Cout, k):
    x = torch.randn(N, H, W, Cin, generator=gen, device="cuda").bfloat16()
    w = torch.randn(Cout, Cin, k, k, generator=gen, device="cuda") * (2.0 / (Cin * k * k)) ** 0.5
    return x, w


def _maxdiff(a, b, chunk=8):
    """max |a - b| and max |b| over big bf16 NHWC tensors without materialising fp32 copies of everything."""
    d = m = 0.0
    for i in range(0, a.shape[0], chunk):
        x, y = a[i:i + chunk].float(), b[i:i + chunk].float()
        d = max(d, float((x - y).abs().max()))
        m = max(m, float(y.abs().max()))
    return d, m


@pytest.mark.parametrize("shape", FWD, ids=lambda s: "x".join(map(str, s)))
def test_forward_at_bench_shape(shape):
    from ryolo_b200 import ops
    H, W, Cin, Cout, k, s = shape
    gen = torch.Generator(device="cuda").manual_seed(sum(shape))
    x, w = _mk(gen, H, W, Cin, Cout, k)
    xa, wp = ops.Act(x), ops.pack_weights(w)
    out = ops.conv2d(xa, wp, Cout, k, s)
    chk = ops.conv2d(xa, wp, Cout, k, s, reference=True)
    torch.cuda.synchronize()
    d, m = _maxdiff(out.buf, chk.buf)
    assert m > 0.5 and d <= 1.6e-2 * m, (d, m)           # bf16 output rounding (2^-8) of both + fp32 summation order
    # the far corner of the last image and the channel tail are where a wrapped tile index would land
    assert float((out.buf[-1, -1, -1].float() - chk.buf[-1, -1, -1].float()).abs().max()) <= 1.6e-2 * m


@pytest.mark.parametrize("shape", [(400, 400, 32, 32, 3, 1), (400, 400, 64, 32, 1, 1), (200, 200, 64, 64, 3, 1),
                                   (100, 100, 128, 128, 3, 1), (50, 50, 256, 512, 3, 1)], ids=lambda s: "x".join(map(str, s)))
def test_dgrad_stride1_at_bench_shape(shape):
    from ryolo_b200 import ops
    H, W, Cin, Cout, k, s = shape
    gen = torch.Generator(device="cuda").manual_seed(1 + sum(shape))
    _, w = _mk(gen, 8, 8, Cin, Cout, k)
    dy = torch.randn(N, H, W, Cout, generator=gen, device="cuda").bfloat16()
    dx = ops.Act(torch.empty(N, H, W, Cin, device="cuda", dtype=torch.bfloat16))
    ops.conv2d_dgrad(ops.Act(dy), ops.pack_weights(w, transpose=True), Cin, k, 1, dx, accumulate=False)
    # dx = conv(dy, w') with w'[ci][co][kh][kw] = w[co][ci][k-1-kh][k-1-kw]
    wflip = w.flip(2, 3).permute(1, 0, 2, 3).contiguous()
    chk = ops.conv2d(ops.Act(dy), ops.pack_weights(wflip), Cin, k, 1, reference=True)
    torch.cuda.synchronize()
    d, m = _maxdiff(dx.buf, chk.buf)
    assert m > 0.5 and d <= 1.6e-2 * m, (d, m)


@pytest.mark.parametrize("shape", [(800, 800, 32, 64, 3, 2), (400, 400, 64, 128, 3, 2)],
                         ids=lambda s: "x".join(map(str, s)))
def test_dgrad_stride2_at_bench_shape(shape):
    from ryolo_b200 import ops
    H, W, Cin, Cout, k, s = shape
    gen = torch.Generator(device="cuda").manual_seed(2 + sum(shape))
    _, w = _mk(gen, 8, 8, Cin, Cout, k)
    dy = torch.randn(N, H // 2, W // 2, Cout, generator=gen, device="cuda").bfloat16()
    dx = ops.Act(torch.full((N, H, W, Cin), 3.0, device="cuda").bfloat16())
    ops.conv2d_dgrad(ops.Act(dy), ops.pack_weights(w, transpose=True), Cin, k, 2, dx, accumulate=False)
    torch.cuda.synchronize()
    wr = w.bfloat16().float().cpu()
    for n in (0, N - 1):
        ref = F.conv_transpose2d(dy[n:n + 1].float().cpu().permute(0, 3, 1, 2), wr, None, 2, 1, output_padding=1)
        ref = ref.permute(0, 2, 3, 1)
        got = dx.buf[n:n + 1].float().cpu()
        assert ref.shape == got.shape
        assert float((got - ref).abs().max()) <= 1.6e-2 * float(ref.abs().max()), n


@pytest.mark.parametrize("shape", [(800, 800, 32, 32, 1, 1), (800, 800, 32, 64, 3, 2), (400, 400, 32, 32, 3, 1),
                                   (400, 400, 64, 32, 1, 1), (200, 200, 64, 64, 3, 1), (100, 100, 128, 128, 3, 1)],
                         ids=lambda s: "x".join(map(str, s)))
def test_wgrad_at_bench_shape(shape):
    from ryolo_b200 import ops
    H, W, Cin, Cout, k, s = shape
    gen = torch.Generator(device="cuda").manual_seed(3 + sum(shape))
    x, _ = _mk(gen, H, W, Cin, Cout, k)
    Ho, Wo = ops.out_hw(H, W, k, s)
    dy = torch.randn(N, Ho, Wo, Cout, generator=gen, device="cuda").bfloat16()
    dwk = torch.zeros(Cout * k * k * Cin, device="cuda")
    ops.conv2d_wgrad(ops.Act(x), ops.Act(dy), Cout, k, s, dwk)
    torch.cuda.synchronize()
    got = dwk.view(Cout, k * k, Cin)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        pad = (k - 1) // 2
        ref = torch.zeros(Cout, k * k, Cin, device="cuda")
        for n0 in range(0, N, 4):                       # chunks of 4 images keep the fp32 copies small
            xp = F.pad(x[n0:n0 + 4].float(), (0, 0, pad, pad, pad, pad))
            dyf = dy[n0:n0 + 4].float().reshape(-1, Cout)
            for kh in range(k):
                for kw in range(k):
                    xs = xp[:, kh:kh + s * Ho:s, kw:kw + s * Wo:s, :].reshape(-1, Cin)
                    ref[:, kh * k + kw, :] += dyf.t() @ xs
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    err = float((got - ref).abs().max() / ref.abs().max())
    assert err < 2e-3, err                               # fp32 accumulation over 5-20 M pixels, split-K atomics
