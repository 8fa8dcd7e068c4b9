"""Does the HBM-bound BatchNorm backward run UNDER a weight-gradient GEMM (same SMs, two streams), or do the two
time-slice?  Times wgrad alone, bn_act_bwd alone, and both launched together (wgrad on a default-priority stream,
bn_act_bwd on a high-priority one) for a few layer shapes of the yolov4 bs=32 step.  overlap = 1 - (t_both - max) / min."""
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import ryolo_b200._lib as L  # noqa: E402
from ryolo_b200 import ops  # noqa: E402

bs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
lo, hi = torch.cuda.Stream(priority=0), torch.cuda.Stream(priority=-1)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")


def t(fn, iters=5):
    fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


# (H=W, Cin, Cout, k) wgrad layer   |   elementwise layer (pixels/img, C)
CASES = [((100, 128, 128, 3), (10000, 128)), ((50, 256, 256, 3), (2500, 256)), ((200, 64, 64, 3), (40000, 64)),
         ((400, 32, 32, 3), (160000, 32)), ((100, 128, 128, 3), (40000, 64))]
for (hw, cin, cout, k), (ppi, C) in CASES:
    x = ops.Act(torch.randn(bs, hw, hw, cin, device="cuda").bfloat16())
    dy = ops.Act(torch.randn(bs, hw, hw, cout, device="cuda").bfloat16())
    dwk = torch.zeros(cout * k * k * cin, device="cuda")
    P = ppi * bs
    raw = ops.Act(torch.randn(1, 1, P, C, device="cuda").bfloat16())
    dout = ops.Act(torch.randn(1, 1, P, C, device="cuda").bfloat16())
    out = ops.Act(torch.empty(1, 1, P, C, device="cuda", dtype=torch.bfloat16))
    sc, sh = torch.rand(C, device="cuda") + 0.5, torch.zeros(C, device="cuda")
    mu, inv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    sums, dg, db = torch.zeros(2 * C, device="cuda"), torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    reps = 3

    def wg():
        for _ in range(reps):
            ops.conv2d_wgrad(x, dy, cout, k, 1, dwk)

    def ew():
        for _ in range(reps):
            ops.bn_act_bwd(dout, raw, sc, sh, mu, inv, "mish", sums, out, dg, db)

    def both():
        cur = torch.cuda.current_stream()
        lo.wait_stream(cur)
        hi.wait_stream(cur)
        with torch.cuda.stream(lo):
            wg()
        with torch.cuda.stream(hi):
            ew()
        cur.wait_stream(lo)
        cur.wait_stream(hi)

    a, b, c = t(wg), t(ew), t(both)
    print(f"wgrad {hw}x{hw} {cin}->{cout} k{k}: {a:.3f} ms | bn_bwd {P}x{C}: {b:.3f} ms | together {c:.3f} ms | "
          f"sum {a + b:.3f} max {max(a, b):.3f} overlap {1 - (c - max(a, b)) / min(a, b):.2f}", flush=True)
