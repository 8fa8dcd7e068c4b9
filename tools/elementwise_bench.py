"""A/B timing of the HBM-bound BatchNorm / activation passes (scale_shift_act, bn_act_bwd reduce + apply) on the layer
shapes of the yolov4 800x800 bs=32 training step, old kernels (knobs ssa=0, bn_bwd=2) against the register
double-buffered ones (ssa=1, bn_bwd=3).  CUDA events on the current stream, operands far larger than L2 or L2 flushed
between launches.  Writes gpurun_out/elementwise_bench.json + a text table.

    python tools/elementwise_bench.py [batch]"""
import json
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import ryolo_b200._lib as L  # noqa: E402
from ryolo_b200 import ops  # noqa: E402

bs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
# (pixels per image, channels, activation, layers of that shape in yolov4/csl/nc2): model/backbone.py:4-36, neck.py:47-81
SHAPES = [(640000, 32, "mish", 1), (160000, 64, "mish", 2), (160000, 32, "mish", 4), (40000, 128, "mish", 2),
          (40000, 64, "mish", 7), (10000, 256, "mish", 2), (10000, 128, "mish", 19), (2500, 512, "mish", 2),
          (2500, 256, "mish", 19), (625, 1024, "mish", 2), (625, 512, "mish", 11),
          (625, 512, "leaky", 9), (625, 1024, "leaky", 6), (2500, 256, "leaky", 9), (2500, 512, "leaky", 5),
          (10000, 128, "leaky", 5), (10000, 256, "leaky", 3)]
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=5):
    fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


rows, tot = [], {}
for ppi, C, act, count in SHAPES:
    P = ppi * bs
    x = torch.randn(P, C, device="cuda").bfloat16().view(1, 1, P, C)
    d = torch.randn(P, C, device="cuda").bfloat16().view(1, 1, P, C)
    o = torch.empty_like(x)
    sc = torch.rand(C, device="cuda") + 0.5
    sh = torch.randn(C, device="cuda") * 0.1
    mu, inv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    sums = torch.zeros(2 * C, device="cuda")
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    rec = dict(pixels=P, C=C, act=act, layers=count, elements=P * C)
    for name, knobs in (("old", dict(ssa=0, bn_bwd=2)), ("new", dict(ssa=1, bn_bwd=3))):
        L.tune(**knobs)
        t_f = timeit(lambda: ops.scale_shift_act(ops.Act(x), sc, sh, act, ops.Act(o)))
        t_r = timeit(lambda: ops.scale_shift_act(ops.Act(x), sc, sh, act, ops.Act(o), residual=ops.Act(d)))
        t_b = timeit(lambda: ops.bn_act_bwd(ops.Act(d), ops.Act(x), sc, sh, mu, inv, act, sums, ops.Act(o), dg, db))
        rec[name] = dict(ssa_ms=t_f, ssa_res_ms=t_r, bwd_ms=t_b, ssa_tbs=P * C * 4 / t_f / 1e9,
                         ssa_res_tbs=P * C * 6 / t_r / 1e9, bwd_tbs=P * C * 10 / t_b / 1e9)
        tot[name + "_fwd"] = tot.get(name + "_fwd", 0.0) + t_f * count
        tot[name + "_bwd"] = tot.get(name + "_bwd", 0.0) + t_b * count
    rows.append(rec)
    print(f"{P:>10} x {C:<5} {act:<6} x{count:<3} fwd {rec['old']['ssa_ms']:.3f} -> {rec['new']['ssa_ms']:.3f} ms "
          f"({rec['old']['ssa_tbs']:.2f} -> {rec['new']['ssa_tbs']:.2f} TB/s) | +res {rec['old']['ssa_res_tbs']:.2f} -> "
          f"{rec['new']['ssa_res_tbs']:.2f} TB/s | bwd {rec['old']['bwd_ms']:.3f} -> {rec['new']['bwd_ms']:.3f} ms "
          f"({rec['old']['bwd_tbs']:.2f} -> {rec['new']['bwd_tbs']:.2f} TB/s)", flush=True)
print("per step (shape time x layer count): ", {k: round(v, 3) for k, v in tot.items()})
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(dict(batch=bs, rows=rows, per_step_ms=tot), open(os.path.join(ROOT, "gpurun_out", "elementwise_bench.json"), "w"))
