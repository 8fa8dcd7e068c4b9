"""Does a TMA box whose inner (channel) extent overhangs the tensor (Cin = 32 under a 64-channel box) run on a slow
path?  Times the same 1x1 / 3x3 convs and weight gradients on 32x400x400 maps with 32 real channels vs 64 real channels
(twice the bytes, no overhang) and with the 32 channels living in a pitch-64 buffer."""
import json, os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
from ryolo_b200 import ops

def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[n // 2]

N, H, W = 32, 400, 400
out = {}
buf64 = torch.randn(N, H, W, 64, device="cuda").bfloat16()
buf32 = torch.randn(N, H, W, 32, device="cuda").bfloat16()
dy32 = torch.randn(N, H, W, 32, device="cuda").bfloat16()
dy64 = torch.randn(N, H, W, 64, device="cuda").bfloat16()
for k in (1, 3):
    for name, x in (("cin32_pitch32", ops.Act(buf32)), ("cin32_pitch64", ops.Act(buf64, 32, 0)), ("cin64", ops.Act(buf64))):
        for Cout, dy in ((32, dy32), (64, dy64)):
            w = ops.pack_weights(torch.randn(Cout, x.C, k, k, device="cuda") * 0.05)
            o = ops.Act.empty(N, H, W, Cout, "cuda")
            ms = t(lambda: ops.conv2d(x, w, Cout, k, 1, out=o))
            gb = (x.P * x.C * 2 + x.P * Cout * 2) / 1e9
            out[f"fwd_k{k}_{name}_cout{Cout}"] = dict(ms=ms, alg_gbs=gb / ms * 1e3)
            dwk = torch.zeros(Cout * k * k * x.C, device="cuda")
            ms = t(lambda: ops.conv2d_wgrad(x, ops.Act(dy), Cout, k, 1, dwk))
            out[f"wgrad_k{k}_{name}_cout{Cout}"] = dict(ms=ms, alg_gbs=gb / ms * 1e3)
for k, v in out.items():
    print(f"{k:40s} {v['ms']:8.3f} ms  {v['alg_gbs']:8.1f} GB/s (algorithmic)")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "oob_probe.json"), "w"), indent=1)
