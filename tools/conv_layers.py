"""Per-layer timing of the conv kernel inside one yolov4 800x800 bs=32 train-mode forward.
Writes gpurun_out/conv_layers.json (M, N, K, ms, TFLOP/s per launch, in execution order)."""
import json, os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
import ryolo_b200 as R
from ryolo_b200 import ops
from bench import CFG, NC, S, weights_init_normal

bs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ver = sys.argv[2] if len(sys.argv) > 2 else "yolov4"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
torch.manual_seed(42)
m = R.Yolo(NC if ver == "yolov4" else 16, CFG, "csl", ver)
m.apply(weights_init_normal)
m = m.cuda().train()
img = torch.rand(bs, 3, S, S, device="cuda")
for _ in range(2):
    m(img, training=True)
torch.cuda.synchronize()
ops.PROFILE = []
for _ in range(steps):
    m(img, training=True)
torch.cuda.synchronize()
prof, ops.PROFILE = ops.PROFILE, None
n = len(prof) // steps
rows = []
for i in range(n):
    tag = prof[i][0]
    ms = sorted(prof[i + s * n][1].elapsed_time(prof[i + s * n][2]) for s in range(steps))[steps // 2]
    _, M, N, K = tag
    rows.append(dict(i=i, M=M, N=N, K=K, ms=ms, tflops=2.0 * M * N * K / ms / 1e9))
tot = sum(r["ms"] for r in rows)
print(f"{ver} bs={bs}: {n} conv launches, {tot:.2f} ms, {sum(2.0*r['M']*r['N']*r['K'] for r in rows)/tot/1e9:.1f} TFLOP/s")
for r in rows:
    print(f"{r['i']:3d} M={r['M']:9d} N={r['N']:5d} K={r['K']:5d} {r['ms']:7.3f} ms {r['tflops']:7.1f} TF/s")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"conv_layers_{ver}.json"), "w"))
