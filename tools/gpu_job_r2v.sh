#!/bin/bash
set +e
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_decode_loss.py tests/test_gpu_model.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
python - <<'PY'
import torch, sys
sys.path.insert(0, '.')
import ryolo_b200 as R
from tests.util import CFG
from oracle import hotpath as hp
layer = R.YoloCSLLayer(2, hp.make_anchors(CFG["anchors"]), [8, 16, 32])
heads = [torch.randn(32, 3, g, g, 187, device="cuda") for g in (100, 50, 25)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3): layer(heads, training=False)
ts = []
for _ in range(7):
    flush.zero_(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); layer(heads, training=False); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ms = sorted(ts)[3]
byts = sum(h.numel() * 4 for h in heads) + 32 * 3 * (100*100 + 50*50 + 25*25) * 8 * 4
print(f"decode_csl 3 levels bs=32: {ms:.3f} ms  {byts/ms/1e9:.2f} TB/s algorithmic")
PY
