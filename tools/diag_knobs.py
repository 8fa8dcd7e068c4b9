"""A/B timing of the library's tuning knobs inside one process (ryolo_tune): per-launch CUDA-event times of the
conv / dgrad / wgrad kernels of a yolov4 800x800 train step under each knob setting (wgrad on the main stream so the
per-launch times are clean).  dbg / wg_dbg settings produce wrong results on purpose: they remove one pipeline stage to
show what the kernel is waiting for.  Writes gpurun_out/diag_knobs.json + a text table."""
import json, os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
import ryolo_b200 as R
from ryolo_b200 import ops, _lib as L
from ryolo_b200.model import backward as BW
from bench import CFG, HYP, S, make_targets, weights_init_normal

bs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
which = sys.argv[2].split(",") if len(sys.argv) > 2 else None
BW.WGRAD_SIDE_STREAM = False
torch.manual_seed(42)
m = R.Yolo(2, CFG, "csl", "yolov4")
m.apply(weights_init_normal)
m = m.cuda().train()
crit = R.ComputeCSLLoss(m, HYP)
crit.sync_items = False
step = R.TrainStep(m, crit)
img = torch.rand(bs, 3, S, S, device="cuda")
tg = make_targets(0, bs, 2).cuda()
flat0 = step.flat.clone()
KNOBS = ["halo", "dbg", "wg_split", "wg_dbg", "epi_tma", "epi_maxbn", "wg_tapgrp", "bn_bwd", "wg_trans", "sw64", "nacc", "pdl", "ssa", "wg_boxes", "ew_regs", "nms_band", "bn_fuse", "kgrp", "pair", "wres", "wg_x32"]
BASE = {k: L.lib().ryolo_knob(i) for i, k in enumerate(KNOBS)}          # the library's defaults
VARIANTS = [
    ("base", {}),
    ("wg_uniform", dict(wg_split=0)),
    ("wg_bytaps", dict(wg_split=2)),
    ("wg_tap1", dict(wg_tapgrp=0)),
    ("wg_mn", dict(wg_trans=0)),
    ("wg_mn_uniform", dict(wg_trans=0, wg_split=0)),
    ("wg_nomma", dict(wg_dbg=1)),
    ("wg_noload", dict(wg_dbg=2)),
    ("bn_old", dict(bn_bwd=0)),
    ("bn_store", dict(bn_bwd=1)),
    ("sw128", dict(sw64=0)),
    ("acc2", dict(nacc=0)),
    ("nopdl", dict(pdl=0)),
    ("epi_direct", dict(epi_tma=0)),
    ("epi_tma1", dict(epi_tma=1)),
    ("epi_bn64", dict(epi_maxbn=64)),
    ("epi_bn256", dict(epi_maxbn=256)),
    ("nostore", dict(dbg=1)),
    ("nostats", dict(dbg=2)),
    ("nostore_nostats", dict(dbg=3)),
    ("nomma", dict(dbg=4)),
    ("noaload", dict(dbg=8)),
    ("nostatstail", dict(dbg=16)),
    ("nobload", dict(dbg=32)),
    ("noload", dict(dbg=40)),
    ("noaload_nostats", dict(dbg=10)),
    ("noaload_nomma", dict(dbg=12)),
    ("halo", dict(halo=1)),
    ("k32_n128", dict(sw64=2)),
    ("k32_all", dict(sw64=3)),
    ("wg_x32", dict(wg_x32=1)),
    ("wg_x64", dict(wg_x32=0)),
    ("wres0", dict(wres=0)),
    ("wres48", dict(wres=48)),
    ("wres96", dict(wres=96)),
    ("wres128", dict(wres=128)),
    ("pair0", dict(pair=0)),
    ("pair256", dict(pair=256)),
    ("pair128", dict(pair=128)),
    ("pair64", dict(pair=64)),
    ("kgrp0", dict(kgrp=0)),
    ("kgrp1", dict(kgrp=1)),
    ("kgrp2", dict(kgrp=2)),
    ("kgrp3", dict(kgrp=3)),
]
res = {}
for name, kn in VARIANTS:
    if which and name not in which:
        continue
    L.tune(**{**BASE, **kn})
    step.flat.copy_(flat0)                      # dbg variants may write garbage: restart from the same weights
    step.buf.zero_(); step.first = True
    m.repack_weights()
    try:
        step(img, tg)
        torch.cuda.synchronize()
        ops.PROFILE = []
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step(img, tg)
        b.record()
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None
    except Exception as e:                      # keep going: one broken variant must not lose the others
        ops.PROFILE = None
        res[name] = {"error": repr(e)}
        print(name, "ERROR", e, flush=True)
        continue
    rows = [(t[0], t[1], t[2], t[3], s.elapsed_time(e)) for t, s, e in prof]
    tot = {}
    for k, M, N, K, ms in rows:
        tot[k] = tot.get(k, 0.0) + ms
    res[name] = {"step_ms": a.elapsed_time(b), "totals": tot, "rows": rows}
    print(f"{name:16s} step {a.elapsed_time(b):7.2f} ms  " + "  ".join(f"{k} {v:6.2f}" for k, v in tot.items()), flush=True)
L.tune(**BASE)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"diag_knobs_bs{bs}.json"), "w"))
# per-layer table for the variants that ran
names = [n for n in res if "rows" in res[n]]
if names:
    n0 = names[0]
    with open(os.path.join(ROOT, "gpurun_out", f"diag_knobs_bs{bs}.txt"), "w") as f:
        f.write("kind        M      N     K  " + " ".join(f"{n[:9]:>9s}" for n in names) + "\n")
        for i, r in enumerate(res[n0]["rows"]):
            f.write(f"{r[0]:6s} {r[1]:9d} {r[2]:5d} {r[3]:5d} " +
                    " ".join(f"{res[n]['rows'][i][4]:9.3f}" if i < len(res[n]['rows']) else "        -" for n in names) + "\n")
