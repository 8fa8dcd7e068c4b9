"""GPU parity of the whole conv stack (Yolo.forward on tcgen05 kernels, bf16 storage / fp32 accumulate)
against the goldens produced by the reference's fp32 nn.Modules.

Tolerance: the 1e-4 bar of BASELINE.json applies to decode/loss/NMS on identical inputs; through ~110
bf16 layers (with train-mode BatchNorm over as few as 8 samples per channel at this fixture size) the
stack is compared by cosine similarity and relative L2 error instead (SURVEY.md §7 hard part 7).
Per-layer bf16 parity is covered by tests/test_gpu_conv.py."""
import json
import os

import pytest
import torch

from tests.util import CFG, ROOT, det_init, load

pytestmark = pytest.mark.gpu
CASES = [("yolov4", "csl", 2), ("yolov4", "kfiou", 2), ("yolov7", "csl", 16), ("yolov5", "csl", 2)]


def _cmp(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten()
    cos = float(torch.dot(a, b) / (a.norm() * b.norm()))
    return cos, float((a - b).norm() / b.norm())


def _log(name, rec):
    d = os.path.join(ROOT, "gpurun_out")
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "model_parity.jsonl"), "a") as f:
        f.write(json.dumps(dict(case=name, **rec)) + "\n")


@pytest.mark.parametrize("ver,mode,nc", CASES)
def test_eval_forward_vs_reference(ver, mode, nc):
    import ryolo_b200 as R
    g = load(f"model_{ver}_{mode}_nc{nc}.pt")
    m = det_init(R.Yolo(nc, CFG, mode, ver))
    sd = m.state_dict()
    for k, v in g["running_after"].items():           # eval golden was taken after one train-mode forward
        sd[k].copy_(v)
    m = m.cuda().eval()
    with torch.no_grad():
        levels, infer = m(g["img"].cuda(), training=False)
    rec = {}
    for i, (a, b) in enumerate(zip(levels, g["eval_levels"])):
        assert a.shape == b.shape and a.dtype == torch.float32
        cos, l2 = _cmp(a, b)
        rec[f"eval_l{i}"] = (cos, l2)
        assert cos > 0.995 and l2 < 0.1, (i, cos, l2)
    assert infer.shape == g["eval_infer"].shape
    cos, l2 = _cmp(infer[..., :4], g["eval_infer"][..., :4])
    rec["infer_box"] = (cos, l2)
    assert cos > 0.995
    _log(f"eval_{ver}_{mode}", rec)


@pytest.mark.parametrize("ver,mode,nc", CASES)
def test_train_forward_vs_reference(ver, mode, nc):
    import ryolo_b200 as R
    g = load(f"model_{ver}_{mode}_nc{nc}.pt")
    m = det_init(R.Yolo(nc, CFG, mode, ver)).cuda().train()
    levels = m(g["img"].cuda(), training=True)
    rec = {}
    for i, (a, b) in enumerate(zip(levels, g["train_levels"])):
        assert a.shape == b.shape
        cos, l2 = _cmp(a, b)
        rec[f"train_l{i}"] = (cos, l2)
        assert cos > 0.4, (i, cos, l2)   # chaotic with random weights: see test_every_layer_teacher_forced
    sd = m.state_dict()
    worst = 0.0
    first = [k for k in g["running_after"] if k.split(".")[1] in ("cbm0", "cbm1", "cbs0", "cbs1")]
    first = [k for k in first if "running" in k]
    assert len(first) == 4
    for k in first:                                   # early layers: before the chaos sets in
        v = g["running_after"][k]
        worst = max(worst, float((sd[k].cpu() - v).abs().max() / v.abs().max().clamp_min(1e-3)))
    rec["running_stats_first_layers_worst_rel"] = worst
    assert worst < 0.02
    assert int(sd["backbone.cbm0.conv.1.num_batches_tracked" if ver == "yolov4" else
                  "backbone.cbs0.conv.1.num_batches_tracked"]) == 1
    _log(f"train_{ver}_{mode}", rec)


def test_forward_api_contract():
    import ryolo_b200 as R
    m = det_init(R.Yolo(2, CFG, "csl", "yolov4")).cuda()
    img = torch.rand(1, 3, 96, 96, device="cuda")
    m.train()
    out = m(img, training=True)
    assert isinstance(out, list) and [tuple(o.shape) for o in out] == [(1, 3, 12, 12, 187), (1, 3, 6, 6, 187),
                                                                       (1, 3, 3, 3, 187)]
    m.eval()
    out, infer = m(img, training=False)
    assert infer.shape == (1, 3 * (144 + 36 + 9), 8)
    dets = R.post_process(infer, 0.0, 0.4)
    assert len(dets) == 1 and dets[0].shape[1] == 7


@pytest.mark.parametrize("ver,mode,nc", [("yolov4", "csl", 2), ("yolov7", "csl", 16), ("yolov5", "csl", 2)])
@pytest.mark.parametrize("train", [True, False])
def test_every_layer_teacher_forced(ver, mode, nc, train):
    """Each Conv / RepConv of the real network, fed the ORACLE's input for that layer, must reproduce the
    oracle's output to bf16 accuracy (the end-to-end train-mode comparison is chaotic with random weights:
    a 1e-7 perturbation of the fp32 reference itself grows to 1e-3 over the stack)."""
    import ryolo_b200 as R
    from oracle import model_cpu
    from ryolo_b200 import ops
    from ryolo_b200.model.blocks import Conv, Ctx, RepConv
    S, bs = 96, 2
    m = det_init(R.Yolo(nc, CFG, mode, ver))
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    img = torch.rand(bs, 3, S, S, generator=torch.Generator().manual_seed(1))
    trace = []
    model_cpu.forward(sd, img, ver, mode, nc, train=train, trace=trace, decode=False)
    m = m.cuda()
    m.train(train)
    mods = dict(m.named_modules())
    ctx = Ctx(m, train, torch.device("cuda"))
    worst = {}
    for pre, x, y in trace:
        mod = mods[pre]
        if isinstance(mod, Conv) and mod.stem:
            xin = ops.stem_im2col(x.cuda(), mod.k, mod.s)
        else:
            xin = ops.Act(x.permute(0, 2, 3, 1).contiguous().bfloat16().cuda())
        if isinstance(mod, Conv) and not mod.has_bn:                       # head: fp32 [B,na,gs,gs,ch]
            na, ch = m.na, m.ch
            if ver == "yolov7":
                continue                                                  # folded Implicit head: checked end to end
            out = mod(ctx, xin, head=(na, ch), head_shift=mod.conv[0].bias.data)
            ref = y.view(bs, na, ch, y.shape[2], y.shape[3]).permute(0, 1, 3, 4, 2)
            got = out.cpu()
        else:
            assert isinstance(mod, (Conv, RepConv))
            got = mod(ctx, xin).torch().float().cpu()
            ref = y.permute(0, 2, 3, 1)
        err = float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-6))
        worst[pre] = err
        assert err < 3e-2, (pre, err)
    _log(f"layers_{ver}_{'train' if train else 'eval'}", dict(n_layers=len(worst), worst=max(worst.values())))
    assert len(worst) >= 60
