"""Generate the committed golden fixtures by EXECUTING THE REFERENCE'S OWN PYTHON.

Runs only in the build container (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference imports detectron2 (absent, un-vendored): the two symbols it needs are provided by a
stub module that forwards to the C++ restatement in oracle/rotated_ops.cpp.  Only
``post_process.pt`` depends on that stub (and is therefore "parity unpinned" at the NMS boundary);
every other fixture is produced by unmodified reference code.

All inputs are seeded and stored next to the reference outputs, so tests never need the reference.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import rotated as orot  # noqa: E402


def _install_stubs():
    d2 = types.ModuleType("detectron2")
    layers = types.ModuleType("detectron2.layers")
    nms = types.ModuleType("detectron2.layers.nms")
    rb = types.ModuleType("detectron2.layers.rotated_boxes")
    nms.nms_rotated = lambda boxes, scores, thr: orot.nms_rotated(boxes, scores, thr)
    rb.pairwise_iou_rotated = lambda a, b: orot.pairwise_iou_rotated(a, b)
    sys.modules.update({"detectron2": d2, "detectron2.layers": layers, "detectron2.layers.nms": nms,
                        "detectron2.layers.rotated_boxes": rb})
    sys.path.insert(0, REF)


HYP = dict(fl_gamma=0.0, box=0.05, obj=1.0, obj_pw=1.0, cls=0.5, cls_pw=1.0)
CFG = dict(anchors=[[12, 16, 19, 36, 40, 28], [36, 75, 76, 55, 72, 146], [142, 110, 192, 243, 459, 401]],
           angles=[-90, -60, -30, 0, 30, 60])


def gaussian_label(label, num_class=180, u=0, sig=6.0):
    # restated from datasets/base_dataset.py:13-31 (label pipeline is out of scope; used for inputs only)
    x = np.arange(-num_class / 2, num_class / 2)
    y = np.exp(-(x - u) ** 2 / (2 * sig ** 2))
    i = int(num_class / 2 - label)
    return np.concatenate([y[i:], y[:i]], axis=0)


def make_targets(gen, bs, per_img, nc, csl, wmax=0.15):
    rows = []
    for b in range(bs):
        for _ in range(per_img):
            w = float(torch.empty(1).uniform_(0.02, wmax, generator=gen))
            h = min(w * float(torch.empty(1).uniform_(1, 3, generator=gen)), 0.9)
            th = float(torch.empty(1).uniform_(-np.pi / 2, np.pi / 2 - 1e-3, generator=gen))
            r = [b, float(torch.randint(0, nc, (1,), generator=gen)),
                 float(torch.empty(1).uniform_(0.05, 0.95, generator=gen)),
                 float(torch.empty(1).uniform_(0.05, 0.95, generator=gen)), w, h, th]
            if csl:
                r += list(gaussian_label(th * 180 / np.pi + 90))
            rows.append(r)
    return torch.tensor(rows, dtype=torch.float32)


class _FakeModel:
    def __init__(self, anchors, nc):
        self.anchors, self.nc = anchors, nc
        self._p = torch.nn.Parameter(torch.zeros(1))

    def parameters(self):
        return iter([self._p])


def make_model_goldens(Yolo):
    """Reference Yolo fwd (train + eval) on weights that tests can regenerate: tests.util.det_init
    seeds every state-dict entry from its NAME (distributions of train.py:28-33), so the fixture does
    not depend on module traversal order or on the unseeded default inits of biases / implicit params."""
    from tests.util import det_init
    torch.set_num_threads(8)
    gen = torch.Generator().manual_seed(4321)
    for ver, mode, nc, S in (("yolov4", "csl", 2, 64), ("yolov4", "kfiou", 2, 64), ("yolov7", "csl", 16, 64),
                             ("yolov5", "csl", 2, 64)):
        model = Yolo(nc, CFG, mode, ver)
        det_init(model)
        img = torch.rand(2, 3, S, S, generator=gen)
        model.train()
        tr = model(img, training=True)
        stats = {k: v.clone() for k, v in model.state_dict().items() if "running" in k}
        keys = list(model.state_dict().keys())
        model.eval()
        with torch.no_grad():
            ev_levels, ev_infer = model(img, training=False)
        probe = {k: float(v.double().sum()) for k, v in model.state_dict().items()}
        torch.save(dict(ver=ver, mode=mode, nc=nc, img=img, train_levels=[t.detach() for t in tr],
                        running_after=stats, eval_levels=ev_levels, eval_infer=ev_infer, keys=keys,
                        weight_probe=probe),
                   os.path.join(HERE, f"model_{ver}_{mode}_nc{nc}.pt"))
        print("model", ver, mode, nc, len(keys), [tuple(t.shape) for t in tr])


def make_metrics_golden():
    """Runs the reference's get_batch_statistics / ap_per_class / compute_ap (test.py:14-149).  test.py cannot be
    imported here (it pulls lib.load -> datasets), so the three function definitions are compiled from its
    source text into a namespace that provides numpy/torch and the restated detectron2 op."""
    import ast
    src = open(os.path.join(REF, "test.py")).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef)
            and n.name in ("get_batch_statistics", "ap_per_class", "compute_ap")]
    if not hasattr(np, "trapz"):
        np.trapz = np.trapezoid
    ns = {"np": np, "torch": torch, "pairwise_iou_rotated": orot.pairwise_iou_rotated}
    exec(compile(ast.Module(body=keep, type_ignores=[]), "reference_test_py", "exec"), ns)
    gen = torch.Generator().manual_seed(99)
    B, nc = 4, 3
    targets, outputs = [], []
    for b in range(B):
        nt = [6, 0, 9, 5][b]
        t = torch.zeros(nt, 7)
        t[:, 0] = b
        t[:, 1] = torch.randint(0, nc, (nt,), generator=gen).float()
        t[:, 2:4] = torch.rand(nt, 2, generator=gen) * 300 + 50
        t[:, 4] = torch.rand(nt, generator=gen) * 40 + 10
        t[:, 5] = t[:, 4] * (1 + 2 * torch.rand(nt, generator=gen))
        t[:, 6] = (torch.rand(nt, generator=gen) - 0.5) * np.pi * 0.99
        targets.append(t)
        npred = [14, 5, 0, 11][b]
        d = torch.zeros(npred, 7)
        for i in range(npred):
            if nt and i < 2 * nt:                 # jittered copies of targets (some duplicates -> claimed once)
                src_t = t[i % nt]
                d[i, :5] = src_t[2:7] + torch.randn(5, generator=gen) * torch.tensor([3, 3, 2, 2, 0.05])
                d[i, 6] = src_t[1] if i % 5 else (src_t[1] + 1) % nc
            else:
                d[i, :2] = torch.rand(2, generator=gen) * 300 + 50
                d[i, 2] = torch.rand(1, generator=gen) * 40 + 10
                d[i, 3] = d[i, 2] * 2
                d[i, 4] = (torch.rand(1, generator=gen) - 0.5) * 3
                d[i, 6] = float(torch.randint(0, nc, (1,), generator=gen))
            d[i, 5] = torch.rand(1, generator=gen)
        d = d[d[:, 5].argsort(descending=True)] if npred else d
        outputs.append(d)
    targets = torch.cat(targets, 0)
    iouv = torch.linspace(0.5, 0.95, 10)
    stats = ns["get_batch_statistics"]([o.clone() for o in outputs], targets.clone(), iouv, 10)
    tp, conf, pcls, tcls = [np.concatenate(x, 0) for x in zip(*stats)]
    p, r, ap, f1, ucls = ns["ap_per_class"](tp, conf, pcls, tcls)
    torch.save(dict(outputs=outputs, targets=targets, iouv=iouv,
                    stats=[(torch.as_tensor(np.asarray(a)), torch.as_tensor(np.asarray(b)), torch.as_tensor(np.asarray(c)), d)
                           for a, b, c, d in stats],
                    p=p, r=r, ap=ap, f1=f1, ucls=ucls), os.path.join(HERE, "metrics.pt"))
    print("metrics: images with stats", len(stats), "mAP50", float(ap[:, 0].mean()))


def main():
    _install_stubs()
    if "--only-metrics" in sys.argv:
        make_metrics_golden()
        return
    if "--only-model" in sys.argv:
        from model.yolo import Yolo
        make_model_goldens(Yolo)
        return
    from lib import general as rgen
    from lib import loss as rloss
    from model.yolo import Yolo
    from model.yololayer import YoloCSLLayer, YoloKFIoULayer

    torch.set_num_threads(1)  # deterministic index_put winner (last writer) in the reference
    gen = torch.Generator().manual_seed(1234)
    strides = [8, 16, 32]
    an_csl = Yolo._make_anchors(strides, CFG["anchors"])
    an_kf = Yolo._make_rotated_anchors(strides, CFG["anchors"], [a * np.pi / 180 for a in CFG["angles"]])

    # ---- decode -------------------------------------------------------------------------
    for mode, nc, S in (("csl", 2, 128), ("kfiou", 2, 128), ("csl", 16, 96), ("kfiou", 16, 96)):
        na, ch = (3, nc + 185) if mode == "csl" else (18, nc + 6)
        heads = [torch.randn(2, na * ch, S // s, S // s, generator=gen) * 2.0 for s in strides]
        layer = (YoloCSLLayer(nc, an_csl, strides) if mode == "csl" else YoloKFIoULayer(nc, an_kf, strides))
        levels, infer = layer([h.clone() for h in heads], training=False)
        torch.save(dict(mode=mode, nc=nc, heads=heads, levels=levels, infer=infer),
                   os.path.join(HERE, f"decode_{mode}_nc{nc}.pt"))

    # ---- box losses ---------------------------------------------------------------------
    N = 512
    p4 = torch.cat((torch.rand(N, 2, generator=gen) * 2 - 0.5, torch.rand(N, 2, generator=gen) * 8 + 0.5), 1)
    t4 = torch.cat((torch.rand(N, 2, generator=gen), torch.rand(N, 2, generator=gen) * 8 + 0.5), 1)
    p4.requires_grad_(True)
    c = rloss.bbox_ciou(p4, t4)
    (1 - c).mean().backward()
    torch.save(dict(pred=p4.detach(), target=t4, ciou=c.detach(), grad=p4.grad.clone()),
               os.path.join(HERE, "ciou.pt"))

    p5 = torch.cat((p4.detach(), (torch.rand(N, 1, generator=gen) - 0.5) * np.pi * 0.999), 1).requires_grad_(True)
    t5 = torch.cat((t4, (torch.rand(N, 1, generator=gen) - 0.5) * np.pi * 0.999), 1)
    kl, kfiou = rloss.KFLoss()(p5, t5)
    kl.backward()
    torch.save(dict(pred=p5.detach(), target=t5, loss=kl.detach(), kfiou=kfiou.detach(), grad=p5.grad.clone()),
               os.path.join(HERE, "kfloss.pt"))

    # ---- full losses --------------------------------------------------------------------
    for mode, nc, S, gamma, tag in (("csl", 2, 96, 0.0, ""), ("kfiou", 2, 96, 0.0, ""), ("csl", 16, 64, 0.0, ""),
                                    ("kfiou", 16, 64, 0.0, ""), ("csl", 1, 64, 0.0, ""),
                                    ("csl", 2, 64, 1.5, "_focal"), ("kfiou", 2, 64, 1.5, "_focal")):
        csl = mode == "csl"
        na, ch = (3, nc + 185) if csl else (18, nc + 6)
        bs = 3
        levels = [(torch.randn(bs, na, S // s, S // s, ch, generator=gen) * 1.5).requires_grad_(True) for s in strides]
        targets = make_targets(gen, bs, 8, nc, csl, wmax=0.45)
        # a target exactly on an integer grid coordinate + two targets in one cell (duplicate rows)
        targets[1, 2:4] = torch.tensor([0.5, 0.25])
        targets[2, 2:6] = targets[3, 2:6]
        targets[2, 0] = targets[3, 0]
        hyp = dict(HYP, fl_gamma=gamma)
        fm = _FakeModel(an_csl if csl else an_kf, nc)
        crit = rloss.ComputeCSLLoss(fm, hyp) if csl else rloss.ComputeKFIoULoss(fm, hyp)
        bt = crit.build_targets(levels, targets)
        loss, items = crit(levels, targets)
        loss.backward()
        if csl:
            tcls, tbox, ta, tg, indices, anch = bt
        else:
            tcls, tbox, indices, anch = bt
        torch.save(dict(mode=mode, nc=nc, hyp=hyp, levels=[l.detach() for l in levels], targets=targets,
                        loss=loss.detach(), items=dict(items), grads=[l.grad.clone() for l in levels],
                        tcls=tcls, tbox=tbox, indices=[tuple(t.clone() for t in ix) for ix in indices], anch=anch),
                   os.path.join(HERE, f"loss_{mode}_nc{nc}{tag}.pt"))
    # empty-target case (lib/loss.py:311-313)
    for mode in ("csl", "kfiou"):
        csl = mode == "csl"
        nc, S, bs = 2, 64, 2
        na, ch = (3, nc + 185) if csl else (18, nc + 6)
        levels = [(torch.randn(bs, na, S // s, S // s, ch, generator=gen)).requires_grad_(True) for s in strides]
        targets = torch.zeros((0, 187 if csl else 7))
        fm = _FakeModel(an_csl if csl else an_kf, nc)
        crit = rloss.ComputeCSLLoss(fm, HYP) if csl else rloss.ComputeKFIoULoss(fm, HYP)
        loss, items = crit(levels, targets)
        loss.backward()
        torch.save(dict(mode=mode, nc=nc, hyp=HYP, levels=[l.detach() for l in levels], targets=targets,
                        loss=loss.detach(), items=dict(items), grads=[l.grad.clone() for l in levels]),
                   os.path.join(HERE, f"loss_{mode}_empty.pt"))

    # ---- post_process (NMS boundary = restated detectron2: parity unpinned) ---------------
    B, R, nc = 4, 3000, 2
    ctr = torch.rand(B, 40, 2, generator=gen) * 400
    which = torch.randint(0, 40, (B, R), generator=gen)
    xy = torch.gather(ctr, 1, which[..., None].expand(-1, -1, 2)) + torch.randn(B, R, 2, generator=gen) * 6
    w = torch.rand(B, R, 1, generator=gen) * 40 + 6
    h = w * (torch.rand(B, R, 1, generator=gen) * 2 + 1)
    th = (torch.rand(B, R, 1, generator=gen) - 0.5) * np.pi * 0.999
    conf = torch.rand(B, R, 1 + nc, generator=gen)
    pred = torch.cat((xy, w, h, th, conf), 2)
    pred[3, :, 5] = 0.0  # an image with no survivors of the confidence filter
    cases = []
    for ct, it in ((0.5, 0.4), (0.25, 0.1), (0.7, 0.2)):
        p = pred.clone()
        outs = rgen.post_process(p, ct, it)
        sc = (pred[:, :, 6:] * pred[:, :, 5:6]).max(2)[0]
        ties = sum(int(len(s[s > ct]) - len(torch.unique(s[s > ct]))) for s in sc)
        cases.append(dict(conf_thres=ct, iou_thres=it, outs=[o.clone() for o in outs], mutated=p, score_ties=ties))
        print("post_process", ct, it, [o.shape[0] for o in outs], "score ties:", ties)
    torch.save(dict(pred=pred, cases=cases), os.path.join(HERE, "post_process.pt"))

    # ---- conv stack ---------------------------------------------------------------------
    make_model_goldens(Yolo)
    make_metrics_golden()
    make_metrics_golden()

    for f in sorted(os.listdir(HERE)):
        if f.endswith(".pt"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
