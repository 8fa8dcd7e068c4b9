"""GPU parity: decode (csrc/decode.cu) and assignment + loss fwd/bwd (csrc/loss.cu) vs golden
fixtures generated from the reference and vs the oracle.  fp32 tolerance: 1e-4 relative
(BASELINE.json north_star); indices bit-exact."""
import numpy as np
import pytest
import torch

from oracle import hotpath as hp
from tests.util import CFG, HYP, load, make_targets, rel_err

pytestmark = pytest.mark.gpu

AN_CSL = hp.make_anchors(CFG["anchors"])
AN_KF = hp.make_rotated_anchors(CFG["anchors"], CFG["angles"])
TOL = 1e-4


class _M:  # what the loss constructor reads from the model (lib/loss.py:155,177,180)
    def __init__(self, anchors, nc):
        self.anchors, self.nc = anchors, nc
        self._p = torch.nn.Parameter(torch.zeros(1, device="cuda"))

    def parameters(self):
        return iter([self._p])


@pytest.mark.parametrize("name", ["decode_csl_nc2", "decode_kfiou_nc2", "decode_csl_nc16", "decode_kfiou_nc16"])
def test_decode_golden(name):
    import ryolo_b200 as R
    g = load(name + ".pt")
    csl = g["mode"] == "csl"
    nc = g["nc"]
    layer = (R.YoloCSLLayer(nc, AN_CSL, [8, 16, 32]) if csl else R.YoloKFIoULayer(nc, AN_KF, [8, 16, 32]))
    heads = [h.clone().cuda() for h in g["heads"]]
    levels, infer = layer(heads, training=False)
    for mine, ref in zip(levels, g["levels"]):
        assert torch.equal(mine.cpu(), ref)
    out, ref = infer.cpu(), g["infer"]
    assert out.shape == ref.shape
    if csl:
        # angle = argmax over sigmoid values; allow a different bin only where the two sigmoid values tie within 2 ulp
        bad = (out[..., 4] != ref[..., 4])
        assert bad.float().mean() < 1e-3
        out[..., 4][bad] = ref[..., 4][bad]
    assert rel_err(out[..., :4], ref[..., :4]) < TOL
    assert (out[..., 4:] - ref[..., 4:]).abs().max() < 1e-5


def _run_loss(R, g, levels_cuda):
    csl = g["mode"] == "csl"
    fn = (R.ComputeCSLLoss if csl else R.ComputeKFIoULoss)(_M(AN_CSL if csl else AN_KF, g["nc"]), g["hyp"])
    assert list(fn.loss_items.keys()) == list(g["items"].keys())
    loss, items = fn(levels_cuda, g["targets"].cuda())
    return fn, loss, items


LOSS_CASES = ["loss_csl_nc2", "loss_kfiou_nc2", "loss_csl_nc16", "loss_kfiou_nc16", "loss_csl_nc1",
              "loss_csl_nc2_focal", "loss_kfiou_nc2_focal", "loss_csl_empty", "loss_kfiou_empty"]


@pytest.mark.parametrize("name", LOSS_CASES)
def test_loss_golden(name):
    import ryolo_b200 as R
    g = load(name + ".pt")
    levels = [l.clone().cuda().requires_grad_(True) for l in g["levels"]]
    fn, loss, items = _run_loss(R, g, levels)
    assert loss.shape == (1,) and loss.requires_grad
    loss.backward()
    assert list(items.keys()) == list(g["items"].keys())
    for k in items:
        assert abs(items[k] - g["items"][k]) <= TOL * max(1e-3, abs(g["items"][k])), (k, items[k], g["items"][k])
    # duplicate objectness cells: the reference's winner is implementation-defined;
    # its goldens were generated single-threaded (last writer wins), which is also our rule.
    for mine, ref in zip(levels, g["grads"]):
        assert rel_err(mine.grad.cpu(), ref) < 2e-4
    if "indices" in g:
        bt = fn.build_targets(levels, g["targets"].cuda())
        indices, tbox, tcls, anch = bt[-2], bt[1], bt[0], bt[-1]
        for ix, rix, tb, rtb, tc, rtc, an, ran in zip(indices, g["indices"], tbox, g["tbox"], tcls, g["tcls"], anch,
                                                      g["anch"]):
            for a, b in zip(ix, rix):
                assert torch.equal(a.cpu(), b)                   # bit-exact, reference order
            assert torch.equal(tb.cpu(), rtb)
            assert torch.equal(tc.cpu(), rtc)
            assert torch.equal(an.cpu(), ran)


def test_loss_no_grad_value_only():
    import ryolo_b200 as R
    g = load("loss_csl_nc2.pt")
    with torch.no_grad():
        _, loss, items = _run_loss(R, g, [l.clone().cuda() for l in g["levels"]])
    assert not loss.requires_grad
    assert abs(items["total_loss"] - g["items"]["total_loss"]) <= TOL * abs(g["items"]["total_loss"])


@pytest.mark.parametrize("mode,nc,S,bs,per_img", [("csl", 2, 416, 2, 20), ("kfiou", 2, 416, 2, 20),
                                                  ("csl", 16, 256, 4, 60), ("kfiou", 16, 256, 4, 60)])
def test_loss_vs_oracle_bigger(mode, nc, S, bs, per_img):
    """BASELINE config 1 shape (416^2, bs 2, 20 targets/img) and a denser case, against the oracle."""
    import ryolo_b200 as R
    csl = mode == "csl"
    na, ch = (3, nc + 185) if csl else (18, nc + 6)
    gen = torch.Generator().manual_seed(7)
    levels = [torch.randn(bs, na, S // s, S // s, ch, generator=gen) for s in (8, 16, 32)]
    targets = make_targets(3, bs, per_img, nc, csl)
    anchors = AN_CSL if csl else AN_KF
    ref_lv = [l.clone().requires_grad_(True) for l in levels]
    ref_loss, ref_items = (hp.csl_loss if csl else hp.kfiou_loss)(ref_lv, targets, anchors, nc, HYP)
    ref_loss.backward()
    fn = (R.ComputeCSLLoss if csl else R.ComputeKFIoULoss)(_M(anchors, nc), HYP)
    lv = [l.clone().cuda().requires_grad_(True) for l in levels]
    loss, items = fn(lv, targets.cuda())
    loss.backward()
    for k in items:
        assert abs(items[k] - ref_items[k]) <= TOL * max(1e-3, abs(ref_items[k])), k
    for mine, ref in zip(lv, ref_lv):
        assert rel_err(mine.grad.cpu(), ref.grad) < 2e-4
    asg = hp.assign_targets(targets, [(l.shape[2], l.shape[3]) for l in levels], anchors, rotated=not csl)
    bt = fn.build_targets(lv, targets.cuda())
    for ix, s in zip(bt[-2], asg):
        for a, b in zip(ix, (s["b"], s["a"], s["gj"], s["gi"])):
            assert torch.equal(a.cpu(), b)


def test_kfloss_standalone_golden_and_scale():
    """KFLoss drop-in (lib/loss.py:81-150): golden from the reference's N x N form; then BASELINE config 4 size
    (50k pairs x 32 images) against the O(N) oracle."""
    import ryolo_b200 as R
    g = load("kfloss.pt")
    p = g["pred"].clone().cuda().requires_grad_(True)
    loss, kfiou = R.KFLoss()(p, g["target"].cuda())
    loss.backward()
    assert rel_err(kfiou.cpu(), g["kfiou"]) < TOL
    assert abs(float(loss) - float(g["loss"])) <= TOL * abs(float(g["loss"]))
    assert rel_err(p.grad.cpu(), g["grad"]) < 2e-4
    gen = torch.Generator().manual_seed(9)
    N = 50000 * 32 + 3                                      # ragged tail on purpose
    pr = torch.cat((torch.rand(N, 2, generator=gen) * 2 - 0.5, torch.rand(N, 2, generator=gen) * 8 + 0.5,
                    (torch.rand(N, 1, generator=gen) - 0.5) * 3.14), 1)
    tg = torch.cat((torch.rand(N, 2, generator=gen), torch.rand(N, 2, generator=gen) * 8 + 0.5,
                    (torch.rand(N, 1, generator=gen) - 0.5) * 3.14), 1)
    ref_loss, ref_k = hp.kf_loss(pr, tg)
    loss, kfiou = R.KFLoss()(pr.cuda(), tg.cuda())
    assert abs(float(loss) - float(ref_loss)) <= TOL * abs(float(ref_loss))
    assert rel_err(kfiou.cpu(), ref_k) < TOL
    empty = R.KFLoss()(torch.zeros(0, 5).cuda(), torch.zeros(0, 5).cuda())
    assert float(empty[0]) == 0.0 and empty[1].numel() == 0


def test_decode_csl_angle_argmax_ties_and_saturation():
    """The CSL angle is argmax over the fp32 SIGMOID of 180 logits, first maximum wins (model/yololayer.py:39,48).  The
    kernel only evaluates the sigmoid of logits that can tie with the largest one: check the cases where that matters —
    a saturated plateau (every logit >= 16.64 maps to 1.0f: the FIRST of them wins, not the largest), exact duplicates,
    near-saturation neighbours that still differ, -inf rows, and plain rows (argmax of the logits)."""
    import ryolo_b200 as R
    nc, gs, B = 2, 4, 2
    gen = torch.Generator().manual_seed(11)
    head = torch.randn(B, 3, gs, gs, nc + 185, generator=gen) * 3
    ang = head[..., 5 + nc:]
    ang.clamp_(max=9.0)
    expect = torch.sigmoid(ang.double()).float().argmax(-1)      # maxima far from saturation (first one on ties)
    cells = [(0, 0, 0, 0), (0, 1, 2, 3), (1, 2, 1, 1), (1, 0, 3, 2), (0, 2, 2, 2), (1, 1, 0, 3)]
    # plateau: three saturated logits, the largest one last -> first index wins
    ang[cells[0]][[40, 100, 150]] = torch.tensor([20.0, 25.0, 30.0]); expect[cells[0]] = 40
    # plateau reached from below the cut-off of a huge maximum: 17 saturates too, and comes first
    ang[cells[1]][[7, 90]] = torch.tensor([17.0, 80.0]); expect[cells[1]] = 7
    # exact duplicates of the maximum
    ang[cells[2]][[33, 34, 170]] = 12.5; expect[cells[2]] = 33
    # neighbours below saturation that still round to different sigmoid values: the larger logit wins
    ang[cells[3]][[10, 20]] = torch.tensor([15.4, 15.9]); expect[cells[3]] = 20
    # 16.0 is not saturated (1 - 1.1e-7), 16.7 is
    ang[cells[4]][[5, 60]] = torch.tensor([16.0, 16.7]); expect[cells[4]] = 60
    # a row of -inf: every sigmoid is 0, index 0 wins
    ang[cells[5]][:] = float("-inf"); expect[cells[5]] = 0
    layer = R.YoloCSLLayer(nc, AN_CSL, [8, 16, 32])
    heads = [head.cuda(), torch.zeros(B, 3, 2, 2, nc + 185).cuda(), torch.zeros(B, 3, 1, 1, nc + 185).cuda()]
    _, infer = layer(heads, training=False)
    got = infer[:, :3 * gs * gs, 4].cpu().view(B, 3, gs, gs)
    want = (expect.float() - 90) / 180 * 3.14159274101257324
    assert torch.equal(got, want.float()), (got - want).abs().max()
