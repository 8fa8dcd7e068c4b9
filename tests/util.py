"""Shared helpers for tests (inputs, constants mirrored from the reference's data/hyp.yaml)."""
import os

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
GOLDEN = os.path.join(ROOT, "tests", "golden")

HYP = dict(fl_gamma=0.0, box=0.05, obj=1.0, obj_pw=1.0, cls=0.5, cls_pw=1.0)   # data/hyp.yaml:11-17
CFG = dict(anchors=[[12, 16, 19, 36, 40, 28], [36, 75, 76, 55, 72, 146], [142, 110, 192, 243, 459, 401]],
           angles=[-90, -60, -30, 0, 30, 60])                                   # data/hyp.yaml:2-7


def load(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def rel_err(a, b):
    a, b = a.detach(), b.detach()
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def gaussian_label(label, num_class=180, u=0, sig=6.0):
    x = np.arange(-num_class / 2, num_class / 2)
    y = np.exp(-(x - u) ** 2 / (2 * sig ** 2))
    i = int(num_class / 2 - label)
    return np.concatenate([y[i:], y[:i]], axis=0)


def make_targets(seed, bs, per_img, nc, csl, wmax=0.15):
    """Synthetic labels as SURVEY.md §8(d): (img, cls, x, y, w, h in [0,1], theta rad[, csl 180])."""
    g = np.random.default_rng(seed)
    rows = []
    for b in range(bs):
        for _ in range(per_img):
            w = g.uniform(0.02, wmax)
            h = min(w * g.uniform(1, 3), 0.9)
            th = g.uniform(-np.pi / 2, np.pi / 2 - 1e-3)
            r = [b, float(g.integers(0, nc)), g.uniform(0.05, 0.95), g.uniform(0.05, 0.95), w, h, th]
            if csl:
                r += list(gaussian_label(th * 180 / np.pi + 90))
            rows.append(r)
    return torch.tensor(np.array(rows, dtype=np.float64), dtype=torch.float32).reshape(-1, 187 if csl else 7)


def winit(m):
    """Reference init (train.py:28-33): conv N(0,.02), BN gamma N(1,.02), beta 0."""
    cn = m.__class__.__name__
    if cn.find("Conv2d") != -1:
        torch.nn.init.normal_(m.weight.data, 0.0, 0.02)
    elif cn.find("BatchNorm2d") != -1:
        torch.nn.init.normal_(m.weight.data, 1.0, 0.02)
        torch.nn.init.constant_(m.bias.data, 0.0)


def det_init(model):
    """Order-independent deterministic init: every state-dict entry is seeded from its NAME.
    Distributions follow the reference's weights_init_normal (train.py:28-33): conv weights N(0,.02),
    BN gamma N(1,.02), BN beta 0; plus small seeded values for the entries the reference leaves to
    unseeded default inits (head-conv biases, ImplicitA ~N(0,.02), ImplicitM ~N(1,.02))."""
    import zlib
    sd = model.state_dict()
    with torch.no_grad():
        for name, t in sd.items():
            g = torch.Generator().manual_seed(zlib.crc32(name.encode()))
            if name.endswith("num_batches_tracked"):
                t.zero_()
            elif name.endswith("running_mean"):
                t.zero_()
            elif name.endswith("running_var"):
                t.fill_(1.0)
            elif name.endswith("implicit"):
                base = 1.0 if ".im" in name else 0.0
                t.copy_(base + 0.02 * torch.randn(t.shape, generator=g))
            elif t.dim() == 4:                                   # conv weight
                t.copy_(0.02 * torch.randn(t.shape, generator=g))
            elif name.endswith(".bias") and ".conv.0." in name:   # head conv bias
                t.copy_(0.05 * torch.randn(t.shape, generator=g))
            elif name.endswith(".weight"):                       # BN gamma
                t.copy_(1.0 + 0.02 * torch.randn(t.shape, generator=g))
            else:                                                # BN beta
                t.zero_()
    return model
