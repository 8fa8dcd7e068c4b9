"""CPU checks of the PRODUCT's loss math (r-yolov4_b200/csrc/loss_math.cuh, assign.cuh — the same
__host__ __device__ source the CUDA kernels use), compiled for the host and compared with the oracle
(oracle/hotpath.py, itself pinned to the reference by tests/test_oracle_golden.py)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import hotpath as hp
from tests.util import CFG, ROOT, load, rel_err

F32P = ctypes.POINTER(ctypes.c_float)


@pytest.fixture(scope="module")
def hm(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hm") / "libhm.so")
    src = os.path.join(ROOT, "tests", "host", "loss_math_host.cpp")
    inc = os.path.join(ROOT, "r-yolov4_b200", "csrc")
    subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-I", inc, src, "-o", out],
                   check=True)
    return ctypes.CDLL(out)


def _p(a):
    return a.ctypes.data_as(F32P)


def test_ciou_value_and_grad(hm):
    g = load("ciou.pt")
    p, t = g["pred"].numpy().copy(), g["target"].numpy().copy()
    n = p.shape[0]
    out, grad = np.zeros(n, np.float32), np.zeros((n, 4), np.float32)
    hm.hm_ciou(ctypes.c_int64(n), _p(p), _p(t), _p(out), _p(grad))
    assert rel_err(torch.from_numpy(out), g["ciou"]) < 1e-5
    # golden grad is d mean(1-ciou)/dp
    assert rel_err(torch.from_numpy(-grad / n), g["grad"]) < 1e-4


def test_kf_value_and_grad(hm):
    g = load("kfloss.pt")
    p, t = g["pred"].numpy().copy(), g["target"].numpy().copy()
    n = p.shape[0]
    xy, kf, k = (np.zeros(n, np.float32) for _ in range(3))
    grad = np.zeros((n, 5), np.float32)
    hm.hm_kf(ctypes.c_int64(n), _p(p), _p(t), _p(xy), _p(kf), _p(k), _p(grad))
    assert rel_err(torch.from_numpy(k), g["kfiou"]) < 1e-5
    loss = xy.astype(np.float64).mean() + kf.astype(np.float64).mean()
    assert abs(loss - float(g["loss"])) / float(g["loss"]) < 1e-5
    assert rel_err(torch.from_numpy(grad / n), g["grad"]) < 1e-4


@pytest.mark.parametrize("gamma,pw", [(0.0, 1.0), (0.0, 2.5), (1.5, 1.0), (2.0, 0.7)])
def test_bce_focal(hm, gamma, pw):
    gen = torch.Generator().manual_seed(3)
    x = (torch.randn(4096, generator=gen) * 4).requires_grad_(True)
    t = torch.rand(4096, generator=gen)
    t[::3] = 0.0
    t[1::7] = 1.0
    ref = hp.focal(x, t, pw, gamma)
    ref.sum().backward()
    loss, dx = np.zeros(4096, np.float32), np.zeros(4096, np.float32)
    xn, tn = x.detach().numpy().copy(), t.numpy().copy()
    hm.hm_bce(ctypes.c_int64(4096), _p(xn), _p(tn), ctypes.c_float(pw), ctypes.c_float(gamma), _p(loss), _p(dx))
    assert rel_err(torch.from_numpy(loss), ref) < 1e-5
    assert rel_err(torch.from_numpy(dx), x.grad) < 1e-5


@pytest.mark.parametrize("name", ["loss_csl_nc2", "loss_kfiou_nc2", "loss_csl_nc16", "loss_kfiou_nc16"])
def test_assignment_bit_exact(hm, name):
    g = load(name + ".pt")
    rotated = g["mode"] == "kfiou"
    anchors = hp.make_rotated_anchors(CFG["anchors"], CFG["angles"]) if rotated else hp.make_anchors(CFG["anchors"])
    tg = g["targets"].numpy().copy()
    T, tcols = tg.shape
    nb = int(hm.hm_pos_bytes())
    assert nb == 48
    hm.hm_assign.restype = ctypes.c_int64
    for lvl, (lv, ix, tb, tc) in enumerate(zip(g["levels"], g["indices"], g["tbox"], g["tcls"])):
        an = np.zeros((len(anchors[lvl]), 3), np.float32)
        an[:, :len(anchors[lvl][0])] = np.array(anchors[lvl], np.float32)
        na, gh, gw = an.shape[0], lv.shape[2], lv.shape[3]
        buf = np.zeros((5 * na * T, 12), np.int32)
        n = hm.hm_assign(_p(tg), ctypes.c_int64(T), tcols, _p(an), na, int(rotated), gh, gw, int(lv.shape[0]),
                         buf.ctypes.data_as(ctypes.c_void_p))
        rec = buf[:n]
        assert n == ix[0].numel()
        for col, ref in zip((0, 1, 2, 3), ix):
            assert np.array_equal(rec[:, col].astype(np.int64), ref.numpy())
        fl = rec.view(np.float32)
        ncol = 5 if rotated else 4
        assert np.array_equal(fl[:, 4:4 + ncol], tb.numpy())
        assert np.array_equal(rec[:, 9].astype(np.int64), tc.numpy())
