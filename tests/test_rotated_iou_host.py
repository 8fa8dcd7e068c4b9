"""CPU checks of the PRODUCT's skew-IoU source (r-yolov4_b200/csrc/rotated_iou.cuh, compiled for the host) against
the oracle (oracle/rotated_ops.cpp = detectron2's algorithm as frozen in SURVEY.md Appendix B), plus the oracle's own
cross-check against OpenCV (reference call sites lib/general.py:177, test.py:135).

Why this exists: the NMS kernel decides most pairs from a cheap fp32 clip (rbox_iou_fast) and from geometric early-outs.
Appendix B is NOT the geometric IoU for degenerate configurations — near-duplicate boxes come out as 1/3, 3, 55 …,
sub-unit boxes "overlap" through the absolute EPS slack — and those values are the spec.  The adversarial generators
below hit exactly those regimes and require that every decision of the product equals the oracle's."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import rotated as orot
from tests.util import ROOT

F32P = ctypes.POINTER(ctypes.c_float)
U8P = ctypes.POINTER(ctypes.c_uint8)
THRESHOLDS = (0.2, 0.4, 0.65)          # detect.py:91, post_process default, test.py:270


@pytest.fixture(scope="module")
def hr(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hr") / "libhr.so")
    src = os.path.join(ROOT, "tests", "host", "rotated_iou_host.cpp")
    inc = os.path.join(ROOT, "r-yolov4_b200", "csrc")
    subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-I", inc,
                    src, "-o", out], check=True)
    return ctypes.CDLL(out)


def run_pairs(hr, a, b, thr):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    n = a.shape[0]
    fast, full = np.zeros(n, np.float32), np.zeros(n, np.float32)
    gate, dec = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
    pair = np.zeros(n, np.float32)
    hr.hr_pairs(ctypes.c_int64(n), a.ctypes.data_as(F32P), b.ctypes.data_as(F32P), ctypes.c_float(thr),
                fast.ctypes.data_as(F32P), full.ctypes.data_as(F32P), gate.ctypes.data_as(U8P),
                dec.ctypes.data_as(U8P), pair.ctypes.data_as(F32P))
    run_pairs.last_pairwise = pair
    return fast, full, gate.astype(bool), dec.astype(bool)


def oracle_pairs(a, b):
    """Oracle IoU of pair i = (a[i], b[i])."""
    lib = orot._load()
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    out = np.zeros(a.shape[0], np.float32)
    one = np.zeros(1, np.float32)
    for i in range(a.shape[0]):
        lib.oracle_pairwise_iou_rotated(a[i].ctypes.data_as(F32P), 1, b[i].ctypes.data_as(F32P), 1,
                                        one.ctypes.data_as(F32P))
        out[i] = one[0]
    return out


def adversarial_pairs(rng, n, kind):
    """Pairs (A, B = perturbed / related copy of A).  kinds:
    normal  4-120 px boxes (post_process's range), aspect <= 4      thin   aspect up to 1:1000
    tiny    sides 1e-3 .. 1e3                                       any    everything mixed
    edge    B shares an edge LINE with A (same axes, one side flush, sizes differ): the area-bound / hull-order trap
    corner  B's corner sits on A's corner or edge, arbitrary relative angle
    Perturbation magnitudes are log-uniform in 1e-7 .. 2; half the pairs carry a 4096*cls offset (lib/general.py:171)."""
    if kind == "cross":
        # UNRELATED boxes of wildly different scales at unrelated places (what an NMS set is mostly made of): a
        # sub-0.01 px box collapses to a point at the magnitude of the midpoint-shifted centres and then "contains"
        # every corner of the other box (see rbox_resolvable in rotated_iou.cuh)
        A1, _ = adversarial_pairs(rng, n, "any")
        B1, _ = adversarial_pairs(rng, n, "any")
        near = rng.random(n) < 0.3                           # some of them on top of each other
        B1[near, :2] = A1[near, :2] + (rng.normal(0, 1, (int(near.sum()), 2)) * A1[near, 2:3]).astype(np.float32)
        return A1, B1
    base = "any" if kind in ("edge", "corner") else kind
    scale = 10 ** rng.uniform(-3, 3, n) if base in ("tiny", "any") else rng.uniform(4, 120, n)
    asp = {"thin": 10 ** rng.uniform(0, 3, n), "any": 10 ** rng.uniform(0, 2.5, n)}.get(base, rng.uniform(1, 4, n))
    if kind in ("edge", "corner"):
        scale = np.where(rng.random(n) < 0.7, rng.uniform(2, 200, n), scale)
    w, h = scale, scale * asp
    sw = rng.random(n) < 0.5
    w, h = np.where(sw, h, w), np.where(sw, w, h)
    cx, cy = rng.uniform(0, 800, n), rng.uniform(0, 800, n)
    off = rng.integers(0, 17, n) * 4096.0 * (rng.random(n) < 0.5)
    cx, cy = cx + off, cy + off
    ang = rng.uniform(-90, 90, n)
    A = np.stack([cx, cy, w, h, ang], 1)
    mag = 10 ** rng.uniform(-7, 0.3, n)
    B = A.copy()
    if kind == "edge":
        fw, fh = 10 ** rng.uniform(-1.2, 0.3, n), 10 ** rng.uniform(-1.2, 0.3, n)
        B[:, 2], B[:, 3] = w * fw, h * fh
        # flush along +w and/or +h side: centre shifts by half the size difference along the box axes
        t = np.deg2rad(ang)
        ux, uy = np.cos(t), -np.sin(t)          # w axis (Appendix B corner formula: +c*w, -s*w)
        vx, vy = np.sin(t), np.cos(t)           # h axis
        sx = rng.choice([-1, 0, 1], n) * (w - B[:, 2]) / 2
        sy = rng.choice([-1, 0, 1], n) * (h - B[:, 3]) / 2
        B[:, 0] += sx * ux + sy * vx
        B[:, 1] += sx * uy + sy * vy
        mag = mag * 1e-3
    elif kind == "corner":
        B[:, 2], B[:, 3] = w * 10 ** rng.uniform(-1, 1, n), h * 10 ** rng.uniform(-1, 1, n)
        B[:, 4] = rng.uniform(-90, 90, n)
        ta, tb = np.deg2rad(ang), np.deg2rad(B[:, 4])
        sa, sb = rng.choice([-1, 1], (2, n)), rng.choice([-1, 1], (2, n))
        fa = np.where(rng.random(n) < 0.5, 1.0, rng.uniform(-1, 1, n))      # corner, or a point on A's edge
        pax = cx + sa[0] * fa * w / 2 * np.cos(ta) + sa[1] * h / 2 * np.sin(ta)
        pay = cy - sa[0] * fa * w / 2 * np.sin(ta) + sa[1] * h / 2 * np.cos(ta)
        B[:, 0] = pax - (sb[0] * B[:, 2] / 2 * np.cos(tb) + sb[1] * B[:, 3] / 2 * np.sin(tb))
        B[:, 1] = pay - (-sb[0] * B[:, 2] / 2 * np.sin(tb) + sb[1] * B[:, 3] / 2 * np.cos(tb))
        mag = mag * 1e-3
    big = np.maximum(w, h)
    B[:, 0] += rng.normal(0, 1, n) * mag * big
    B[:, 1] += rng.normal(0, 1, n) * mag * big
    B[:, 2] *= np.exp(rng.normal(0, 1, n) * mag)
    B[:, 3] *= np.exp(rng.normal(0, 1, n) * mag)
    dang = rng.normal(0, 1, n) * mag * 90
    mode = rng.integers(0, 4, n)
    B[:, 4] += np.where(mode == 0, 0, np.where(mode == 1, 90 + dang, np.where(mode == 2, 180 + dang, dang)))
    return A.astype(np.float32), B.astype(np.float32)


def test_host_build_of_product_iou_is_bitwise_the_oracle(hr):
    """rbox_iou_full (product source, host build) == oracle/rotated_ops.cpp bit for bit, degenerate regimes included."""
    rng = np.random.default_rng(1)
    for kind in ("normal", "thin", "tiny", "edge", "corner", "cross"):
        A, B = adversarial_pairs(rng, 4000, kind)
        _, full, _, _ = run_pairs(hr, A, B, 0.4)
        ref = oracle_pairs(A, B)
        assert np.array_equal(full.view(np.uint32), ref.view(np.uint32)), kind


@pytest.mark.parametrize("kind", ["normal", "thin", "tiny", "any", "edge", "corner", "cross"])
def test_nms_pair_decisions_match_oracle_on_adversarial_pairs(hr, kind):
    """Every pair decision of nms_mask_kernel's logic (bounding circle -> area / separating-axis bound -> gated fast
    estimate with the exact path inside the band) equals `oracle IoU > thr`; the gated estimate stays >= 20x inside
    the band.  Reports how many pairs sat within 1e-6 of a threshold (SURVEY.md §7 hard part 5)."""
    rng = np.random.default_rng(7)
    A, B = adversarial_pairs(rng, 300000, kind)
    worst, near = 0.0, 0
    for thr in THRESHOLDS:
        fast, full, gate, dec = run_pairs(hr, A, B, thr)
        want = full > thr                      # == the oracle (previous test)
        bad = np.nonzero(dec != want)[0]
        assert bad.size == 0, (kind, thr, bad.size, A[bad[:3]], B[bad[:3]], fast[bad[:3]], full[bad[:3]])
        pw = run_pairs.last_pairwise           # pairwise_iou_kernel's value (bounding-circle reject, else exact)
        assert np.array_equal(pw.view(np.uint32), full.view(np.uint32)), (kind, int((pw != full).sum()))
        if gate.any():
            worst = max(worst, float(np.abs(fast - full)[gate].max()))
        near += int((np.abs(full - thr) <= 1e-6).sum())
    assert worst < 1e-4, worst                 # kFastBand = 2e-3
    print(f"{kind}: gated max |fast - exact| = {worst:.2e}, pairs within 1e-6 of a threshold: {near}")


def test_appendix_b_is_not_geometric_for_near_duplicates(hr):
    """Documents the trap: a near-duplicate pair whose geometric IoU is ~1 but whose frozen-spec IoU is not, and which
    the product therefore sends down the exact path (gate False)."""
    a = np.array([[678.7636, 197.85667, 49.427803, 129.10036, -79.52609]], np.float32)
    b = np.array([[678.7636, 197.85666, 49.427887, 129.10034, -79.52609]], np.float32)
    fast, full, gate, dec = run_pairs(hr, a, b, 0.4)
    assert fast[0] > 0.9999
    assert full[0] == oracle_pairs(a, b)[0]
    assert not gate[0]
    assert dec[0] == (full[0] > 0.4)


def _cv2_iou(a, b):
    import cv2
    ra = ((float(a[0]), float(a[1])), (float(a[2]), float(a[3])), -float(a[4]))   # cv2 angles are clockwise
    rb = ((float(b[0]), float(b[1])), (float(b[2]), float(b[3])), -float(b[4]))
    rc, pts = cv2.rotatedRectangleIntersection(ra, rb)
    if rc == 0 or pts is None or len(pts) < 3:
        return 0.0
    inter = cv2.contourArea(cv2.convexHull(pts))
    return inter / (a[2] * a[3] + b[2] * b[3] - inter)


def test_oracle_rotated_iou_vs_opencv():
    """oracle/rotated_ops.cpp against cv2.rotatedRectangleIntersection on generic (well-conditioned) pairs.  OpenCV is
    fp32 inside, so this is a sanity cross-check of the restatement (SURVEY.md §8c measured 3.2e-4), not a bit-exact one;
    known answers pin the exact cases."""
    pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    n = 20000
    A = np.stack([rng.uniform(0, 100, n), rng.uniform(0, 100, n), rng.uniform(4, 60, n), rng.uniform(4, 60, n),
                  rng.uniform(-90, 90, n)], 1).astype(np.float32)
    B = A.copy()
    B[:, :2] += rng.normal(0, 12, (n, 2))
    B[:, 2:4] *= np.exp(rng.normal(0, 0.4, (n, 2)))
    B[:, 4] = rng.uniform(-90, 90, n)
    B = B.astype(np.float32)
    ref = oracle_pairs(A, B)
    cv = np.array([_cv2_iou(a, b) for a, b in zip(A, B)], np.float32)
    assert (ref > 0.05).mean() > 0.3                       # the sample does overlap
    err = np.abs(ref - cv)
    assert float(err.max()) < 2e-3, float(err.max())
    assert float(np.median(err)) < 1e-6
    # known answers: identical, 4x2 cross at 90 degrees, contained, disjoint, 180-degree wrap, zero area
    kat = torch.tensor([[0, 0, 4, 2, 0.0], [0, 0, 4, 2, 90.0], [0, 0, 4, 2, 180.0], [100, 100, 4, 2, 0.0],
                        [0, 0, 2, 1, 0.0], [0, 0, 0, 0, 0.0]])
    m = orot.pairwise_iou_rotated(kat, kat)
    assert abs(m[0, 0] - 1) < 1e-6 and abs(m[0, 2] - 1) < 1e-6 and abs(m[0, 1] - 1 / 3) < 1e-6
    assert m[0, 3] == 0 and abs(m[0, 4] - 0.25) < 1e-6 and m[0, 5] == 0 and m[5, 5] == 0


def test_oracle_nms_is_greedy_over_its_own_iou():
    """oracle nms_rotated == an independent numpy greedy loop over the oracle's IoU matrix (strict `>`, stable order)."""
    rng = np.random.default_rng(5)
    n = 400
    b = np.stack([rng.uniform(0, 200, n), rng.uniform(0, 200, n), rng.uniform(4, 60, n), rng.uniform(4, 60, n),
                  rng.uniform(-90, 90, n)], 1).astype(np.float32)
    s = rng.random(n).astype(np.float32)
    s[::9] = s[0]
    iou = orot.pairwise_iou_rotated(torch.from_numpy(b), torch.from_numpy(b)).numpy()
    order = np.argsort(-s, kind="stable")
    dead, keep = np.zeros(n, bool), []
    for i in order:
        if dead[i]:
            continue
        keep.append(i)
        dead |= iou[i] > 0.3
    got = orot.nms_rotated(torch.from_numpy(b), torch.from_numpy(s), 0.3).numpy()
    assert np.array_equal(got, np.array(keep))
