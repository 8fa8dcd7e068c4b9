import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(params=["direct", "tma", "tma_red"])
def epi(request):
    """Runs a conv test under each bf16 epilogue of the conv kernel: per-thread 16-byte stores, TMA slab stores,
    TMA slab stores + TMA reduce-add for dgrad's accumulation (library knobs epi_tma / epi_maxbn)."""
    from ryolo_b200 import _lib as L
    mode = {"direct": 0, "tma": 1, "tma_red": 2}[request.param]
    old = (L.lib().ryolo_knob(4), L.lib().ryolo_knob(5))
    L.tune(epi_tma=mode, epi_maxbn=256)
    yield request.param
    L.tune(epi_tma=old[0], epi_maxbn=old[1])


@pytest.fixture(params=["pair_any", "kgrp_wide", "plain", "wres"])
def issue(request):
    """Runs a conv test under each operand-ring variant of the conv kernel (library knobs pair / kgrp): cluster pairs
    with weight-tile multicast forced on every eligible layer, up to three K blocks per ring stage on any tile width,
    one K block per stage without pairing, and weight matrices of up to 128 KB resident in shared memory."""
    from ryolo_b200 import _lib as L
    names = ["halo", "dbg", "wg_split", "wg_dbg", "epi_tma", "epi_maxbn", "wg_tapgrp", "bn_bwd", "wg_trans", "sw64", "nacc",
             "pdl", "ssa", "wg_boxes", "ew_regs", "nms_band", "bn_fuse", "kgrp", "pair", "wres", "wg_x32"]
    old = [L.lib().ryolo_knob(names.index(k)) for k in ("kgrp", "pair", "wres")]
    L.tune(**{"pair_any": dict(pair=1024 + 32, kgrp=0, wres=0), "kgrp_wide": dict(pair=0, kgrp=3, wres=0),
              "plain": dict(pair=0, kgrp=0, wres=0), "wres": dict(pair=0, kgrp=1, wres=1024 + 128)}[request.param])
    yield request.param
    L.tune(kgrp=old[0], pair=old[1], wres=old[2])
