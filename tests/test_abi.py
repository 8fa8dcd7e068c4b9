"""The C-ABI library loads on a CPU-only box and exports every symbol include/ryolo_b200.h declares."""
import ctypes
import os
import re

import pytest
import torch

from tests.util import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "ryolo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ryolo_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    import importlib.util
    spec = importlib.util.spec_from_file_location("ryolo_build", os.path.join(ROOT, "r-yolov4_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    so = mod.build()
    h = ctypes.CDLL(so)
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(h, n), f"{n} declared in include/ryolo_b200.h but not exported"
    h.ryolo_abi_version.restype = ctypes.c_int
    assert h.ryolo_abi_version() >= 1


def test_binding_table_matches_header():
    import ryolo_b200
    from ryolo_b200 import _lib
    assert set(_declared()) == set(_lib.SIGNATURES.keys())
    assert ryolo_b200.lib() is not None


def test_no_cpu_fallback():
    import ryolo_b200
    with pytest.raises(ryolo_b200.RyoloError):
        ryolo_b200.post_process(torch.zeros(1, 8, 8))
    with pytest.raises(ryolo_b200.RyoloError):
        ryolo_b200.nms_rotated(torch.zeros(4, 5), torch.zeros(4), 0.5)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "r-yolov4_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                s = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", s, flags=re.M), os.path.join(dp, f)
                assert "oracle/" not in s or f.endswith((".cuh", ".cu")), os.path.join(dp, f)


def test_tuning_knobs_host_api():
    """ryolo_tune / ryolo_knob are host-only: defaults, overrides, unknown keys (include/ryolo_b200.h)."""
    from ryolo_b200 import _lib as L
    lib = L.lib()
    names = ["halo", "dbg", "wg_split", "wg_dbg", "epi_tma", "epi_maxbn", "wg_tapgrp", "bn_bwd", "wg_trans", "sw64",
             "nacc", "pdl", "ssa", "wg_boxes", "ew_regs", "nms_band", "bn_fuse", "kgrp", "pair", "wres", "wg_x32"]
    before = [lib.ryolo_knob(i) for i in range(len(names))]
    if not any(os.environ.get("RYOLO_" + n.upper()) for n in ("kgrp", "pair", "wres", "wg_x32")):
        # operand-ring variants of the conv / wgrad kernels as shipped (csrc/lib.cu)
        assert [before[names.index(n)] for n in ("kgrp", "pair", "wres", "wg_x32")] == [1, 256, 96, 1]
    assert before[names.index("dbg")] == 0 and before[names.index("wg_dbg")] == 0, "timing experiments must be off"
    assert before[names.index("epi_tma")] == 2 and before[names.index("wg_trans")] == 0 and before[names.index("bn_bwd")] == 3
    try:
        L.tune(epi_tma=0, wg_split=2)
        assert lib.ryolo_knob(names.index("epi_tma")) == 0 and lib.ryolo_knob(names.index("wg_split")) == 2
        with pytest.raises(L.RyoloError):
            L.tune(no_such_knob=1)
        assert lib.ryolo_knob(-1) == 0 and lib.ryolo_knob(999) == 0
    finally:
        L.tune(**dict(zip(names, before)))
    assert [lib.ryolo_knob(i) for i in range(len(names))] == before


def test_label_encoder_rejects_cpu_tensors():
    import ryolo_b200
    with pytest.raises(ryolo_b200.RyoloError):
        ryolo_b200.encode_labels(torch.zeros(3, 10))
