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
