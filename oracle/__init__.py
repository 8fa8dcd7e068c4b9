"""ORACLE — test infrastructure only.

CPU restatements of the reference (yingkunwu/R-YOLOv4) hot path used as the parity checker and
as the `cpu_baseline` / `--impl reference` arm of bench.py.  Nothing under ``r-yolov4_b200/`` may
import this package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` do.

Pinning status:
  * decode / build_targets / CSL loss / KFIoU loss / post_process glue / conv stack: PINNED against
    outputs of the reference's own Python, executed in the build container from /root/reference by
    tests/golden/make_golden.py (fixtures committed under tests/golden/).
  * rotated IoU / rotated NMS (detectron2, un-vendored, unpinned git HEAD): PARITY UNPINNED —
    restated from the published algorithm, anchored by known-answer tests and
    an fp32 cross-check against OpenCV.
"""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(verbose: bool = False) -> str:
    """Compile the C++ part of the oracle (rotated IoU / NMS). Returns the .so path."""
    so = os.path.join(_HERE, "liboracle_rotated.so")
    src = os.path.join(_HERE, "rotated_ops.cpp")
    if (not os.path.exists(so)) or os.path.getmtime(so) < os.path.getmtime(src):
        out = subprocess.run(["make", "-C", _HERE, "liboracle_rotated.so"], capture_output=True, text=True)
        if out.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
        if verbose:
            print(out.stdout)
    return so
