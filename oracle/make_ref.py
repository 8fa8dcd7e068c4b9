"""ORACLE — TEST INFRASTRUCTURE ONLY.  Recipe that stages the REFERENCE'S OWN Python for the hot path under
oracle/_ref/ (git-ignored, travels to the GPU box like a built .so) so that bench.py's `--impl reference` /
`cpu_baseline` legs time the unmodified reference code — `Yolo`, `ComputeCSLLoss`, `KFLoss`, `post_process` — on the
box's host cores instead of the oracle port.

    python oracle/make_ref.py            (also run by __graft_entry__.build() when /root/reference is present)

Nothing is copied into the repository's history: the files are read from /root/reference where they lie and written
only below oracle/_ref/.  The reference has no setup.py / pyproject.toml (so `pip install --target` does not apply) and
imports detectron2, which is absent and un-vendored: the two symbols it needs (`nms_rotated`, `pairwise_iou_rotated`)
are provided by a stub package that forwards to oracle/rotated_ops.cpp — the NMS boundary therefore stays
"parity unpinned" in this arm too."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
DST = os.path.join(HERE, "_ref")
FILES = ["model/__init__.py", "model/yolo.py", "model/backbone.py", "model/neck.py", "model/utils.py",
         "model/yololayer.py", "lib/__init__.py", "lib/loss.py", "lib/general.py"]

STUB = '''"""detectron2 stub (see oracle/make_ref.py): forwards to the oracle's C++ restatement."""
'''


def build():
    """Returns the staging dir, or None when /root/reference is absent (GPU box: uses what was staged here)."""
    if not os.path.isdir(REF):
        return DST if os.path.isdir(os.path.join(DST, "rref")) else None
    pkg = os.path.join(DST, "rref")
    for f in FILES:
        src, dst = os.path.join(REF, f), os.path.join(pkg, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.exists(src):
            shutil.copyfile(src, dst)
        elif f.endswith("__init__.py"):
            open(dst, "w").close()
    d2 = os.path.join(DST, "stubs", "detectron2", "layers")
    os.makedirs(d2, exist_ok=True)
    open(os.path.join(DST, "stubs", "detectron2", "__init__.py"), "w").write(STUB)
    open(os.path.join(d2, "__init__.py"), "w").write(STUB)
    open(os.path.join(d2, "nms.py"), "w").write(
        STUB + "from oracle.rotated import nms_rotated  # noqa: F401\n")
    open(os.path.join(d2, "rotated_boxes.py"), "w").write(
        STUB + "from oracle.rotated import pairwise_iou_rotated  # noqa: F401\n")
    return DST


def load():
    """Imports the staged reference; returns a namespace with Yolo, ComputeCSLLoss, ComputeKFIoULoss, KFLoss,
    post_process — or None if nothing is staged."""
    d = build()
    if d is None:
        return None
    root = os.path.dirname(HERE)
    for p in (root, os.path.join(d, "stubs"), os.path.join(d, "rref")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import types
    for name in ("model", "lib"):                       # the reference's top-level package names
        m = sys.modules.get(name)
        if m is not None and not getattr(m, "__file__", "").startswith(d):
            del sys.modules[name]
    from lib import general as rgen
    from lib import loss as rloss
    from model.yolo import Yolo
    ns = types.SimpleNamespace(Yolo=Yolo, ComputeCSLLoss=rloss.ComputeCSLLoss, ComputeKFIoULoss=rloss.ComputeKFIoULoss,
                               KFLoss=rloss.KFLoss, post_process=rgen.post_process, root=d)
    return ns


if __name__ == "__main__":
    print(build())
