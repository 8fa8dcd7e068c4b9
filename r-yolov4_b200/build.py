"""Builds libryolo_b200.so (hand-written sm_100a CUDA kernels behind a C ABI) in-tree with nvcc.

    python r-yolov4_b200/build.py [--force] [--verbose]

Cross-compiles without a GPU.  Translation units whose results are compared bit-for-bit with the
CPU oracle are compiled with --fmad=false (no FMA contraction); the tensor-core conv stack is not.
"""
import argparse
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
SO = os.path.join(HERE, "libryolo_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", CSRC,
          "-I", os.path.join(os.path.dirname(HERE), "include")]
# file -> extra flags
SOURCES = {
    "lib.cu": [],
    "nms.cu": ["--fmad=false"],
    "decode.cu": ["--fmad=false"],
    "loss.cu": ["--fmad=false"],
    "labels.cu": ["--fmad=false"],
}
for _opt in ("conv.cu", "elementwise.cu", "wgrad.cu", "backward.cu", "kfloss.cu"):
    if os.path.exists(os.path.join(CSRC, _opt)):
        SOURCES[_opt] = []


def _newer(dst, deps):
    if not os.path.exists(dst):
        return False
    t = os.path.getmtime(dst)
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    inc = os.path.join(os.path.dirname(HERE), "include")
    if os.path.isdir(inc):
        headers += [os.path.join(inc, f) for f in os.listdir(inc) if f.endswith(".h")]
    objs, procs = [], []
    for src, extra in SOURCES.items():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        if not force and _newer(o, [s] + headers + [__file__]):
            continue
        cmd = ["nvcc"] + ARCH + COMMON + extra + ["-c", s, "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out:
            print(out)
    if force or procs or not _newer(SO, objs):
        cmd = ["nvcc"] + ARCH + ["-shared", "-Xcompiler", "-fPIC", "-o", SO] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return SO


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
