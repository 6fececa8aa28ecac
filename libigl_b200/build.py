"""Builds libigl_b200/libiglb200.so in-tree with nvcc for sm_100a only.

    python -m libigl_b200.build [--force]

build.cu is compiled with -fmad=false: every fp32 value that steers a branch of
the reference (split costs, cluster centres/radii) must be reproduced bit for
bit.  iglb200.cu (traversals) keeps FMA contraction; the places that must stay
un-fused use explicit __fmul_rn/__fadd_rn/__dmul_rn... intrinsics.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libiglb200.so")
OBJ = os.path.join(HERE, "csrc", "_obj")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
          "--expt-relaxed-constexpr"]
UNITS = {
    "build.cu": ["-fmad=false"],
    "iglb200.cu": [],
}


def nvcc():
    exe = shutil.which("nvcc") or os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def _deps():
    d = [os.path.join(SRC, f) for f in os.listdir(SRC) if f.endswith((".h", ".cu", ".cuh"))]
    d.append(os.path.join(os.path.dirname(HERE), "include", "iglb200.h"))
    d.append(os.path.abspath(__file__))
    return d


def build(force: bool = False, verbose: bool = False) -> str:
    newest = max(os.path.getmtime(p) for p in _deps())
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= newest:
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    objs = []
    procs = []
    for unit, extra in UNITS.items():
        o = os.path.join(OBJ, unit.replace(".cu", ".o"))
        objs.append(o)
        cmd = [nvcc()] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + \
              ["-c", os.path.join(SRC, unit), "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    link = [nvcc()] + ARCH + ["-shared", "-o", OUT] + objs + ["-lcudart"]
    subprocess.run(link, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
