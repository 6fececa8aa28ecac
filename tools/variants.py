#!/usr/bin/env python
"""Tuning aid: builds variants of the query kernels (extra -D flags on iglb200.cu) into build/variants/ and prints a shell
loop that runs bench.py against each on the GPU box (IGLB200_LIBRARY selects the library).

    python tools/variants.py name1:-DCP_MIN_BLOCKS=16 name2:-DCP_MIN_BLOCKS=24,-DCP_CHUNK=8 ...
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libigl_b200 import build as B  # noqa: E402

OUT = os.path.join(ROOT, "build", "variants")


def one(spec):
    name, _, defs = spec.partition(":")
    defs = [d for d in defs.split(",") if d]
    o = os.path.join(OUT, f"iglb200_{name}.o")
    so = os.path.join(OUT, f"libiglb200_{name}.so")
    cmd = [B.nvcc()] + B.ARCH + B.COMMON + defs + ["-c", os.path.join(B.SRC, "iglb200.cu"), "-o", o]
    subprocess.run(cmd, check=True)
    subprocess.run([B.nvcc()] + B.ARCH + ["-shared", "-o", so, o, os.path.join(B.OBJ, "build.o"), "-lcudart"], check=True)
    os.remove(o)
    return name, so


def main():
    B.build()
    os.makedirs(OUT, exist_ok=True)
    for f in os.listdir(OUT):
        os.remove(os.path.join(OUT, f))
    with ThreadPoolExecutor(4) as ex:
        res = list(ex.map(one, sys.argv[1:]))
    for name, so in res:
        print(name, os.path.relpath(so, ROOT))


if __name__ == "__main__":
    main()
