"""Host emulation of the device kernels — TEST SCAFFOLDING ONLY.

Compiles libigl_b200/csrc/*.cu with g++ and -DIGLB200_EMUL (every kernel launch
becomes a loop, see csrc/platform.h) into tests/_build/libiglb200_emul.so so that
`pytest -m "not gpu"` can check the kernels' logic against the oracle in a
container without a GPU.  The product package never loads this library and has
no CPU path; GPU parity is established by the `-m gpu` tests through the real
libigl_b200/libiglb200.so.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_SRC = os.path.join(_ROOT, "libigl_b200", "csrc")
_OUT = os.path.join(_HERE, "_build", "libiglb200_emul.so")


def build():
    srcs = [os.path.join(_SRC, f) for f in ("build.cu", "iglb200.cu")]
    deps = [os.path.join(_SRC, f) for f in os.listdir(_SRC)] + [os.path.join(_ROOT, "include", "iglb200.h")]
    if os.path.exists(_OUT) and os.path.getmtime(_OUT) >= max(os.path.getmtime(d) for d in deps):
        return _OUT
    os.makedirs(os.path.dirname(_OUT), exist_ok=True)
    cuda_inc = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-DIGLB200_EMUL",
           "-I" + cuda_inc, "-x", "c++"] + srcs + ["-o", _OUT]
    subprocess.run(cmd, check=True)
    return _OUT


_api = None


def api():
    global _api
    if _api is None:
        from libigl_b200.api import Api
        _api = Api(ctypes.CDLL(build()))
    return _api
