"""The Eigen-signature C++ wrappers (include/iglb200/*.h) compiled against the test-only
mini_eigen stand-in and run on the reference's own known-answer tests.
CPU: linked against the host-emulated kernels (logic check).  GPU: linked against the product library."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "_build")


def _compile(libdir, libname, out):
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, out)
    cmd = ["g++", "-O1", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "tests", "cpp"),
           os.path.join(ROOT, "tests", "cpp", "test_headers.cpp"), "-o", exe, "-L" + libdir, "-l" + libname,
           "-Wl,-rpath," + libdir]
    subprocess.run(cmd, check=True)
    return exe


def test_headers_against_emulated_kernels():
    import emul
    emul.build()
    exe = _compile(BUILD, "iglb200_emul", "test_headers_emul")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_headers_on_gpu():
    import libigl_b200
    libigl_b200.lib()
    exe = _compile(os.path.join(ROOT, "libigl_b200"), "iglb200", "test_headers_gpu")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr


def test_work_donation_protocol_model():
    """The mailbox / slot protocol the search kernel uses to cut the last long walks of a launch into pieces
    (csrc/iglb200.cu: cp_donate, cp_take_task, cp_finish_shared), modelled with one host thread per warp: every thread ends,
    every donated task runs once, every packet's result is the true minimum and is stored exactly once."""
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "donation_model")
    subprocess.run(["g++", "-O1", "-std=c++17", "-pthread", os.path.join(ROOT, "tests", "cpp", "donation_model.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe, "60"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr
