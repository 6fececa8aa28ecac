"""CPU tests of the host side: the C ABI library loads and exports every symbol the header
declares, fails loudly without a GPU, and the multi-GPU sharding logic works (gloo, world size 2)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "iglb200.h")).read()
    return sorted(set(re.findall(r"\b(iglb200_[a-z0-9_]+)\s*\(", hdr)))


def test_product_library_exports_the_whole_abi():
    import ctypes

    import libigl_b200
    from libigl_b200 import build
    build.build()
    lib = ctypes.CDLL(libigl_b200.LIB_PATH)
    syms = _declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/iglb200.h but not exported"
    # the python binding binds exactly these too
    from libigl_b200.api import Api
    missing = set(Api.SYMBOLS) - set(syms)
    assert not missing, missing


def test_no_cpu_fallback_without_a_gpu():
    import libigl_b200
    from libigl_b200 import synth
    api = libigl_b200.lib()
    if api.device_count() > 0:
        pytest.skip("a GPU is present")
    assert "EMULATION" not in api.version()
    V, F = synth.tetrahedron()
    with pytest.raises(libigl_b200.IglB200Error) as e:
        api.Mesh(V, F)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
    with pytest.raises(libigl_b200.IglB200Error):
        libigl_b200.signed_distance(np.zeros((1, 3)), V, F)
    # every other compute entry point refuses the same way: cloud, grids, flood fill, marching cubes, winding number
    calls = [
        lambda: api.Cloud(V, V, np.ones(4)),
        lambda: api.voxel_grid(V, 0.1, 8, 1),
        lambda: api.flood_fill([2, 2, 2], np.zeros(8)),
        lambda: api.marching_cubes(np.zeros(8), np.zeros((8, 3)), 2, 2, 2),
        lambda: api.offset_surface(V, F, 0.1, 8),
        lambda: api.winding_number(V, F, np.zeros((1, 3))),
        lambda: api.check(api.lib.iglb200_init(1)),
    ]
    for c in calls:
        with pytest.raises(libigl_b200.IglB200Error) as e:
            c()
        assert e.value.code == -2, str(e.value)


def test_product_never_references_the_oracle_or_the_emulation():
    pkg = os.path.join(ROOT, "libigl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".h", ".cu")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "libiglb200_emul" not in src, f


def test_shard_range():
    from libigl_b200.dist import shard_range
    for n in (0, 1, 7, 16, 1000003):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch, torch.distributed as dist
import emul
from libigl_b200 import synth
from libigl_b200.dist import shard_range, broadcast_mesh, gather_rows
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
api = emul.api()
V, F = synth.icosphere(3)
P = synth.grid_points(*synth.bbox(V), 16, 1.5)
mesh = api.Mesh(V, F, 2) if rank == 0 else None          # trees are built once, on rank 0
mesh, keep = broadcast_mesh(api, mesh, src=0)
lo, hi = shard_range(P.shape[0], rank, world)
S, I, C, W = mesh.signed_distance(P[lo:hi], return_W=True)
S_all = gather_rows(S, P.shape[0]); I_all = gather_rows(I, P.shape[0]); C_all = gather_rows(np.asarray(C), P.shape[0])
if rank == 0:
    S1, I1, C1, W1 = mesh.signed_distance(P, return_W=True)
    assert np.array_equal(S_all, S1, equal_nan=True) and np.array_equal(I_all, I1) and np.array_equal(C_all, C1)
    print("SHARDED_OK", world, P.shape[0])
dist.barrier()
dist.destroy_process_group()
'''


def test_query_sharding_world_size_2_gloo(tmp_path):
    import emul
    emul.build()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "SHARDED_OK 2" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_c_abi_header_is_plain_c(tmp_path):
    """include/iglb200.h is what an FFI binds: it must compile as C99 and link against the library from C."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    src = tmp_path / "abi.c"
    src.write_text('#include "iglb200.h"\n#include <stdio.h>\n'
                   'int main(void) { printf("%s %d\\n", iglb200_version(), iglb200_device_count() >= 0); return 0; }\n')
    import libigl_b200
    from libigl_b200 import build
    build.build()
    libdir = os.path.dirname(libigl_b200.LIB_PATH)
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L" + libdir, "-liglb200", "-Wl,-rpath," + libdir], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("iglb200"), r.stdout + r.stderr


def _sd_arrays(n):
    return (np.empty(n), np.empty(n, np.int32), np.empty((n, 3), order="F"), np.empty(n))


def test_device_group_deals_chunks_to_every_device(emul_api, monkeypatch):
    """iglb200_init(n): the mesh is replicated, host queries are dealt chunk by chunk to the workers of every device,
    and the result is bit-identical to the single-device call — in every host layout the chunk pipeline takes."""
    from libigl_b200 import synth
    api = emul_api
    V, F = synth.icosphere(3)
    P = synth.grid_points(*synth.bbox(V), 36, 1.5)           # 46 656 rows: above the inline threshold
    api.init(1)
    m1 = api.Mesh(V, F, 2)
    assert m1.info()["n_devices"] == 1
    ref = m1.signed_distance(P, return_W=True)
    D1 = m1.squared_distance(P)
    W1 = m1.fast_winding_number(P)
    m1.close()
    try:
        api.init(3)
        assert api.group_info()["devices"] == 3
        m = api.Mesh(V, F, 2)
        assert m.info()["n_devices"] == 3
        for layout in ("C", "F"):
            for dt in (np.float64, np.float32):
                Pl = np.array(P, dtype=dt, order=layout)
                out = m.signed_distance(Pl, return_W=True)
                if dt == np.float64:
                    for a, b in zip(out, ref):
                        assert np.array_equal(a, b, equal_nan=True), (layout, dt)
                else:
                    out64 = m.signed_distance(Pl.astype(np.float64), return_W=True)
                    for a, b in zip(out, out64):
                        assert np.array_equal(a, b, equal_nan=True), (layout, dt)
        for a, b in zip(m.squared_distance(np.ascontiguousarray(P)), D1):
            assert np.array_equal(a, b)
        assert np.array_equal(m.fast_winding_number(P), W1)
        m.close()
    finally:
        api.init(1)


def test_row_major_outputs_and_float_outputs(emul_api):
    """C / N as dense row-major matrices and S / C / W as float32 go through the same chunk pipeline."""
    import ctypes as C
    from libigl_b200 import synth
    api = emul_api
    V, F = synth.icosphere(2)
    P = synth.random_points(40000, -1.5, 1.5, 3)
    m = api.Mesh(V, F, 2, pseudonormals=True)
    S0, I0, C0, N0 = m.signed_distance(P, sign_type=0)
    n = P.shape[0]
    vp = C.c_void_p
    for dt, f64 in ((np.float64, 1), (np.float32, 0)):
        S = np.empty(n, dt); I = np.empty(n, np.int32); Cr = np.empty((n, 3), dt); Nr = np.empty((n, 3), dt)
        Pr = np.ascontiguousarray(P)
        api.check(api.lib.iglb200_signed_distance(m.h, Pr.ctypes.data_as(vp), 1, n, 3, 1, 0, float(np.finfo(float).tiny),
                                                  float(np.finfo(float).max), S.ctypes.data_as(vp), f64, I.ctypes.data_as(vp),
                                                  Cr.ctypes.data_as(vp), f64, 3, 1, None, 1, Nr.ctypes.data_as(vp), f64, 3, 1))
        assert np.array_equal(I, I0)
        assert np.array_equal(S, S0.astype(dt)) and np.array_equal(Cr, np.asarray(C0).astype(dt)) and np.array_equal(Nr, np.asarray(N0).astype(dt))
    # N is an output of PSEUDONORMAL only: other sign types leave the caller's array untouched
    Nr = np.full((n, 3), 7.0)
    S = np.empty(n); I = np.empty(n, np.int32); Cr = np.empty((n, 3))
    api.check(api.lib.iglb200_signed_distance(m.h, Pr.ctypes.data_as(vp), 1, n, 3, 1, 4, float(np.finfo(float).tiny),
                                              float(np.finfo(float).max), S.ctypes.data_as(vp), 1, I.ctypes.data_as(vp),
                                              Cr.ctypes.data_as(vp), 1, 3, 1, None, 1, Nr.ctypes.data_as(vp), 1, 3, 1))
    assert (Nr == 7.0).all()
    m.close()


COOP_WORKER = r'''
import os, sys, ctypes as C
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch, torch.distributed as dist
import emul
from libigl_b200 import synth
from libigl_b200.dist import broadcast_mesh, SharedArrays
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
api = emul.api()
V, F = synth.icosphere(3)
P = synth.grid_points(*synth.bbox(V), 30, 1.5)
n = P.shape[0]
mesh = api.Mesh(V, F, 2) if rank == 0 else None
mesh, keep = broadcast_mesh(api, mesh, src=0)
# ONE query set in shared host memory; every rank takes chunks off one shared counter
sh = SharedArrays("iglb200_coop_test", n, create=(rank == 0))
if rank == 0:
    sh.P[:] = P.T
    sh.counter[0] = 0
    sh.S[:] = -777.0
dist.barrier()
chunk = 2048
mesh.signed_distance_coop(sh, chunk)
dist.barrier()
if rank == 0:
    S1, I1, C1, W1 = mesh.signed_distance(P, return_W=True)
    assert np.array_equal(sh.S, S1, equal_nan=True) and np.array_equal(sh.I, I1) and np.array_equal(sh.C.T, C1)
    assert np.array_equal(sh.W, W1, equal_nan=True)
    assert sh.counter[0] >= (n + chunk - 1) // chunk + world - 1     # every rank saw the end of the rows
    print("COOP_OK", world, n)
dist.barrier()
sh.close(unlink=(rank == 0))
dist.destroy_process_group()
'''


def test_cooperative_processes_share_one_query_set_gloo(tmp_path):
    import emul
    emul.build()
    script = tmp_path / "coop.py"
    script.write_text(COOP_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29613", str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "COOP_OK 2" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_sharded_device_queries_match_the_fused_path(emul_api):
    """begin / run (two devices, super-chunks off one counter) / finish == the single-device fused kernels, bit for bit,
    for signed distance (FWN and UNSIGNED, bounded and unbounded), squared distance and the winding number alone."""
    import ctypes as C
    from libigl_b200 import synth
    api, L = emul_api, emul_api.lib
    vp = C.c_void_p
    V, F = synth.icosphere(3)
    P = synth.grid_points(*synth.bbox(V), 40, 1.6)
    n = P.shape[0]
    Pc = np.ascontiguousarray(P.T)           # [3, n] = column-major n x 3
    tiny, huge = float(np.finfo(float).tiny), float(np.finfo(float).max)
    ptr = lambda a: a.ctypes.data_as(vp)
    try:
        api.init(2)
        m = api.Mesh(V, F, 2)
        scratch = np.zeros(int(L.iglb200_shard_scratch_bytes(n)), np.uint8)
        qs = np.zeros(int(L.iglb200_query_scratch_bytes(n)), np.uint8)
        sup = int(L.iglb200_shard_super_rows(n, 2))
        assert sup % 1024 == 0 and sup >= 8192
        sup = 8192                                    # the smallest super-chunk: 64 000 rows make eight
        for mode, sign, lo, hi in ((0, 4, tiny, huge), (0, 4, 0.05, 0.4), (0, 3, tiny, huge), (1, 0, 0, 0), (2, 0, 0, 0)):
            S, I, Cc, W = np.full(n, -5.0), np.full(n, -5, np.int32), np.full((3, n), -5.0), np.full(n, -5.0)
            api.set_device(0)
            api.check(L.iglb200_shard_begin_dev(m.h, ptr(Pc), n, ptr(scratch), 1, sup, None))
            outs = (ptr(S), ptr(I), ptr(Cc), ptr(W))
            api.check(L.iglb200_shard_run_dev(m.h, ptr(Pc), n, ptr(scratch), mode, sign, lo, hi, 2.0, 1, *outs, sup, 3, 0, 2, None))   # device 0: 3 pieces of its slice
            api.set_device(1)
            api.check(L.iglb200_shard_run_dev(m.h, ptr(Pc), n, ptr(scratch), mode, sign, lo, hi, 2.0, 1, *outs, sup, 0, 1, 2, None))   # device 1: its slice, then the rest of the other
            counters = scratch[:1024].view(np.uint64)   # [0] winding-number pieces, [2 + k] super-chunks of slice k
            assert int(counters[0] + counters[2:4].sum()) >= 8
            api.set_device(0)
            api.check(L.iglb200_shard_finish_dev(m.h, n, ptr(scratch), mode, sign, lo, hi, 1, *outs, None))
            S1, I1, C1, W1 = np.empty(n), np.empty(n, np.int32), np.empty((3, n)), np.empty(n)
            if mode == 0:
                api.check(L.iglb200_signed_distance_dev(m.h, ptr(Pc), n, sign, lo, hi, ptr(S1), ptr(I1), ptr(C1), ptr(W1), None, ptr(qs), 1, None))
                assert np.array_equal(S, S1, equal_nan=True) and np.array_equal(I, I1) and np.array_equal(Cc, C1) and np.array_equal(W, W1, equal_nan=True)
                if lo > tiny:
                    assert np.isnan(S).any() and not np.isnan(S).all()
            elif mode == 1:
                api.check(L.iglb200_squared_distance_dev(m.h, ptr(Pc), n, ptr(S1), ptr(I1), ptr(C1), ptr(qs), 1, None))
                assert np.array_equal(S, S1) and np.array_equal(I, I1) and np.array_equal(Cc, C1)
            else:
                api.check(L.iglb200_fwn_eval_dev(m.h, ptr(Pc), n, 2.0, ptr(W1), ptr(qs), 1, None))
                assert np.array_equal(W, W1)
        m.close()
    finally:
        api.set_device(0)
        api.init(1)


def test_sharded_slices_are_handed_out_once_whoever_comes_first(emul_api):
    """Three slices of the sorted order, callers arriving in every order and with every mix of caps: each super-chunk is
    taken exactly once (own slice first, then the slice with the most left), an unknown slice index is refused, and the
    result is the single call's bit for bit."""
    import ctypes as C
    from libigl_b200 import synth
    api, L = emul_api, emul_api.lib
    vp = C.c_void_p
    V, F = synth.icosphere(2)
    P = synth.random_points(90_000, -1.4, 1.4, 5)
    n = P.shape[0]
    Pc = np.ascontiguousarray(P.T)
    ptr = lambda a: a.ctypes.data_as(vp)
    sup = 8192
    nsup = -(-n // sup)                                # 11 super-chunks: slices of 3, 4 and 4
    m = api.Mesh(V, F, 2)
    scratch = np.zeros(int(L.iglb200_shard_scratch_bytes(n)), np.uint8)
    qs = np.zeros(int(L.iglb200_query_scratch_bytes(n)), np.uint8)
    S1, I1, C1 = np.empty(n), np.empty(n, np.int32), np.empty((3, n))
    api.check(L.iglb200_squared_distance_dev(m.h, ptr(Pc), n, ptr(S1), ptr(I1), ptr(C1), ptr(qs), 1, None))
    # (shard, cap) in calling order; cap 0 = no cap
    plans = [[(2, 0), (0, 0), (1, 0)],                 # the last slice's owner comes first and takes everything
             [(0, 1), (1, 1), (2, 1), (1, 0)],         # one each, then shard 1 finishes its slice and both others
             [(1, 5), (1, 0)],                         # one caller past its own slice (4) into the fullest other one
             [(0, 2), (2, 2), (0, 2), (2, 0), (1, 0)]]
    for plan in plans:
        S, I, Cc, W = np.full(n, -5.0), np.full(n, -5, np.int32), np.full((3, n), -5.0), np.full(n, -5.0)
        outs = (ptr(S), ptr(I), ptr(Cc), ptr(W))
        api.check(L.iglb200_shard_begin_dev(m.h, ptr(Pc), n, ptr(scratch), 1, sup, None))
        for shard, cap in plan:
            api.check(L.iglb200_shard_run_dev(m.h, ptr(Pc), n, ptr(scratch), 1, 0, 0.0, 0.0, 2.0, 1, *outs, sup, cap, shard, 3, None))
        api.check(L.iglb200_shard_finish_dev(m.h, n, ptr(scratch), 1, 0, 0.0, 0.0, 1, *outs, None))
        counters = scratch[:1024].view(np.uint64)[2:5]
        sizes = [nsup * (k + 1) // 3 - nsup * k // 3 for k in range(3)]
        assert all(int(c) >= sz for c, sz in zip(counters, sizes)), (plan, counters, sizes)
        assert np.array_equal(S, S1) and np.array_equal(I, I1) and np.array_equal(Cc, C1), plan
    for shard, nshards in ((3, 3), (-1, 3), (0, 0), (0, 65)):
        rc = L.iglb200_shard_run_dev(m.h, ptr(Pc), n, ptr(scratch), 1, 0, 0.0, 0.0, 2.0, 1, ptr(S1), ptr(I1), ptr(C1), None, sup, 0, shard, nshards, None)
        assert rc != 0
    m.close()

