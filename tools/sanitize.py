"""Small end-to-end run of every device path for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import libigl_b200  # noqa: E402
from libigl_b200 import synth  # noqa: E402

api = libigl_b200.lib()
V, F = synth.torus_knot(300, 16)
P = synth.grid_points(*synth.bbox(V), 20, 1.1)
m = api.Mesh(V, F, 2, pseudonormals=True)
S, I, C, W = m.signed_distance(np.asfortranarray(P), return_W=True)          # streamed path
S2, I2, C2, N2 = m.signed_distance(P, 0, -0.2, 0.3)                          # pseudonormal + bounds, whole-batch path
D, I3, C3 = m.squared_distance(P[:1000])
Wq = m.fast_winding_number(P[:1000], 2.0)
print("visits", m.visit_counts(P[:2048]))
m.close()
Pc, N, A = synth.oriented_cloud(V, F, 3000, 2)
c = api.Cloud(Pc, N, A, 2)
Wc = c.eval(P[:2000], 2.0)
c.close()
# exact winding number (open soup: chains with every kind of multiplicity) and the grid pipeline
rng = np.random.default_rng(0)
Fo = np.concatenate([F[rng.random(F.shape[0]) > 0.2], F[:50], F[:30][:, ::-1]])
mw = api.Mesh(V, Fo, 2, winding_number=True)
Wx = mw.winding_number(P[:3000])
Sx = mw.signed_distance(P[:3000], 2)[0]
mw.close()
SV, SF, GV, side, Sg = api.offset_surface(V, F, 0.05, 24)
Sn = Sg.copy(); Sn[rng.random(Sn.size) < 0.5] = np.nan
Sf = api.flood_fill(side, Sn)
V2, F2, E2 = api.marching_cubes(Sf, GV, side[0], side[1], side[2], 0.01, return_E2V=True)
# iterative closest point (sampling, search, reduction)
mi = api.Mesh(V, F, 0)
Ricp, ticp = mi.iterative_closest_point(V + 0.01, F, 300, 2, seed=3)
mi.close()
# sharded device-resident path (begin / run / finish) on device buffers of the IPC allocator
import ctypes  # noqa: E402
import torch  # noqa: E402
from libigl_b200.dist import _arena_tensor  # noqa: E402
L = api.lib
nq = P.shape[0]
ms = api.Mesh(V, F, 2)
al = lambda x: (x + 255) & ~255
sb = al(int(L.iglb200_shard_scratch_bytes(nq)))
offs = [sb]
for b in (24, 8, 4, 24, 8):
    offs.append(offs[-1] + al(b * nq))
base = ctypes.c_void_p()
handle = ctypes.create_string_buffer(64)
api.check(L.iglb200_ipc_alloc(offs[-1], ctypes.byref(base), handle))
whole = _arena_tensor(api, base.value, offs[-1], torch.device("cuda", 0))
whole[offs[0]:offs[0] + 24 * nq].view(torch.float64).view(3, nq).copy_(torch.from_numpy(np.ascontiguousarray(P.T)))
vp = ctypes.c_void_p
ptr = [vp(base.value + o) for o in offs[:5]]
tiny, huge = float(np.finfo(float).tiny), float(np.finfo(float).max)
for mode in (0, 1, 2):
    api.check(L.iglb200_shard_begin_dev(ms.h, ptr[0], nq, base, 1, 8192, None))
    api.check(L.iglb200_shard_run_dev(ms.h, ptr[0], nq, base, mode, 4, tiny, huge, 2.0, 1, ptr[1], ptr[2], ptr[3], ptr[4], 8192, 0, 0, 1, None))
    api.check(L.iglb200_shard_finish_dev(ms.h, nq, base, mode, 4, tiny, huge, 1, ptr[1], ptr[2], ptr[3], ptr[4], None))
    torch.cuda.synchronize()
    api.check(L.iglb200_poll_device_error())
Ssh = whole[offs[1]:offs[1] + 8 * nq].view(torch.float64).cpu().numpy()
del whole
L.iglb200_ipc_free(base)
ms.close()
print("icp", float(np.abs(ticp).max()), "shard", float(np.nanmax(np.abs(Ssh))))
print("grid", side.tolist(), SV.shape, SF.shape, V2.shape, float(np.abs(Wx).max()), float(np.nanmax(Sx)))
print("ok", float(np.nanmax(np.abs(S))), int(I.max()), float(Wc.max()))
