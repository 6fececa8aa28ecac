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
print("grid", side.tolist(), SV.shape, SF.shape, V2.shape, float(np.abs(Wx).max()), float(np.nanmax(Sx)))
print("ok", float(np.nanmax(np.abs(S))), int(I.max()), float(Wc.max()))
