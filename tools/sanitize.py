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
print("ok", float(np.nanmax(np.abs(S))), int(I.max()), float(Wc.max()))
