#!/usr/bin/env python
"""Times the exact hierarchical winding number (sign types WINDING_NUMBER / DEFAULT, igl::winding_number) on one GPU:
build of the boundary chains with the tree, iglb200_winding_number and signed_distance(DEFAULT) on a grid."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libigl_b200 as ig
from libigl_b200 import synth
api = ig.lib()
api.profile_enable(True)
for nf, g in ((100_000, 96), (1_000_000, 128)):
    V, F = synth.torus_knot(*synth.knot_dims(nf))
    t0 = time.perf_counter(); m = api.Mesh(V, F, 2, winding_number=True); tb = time.perf_counter() - t0
    info = m.info()
    P = synth.grid_points(*synth.bbox(V), g, 1.1)
    m.winding_number(P[:1000])
    api.profile_collect()
    t0 = time.perf_counter(); W = m.winding_number(P); t1 = time.perf_counter() - t0
    prof = api.profile_collect()
    S, I, C = m.signed_distance(P, 2)
    prof2 = api.profile_collect()
    print(nf, g, "build_s", round(tb, 3), "arena_MB", info["device_bytes"] / 1e6, "wn wall_s", round(t1, 3), prof, "sd default", prof2,
          "q/s (wall clock, host arrays in and out)", P.shape[0] / t1,
          "q/s (kernel; spans are per thread, so only calls small enough to run on the calling thread report one)",
          (P.shape[0] / (prof["exact_winding"] / 1e3)) if "exact_winding" in prof else None, "Wrange", W.min(), W.max(), flush=True)
    m.close()
