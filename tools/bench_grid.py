#!/usr/bin/env python
"""Times the device-resident grid pipeline on one GPU:
    igl::offset_surface = voxel_grid -> signed_distance(bounded) -> flood_fill -> marching_cubes
for the BASELINE c2 mesh (1 M-face torus knot), and the same stages of the CPU checker (oracle port,
sequential flood fill / marching cubes exactly like the reference) on a smaller grid for comparison.
Prints one JSON line per grid size."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import libigl_b200  # noqa: E402
from libigl_b200 import synth  # noqa: E402
from libigl_b200.api import Grid, Mesh, SIGNED_DISTANCE_TYPE_FAST_WINDING_NUMBER as FWN  # noqa: E402


def gpu_run(api, V, F, isolevel, s, reps):
    m = Mesh(api, V, F, 2)
    out = None
    api.profile_enable(True)
    api.profile_collect()
    wall = []
    for r in range(reps + 1):
        t0 = time.perf_counter()
        g = Grid(api, V.min(0) - isolevel, V.max(0) + isolevel, s, 1)
        c = g.corners()
        h = (c[3] - c[0]) / float(g.side[0] - 1)
        g.signed_distance(m, FWN, isolevel - np.sqrt(3.0) * h, isolevel + np.sqrt(3.0) * h)
        g.flood_fill()
        sf = g.marching_cubes(0.0, shift=isolevel, keep=True)
        nv, nf = sf.sizes()
        wall.append(time.perf_counter() - t0)
        if r == 0:
            api.profile_collect()      # drop the warm-up
        out = dict(side=g.side.tolist(), grid_vertices=g.n, nV=nv, nF=nf)
        sf.close(); g.close()
    prof = {k: v / reps for k, v in api.profile_collect().items()}
    api.profile_enable(False)
    m.close()
    out.update(stage_ms=prof, wall_ms=1e3 * float(np.mean(wall[1:])))
    return out


def cpu_run(V, F, isolevel, s):
    import oracle
    t = {}
    t0 = time.perf_counter(); GV, side = oracle.voxel_grid_mesh(V, isolevel, s, 1); t["voxel_grid"] = time.perf_counter() - t0
    h = (GV[:, 0].max() - GV[:, 0].min()) / float(side[0] - 1)
    tree = oracle.PortAABB(V, F)
    fw = oracle.RefFwn(V, F)
    t0 = time.perf_counter()
    S = oracle.signed_distance_fwn(GV, V, F, isolevel - np.sqrt(3.0) * h, isolevel + np.sqrt(3.0) * h, tree=tree, fwn=fw)[0]
    t["signed_distance"] = time.perf_counter() - t0
    t0 = time.perf_counter(); S = oracle.flood_fill(side, S); t["flood_fill"] = time.perf_counter() - t0
    t0 = time.perf_counter(); SV, SF, _ = oracle.marching_cubes(S - isolevel, GV, side[0], side[1], side[2], 0.0); t["marching_cubes"] = time.perf_counter() - t0
    return dict(side=side.tolist(), grid_vertices=int(GV.shape[0]), nV=int(SV.shape[0]), nF=int(SF.shape[0]),
                stage_ms={k: 1e3 * v for k, v in t.items()}, threads=oracle.num_threads())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--faces", type=int, default=1_000_000)
    ap.add_argument("--sizes", default="128,256,512")
    ap.add_argument("--cpu-size", type=int, default=128)
    ap.add_argument("--isolevel", type=float, default=0.02)
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    api = libigl_b200.lib()
    V, F = synth.torus_knot(*synth.knot_dims(a.faces))
    for s in [int(x) for x in a.sizes.split(",")]:
        r = gpu_run(api, V, F, a.isolevel, s, a.reps)
        r.update(what="offset_surface on the device", faces=int(F.shape[0]), s=s, isolevel=a.isolevel)
        print(json.dumps(r), flush=True)
    if a.cpu_size > 0:
        r = cpu_run(V, F, a.isolevel, a.cpu_size)
        r.update(what="CPU checker (oracle), same stages", faces=int(F.shape[0]), s=a.cpu_size, isolevel=a.isolevel)
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
