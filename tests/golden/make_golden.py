"""Generates the golden fixtures in this directory from the GENUINE reference code.

Run in the build container (needs /root/reference to build oracle/_ref):
    python tests/golden/make_golden.py
The winding numbers, tree topology and per-node moments come from the unmodified
UT_SolidAngle<float,float> / UT_BVH<4> in /root/reference/include/igl/FastWindingNumberForSoups.h
(through oracle/ref_fwn.cpp); distances come from the restated f64 AABB cross-checked against
brute force; grid.npz holds flood fill / marching cubes of the restated sequential sweeps (case tables from the
genuine header) and the face-by-face exact winding number.  The fixtures travel to the GPU box,
/root/reference does not.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from libigl_b200 import synth  # noqa: E402


def main():
    assert os.path.isdir("/root/reference/include/igl"), "the reference tree is needed to regenerate fixtures"
    oracle.build(force=True)
    # 1. soup FWN: icosphere level 2 and a small torus knot
    for name, (V, F), res, scale in (("ico2", synth.icosphere(2), 14, 1.5), ("knot", synth.torus_knot(90, 10), 14, 1.1)):
        P = synth.grid_points(*synth.bbox(V), res, scale)
        ref = oracle.RefFwn(V, F, 2)
        W = ref.winding(P)
        W1 = ref.winding(P, 1.0)
        S, I, C, Wf = oracle.signed_distance_fwn(P, V, F, fwn=ref)
        Db, Ib, Cb = oracle.brute_closest(V, F, P)
        assert np.array_equal(Db, S * 0 + Db)  # finite
        np.savez_compressed(os.path.join(HERE, f"soup_{name}.npz"), V=V, F=F, P=P, W_beta2=W, W_beta1=W1, S=S, I=I, C=C,
                            W_sd=Wf, sqrD_brute=Db, I_brute=Ib, nodes=ref.nodes(), boxdata=ref.boxdata())
    # 2. bounded signed distance (NaN sentinels)
    V, F = synth.icosphere(2)
    P = synth.grid_points(*synth.bbox(V), 10, 2.0)
    S, I, C, W = oracle.signed_distance_fwn(P, V, F, -0.2, 0.3)
    np.savez_compressed(os.path.join(HERE, "soup_bounds.npz"), V=V, F=F, P=P, lower=-0.2, upper=0.3, S=S, I=I, C=C)
    # 3. point cloud
    V, F = synth.icosphere(3)
    Pc, N, A = synth.oriented_cloud(V, F, 400, 2)
    Q = synth.random_points(300, -1.5, 1.5, 4)
    out = dict(P=Pc, N=N, A=A, Q=Q)
    for order in (0, 1, 2):
        o = oracle.PortCloud(Pc, N, A, order)
        out[f"WN_order{order}"] = o.eval(Q, 2.0)
        out[f"ncells_order{order}"] = o.num_cells()
    out["WN_brute"] = oracle.PortCloud(Pc, N, A, 2).eval(Q, 0.0)
    np.savez_compressed(os.path.join(HERE, "cloud.npz"), **out)
    # 4. grids: flood fill and marching cubes of the restated sequential sweeps (case tables from the genuine header),
    #    exact winding number (face-by-face sum) on an open soup
    rng = np.random.default_rng(9)
    dims = (9, 7, 8)
    n = int(np.prod(dims))
    GV = oracle.grid(list(dims)) * np.array([1.0, 1.5, 0.8]) + 0.01 * rng.standard_normal((n, 3))
    S = rng.standard_normal(n)
    Sn = S.copy()
    Sn[rng.random(n) < 0.6] = np.nan
    Sn.reshape(dims[2], dims[1], dims[0])[2] = np.nan
    Vm, Fm, Em = oracle.marching_cubes(S, GV, dims[0], dims[1], dims[2], 0.1)
    V, F = synth.icosphere(2)
    Fo = F[V[F].mean(1)[:, 2] > -0.2]
    O = synth.random_points(200, -1.5, 1.5, 6)
    np.savez_compressed(os.path.join(HERE, "grid.npz"), dims=np.array(dims), GV=GV, S=S, S_nan=Sn, S_filled=oracle.flood_fill(dims, Sn),
                        iso=0.1, mc_V=Vm, mc_F=Fm, mc_E=Em, wn_V=V, wn_F=Fo, wn_O=O, wn_W=oracle.winding_number(V, Fo, O))
    print("fixtures written to", HERE)


if __name__ == "__main__":
    main()
