"""Pins the oracle: the reference's own known-answer tests, the committed golden fixtures
(generated from the genuine reference code, tests/golden/make_golden.py) and brute force."""
import os

import numpy as np

import oracle
from libigl_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_single_tet_all_queries():
    # tests/include/igl/signed_distance.cpp:4-43 (FWN tolerance 1e-7)
    V, F = synth.tetrahedron()
    S, I, C, W = oracle.signed_distance_fwn(np.array([[0.5, 0.5, 0.5]]), V, F)
    assert abs(S[0] - np.sqrt(1 / 12)) < 1e-7 and I[0] == 3
    assert abs(W[0] - 1.897e-8) < 1e-10          # SURVEY [probe] of the genuine core
    S, _, _, W = oracle.signed_distance_fwn(np.array([[0.1, 0.1, 0.1]]), V, F)
    assert abs(W[0] - 1.0000001) < 1e-6 and S[0] < 0


def test_one_point_cloud_golden():
    # tests/include/igl/fast_winding_number.cpp:17-53 (1e-15)
    P = np.array([[0.1, 0.2, 0.3]])
    N = np.array([[4.0, 5.0, 6.0]])
    N /= np.linalg.norm(N)
    Q = np.array([[0, 0, 0], [4, -3, 2], [-1, 1, 3], [0, 5, -2.0]])
    gold = np.array([0.38779369004261133, -0.00041235296362485, -0.00362978253577090, -0.00041235296362485])
    assert np.abs(oracle.PortCloud(P, N, np.array([0.7]), 2).eval(Q, 2) - gold).max() < 1e-15


def test_soup_barycentres():
    # tests/include/igl/AABB.cpp:100-113 on a synthetic soup
    V, F = synth.soupify(*synth.icosphere(3))
    BC = (V[F[:, 0]] + V[F[:, 1]] + V[F[:, 2]]) / 3
    t = oracle.PortAABB(V, F)
    D, I, C = t.squared_distance(BC)
    assert np.array_equal(I, np.arange(len(F))) and np.abs(D).max() < 1e-15
    assert t.num_nodes() == 2 * len(F) - 1


def exact_winding(V, F, Q):
    a, b, c = (V[F[:, k]][None] - Q[:, None] for k in range(3))
    la, lb, lc = (np.linalg.norm(x, axis=2) for x in (a, b, c))
    num = np.einsum("qfi,qfi->qf", a, np.cross(b, c))
    den = la * lb * lc + np.einsum("qfi,qfi->qf", a, b) * lc + np.einsum("qfi,qfi->qf", b, c) * la + np.einsum("qfi,qfi->qf", c, a) * lb
    return (2 * np.arctan2(num, den)).sum(1) / (4 * np.pi)


def test_fwn_against_exact_winding_number():
    # tests/include/igl/fast_winding_number.cpp:55-107 re-created on synthetic closed meshes (tol 1e-2)
    for V, F in (synth.icosphere(3), synth.torus_knot(150, 12)):
        Q = np.concatenate([V.mean(0, keepdims=True), synth.random_points(200, V.min(), V.max(), 9)])
        Wx = exact_winding(V, F, Q)
        assert np.abs(oracle.RefFwn(V, F).winding(Q) - Wx).max() < 1e-2
        # the cloud of the reference test: face barycentres, unit face normals, face areas
        a, b, c = V[F[:, 0]], V[F[:, 1]], V[F[:, 2]]
        n = np.cross(b - a, c - a)
        dbl = np.linalg.norm(n, axis=1)
        Wc = oracle.PortCloud((a + b + c) / 3, n / dbl[:, None], dbl / 2, 2).eval(Q[:1], 2.0)
        assert abs(Wc[0] - Wx[0]) < 1e-2


def test_restated_aabb_equals_brute_force():
    for V, F in (synth.icosphere(2), synth.torus_knot(60, 8), synth.noisy_sphere(2, 0.05, 1)):
        P = synth.random_points(800, -3, 3, 2)
        D, I, C = oracle.PortAABB(V, F).squared_distance(P)
        Db, Ib, Cb = oracle.brute_closest(V, F, P)
        assert np.array_equal(D, Db)
        mism = I != Ib
        assert np.array_equal(oracle.face_distance(V, F, P[mism], I[mism]), Db[mism])   # ties only


def test_oracle_reproduces_golden_fixtures():
    for name in ("ico2", "knot"):
        g = np.load(os.path.join(GOLD, f"soup_{name}.npz"))
        ref = oracle.RefFwn(g["V"], g["F"], 2)
        assert np.array_equal(ref.nodes(), g["nodes"])
        assert np.array_equal(ref.boxdata().view(np.uint32), g["boxdata"].view(np.uint32))
        assert np.array_equal(ref.winding(g["P"]), g["W_beta2"])
        assert np.array_equal(ref.winding(g["P"], 1.0), g["W_beta1"])
        S, I, C, W = oracle.signed_distance_fwn(g["P"], g["V"], g["F"], fwn=ref)
        assert np.array_equal(S, g["S"]) and np.array_equal(I, g["I"]) and np.array_equal(C, g["C"])
    g = np.load(os.path.join(GOLD, "soup_bounds.npz"))
    S, I, C, W = oracle.signed_distance_fwn(g["P"], g["V"], g["F"], float(g["lower"]), float(g["upper"]))
    assert np.array_equal(S, g["S"], equal_nan=True) and np.array_equal(I, g["I"]) and np.isnan(S).sum() > 0
    g = np.load(os.path.join(GOLD, "cloud.npz"))
    for order in (0, 1, 2):
        o = oracle.PortCloud(g["P"], g["N"], g["A"], order)
        assert o.num_cells() == int(g[f"ncells_order{order}"])
        assert np.abs(o.eval(g["Q"], 2.0) - g[f"WN_order{order}"]).max() < 1e-15
    g = np.load(os.path.join(GOLD, "grid.npz"))
    dims = [int(x) for x in g["dims"]]
    assert np.array_equal(oracle.flood_fill(dims, g["S_nan"]), g["S_filled"], equal_nan=True)
    V, F, E = oracle.marching_cubes(g["S"], g["GV"], dims[0], dims[1], dims[2], float(g["iso"]))
    assert np.array_equal(V, g["mc_V"]) and np.array_equal(F, g["mc_F"]) and np.array_equal(E, g["mc_E"])
    assert np.abs(oracle.winding_number(g["wn_V"], g["wn_F"], g["wn_O"]) - g["wn_W"]).max() < 1e-15


# ---- the checker's grid functions and the product's marching-cubes tables, pinned to the reference -------------
def test_grid_known_answer():
    """tests/include/igl/grid.cpp:5-41 ("grid: 3d") re-created against the restated igl::grid."""
    GV = oracle.grid([3, 3, 3])
    want = np.array([[x, y, z] for z in (0, .5, 1) for y in (0, .5, 1) for x in (0, .5, 1)], float)
    assert np.array_equal(GV, want)


def test_marching_cubes_tables_are_the_reference_s():
    """libigl_b200/csrc/mc_tables.h (packed, generated) and the corner/edge numbering in csrc/grid.h against
    the genuine header as march_cube.cpp:44 compiles it (exported by oracle/_ref)."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ef, conn, tri = oracle.mc_tables()
    txt = open(os.path.join(root, "libigl_b200", "csrc", "mc_tables.h")).read()
    words = [int(w, 16) for w in re.findall(r"0x([0-9A-Fa-f]{16})ULL", txt)]
    assert len(words) == 256
    for c in range(256):
        got = [(words[c] >> (4 * k)) & 0xF for k in range(16)]
        assert got == [0xF if e < 0 else int(e) for e in tri[c]], c
    g = open(os.path.join(root, "libigl_b200", "csrc", "grid.h")).read()

    def arr(name):
        m = re.search(name + r"\[\d+\] = \{([^}]*)\}", g)
        return [int(x) for x in m.group(1).split(",")]
    assert list(zip(arr("MC_EA"), arr("MC_EB"))) == [tuple(r) for r in conn.tolist()]
    # an edge is cut iff its corners lie on different sides: the header's edge-flag table is exactly that
    derived = [sum(1 << e for e, (a, b) in enumerate(conn) if ((c >> a) ^ (c >> b)) & 1) for c in range(256)]
    assert derived == ef.tolist()
    # corner offsets in the order of marching_cubes.cpp:52 (ioffset = {0,1,1+nx,nx,nx*ny,...})
    cx, cy, cz = arr("MC_CX"), arr("MC_CY"), arr("MC_CZ")
    nx, ny = 7, 5
    assert [x + nx * (y + ny * z) for x, y, z in zip(cx, cy, cz)] == [0, 1, 1 + nx, nx, nx * ny, 1 + nx * ny, 1 + nx + nx * ny, nx + nx * ny]


def test_restated_flood_fill_and_marching_cubes_properties():
    g = 20
    GV = oracle.grid([g, g, g]) * 2 - 1
    S = np.linalg.norm(GV, axis=1) - 0.6
    Sn = S.copy()
    Sn[np.abs(S) > 0.15] = np.nan
    Sf = oracle.flood_fill([g, g, g], Sn)
    assert not np.isnan(Sf).any() and np.array_equal(np.sign(Sf), np.sign(S))      # the fill keeps every sign
    V, F, E = oracle.marching_cubes(Sf, GV, g, g, g, 0.0)
    assert V.shape[0] - F.shape[0] * 3 // 2 + F.shape[0] == 2                        # closed, genus 0
    assert np.abs(np.linalg.norm(V, axis=1) - 0.6).max() < 0.02
    assert (E[:, 0] < E[:, 1]).all() and len({tuple(e) for e in E.tolist()}) == E.shape[0]   # one vertex per grid edge


def test_exact_winding_number_known_values():
    """winding_number.cpp + solid_angle.cpp restated: 1 inside / 0 outside a closed mesh, 1/8 for one octant."""
    V = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], float)
    F = np.array([[0, 1, 3], [0, 2, 1], [0, 3, 2], [1, 2, 3]], np.int32)
    W = oracle.winding_number(V, F, np.array([[0.1, 0.1, 0.1], [2, 2, 2], [-0.2, -0.2, -0.2]], float))
    assert np.abs(W - [1, 0, 0]).max() < 1e-15
    w = oracle.winding_number(V, F[3:], np.zeros((1, 3)))
    assert abs(abs(w[0]) - 0.125) < 1e-15
    S, I, C, Wd = oracle.signed_distance_winding_number(np.array([[0.5, 0.5, 0.5]]), V, F)
    assert abs(S[0] - np.sqrt(1. / 12.)) < 1e-15 and I[0] == 3        # tests/include/igl/signed_distance.cpp:4-43


# ---------------------------------------------------------------------------------------------------------------
# Pinning the CLOUD oracle beyond the dipole term.  The only vector the reference's tests hold for the point-cloud
# fast winding number (tests/include/igl/fast_winding_number.cpp:17-53) is a one-point cloud: a leaf, direct_eval only.
# The order-1/2 Taylor terms (fast_winding_number.cpp:94-110 precompute, :190-224 evaluation) are pinned here against
# mathematics instead: (a) the brute-force dipole sum, (b) finite differences of the dipole kernel, (c) the coefficient
# sums written as plain tensor contractions, (d) the evaluation as the Taylor polynomial of (b) contracted with (c).
def _dipole_K(x):
    """K(x) = x / (4 pi |x|^3): the dipole at p with strength m = a*n contributes K(p - q) . m  (:159-172)."""
    r = np.linalg.norm(x, axis=-1, keepdims=True)
    return x / (4.0 * np.pi * r ** 3)


def _dK(L):
    """dK[i, j] = d_i K_j at L (SecondDerivative, :190-197)."""
    r = np.linalg.norm(L)
    return (np.eye(3) / r ** 3 - 3.0 * np.outer(L, L) / r ** 5) / (4.0 * np.pi)


def _ddK(L):
    """ddK[k, i, j] = d_k d_i K_j at L (ThirdDerivative, :203-222)."""
    r = np.linalg.norm(L)
    e = np.eye(3)
    t = 15.0 * np.einsum("k,i,j->kij", L, L, L) / r ** 7
    t -= 3.0 * (np.einsum("ij,k->kij", e, L) + np.einsum("ik,j->kij", e, L) + np.einsum("jk,i->kij", e, L)) / r ** 5
    return t / (4.0 * np.pi)


def _cloud(n, seed=3):
    rng = np.random.default_rng(seed)
    P = rng.normal(size=(n, 3))
    P /= np.linalg.norm(P, axis=1)[:, None]
    P *= 1.0 + 0.05 * rng.normal(size=(n, 1))
    N = P / np.linalg.norm(P, axis=1)[:, None]
    A = np.full(n, 4.0 * np.pi / n) * (1.0 + 0.3 * rng.random(n))
    return P, N, A


def _members(cells, P):
    """Points of every octree cell, re-derived from the cell centres with the reference's octant rule (octree.cpp:43-51)."""
    CH, CN = cells["CH"], cells["CN"]
    mem = {0: np.arange(P.shape[0])}
    todo = [0]
    while todo:
        c = todo.pop()
        if CH[c, 0] == -1:
            continue
        idx = mem[c]
        octant = (P[idx, 0] >= CN[c, 0]) * 1 + (P[idx, 1] >= CN[c, 1]) * 2 + (P[idx, 2] >= CN[c, 2]) * 4
        for o in range(8):
            mem[int(CH[c, o])] = idx[octant == o]
            todo.append(int(CH[c, o]))
    return mem


def test_cloud_oracle_brute_force_dipole_sum():
    P, N, A = _cloud(1500)
    Q = synth.random_points(300, -2.0, 2.0, 9)
    M = A[:, None] * N
    brute = np.array([np.einsum("ij,ij->", _dipole_K(P - q), M) for q in Q])
    for order in (0, 1, 2):
        o = oracle.PortCloud(P, N, A, order)
        assert np.abs(o.eval(Q, 0.0) - brute).max() < 1e-12            # beta <= 0: the plain sum (:300-309)
        assert np.abs(o.eval(Q, 1e12) - brute).max() < 1e-12           # beta -> inf: every cell descends to its leaves
    # the expansion converges to it with the order
    errs = [np.abs(oracle.PortCloud(P, N, A, k).eval(Q, 2.0) - brute).max() for k in (0, 1, 2)]
    assert errs[2] < errs[1] < errs[0] and errs[2] < 0.1 and np.median(np.abs(oracle.PortCloud(P, N, A, 2).eval(Q, 2.0) - brute)) < 2e-3


def test_cloud_kernel_derivatives_by_finite_differences():
    rng = np.random.default_rng(1)
    for _ in range(5):
        L = rng.normal(size=3) * 2.0
        h = 1e-4
        e = np.eye(3)
        fd1 = np.array([(_dipole_K(L + h * e[i]) - _dipole_K(L - h * e[i])) / (2 * h) for i in range(3)])
        assert np.abs(fd1 - _dK(L)).max() < 1e-7 * np.abs(_dK(L)).max() + 1e-9
        fd2 = np.array([(_dK(L + h * e[k]) - _dK(L - h * e[k])) / (2 * h) for k in range(3)])
        assert np.abs(fd2 - _ddK(L)).max() < 1e-7 * np.abs(_ddK(L)).max() + 1e-9


def test_cloud_oracle_coefficients_and_expansion_are_the_taylor_series():
    P, N, A = _cloud(1200, seed=5)
    o = oracle.PortCloud(P, N, A, 2)
    cells = o.cells()
    mem = _members(cells, P)
    assert all(len(mem[c]) == cells["counts"][c] for c in mem)
    internal = [c for c in mem if cells["CH"][c, 0] != -1 and len(mem[c]) > 1]
    assert len(internal) > 50
    for c in internal[:200]:
        idx = mem[c]
        a, p, n = A[idx], P[idx], N[idx]
        cm = (a[:, None] * p).sum(0) / a.sum()
        d = p - cm
        assert np.abs(cells["CM"][c] - cm).max() < 1e-13
        assert abs(cells["R"][c] - np.linalg.norm(d, axis=1).max()) < 1e-13
        ec = cells["EC"][c]
        assert np.abs(ec[:3] - (a[:, None] * n).sum(0)).max() < 1e-13
        T1 = np.einsum("p,pi,pj->ij", a, d, n)                      # sum a * delta (x) n
        assert np.abs(ec[3:12].reshape(3, 3).T - T1).max() < 1e-13    # Eigen::Map of a column-major 3x3 (:98-101)
        T2 = 0.5 * np.einsum("p,pk,pi,pj->kij", a, d, d, n)
        for k in range(3):
            assert np.abs(ec[12 + 9 * k:21 + 9 * k].reshape(3, 3).T - T2[k]).max() < 1e-13
    # evaluation: a query far from every child of the root takes expansion_eval for each of them (:262-282)
    q = np.array([[30.0, -20.0, 25.0]])
    kids = [int(k) for k in cells["CH"][0] if cells["counts"][int(k)] > 0]
    assert all(cells["CH"][k, 0] != -1 for k in kids)
    assert all(np.linalg.norm(cells["CM"][k] - q[0]) > 2.0 * cells["R"][k] for k in kids)
    for order, nterms in ((0, 3), (1, 12), (2, 39)):
        oo = oracle.PortCloud(P, N, A, order)
        cc = oo.cells()
        want = 0.0
        for k in kids:
            L = cc["CM"][k] - q[0]
            ec = cc["EC"][k]
            w = _dipole_K(L) @ ec[:3]
            if order >= 1:
                w += np.einsum("ij,ij->", _dK(L), ec[3:12].reshape(3, 3).T)
            if order >= 2:
                w += sum(np.einsum("ij,ij->", _ddK(L)[kk], ec[12 + 9 * kk:21 + 9 * kk].reshape(3, 3).T) for kk in range(3))
            want += w
        got = oo.eval(q, 2.0)[0]
        assert abs(got - want) < 1e-15 + 1e-12 * abs(want), (order, got, want)
    # ... and the truncated series approaches the exact sum of the cell like (R/r)^(order+1)
    exact = np.einsum("ij,ij->", _dipole_K(P - q[0]), A[:, None] * N)
    errs = [abs(oracle.PortCloud(P, N, A, k).eval(q, 2.0)[0] - exact) for k in (0, 1, 2)]
    assert errs[2] < 0.2 * errs[1] < 0.04 * errs[0] or errs[2] < 1e-12
