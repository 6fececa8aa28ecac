"""Parity cases run twice: on the host-emulated kernels (CPU, small) and on the
real CUDA library (`-m gpu`, up to full size)."""
import numpy as np
import pytest

import oracle
import parity
from libigl_b200 import synth
from libigl_b200.api import (IglB200Error, SIGNED_DISTANCE_TYPE_FAST_WINDING_NUMBER, SIGNED_DISTANCE_TYPE_PSEUDONORMAL,
                             SIGNED_DISTANCE_TYPE_UNSIGNED, AABB)


def meshes_small():
    return {
        "tet": synth.tetrahedron(),
        "ico1": synth.icosphere(1),
        "ico3": synth.icosphere(3),
        "knot": synth.torus_knot(120, 12),
        "noisy": synth.noisy_sphere(3, 0.02, 7),
    }


def axis_aligned_box_mesh(n=6):
    """Axis-aligned, highly symmetric grid mesh: many exactly equal box centres (ties,
    elements equal to the pivot) — the hard case for a topology-exact build."""
    xs = np.linspace(-1, 1, n + 1)
    Vs, Fs = [], []

    def face(ax, val, flip):
        u, v = [a for a in range(3) if a != ax]
        base = sum(len(x) for x in Vs)
        g = np.zeros(((n + 1) * (n + 1), 3))
        U, W = np.meshgrid(xs, xs, indexing="ij")
        g[:, u] = U.ravel()
        g[:, v] = W.ravel()
        g[:, ax] = val
        Vs.append(g)
        for i in range(n):
            for j in range(n):
                a = base + i * (n + 1) + j
                b = a + 1
                c = a + (n + 1)
                d = c + 1
                t = [[a, c, d], [a, d, b]]
                if flip:
                    t = [[x[0], x[2], x[1]] for x in t]
                Fs.extend(t)

    for ax in range(3):
        face(ax, -1.0, ax != 1)
        face(ax, 1.0, ax == 1)
    return np.concatenate(Vs), np.array(Fs, np.int32)


def case_known_answers(api):
    # tests/include/igl/signed_distance.cpp:4-43
    V, F = synth.tetrahedron()
    S, I, C = api.signed_distance(np.array([[0.5, 0.5, 0.5]]), V, F, SIGNED_DISTANCE_TYPE_FAST_WINDING_NUMBER)
    assert abs(S[0] - np.sqrt(1.0 / 12.0)) < 1e-7
    S, I, C = api.signed_distance(np.array([[0.5, 0.5, 0.5]]), V, F, SIGNED_DISTANCE_TYPE_UNSIGNED)
    assert abs(S[0] - np.sqrt(1.0 / 12.0)) < 1e-15
    assert I[0] == 3 and np.allclose(C[0], 1.0 / 3.0, atol=1e-15)
    # tests/include/igl/AABB.cpp:100-113 on a synthetic soup
    V, F = synth.soupify(*synth.icosphere(2))
    BC = (V[F[:, 0]] + V[F[:, 1]] + V[F[:, 2]]) / 3.0
    tree = AABB(api)
    tree.init(V, F)
    D, I, C = tree.squared_distance(V, F, BC)
    assert np.array_equal(I, np.arange(F.shape[0]))
    assert np.abs(D).max() < 1e-15
    D2, I2, C2 = api.point_mesh_squared_distance(BC, V, F)
    assert np.array_equal(I2, I) and np.array_equal(D2, D)


def case_tree_exact(api, meshes):
    for name, (V, F) in meshes.items():
        m = api.Mesh(V, F, 2)
        parity.check_tree_against_reference(m, V, F, what=name)
        m.close()


def case_signed_distance(api, V, F, P, what):
    m = api.Mesh(V, F, 2)
    S, I, C, W = m.signed_distance(P, return_W=True)
    r = parity.check_signed_distance(V, F, P, S, I, C, W, what=what)
    # S-only overload (signed_distance.h:283-312)
    assert np.array_equal(m.signed_distance_fast_winding_number(P), S, equal_nan=True)
    m.close()
    return r


def case_closest_vs_brute(api, V, F, P, what):
    m = api.Mesh(V, F, 0)
    D, I, C = m.squared_distance(P)
    Db, Ib, Cb = oracle.brute_closest(V, F, P)
    assert np.array_equal(D, Db), what
    assert np.array_equal(I, Ib), what   # lowest-index tie rule == brute force order
    assert np.array_equal(C, Cb), what
    Do, Io, Co = oracle.PortAABB(V, F).squared_distance(P)
    parity.check_closest(V, F, P, D, I, C, Do, Io, Co, what)
    m.close()


def case_closest_extreme_scales(api):
    """The fp32 stage of the search works in coordinates scaled by a power of two, so the result must equal the all-fp64
    brute force whatever the units of the mesh: 3e19 (squares overflow fp32), 1e-20 (squares are denormal), queries far
    outside the mesh, and a mesh far from the origin."""
    V, F = synth.icosphere(2)
    P = synth.random_points(600, -2.0, 2.0, 5)
    for scale in (3e19, 1e-20, 7.0e-3, 1.0):
        m = api.Mesh(V * scale, F, 0)
        for Q in (P * scale, P * scale * 1e6, (P + 40.0) * scale):
            D, I, C = m.squared_distance(Q)
            Db, Ib, Cb = oracle.brute_closest(V * scale, F, Q)
            assert np.array_equal(D, Db) and np.array_equal(I, Ib) and np.array_equal(C, Cb), scale
        m.close()
    # astronomically far queries (the fp32 distances are inf/NaN): still the exact answer
    m = api.Mesh(V, F, 0)
    Q = P * 1e25
    D, I, C = m.squared_distance(Q)
    Db, Ib, Cb = oracle.brute_closest(V, F, Q)
    assert np.array_equal(D, Db) and np.array_equal(I, Ib) and np.array_equal(C, Cb)
    m.close()
    # a mesh far from the origin (offset 1e4 x its size)
    Vo = V + np.array([1.0e4, -2.0e4, 3.0e4])
    m = api.Mesh(Vo, F, 0)
    Q = P + np.array([1.0e4, -2.0e4, 3.0e4])
    D, I, C = m.squared_distance(Q)
    Db, Ib, Cb = oracle.brute_closest(Vo, F, Q)
    assert np.array_equal(D, Db) and np.array_equal(I, Ib) and np.array_equal(C, Cb)
    m.close()


def _icp_numpy(VX, FX, VY, FY, num_samples, max_iters, seed):
    """igl::iterative_closest_point restated (iterative_closest_point.cpp:40-130, random_points_on_mesh_intrinsic.cpp:21-45,
    rigid_alignment.cpp:24-108) with the device's sample stream (csrc/icp.h: u01) and the oracle's closest points."""
    M = (1 << 64) - 1

    def mix(z):
        z = (z + 0x9E3779B97F4A7C15) & M
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        return z ^ (z >> 31)

    def u01(it, k, j):
        h = mix(mix(mix(seed) ^ ((it * 0x632BE59BD9B4E019) & M)) ^ ((k * 3 + j) & M))
        return (h >> 11) / 9007199254740992.0

    tree = oracle.PortAABB(VY, FY)
    NY = np.cross(VY[FY[:, 1]] - VY[FY[:, 0]], VY[FY[:, 2]] - VY[FY[:, 0]])
    NY /= np.linalg.norm(NY, axis=1)[:, None]
    R, t = np.eye(3), np.zeros(3)
    for it in range(max_iters):
        V = VX @ R + t
        r_, s_ = V[FX[:, 0]] - V[FX[:, 2]], V[FX[:, 1]] - V[FX[:, 2]]
        dblA = np.sqrt(sum((r_[:, d] * s_[:, (d + 1) % 3] - r_[:, (d + 1) % 3] * s_[:, d]) ** 2 for d in range(3)))
        C = np.cumsum(dblA)
        Cn = C / C[-1]
        X = np.empty((num_samples, 3))
        for k in range(num_samples):
            r, s, tt = u01(it, k, 0), u01(it, k, 1), u01(it, k, 2)
            f = min(int(np.searchsorted(Cn, r, side="right")), FX.shape[0] - 1)
            st = np.sqrt(tt)
            X[k] = (1 - st) * V[FX[f, 0]] + (1 - s) * st * V[FX[f, 1]] + s * st * V[FX[f, 2]]
        _, I, P = tree.squared_distance(X)
        N = NY[I]
        Rr, tr, w = np.eye(3), np.zeros(3), 1e-5
        for _ in range(5):
            Xc = X @ Rr + tr
            NA = np.hstack([np.cross(Xc, N), N])
            NB = np.einsum("ij,ij->i", N, P - Xc)
            Z, O = np.zeros(len(Xc)), np.ones(len(Xc))
            A = np.vstack([np.column_stack([Z, Xc[:, 2], -Xc[:, 1], O, Z, Z]), np.column_stack([-Xc[:, 2], Z, Xc[:, 0], Z, O, Z]),
                           np.column_stack([Xc[:, 1], -Xc[:, 0], Z, Z, Z, O])])
            B = np.concatenate([P[:, 0] - Xc[:, 0], P[:, 1] - Xc[:, 1], P[:, 2] - Xc[:, 2]])
            Q = (1 - w) * NA.T @ NA + w * A.T @ A
            f = (1 - w) * NA.T @ NB + w * A.T @ B
            if np.linalg.matrix_rank(Q, tol=np.abs(np.linalg.eigvalsh(Q)).max() * 6 * np.finfo(float).eps) < 6:
                w = 1 - (1 - w) / 2
            u = np.linalg.pinv(Q, rcond=6 * np.finfo(float).eps, hermitian=True) @ f
            W = np.array([[0, u[2], -u[1]], [-u[2], 0, u[0]], [u[1], -u[0], 0]])
            x = np.linalg.norm(u[:3])
            Ri = np.eye(3) if x == 0 else np.eye(3) + np.sin(x) / x * W + (1 - np.cos(x)) / (x * x) * W @ W
            Rr, tr = Rr @ Ri, tr @ Ri + u[3:]
        R, t = R @ Rr, t @ Rr + tr
    return R, t


def case_icp(api):
    """iterative_closest_point: the device run equals the numpy restatement on the same sample stream, and it registers a
    rotated, shifted copy of a mesh back onto the original."""
    VY, FY = synth.torus_knot(160, 14)
    ax = np.array([0.3, -0.5, 0.8]); ax /= np.linalg.norm(ax)
    ang = 0.12
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    Rt = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
    VX = VY @ Rt + np.array([0.05, -0.03, 0.04])
    my = api.Mesh(VY, FY, 0)
    R, t = my.iterative_closest_point(VX, FY, 400, 6, seed=12345)
    Rn, tn = _icp_numpy(VX, FY, VY, FY, 400, 6, 12345)
    assert np.abs(R - Rn).max() < 1e-8 and np.abs(t - tn).max() < 1e-8, (np.abs(R - Rn).max(), np.abs(t - tn).max())
    assert np.abs(R @ R.T - np.eye(3)).max() < 1e-12 and np.linalg.det(R) > 0.999
    d0 = np.sqrt(my.squared_distance(VX)[0]).mean()
    d1 = np.sqrt(my.squared_distance(VX @ R + t)[0]).mean()
    assert d1 < 0.05 * d0, (d0, d1)
    # zero iterations: identity
    R0, t0 = my.iterative_closest_point(VX, FY, 10, 0)
    assert np.array_equal(R0, np.eye(3)) and np.array_equal(t0, np.zeros(3))
    my.close()


def case_fwn_orders(api, V, F, Q):
    """Taylor orders 0 and 1 (fast_winding_number(V,F,order,bvh), order < 2) against the genuine core."""
    for order in (0, 1):
        m = api.Mesh(V, F, order)
        ref = oracle.RefFwn(V, F, order)
        a = m.download()
        pr = parity.match_nodes(a["fwn_child"], ref.nodes())
        assert pr is not None
        W, Wr = m.fast_winding_number(Q, 2.0), ref.winding(Q, 2.0)
        assert np.abs(W - Wr).max() <= 1e-5, f"order {order}: {np.abs(W - Wr).max():.3e}"
        m.close()


def case_fwn(api, V, F, Q, what):
    case_fwn_orders(api, V, F, Q[:2000])
    m = api.Mesh(V, F, 2)
    ref = oracle.RefFwn(V, F)
    for beta in (2.0, 1.0, 3.5):
        W = m.fast_winding_number(Q, beta)
        Wr = ref.winding(Q, beta)
        d = np.abs(W - Wr)
        assert (d <= 1e-5 * np.maximum(1.0, np.abs(Wr))).all(), f"{what} beta={beta}: W off by {d.max():.3e}"
    W1 = api.fast_winding_number(V, F, Q)
    assert np.array_equal(W1, m.fast_winding_number(Q, 2.0))
    m.close()


def case_bounds_and_sentinels(api):
    V, F = synth.icosphere(2)
    P = synth.grid_points(*synth.bbox(V), 12, 2.0)
    m = api.Mesh(V, F, 2)
    for lower, upper in ((-0.2, 0.3), (0.1, 0.5), (-0.5, -0.1), (0.0, 0.25)):
        S, I, C, W = m.signed_distance(P, SIGNED_DISTANCE_TYPE_FAST_WINDING_NUMBER, lower, upper, True)
        r = parity.check_signed_distance(V, F, P, S, I, C, W, lower, upper, what=f"bounds {lower},{upper}")
        assert r["nan"] > 0
        assert (I[np.isnan(S)] == F.shape[0] + 1).all() and (C[np.isnan(S)] == 0).all()
    m.close()
    # #F == 0 (signed_distance.cpp:398-405)
    S, I, C = api.signed_distance(P[:5], V, np.zeros((0, 3), np.int32))
    assert np.isnan(S).all() and (I == -1).all() and np.isnan(C).all()
    # empty query set
    m = api.Mesh(V, F, 2)
    S, I, C = m.signed_distance(np.zeros((0, 3)))
    assert S.shape == (0,) and I.shape == (0,) and C.shape == (0, 3)
    # sign types whose tables the mesh was built without fail loudly (no CPU fallback); unknown types too
    for t in (1, 2):   # SIGNED_DISTANCE_TYPE_WINDING_NUMBER, _DEFAULT
        with pytest.raises(IglB200Error) as e:
            m.signed_distance(P[:5], t)
        assert e.value.code == -1
    with pytest.raises(IglB200Error) as e:
        m.signed_distance(P[:5], 7)
    assert e.value.code == -4
    with pytest.raises(IglB200Error) as e:   # pseudonormals need the precomputed normals
        m.signed_distance(P[:5], SIGNED_DISTANCE_TYPE_PSEUDONORMAL)
    assert e.value.code == -1
    m.close()
    with pytest.raises(IglB200Error):
        api.Mesh(V, np.array([[0, 1, 999999]], np.int32))


def case_layouts(api):
    """float32 / float64, row- and column-major, strided views give the same answers."""
    V, F = synth.icosphere(2)
    P = synth.random_points(500, -1.5, 1.5, 3)
    base = api.Mesh(V, F, 2).signed_distance(P, return_W=True)
    variants = [
        (np.asfortranarray(V), np.asfortranarray(F), np.asfortranarray(P)),
        (V, F, np.ascontiguousarray(np.concatenate([P, P], 1))[:, 3:]),
        (np.ascontiguousarray(np.concatenate([V, V], 1))[:, :3], F, P.astype(np.float32)),
    ]
    for Vv, Fv, Pv in variants:
        out = api.Mesh(Vv, Fv, 2).signed_distance(Pv, return_W=True)
        for a, b in zip(base, out):
            assert np.array_equal(a, b, equal_nan=True)
    # float32 vertices: the mesh IS the float mesh, compare against the oracle on that mesh
    Vf = V.astype(np.float32)
    S, I, C, W = api.Mesh(Vf, F, 2).signed_distance(P, return_W=True)
    parity.check_signed_distance(Vf.astype(np.float64), F, P, S, I, C, W, what="f32 V")


def case_degenerate(api):
    """Degenerate faces [i j j]/[i i i] behave as segments/points (point_mesh_squared_distance.h:23-30);
    NaN vertices are dropped from the winding tree (FastWindingNumberForSoups.h:4606-4615)."""
    V, F = synth.icosphere(1)
    F = F.copy()
    F[3] = [F[3, 0], F[3, 1], F[3, 1]]
    F[7] = [F[7, 0], F[7, 0], F[7, 0]]
    P = synth.random_points(400, -1.5, 1.5, 11)
    case_closest_vs_brute(api, V, F, P, "degenerate")
    m = api.Mesh(V, F, 2)
    parity.check_tree_against_reference(m, V, F, what="degenerate")
    W = m.fast_winding_number(P)
    Wr = oracle.RefFwn(V, F).winding(P)
    assert np.abs(W - Wr).max() <= 1e-5
    m.close()
    # duplicated faces / coincident boxes: exercises the "equal to pivot" partition path
    V2, F2 = axis_aligned_box_mesh(6)
    F2 = np.concatenate([F2, F2[::3]], 0)
    m = api.Mesh(V2, F2, 2)
    parity.check_tree_against_reference(m, V2, F2, what="axis-aligned + duplicates")
    m.close()


def case_arena_roundtrip(api):
    V, F = synth.icosphere(3)
    P = synth.random_points(2000, -1.5, 1.5, 5)
    m = api.Mesh(V, F, 2)
    ref = m.signed_distance(P, return_W=True)
    ptr, nbytes = m.arena()
    from libigl_b200.api import Mesh
    for copy in (True, False):
        m2 = Mesh.from_arena(api, ptr, nbytes, copy)
        out = m2.signed_distance(P, return_W=True)
        for a, b in zip(ref, out):
            assert np.array_equal(a, b, equal_nan=True)
        assert m2.info()["n_nodes"] == m.info()["n_nodes"]
        m2.close()
    m.close()


def adversarial_soup(rng, n, kind):
    c = rng.uniform(-1, 1, (n, 3))
    if kind == "normal":
        a, b, d = (c + rng.normal(0, 0.05, (n, 3)) for _ in range(3))
    elif kind == "sliver":
        e = rng.normal(0, 0.1, (n, 3))
        t = rng.uniform(0.2, 0.8, (n, 1))
        a, b, d = c, c + e, c + t * e + rng.normal(0, 1, (n, 3)) * (10.0 ** rng.uniform(-9, -2, (n, 1)))
    elif kind == "needle":
        a, b, d = c, c + rng.normal(0, 0.3, (n, 3)), c + rng.normal(0, 1, (n, 3)) * (10.0 ** rng.uniform(-8, -3, (n, 1)))
    elif kind == "tiny":
        s = 10.0 ** rng.uniform(-7, -3, (n, 1))
        a, b, d = (c + rng.normal(0, 1, (n, 3)) * s for _ in range(3))
    elif kind == "offset":
        off = np.array([1000.0, -2000.0, 500.0])
        a = c * 0.1 + off + rng.normal(0, 0.01, (n, 3))
        b, d = a + rng.normal(0, 0.01, (n, 3)), a + rng.normal(0, 0.01, (n, 3))
    else:  # degenerate: segments and points
        a, b = c, c + rng.normal(0, 0.1, (n, 3))
        d = b.copy()
        d[::2] = a[::2]
    V = np.concatenate([a, b, d])
    F = np.stack([np.arange(n), np.arange(n) + n, np.arange(n) + 2 * n], 1).astype(np.int32)
    return V, F


def case_adversarial_triangles(api, n_tri=300, reps=2):
    """The fp32 stage of the closest-triangle search must never lose the fp64 winner:
    slivers, needles, tiny faces, large coordinate offsets, degenerate faces; queries on,
    near and far from the surface.  Result must equal the fp64 brute force bit for bit."""
    rng = np.random.default_rng(123)
    for kind in ("normal", "sliver", "needle", "tiny", "offset", "degenerate"):
        for _ in range(reps):
            V, F = adversarial_soup(rng, n_tri, kind)
            lo, hi = V.min(0), V.max(0)
            ext = (hi - lo).max()
            near = V[rng.integers(0, len(V), 1200)] + rng.normal(0, 1, (1200, 3)) * (10.0 ** rng.uniform(-9, -1, (1200, 1))) * ext
            P = np.concatenate([rng.uniform(lo - 0.5 * ext, hi + 0.5 * ext, (1200, 3)), near, V[:200],
                                (lo + hi) / 2 + rng.normal(0, 10 * ext, (150, 3))])
            m = api.Mesh(V, F, 0)
            D, I, C = m.squared_distance(P)
            Db, Ib, Cb = oracle.brute_closest(V, F, P)
            assert np.array_equal(D, Db), kind
            assert np.array_equal(I, Ib), kind
            assert np.array_equal(C, Cb), kind
            m.close()


def case_cloud_golden(api):
    """tests/include/igl/fast_winding_number.cpp:17-53 (1e-15)."""
    P = np.array([[0.1, 0.2, 0.3]])
    N = np.array([[4.0, 5.0, 6.0]])
    N /= np.linalg.norm(N)
    A = np.array([0.7])
    Q = np.array([[0, 0, 0], [4, -3, 2], [-1, 1, 3], [0, 5, -2.0]])
    gold = np.array([0.38779369004261133, -0.00041235296362485, -0.00362978253577090, -0.00041235296362485])
    W = api.fast_winding_number(P, N, A, Q, 2, 2.0)
    assert np.abs(W - gold).max() < 1e-15
    assert np.abs(api.fast_winding_number(P, N, A, Q) - gold).max() < 1e-15   # default order 2, beta 2


def case_cloud(api, n_points, n_queries, orders=(0, 1, 2), betas=(2.0, 1.0, 0.0)):
    """Point-cloud FWN against the restated octree + expansion oracle: same cells, W to 1e-5
    (observed ~1e-14: both sides are double precision and differ only in summation order)."""
    V, F = synth.icosphere(4)
    P, N, A = synth.oriented_cloud(V, F, n_points, 2)
    Q = synth.random_points(n_queries, -1.5, 1.5, 4)
    for order in orders:
        c = api.Cloud(P, N, A, order)
        o = oracle.PortCloud(P, N, A, order)
        assert c.info()["n_cells"] == o.num_cells()
        for beta in betas:
            W = c.eval(Q, beta)
            Wo = o.eval(Q, beta)
            d = np.abs(W - Wo)
            assert (d <= 1e-5 * np.maximum(1.0, np.abs(Wo))).all(), f"order {order} beta {beta}: {d.max():.3e}"
            assert d.max() < 1e-10
        c.close()
    # the expansion is an approximation of the exact dipole sum: agree to ~1e-2 inside/outside
    c = api.Cloud(P, N, A, 2)
    W = c.eval(Q, 2.0)
    Wb = c.eval(Q[:300], 0.0)
    assert np.abs(W[:300] - Wb).max() < 2e-2
    r = np.linalg.norm(Q, axis=1)
    assert (W[r < 0.8] > 0.9).all() and (np.abs(W[r > 1.3]) < 0.1).all()
    # float32 inputs and empty query set
    assert c.eval(np.zeros((0, 3))).shape == (0,)
    assert np.abs(c.eval(Q.astype(np.float32)) - W).max() < 1e-12
    c.close()
    # duplicate points: the depth cap (42 levels) keeps the build finite
    P2 = np.concatenate([P[:50], P[:50]])
    c2 = api.Cloud(P2, np.concatenate([N[:50], N[:50]]), np.concatenate([A[:50], A[:50]]), 2)
    W2 = c2.eval(Q[:100], 2.0)
    assert np.isfinite(W2).all()
    c2.close()


def case_golden_fixtures(api):
    """Committed outputs of the genuine reference code (tests/golden/make_golden.py)."""
    import os
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    for name in ("ico2", "knot"):
        g = np.load(os.path.join(gold, f"soup_{name}.npz"))
        V, F, P = g["V"], g["F"], g["P"]
        m = api.Mesh(V, F, 2)
        a = m.download()
        pr = parity.match_nodes(a["fwn_child"], g["nodes"])
        assert pr is not None and len(pr) == g["nodes"].shape[0], "topology differs from the golden reference tree"
        da, db = a["fwn_data"][pr[:, 0]], g["boxdata"][pr[:, 1]]
        assert np.array_equal(da[:, :4].view(np.uint32), db[:, :4].view(np.uint32))
        fin = np.isfinite(db)
        assert np.abs(da[fin] - db[fin]).max() <= 1e-6 * np.abs(db[fin]).max()
        for beta, key in ((2.0, "W_beta2"), (1.0, "W_beta1")):
            assert np.abs(m.fast_winding_number(P, beta) - g[key]).max() <= 1e-5
        S, I, C, W = m.signed_distance(P, return_W=True)
        assert np.abs(S - g["S"]).max() <= 1e-5 * np.abs(g["S"]).max()
        assert np.abs(W - g["W_sd"]).max() <= 1e-5
        D, Ib, Cb = m.squared_distance(P)
        assert np.array_equal(D, g["sqrD_brute"]) and np.array_equal(Ib, g["I_brute"])
        m.close()
    g = np.load(os.path.join(gold, "soup_bounds.npz"))
    S, I, C = api.Mesh(g["V"], g["F"], 2).signed_distance(g["P"], SIGNED_DISTANCE_TYPE_FAST_WINDING_NUMBER, float(g["lower"]), float(g["upper"]))
    assert np.array_equal(np.isnan(S), np.isnan(g["S"]))
    ok = ~np.isnan(S)
    assert np.abs(S[ok] - g["S"][ok]).max() <= 1e-5 and np.array_equal(I[~ok], g["I"][~ok]) and np.array_equal(C[~ok], g["C"][~ok])
    g = np.load(os.path.join(gold, "cloud.npz"))
    for order in (0, 1, 2):
        c = api.Cloud(g["P"], g["N"], g["A"], order)
        assert c.info()["n_cells"] == int(g[f"ncells_order{order}"])
        assert np.abs(c.eval(g["Q"], 2.0) - g[f"WN_order{order}"]).max() < 1e-12
        c.close()
    g = np.load(os.path.join(gold, "grid.npz"))
    dims = [int(x) for x in g["dims"]]
    assert np.array_equal(api.flood_fill(dims, g["S_nan"]), g["S_filled"], equal_nan=True)
    Vm, Fm, Em = api.marching_cubes(g["S"], g["GV"], dims[0], dims[1], dims[2], float(g["iso"]), return_E2V=True)
    assert np.array_equal(Vm, g["mc_V"]) and np.array_equal(Fm, g["mc_F"]) and np.array_equal(Em, g["mc_E"])
    assert np.abs(api.winding_number(g["wn_V"], g["wn_F"], g["wn_O"]) - g["wn_W"]).max() < 1e-12


def _soup_at(rng, centers, size):
    n = len(centers)
    a, b, c = (centers + rng.normal(0, size, (n, 3)) for _ in range(3))
    return np.concatenate([a, b, c]), np.stack([np.arange(n), np.arange(n) + n, np.arange(n) + 2 * n], 1).astype(np.int32)


def case_builder_corner_paths(api):
    """Builder paths natural meshes rarely reach, each against the genuine reference tree:
    no admissible plane (a cluster plus far outliers: all boxes in one of the 16 spans) -> the
    quick-select fallback (FastWindingNumberForSoups.h:5708-5754); centre-sorted and reverse-sorted
    inputs (long chains in the closed-form partition); NaN/Inf vertices (boxes excluded, :4606-4615)."""
    rng = np.random.default_rng(3)
    for ncl, nout in ((500, 1), (2000, 3), (300, 40)):
        cen = np.concatenate([rng.normal(0, 0.01, (ncl, 3)), rng.uniform(-100, 100, (nout, 3))])
        V, F = _soup_at(rng, cen[rng.permutation(len(cen))], 0.002)
        m = api.Mesh(V, F, 2)
        parity.check_tree_against_reference(m, V, F, what=f"cluster {ncl} + {nout} outliers")
        m.close()
    for order in (1, -1):
        x = np.linspace(0, 1, 3000)[::order]
        V, F = _soup_at(rng, np.stack([x, 0.01 * np.sin(40 * x), 0 * x], 1), 0.0005)
        m = api.Mesh(V, F, 2)
        parity.check_tree_against_reference(m, V, F, what="sorted centres")
        m.close()
    V, F = synth.icosphere(2)
    Vn = np.concatenate([V, [[np.nan, 0, 0], [np.inf, 1, 1], [0, 0, 0.5]]])
    Fn = np.concatenate([F, [[len(V), 0, 1], [len(V) + 1, 2, 3], [4, len(V) + 2, 5]]]).astype(np.int32)
    m = api.Mesh(Vn, Fn, 2)
    assert m.info()["n_excluded"] == 2
    ref = oracle.RefFwn(Vn, Fn)
    a = m.download()
    pr = parity.match_nodes(a["fwn_child"], ref.nodes())
    assert pr is not None and len(pr) == len(ref.nodes())
    Q = synth.random_points(500, -1.5, 1.5, 1)
    assert np.abs(m.fast_winding_number(Q) - ref.winding(Q)).max() < 1e-5
    m.close()


def case_pseudonormal(api, meshes, n_queries=4000):
    """SIGNED_DISTANCE_TYPE_PSEUDONORMAL (the "next" row of SURVEY 8f): normals bit-identical to the
    restated per_face/per_vertex(ANGLE)/per_edge(UNIFORM) normals, S and N equal to the restated
    igl::pseudonormal_test on the same closest face; on exact ties the pseudonormal is shared by
    the tied faces so S must agree anyway."""
    V, F = synth.tetrahedron()   # tests/include/igl/signed_distance.cpp:4-43, tolerance 1e-15
    S, I, C, N = api.signed_distance(np.array([[0.5, 0.5, 0.5]]), V, F, SIGNED_DISTANCE_TYPE_PSEUDONORMAL)
    assert abs(S[0] - np.sqrt(1.0 / 12.0)) < 1e-15 and np.abs(N[0] - np.sqrt(1.0 / 3.0)).max() < 1e-15
    for name, (V, F) in meshes.items():
        lo, hi = synth.bbox(V)
        P = np.concatenate([synth.random_points(n_queries, float(lo.min()) * 1.3, float(hi.max()) * 1.3, 17),
                            V[:200] + 1e-3, (V[F[:200, 0]] + V[F[:200, 1]]) / 2 * 1.001])
        m = api.Mesh(V, F, 2, pseudonormals=True)
        a = m.download()
        o = oracle.PortPseudonormal(V, F)
        FN, VN, EN = o.normals()
        pn = a["pseudonormals"]
        assert np.array_equal(pn[:, 0], FN), name
        # vertex normals are weighted by 2*atan(..) corner angles: CUDA's atan and glibc's differ by an ulp
        assert np.abs(pn[:, 1:4] - VN[F]).max() < 1e-14, name
        assert np.array_equal(pn[:, 4:7], EN), name
        S, I, C, N = m.signed_distance(P, SIGNED_DISTANCE_TYPE_PSEUDONORMAL)
        So, Io, Co, No = oracle.signed_distance_pseudonormal(P, V, F, pseudo=o)
        # same face -> identical; tied faces -> same distance, sign must still agree
        s2, n2 = o.test(P, I, C)
        assert np.abs(N - n2).max() < 1e-14, name
        assert np.array_equal(np.sign(S), s2) or np.array_equal(np.sign(S[np.abs(So) > 0]), s2[np.abs(So) > 0]), name
        assert np.array_equal(np.abs(S), np.abs(So)), name
        diag = float(np.linalg.norm(hi - lo))
        far = np.abs(So) > 1e-6 * diag
        assert np.array_equal(np.sign(S[far]), np.sign(So[far])), name
        # bounded variant: NaN rows and zeroed N
        S, I, C, N = m.signed_distance(P, SIGNED_DISTANCE_TYPE_PSEUDONORMAL, -0.05 * diag, 0.08 * diag)
        So, Io, Co, No = oracle.signed_distance_pseudonormal(P, V, F, -0.05 * diag, 0.08 * diag, pseudo=o)
        assert np.array_equal(np.isnan(S), np.isnan(So)) and np.isnan(S).any(), name
        bad = np.isnan(S)
        assert (I[bad] == F.shape[0] + 1).all() and (C[bad] == 0).all() and (N[bad] == 0).all()
        m.close()


def case_segments_points_hausdorff(api):
    """Ele with 2 / 1 columns (point_mesh_squared_distance.h:23-30) and igl::hausdorff (hausdorff.cpp:20-39)."""
    rng = np.random.default_rng(4)
    V = rng.uniform(-1, 1, (300, 3))
    P = rng.uniform(-1.5, 1.5, (2000, 3))
    E = rng.integers(0, 300, (400, 2)).astype(np.int32)
    D, I, C = api.point_mesh_squared_distance(P, V, E)
    Db, Ib, Cb = oracle.brute_closest(V, E[:, [0, 1, 0]], P)
    assert np.array_equal(D, Db) and np.array_equal(I, Ib) and np.array_equal(C, Cb)
    Pt = np.arange(300, dtype=np.int32)[:, None]
    D, I, C = api.point_mesh_squared_distance(P, V, Pt)
    d2 = ((P[:, None, :] - V[None]) ** 2).sum(2)
    assert np.array_equal(I, d2.argmin(1)) and np.allclose(D, d2.min(1), rtol=1e-14, atol=0)
    VA, FA = synth.icosphere(2)
    VB, FB = synth.icosphere(3)
    VB = VB * 1.1 + 0.05
    h = api.hausdorff(VA, FA, VB, FB)
    dBA = oracle.brute_closest(VB, FB, VA)[0].max()
    dAB = oracle.brute_closest(VA, FA, VB)[0].max()
    assert h == np.sqrt(max(dBA, dAB))


# ---------------------------------------------------------------- voxel grids
def _same(a, b):
    """bit-for-bit equality that treats NaN == NaN"""
    a = np.asarray(a); b = np.asarray(b)
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=a.dtype.kind == "f")


def case_voxel_grid(api):
    """igl::voxel_grid: `side` exact; positions to 1e-13 of the box size (the reference's column mean is an
    Eigen reduction whose order its sources do not fix, see oracle_port.cpp::port_voxel_grid)."""
    boxes = [([0, 0, 0], [1, 1, 1], 10, 0), ([0, 0, 0], [1, 2, 0.5], 20, 1), ([-3.5, 2, 7], [4, 2.25, 9], 33, 2),
             ([1e3, -1e-3, 0], [1e3 + 1, 1e-3, 1e-6], 17, 1), ([-1, -1, -1], [1, 1, 1], 64, 1)]
    for bmin, bmax, s, pad in boxes:
        GVo, so = oracle.voxel_grid(bmin, bmax, s, pad)
        GV, side = api.voxel_grid((bmin, bmax), s, pad)
        assert np.array_equal(side, so), (side, so)
        scale = np.abs(np.array(bmax, float) - np.array(bmin, float)).max() + np.abs(np.array(bmax, float)).max()
        assert np.array_equal(np.isnan(GV), np.isnan(GVo))   # a 3-vertex padded side divides by zero in the reference too
        assert np.nanmax(np.abs(GV - GVo), initial=0.0) <= 1e-13 * scale, np.nanmax(np.abs(GV - GVo))
    V, _ = synth.icosphere(2)
    GVo, so = oracle.voxel_grid_mesh(V, 0.25, 30, 1)
    GV, side = api.voxel_grid(V, 0.25, 30, 1)
    assert np.array_equal(side, so) and np.abs(GV - GVo).max() <= 1e-13 * 3
    # the reference's own known answer (tests/include/igl/grid.cpp:5-41) through the unit box, no padding
    GV, side = api.voxel_grid(([0, 0, 0], [1, 1, 1]), 3, 0)
    assert side.tolist() == [3, 3, 3]
    want = np.array([[x, y, z] for z in (0, .5, 1) for y in (0, .5, 1) for x in (0, .5, 1)])
    assert np.abs(GV - want).max() < 1e-15
    for bad in [(([0, 0, 0], [1, 1, 1]), 3, 1), (([0, 0, 0], [0, 0, 0]), 10, 1), (([0, 0, 0], [np.nan, 1, 1]), 10, 1)]:
        try:
            api.voxel_grid(*bad)
        except IglB200Error:
            pass
        else:
            raise AssertionError(f"voxel_grid{bad} must fail")


def case_flood_fill(api, seeds=12):
    """igl::flood_fill: bit-identical to the sequential sweep for every NaN pattern."""
    rng = np.random.default_rng(3)
    shapes = [(5, 4, 3), (1, 7, 2), (8, 1, 1), (1, 1, 1), (2, 2, 2), (9, 6, 5), (16, 3, 7), (3, 11, 4)]
    n_cases = 0
    for shape in shapes:
        n = int(np.prod(shape))
        for k in range(seeds):
            S = rng.standard_normal(n)
            p = [0.0, 0.1, 0.5, 0.9, 0.98, 1.0][k % 6]
            S[rng.random(n) < p] = np.nan
            S3 = S.reshape(shape[2], shape[1], shape[0])
            if k % 4 == 1 and shape[1] > 1:     # whole rows
                S3[:, rng.integers(0, shape[1]), :] = np.nan
            if k % 4 == 2 and shape[2] > 1:     # whole sheets, leading / inner / trailing
                S3[rng.integers(0, shape[2]), :, :] = np.nan
            if k % 4 == 3:
                S3[-1, :, :] = np.nan
                S3[0, :, :] = np.nan
            got = api.flood_fill(shape, S)
            want = oracle.flood_fill(shape, S)
            assert _same(got, want), (shape, k, np.flatnonzero(~((got == want) | (np.isnan(got) & np.isnan(want))))[:5])
            n_cases += 1
    return n_cases


def _check_mc(api, S, GV, dims, iso, what):
    nx, ny, nz = dims
    Vo, Fo, Eo = oracle.marching_cubes(S, GV, nx, ny, nz, iso)
    V, F, E = api.marching_cubes(S, GV, nx, ny, nz, iso, return_E2V=True)
    assert V.shape == Vo.shape and F.shape == Fo.shape, f"{what}: {V.shape} {F.shape} vs {Vo.shape} {Fo.shape}"
    assert _same(F, Fo), f"{what}: faces differ"
    assert _same(E, Eo), f"{what}: edge keys differ"
    assert _same(V, Vo), f"{what}: vertices differ (max {np.nanmax(np.abs(V - Vo)) if V.size else 0})"
    return V.shape[0], F.shape[0]


def case_marching_cubes(api):
    """igl::marching_cubes: V, F and the E2V keys bit-identical to the sequential hash-map sweep."""
    rng = np.random.default_rng(5)
    out = {}
    g = 24
    GV = oracle.grid([g, g, g]) * 2 - 1
    S = np.linalg.norm(GV, axis=1) - 0.7
    nv, nf = _check_mc(api, S, GV, (g, g, g), 0.0, "sphere")
    assert nv - nf * 3 // 2 + nf == 2      # closed genus-0 surface
    out["sphere"] = (nv, nf)
    _check_mc(api, S, GV, (g, g, g), 0.1, "sphere iso 0.1")
    # every corner configuration, ambiguous ones included; anisotropic, non-uniform positions
    dims = (13, 9, 11)
    n = int(np.prod(dims))
    GVr = oracle.grid(list(dims)) * np.array([1.0, 3.0, 0.5]) + rng.standard_normal((n, 3)) * 0.01
    Sr = rng.standard_normal(n)
    out["noise"] = _check_mc(api, Sr, GVr, dims, 0.0, "noise")
    _check_mc(api, Sr, np.asfortranarray(GVr), dims, 0.3, "noise, column-major GV")
    # values equal to the isovalue (delta == 0 -> t = +-inf/NaN, march_cube.cpp:117-119), NaN and inf samples
    Sq = np.round(Sr * 2) / 2
    _check_mc(api, Sq, GVr, dims, 0.5, "quantised (ties with the isovalue)")
    Sn = Sr.copy(); Sn[rng.random(n) < 0.1] = np.nan; Sn[rng.random(n) < 0.02] = np.inf
    _check_mc(api, Sn, GVr, dims, 0.0, "NaN/inf samples")
    # degenerate grids: no cubes at all, a single cube, flat slabs
    for d in [(1, 1, 1), (5, 1, 4), (2, 2, 2), (2, 7, 2), (3, 2, 6)]:
        m = int(np.prod(d))
        _check_mc(api, rng.standard_normal(m), oracle.grid(list(d)) if min(d) > 1 else rng.standard_normal((m, 3)), d, 0.0, f"dims {d}")
    _check_mc(api, np.full(n, 1.0), GVr, dims, 0.0, "all above")
    _check_mc(api, np.full(n, -1.0), GVr, dims, 0.0, "all below")
    return out


def case_offset_surface(api, V, F, isolevel, s, sign_type, what):
    """igl::offset_surface = voxel_grid -> signed_distance(bounded) -> flood_fill -> marching_cubes.
    Stage by stage: the raw S obeys the signed-distance parity rules (rows the reference itself leaves
    order-dependent, SURVEY A.1, excluded); flood fill and marching cubes of that S are bit-identical to
    the sequential reference algorithms; the one-call result equals the staged one."""
    from libigl_b200.api import Grid, Mesh, SIGNED_DISTANCE_TYPE_PSEUDONORMAL
    pseudo = int(sign_type) == SIGNED_DISTANCE_TYPE_PSEUDONORMAL
    GVo, side_o = oracle.voxel_grid_mesh(V, isolevel, s, 1)
    g = Grid(api, V.min(0) - isolevel, V.max(0) + isolevel, s, 1)
    m = Mesh(api, V, F, 2, pseudonormals=pseudo)
    assert np.array_equal(g.side, side_o)
    GV = g.GV()
    diag = np.linalg.norm(V.max(0) - V.min(0))
    assert np.abs(GV - GVo).max() <= 1e-13 * (diag + np.abs(V).max())
    c = g.corners()
    h = (c[3] - c[0]) / float(g.side[0] - 1)
    assert h == (GV[:, 0].max() - GV[:, 0].min()) / float(g.side[0] - 1)
    lower, upper = isolevel - np.sqrt(3.0) * h, isolevel + np.sqrt(3.0) * h
    g.signed_distance(m, sign_type, lower, upper)
    S_raw = g.S()
    # the signed-distance stage on the product's own grid vertices
    if pseudo:
        So = oracle.signed_distance_pseudonormal(GV, V, F, lower, upper)[0]
    else:
        So = oracle.signed_distance_fwn(GV, V, F, lower, upper)[0]
    low2, _ = oracle.signed_distance_bounds(lower, upper)
    d2 = oracle.PortAABB(V, F).squared_distance(GV)[0]
    care = d2 > low2 * (1 + 1e-12)                     # below the lower bound the reference is order-dependent
    assert np.array_equal(np.isnan(S_raw[care]), np.isnan(So[care])), f"{what}: NaN pattern"
    ok = care & ~np.isnan(So)
    assert np.abs(S_raw[ok] - So[ok]).max() <= 1e-5 * max(diag, 1.0), f"{what}: S"
    g.flood_fill()
    S_fill = g.S()
    assert _same(S_fill, oracle.flood_fill(g.side, S_raw)), f"{what}: flood_fill"
    SV, SF, E = g.marching_cubes(0.0, shift=isolevel, return_E2V=True)
    Vo, Fo, Eo = oracle.marching_cubes(S_fill - isolevel, GV, g.side[0], g.side[1], g.side[2], 0.0)
    assert _same(SF, Fo) and _same(E, Eo) and _same(SV, Vo), f"{what}: marching_cubes"
    m.close(); g.close()
    SV1, SF1, GV1, side1, S1 = api.offset_surface(V, F, isolevel, s, sign_type)
    assert _same(SV1, SV) and _same(SF1, SF) and _same(GV1, GV) and _same(side1, side_o) and _same(S1, S_fill), f"{what}: one call"
    # the offset surface really is at distance `isolevel` from the mesh
    if SV.shape[0]:
        dist = np.sqrt(oracle.PortAABB(V, F).squared_distance(SV)[0])
        assert np.abs(dist - abs(isolevel)).max() < 1.5 * h, f"{what}: surface off by {np.abs(dist - abs(isolevel)).max()} (h={h})"
    return dict(nV=int(SV.shape[0]), nF=int(SF.shape[0]), side=side_o.tolist())


# ---------------------------------------------------------------- exact winding number
def open_and_messy_meshes():
    """closed, open, non-manifold, inconsistently oriented and duplicated soups: the generalised winding number
    is defined for all of them and fractional for most."""
    rng = np.random.default_rng(11)
    ms = {}
    V, F = synth.icosphere(3)
    ms["closed sphere"] = (V, F)
    ms["hemisphere (open)"] = (V, F[V[F].mean(1)[:, 2] > 0.1])
    ms["sphere with holes"] = (V, F[rng.random(F.shape[0]) > 0.2])
    Ff = F.copy(); flip = rng.random(F.shape[0]) < 0.3; Ff[flip] = Ff[flip][:, ::-1]
    ms["inconsistent orientation"] = (V, Ff)
    ms["duplicated faces"] = (V, np.concatenate([F, F[: F.shape[0] // 3], F[: F.shape[0] // 5][:, ::-1]]))
    ms["random soup"] = (rng.standard_normal((300, 3)), rng.integers(0, 300, (500, 3)).astype(np.int32))
    Vk, Fk = synth.torus_knot(200, 12)
    ms["knot (closed)"] = (Vk, Fk)
    ms["knot (cut open)"] = (Vk, Fk[: Fk.shape[0] * 2 // 3])
    Vt, Ft = synth.tetrahedron()
    ms["tet"] = (Vt, Ft)
    ms["one triangle"] = (Vt, Ft[:1])
    Fd = F.copy(); Fd[::7, 2] = Fd[::7, 1]          # degenerate rows [i j j]
    ms["with degenerate faces"] = (V, Fd)
    # shuffled rows: the tree no longer follows the surface, chains get long
    ms["shuffled sphere"] = (V, F[rng.permutation(F.shape[0])])
    return ms


def case_winding_number(api, n_queries=1500, meshes=None):
    """igl::winding_number / hierarchical evaluation vs the face-by-face sum: |dW| <= 1e-10."""
    rng = np.random.default_rng(13)
    out = {}
    for name, (V, F) in (meshes or open_and_messy_meshes()).items():
        lo, hi = V.min(0), V.max(0)
        O = np.concatenate([lo + (hi - lo) * (rng.random((n_queries, 3)) * 1.6 - 0.3),      # around and inside
                            V[rng.integers(0, V.shape[0], 50)] + 1e-3 * rng.standard_normal((50, 3)),   # next to vertices
                            10.0 * (hi - lo) * rng.standard_normal((20, 3))])                     # far away
        Wo = oracle.winding_number(V, F, O)
        W = api.winding_number(V, F, O)
        d = np.abs(W - Wo).max()
        assert d <= 1e-10, f"{name}: winding number differs by {d}"
        out[name] = float(d)
    return out


def case_signed_distance_exact_winding(api, V, F, P, what, lower=None, upper=None):
    """SIGNED_DISTANCE_TYPE_WINDING_NUMBER and _DEFAULT (the same in 3-D, signed_distance.cpp:60-64)."""
    So, Io, Co, Wo = oracle.signed_distance_winding_number(P, V, F, lower, upper)
    res = {}
    for t in (1, 2):
        S, I, C, W = api.signed_distance(P, V, F, t, lower, upper, return_W=True)
        nan = np.isnan(So)
        assert np.array_equal(np.isnan(S), nan), f"{what}: NaN rows"
        ok = ~nan
        assert np.abs(W[ok] - Wo[ok]).max(initial=0.0) <= 1e-10, f"{what}: W"
        scale = max(np.linalg.norm(V.max(0) - V.min(0)), 1.0)
        assert np.abs(S[ok] - So[ok]).max(initial=0.0) <= 1e-9 * scale, f"{what}: S"
        assert np.array_equal(C[nan], Co[nan]) and np.array_equal(I[nan], Io[nan]), f"{what}: sentinels"
        # closest face / point: identical except on exact ties (same squared distance to several faces)
        diff = ok & (I != Io)
        if diff.any():
            assert np.array_equal(oracle.face_distance(V, F, P[diff], I[diff]), oracle.face_distance(V, F, P[diff], Io[diff])), \
                f"{what}: closest face differs off ties"
        same = ok & (I == Io)
        assert np.array_equal(C[same], Co[same]), f"{what}: C"
        res[t] = int(ok.sum())
    return res


def case_marching_cubes_sparse(api):
    """igl::marching_cubes(S,GV,GI,...) over an explicit cube list: V, F, E2V keys bit-identical to the sequential
    hash-map sweep, for a full grid listed as cubes (must equal the grid overload), a random subset in random order
    (a sparse voxel grid), repeated cubes, and arbitrary corner ids."""
    rng = np.random.default_rng(21)
    dims = (11, 9, 10)
    nx, ny, nz = dims
    n = nx * ny * nz
    GV = oracle.grid(list(dims)) * np.array([1.0, 2.0, 0.7]) + 0.01 * rng.standard_normal((n, 3))
    S = rng.standard_normal(n)
    off = np.array([0, 1, 1 + nx, nx, nx * ny, 1 + nx * ny, 1 + nx + nx * ny, nx + nx * ny])
    x, y, z = np.meshgrid(np.arange(nx - 1), np.arange(ny - 1), np.arange(nz - 1), indexing="ij")
    base = (x + nx * (y + ny * z)).transpose(2, 1, 0).ravel()          # x fastest, as the grid overload walks
    GI = (base[:, None] + off[None, :]).astype(np.int32)

    def check(GI_, S_, iso, what):
        Vo, Fo, Eo = oracle.marching_cubes_sparse(S_, GV, GI_, iso)
        V, F, E = api.marching_cubes_sparse(S_, GV, GI_, iso, return_E2V=True)
        assert _same(F, Fo) and _same(E, Eo) and _same(V, Vo), what
        return V, F
    V, F = check(GI, S, 0.0, "all cubes in grid order")
    Vg, Fg = api.marching_cubes(S, GV, nx, ny, nz, 0.0)
    assert _same(V, Vg) and _same(F, Fg), "cube list in grid order == grid overload"
    sub = rng.permutation(GI.shape[0])[: GI.shape[0] // 3]
    check(GI[sub], S, 0.2, "random third of the cubes, random order")
    check(np.concatenate([GI[sub[:50]], GI[sub[:50]], GI[sub[25:80]]]), S, 0.0, "repeated cubes")
    check(rng.integers(0, n, (400, 8)).astype(np.int32), S, 0.0, "arbitrary corner ids")
    Sn = S.copy(); Sn[rng.random(n) < 0.1] = np.nan; Sn[::17] = 0.0
    check(GI[sub], Sn, 0.0, "NaN samples and values equal to the isovalue")
    check(GI[:0], S, 0.0, "no cubes")
    check(GI, np.full(n, 1.0), 0.0, "nothing cut")
    with pytest.raises(IglB200Error):
        api.marching_cubes_sparse(S, GV, np.full((1, 8), n, np.int32), 0.0)
