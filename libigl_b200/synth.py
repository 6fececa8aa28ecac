"""Seeded synthetic inputs for the configs in BASELINE.json.

Shared by the oracle, the tests and bench.py so both sides see identical bytes.
Everything is numpy on the host; meshes are (V float64 [nV,3] C-order,
F int32 [nF,3] C-order), closed and outward oriented unless stated.

Recipes follow the reference's own generators where one exists:
  icosphere : igl/icosahedron.cpp:5-44 vertices/faces + igl/upsample.cpp:17-98
              midpoint subdivision, then normalised to the unit sphere.
"""
from __future__ import annotations

import numpy as np


def f32_exact(a: np.ndarray) -> np.ndarray:
    """float64 array whose values are exactly representable in float32.

    The reference casts each query to float for the winding number
    (signed_distance.cpp:121); with such inputs the cast is exact.
    """
    return np.ascontiguousarray(a, dtype=np.float32).astype(np.float64)


# ---------------------------------------------------------------------------
# meshes
# ---------------------------------------------------------------------------
def icosahedron():
    t = (1.0 + np.sqrt(5.0)) / 2.0
    V = np.array(
        [
            [-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0],
            [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
            [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1],
        ],
        dtype=np.float64,
    )
    V /= np.linalg.norm(V, axis=1, keepdims=True)
    F = np.array(
        [
            [0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11],
            [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
            [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9],
            [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1],
        ],
        dtype=np.int32,
    )
    return V, F


def upsample(V, F):
    """One level of 1->4 midpoint subdivision (pattern of igl/upsample.cpp)."""
    nV = V.shape[0]
    E = np.concatenate([F[:, [1, 2]], F[:, [2, 0]], F[:, [0, 1]]], axis=0).astype(np.int64)
    Es = np.sort(E, axis=1)
    key = Es[:, 0] * nV + Es[:, 1]
    uniq, inv = np.unique(key, return_inverse=True)
    mid = 0.5 * (V[uniq // nV] + V[uniq % nV])
    nF = F.shape[0]
    m0 = (nV + inv[0:nF]).astype(np.int32)        # opposite corner 0 (edge 1-2)
    m1 = (nV + inv[nF:2 * nF]).astype(np.int32)    # edge 2-0
    m2 = (nV + inv[2 * nF:3 * nF]).astype(np.int32)  # edge 0-1
    a, b, c = F[:, 0], F[:, 1], F[:, 2]
    NF = np.concatenate(
        [
            np.stack([a, m2, m1], 1),
            np.stack([b, m0, m2], 1),
            np.stack([c, m1, m0], 1),
            np.stack([m0, m1, m2], 1),
        ],
        axis=0,
    ).astype(np.int32)
    return np.concatenate([V, mid], axis=0), np.ascontiguousarray(NF)


def icosphere(level: int):
    """Unit icosphere, 20*4^level faces (level 4 = 5 120 faces: config c1)."""
    V, F = icosahedron()
    for _ in range(level):
        V, F = upsample(V, F)
        V /= np.linalg.norm(V, axis=1, keepdims=True)
    return np.ascontiguousarray(V), np.ascontiguousarray(F)


def noisy_sphere(level: int, sigma: float = 0.01, seed: int = 0):
    """Icosphere with seeded radial noise (config c3: level 9 = 5 242 880 faces)."""
    V, F = icosphere(level)
    rng = np.random.default_rng(seed)
    r = 1.0 + sigma * rng.standard_normal(V.shape[0])
    return np.ascontiguousarray(V * r[:, None]), F


def torus_knot(nu: int, nv: int, p: int = 2, q: int = 3, tube: float = 0.35):
    """Closed tube around a (p,q) torus knot, 2*nu*nv faces, outward oriented.

    Config c2: nu*nv = 500 000 (1 M faces); c5: 5 000 000 (10 M faces).
    """
    t = np.linspace(0.0, 2.0 * np.pi, nu, endpoint=False)

    def curve(t):
        r = 2.0 + np.cos(q * t)
        return np.stack([r * np.cos(p * t), r * np.sin(p * t), np.sin(q * t)], 1)

    def d1(t):
        r = 2.0 + np.cos(q * t)
        dr = -q * np.sin(q * t)
        return np.stack(
            [dr * np.cos(p * t) - p * r * np.sin(p * t),
             dr * np.sin(p * t) + p * r * np.cos(p * t),
             q * np.cos(q * t)], 1)

    def d2(t):
        r = 2.0 + np.cos(q * t)
        dr = -q * np.sin(q * t)
        ddr = -q * q * np.cos(q * t)
        return np.stack(
            [ddr * np.cos(p * t) - 2 * p * dr * np.sin(p * t) - p * p * r * np.cos(p * t),
             ddr * np.sin(p * t) + 2 * p * dr * np.cos(p * t) - p * p * r * np.sin(p * t),
             -q * q * np.sin(q * t)], 1)

    C = curve(t)
    T = d1(t)
    T /= np.linalg.norm(T, axis=1, keepdims=True)
    A = d2(t)
    Nn = A - np.sum(A * T, 1, keepdims=True) * T   # Frenet normal (curvature never vanishes)
    Nn /= np.linalg.norm(Nn, axis=1, keepdims=True)
    B = np.cross(T, Nn)
    s = np.linspace(0.0, 2.0 * np.pi, nv, endpoint=False)
    cs, sn = np.cos(s), np.sin(s)
    V = (C[:, None, :] + tube * (cs[None, :, None] * Nn[:, None, :] + sn[None, :, None] * B[:, None, :]))
    V = V.reshape(nu * nv, 3)
    i = np.arange(nu)[:, None]
    j = np.arange(nv)[None, :]
    i1 = (i + 1) % nu
    j1 = (j + 1) % nv
    v00 = (i * nv + j).ravel()
    v10 = (i1 * nv + j).ravel()
    v01 = (i * nv + j1).ravel()
    v11 = (i1 * nv + j1).ravel()
    # (N, B, T) is right handed: going +s then +t gives outward normal with (v00, v01, v11)
    F = np.concatenate([np.stack([v00, v01, v11], 1), np.stack([v00, v11, v10], 1)], 0).astype(np.int32)
    # interleave the two triangles of each quad so face order follows the surface
    nq = nu * nv
    order = np.empty(2 * nq, dtype=np.int64)
    order[0::2] = np.arange(nq)
    order[1::2] = np.arange(nq) + nq
    F = np.ascontiguousarray(F[order])
    # orient outward (signed volume must be positive)
    a, b, c = V[F[:, 0]], V[F[:, 1]], V[F[:, 2]]
    vol = np.sum(a * np.cross(b, c)) / 6.0
    if vol < 0:
        F = np.ascontiguousarray(F[:, [0, 2, 1]])
    return np.ascontiguousarray(V), F


def knot_dims(n_faces: int):
    """(nu, nv) with 2*nu*nv == n_faces and a tube cross-section of ~1/40 of the length."""
    nq = n_faces // 2
    nv = max(8, int(round(np.sqrt(nq / 40.0))))
    while nq % nv:
        nv -= 1
    return nq // nv, nv


def tetrahedron():
    """The reference's single_tet fixture (tests/include/igl/signed_distance.cpp:6-18)."""
    V = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=np.float64)
    F = np.array([[0, 1, 3], [0, 2, 1], [0, 3, 2], [1, 2, 3]], dtype=np.int32)
    return V, F


def soupify(V, F):
    """Every face gets its own three vertices (tests/include/igl/AABB.cpp:96-98)."""
    V2 = np.ascontiguousarray(V[F.reshape(-1)])
    F2 = np.arange(V2.shape[0], dtype=np.int32).reshape(-1, 3)
    return V2, np.ascontiguousarray(F2)


# ---------------------------------------------------------------------------
# queries
# ---------------------------------------------------------------------------
def bbox(V):
    return V.min(0), V.max(0)


def grid_points(lo, hi, res: int, scale: float = 1.0):
    """res^3 grid over the bbox scaled about its centre; f32-exact doubles, x fastest."""
    lo = np.asarray(lo, np.float64)
    hi = np.asarray(hi, np.float64)
    c = 0.5 * (lo + hi)
    h = 0.5 * (hi - lo) * scale
    ax = [np.linspace(c[d] - h[d], c[d] + h[d], res) for d in range(3)]
    Z, Y, X = np.meshgrid(ax[2], ax[1], ax[0], indexing="ij")
    P = np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1)
    return f32_exact(P)


def random_points(n: int, lo=-1.5, hi=1.5, seed: int = 1):
    rng = np.random.default_rng(seed)
    return f32_exact(rng.uniform(lo, hi, size=(n, 3)))


def oriented_cloud(V, F, n: int, seed: int = 2):
    """n distinct points on the surface with outward unit normals and equal areas
    (config c4; what copyleft/cgal/point_areas would estimate is supplied directly)."""
    rng = np.random.default_rng(seed)
    a, b, c = V[F[:, 0]], V[F[:, 1]], V[F[:, 2]]
    fn = np.cross(b - a, c - a)
    dbl = np.linalg.norm(fn, axis=1)
    cdf = np.cumsum(dbl)
    total = cdf[-1] * 0.5
    f = np.searchsorted(cdf, rng.uniform(0, cdf[-1], n)).clip(0, F.shape[0] - 1)
    u = rng.uniform(size=n)
    v = rng.uniform(size=n)
    flip = u + v > 1
    u[flip] = 1 - u[flip]
    v[flip] = 1 - v[flip]
    P = a[f] + u[:, None] * (b[f] - a[f]) + v[:, None] * (c[f] - a[f])
    N = fn[f] / dbl[f][:, None]
    A = np.full(n, total / n)
    return np.ascontiguousarray(P), np.ascontiguousarray(N), A
