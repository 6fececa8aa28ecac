"""Parity tests proper: the CUDA library (libigl_b200/libiglb200.so, through the
C ABI) against the oracle, on a real B200.  Same gates as the CPU logic tests
plus the BASELINE.json configs at full size."""
import numpy as np
import pytest

import cases
import oracle
import parity
from libigl_b200 import synth

pytestmark = pytest.mark.gpu


def test_known_answers(gpu_api):
    cases.case_known_answers(gpu_api)


def test_tree_topology_and_moments_match_reference(gpu_api):
    ms = cases.meshes_small()
    ms["axis"] = cases.axis_aligned_box_mesh(8)
    ms["ico5"] = synth.icosphere(5)
    ms["knot40k"] = synth.torus_knot(1000, 20)
    ms["knot1M"] = synth.torus_knot(*synth.knot_dims(1_000_000))
    cases.case_tree_exact(gpu_api, ms)


def test_tiny_meshes_every_size(gpu_api):
    V, F = synth.icosphere(2)
    rng = np.random.default_rng(0)
    ms = {f"n{n}": (V, F[rng.permutation(F.shape[0])[:n]]) for n in list(range(1, 41)) + [64, 65, 100]}
    cases.case_tree_exact(gpu_api, ms)


def test_c1_icosphere_grid64(gpu_api):
    """BASELINE config c1: 5 120-face icosphere, 64^3 grid over bbox x 1.5."""
    V, F = synth.icosphere(4)
    P = synth.grid_points(*synth.bbox(V), 64, 1.5)
    r = cases.case_signed_distance(gpu_api, V, F, P, "c1")
    assert r["n"] == 64 ** 3 and r["nan"] == 0


def test_signed_distance_knot200k(gpu_api):
    V, F = synth.torus_knot(2000, 50)
    P = synth.grid_points(*synth.bbox(V), 48, 1.1)
    cases.case_signed_distance(gpu_api, V, F, P, "knot200k grid")


def test_closest_vs_brute(gpu_api):
    for name, (V, F) in cases.meshes_small().items():
        P = synth.random_points(3000, -2.0, 2.0, 1)
        cases.case_closest_vs_brute(gpu_api, V, F, P, name)


def test_closest_extreme_scales(gpu_api):
    cases.case_closest_extreme_scales(gpu_api)


def test_iterative_closest_point(gpu_api):
    cases.case_icp(gpu_api)


def test_fwn(gpu_api):
    V, F = synth.noisy_sphere(4, 0.02, 7)
    cases.case_fwn(gpu_api, V, F, synth.random_points(50000, -1.5, 1.5, 2), "noisy")


def test_bounds_and_sentinels(gpu_api):
    cases.case_bounds_and_sentinels(gpu_api)


def test_layouts(gpu_api):
    cases.case_layouts(gpu_api)


def test_degenerate(gpu_api):
    cases.case_degenerate(gpu_api)


def test_arena_roundtrip(gpu_api):
    cases.case_arena_roundtrip(gpu_api)


def test_c2_full_size_sample_and_properties(gpu_api):
    """BASELINE config c2 at full size: 1 M-face torus knot, 256^3 grid.
    The oracle checks a 200 k-query strided sample; the whole grid is checked
    through size-independent properties."""
    V, F = synth.torus_knot(*synth.knot_dims(1_000_000))
    assert F.shape[0] == 1_000_000
    m = gpu_api.Mesh(V, F, 2)
    P = synth.grid_points(*synth.bbox(V), 256, 1.1)
    # column-major doubles (Eigen's default layout) take the streamed path: 8 chunks on two streams
    S, I, C, W = m.signed_distance(np.asfortranarray(P), return_W=True)
    assert not np.isnan(S).any()
    # the row-major call (whole-batch path) must give the same bits
    kk = slice(5_000_000, 5_300_000)
    S2, I2, C2, W2 = m.signed_distance(P[kk], return_W=True)
    assert np.array_equal(S2, S[kk]) and np.array_equal(I2, I[kk]) and np.array_equal(C2, C[kk]) and np.array_equal(W2, W[kk])
    # properties on all 16.7 M rows
    assert (I >= 0).all() and (I < F.shape[0]).all()
    d = np.linalg.norm(P - C, axis=1)
    assert np.abs(np.abs(S) - d * np.abs(1 - 2 * np.abs(W))).max() < 1e-12      # S = (1-2|w|) sqrt(d2)
    inside = W > 0.5
    assert (S[inside] <= 0).all() and (S[~inside] >= 0).all()
    assert 0.001 < inside.mean() < 0.2                                           # the tube is thin
    # closest point lies on the reported face: re-evaluating that face gives the same distance
    k = np.arange(0, P.shape[0], 83)
    assert np.array_equal(oracle.face_distance(V, F, P[k], I[k]), (d[k] ** 2)) or \
        np.abs(oracle.face_distance(V, F, P[k], I[k]) - d[k] ** 2).max() < 1e-12
    # oracle on a strided sample
    ref = oracle.RefFwn(V, F)
    tree = oracle.PortAABB(V, F)
    r = parity.check_signed_distance(V, F, P[k], S[k], I[k], C[k], W[k], what="c2 sample", fwn=ref, tree=tree)
    assert r["n"] == len(k)
    # barycenters of a face sample: I == face, d2 == 0 (tests/include/igl/AABB.cpp:111-112 at 1 M faces)
    fs = np.arange(0, F.shape[0], 97)
    BC = (V[F[fs, 0]] + V[F[fs, 1]] + V[F[fs, 2]]) / 3.0
    D, Ib, Cb = m.squared_distance(BC)
    assert np.abs(D).max() < 1e-24
    same = Ib == fs
    assert same.mean() > 0.99   # a barycentre rounded onto a neighbouring face's plane may tie at d2==0
    assert (oracle.face_distance(V, F, BC[~same], Ib[~same]) <= D[~same] + 1e-30).all()
    m.close()


def test_c3_noisy_sphere_random_queries(gpu_api):
    """BASELINE config c3 (reduced to 1.3 M faces / 2 M queries for the oracle's sake):
    squared_distance only, bit-exact distances against the restated AABB."""
    V, F = synth.noisy_sphere(8, 0.01, 0)
    P = synth.random_points(2_000_000, -1.5, 1.5, 1)
    m = gpu_api.Mesh(V, F, 0)
    D, I, C = m.squared_distance(P)
    k = np.arange(0, P.shape[0], 10)
    Do, Io, Co = oracle.PortAABB(V, F).squared_distance(P[k])
    parity.check_closest(V, F, P[k], D[k], I[k], C[k], Do, Io, Co, "c3 sample")
    assert np.abs(np.linalg.norm(P - C, axis=1) ** 2 - D).max() < 1e-12
    m.close()


def test_c3_full_size(gpu_api):
    """BASELINE config c3 at full size: 5.2 M-face noisy sphere, 10 M random queries, AABB::squared_distance.
    Oracle (restated f64 AABB) on a strided 100 k-row sample: d2 and C bit-identical; properties on all rows."""
    V, F = synth.noisy_sphere(9, 0.01, 0)
    assert F.shape[0] == 5_242_880
    P = synth.random_points(10_000_000, -1.5, 1.5, 1)
    m = gpu_api.Mesh(V, F, 0)
    D, I, C = m.squared_distance(np.asfortranarray(P))
    m.close()
    assert (I >= 0).all() and (I < F.shape[0]).all()
    assert np.abs(((P - C) ** 2).sum(1) - D).max() < 1e-12
    k = np.arange(0, P.shape[0], 100)
    Do, Io, Co = oracle.PortAABB(V, F).squared_distance(P[k])
    parity.check_closest(V, F, P[k], D[k], I[k], C[k], Do, Io, Co, "c3 full size")
    k2 = np.arange(7, P.shape[0], 997)
    assert np.array_equal(oracle.face_distance(V, F, P[k2], I[k2]), D[k2])   # the reported face attains the reported distance


def test_c4_full_size(gpu_api):
    """BASELINE config c4 at full size: 2 M oriented points, 20 M queries, order 2, beta 2.  Oracle (restated octree +
    expansion, pinned against the Taylor series in test_oracle.py) on a strided 40 k-row sample."""
    V, F = synth.icosphere(7)
    Pc, N, A = synth.oriented_cloud(V, F, 2_000_000, 2)
    Q = synth.random_points(20_000_000, -1.5, 1.5, 2)
    c = gpu_api.Cloud(Pc, N, A, 2)
    W = c.eval(np.asfortranarray(Q), 2.0)
    o = oracle.PortCloud(Pc, N, A, 2)
    assert c.info()["n_cells"] == o.num_cells()
    c.close()
    k = np.arange(0, Q.shape[0], 500)
    Wo = o.eval(Q[k], 2.0)
    d = np.abs(W[k] - Wo)
    assert (d <= 1e-5 * np.maximum(1.0, np.abs(Wo))).all() and d.max() < 1e-10, d.max()
    r = np.linalg.norm(Q, axis=1)
    assert (W[r < 0.9] > 0.9).all() and (np.abs(W[r > 1.2]) < 0.1).all()          # inside / outside the unit sphere
    assert np.isfinite(W).all()


def test_c5_full_size(gpu_api):
    """BASELINE config c5 at full size: 10 M-face torus knot, 512^3 = 134 M grid queries, FAST_WINDING_NUMBER sign.
    Oracle (genuine FWN core + restated AABB) on a strided ~33 k-row sample; size-independent properties on all rows."""
    V, F = synth.torus_knot(*synth.knot_dims(10_000_000))
    assert F.shape[0] == 10_000_000
    m = gpu_api.Mesh(V, F, 2)
    P = synth.grid_points(*synth.bbox(V), 512, 1.1)
    n = P.shape[0]
    assert n == 512 ** 3
    Pf = np.asfortranarray(P)
    S, I, C, W = m.signed_distance(Pf, return_W=True)
    m.close()
    assert not np.isnan(S).any()
    assert (I >= 0).all() and (I < F.shape[0]).all()
    d2 = np.zeros(n)
    for j in range(3):                                   # column by column: 134 M rows
        d2 += (Pf[:, j] - C[:, j]) ** 2
    assert np.abs(np.abs(S) - np.sqrt(d2) * np.abs(1 - 2 * np.abs(W))).max() < 1e-12      # S = (1-2|w|) sqrt(d2)
    inside = W > 0.5
    assert (S[inside] <= 0).all() and (S[~inside] >= 0).all()
    assert 0.001 < inside.mean() < 0.2
    k = np.arange(0, n, 4099)
    Pk = np.ascontiguousarray(P[k])
    assert np.abs(oracle.face_distance(V, F, Pk, I[k]) - d2[k]).max() < 1e-12
    r = parity.check_signed_distance(V, F, Pk, S[k], I[k], np.ascontiguousarray(C[k]), W[k], what="c5 sample",
                                     fwn=oracle.RefFwn(V, F), tree=oracle.PortAABB(V, F))
    assert r["n"] == len(k)


def test_adversarial_triangles(gpu_api):
    cases.case_adversarial_triangles(gpu_api)


def test_cloud_golden(gpu_api):
    cases.case_cloud_golden(gpu_api)


def test_cloud_vs_oracle(gpu_api):
    cases.case_cloud(gpu_api, 3000, 2000)


def test_c4_cloud_reduced(gpu_api):
    """BASELINE config c4 reduced to what the CPU oracle finishes quickly: 200 k points, 100 k queries."""
    cases.case_cloud(gpu_api, 200000, 100000, orders=(2,), betas=(2.0,))


def test_golden_fixtures(gpu_api):
    cases.case_golden_fixtures(gpu_api)


def test_builder_corner_paths(gpu_api):
    cases.case_builder_corner_paths(gpu_api)


def test_pseudonormal_sign(gpu_api):
    cases.case_pseudonormal(gpu_api, dict(cases.meshes_small(), knot200k=synth.torus_knot(2000, 50)))


def test_segments_points_hausdorff(gpu_api):
    cases.case_segments_points_hausdorff(gpu_api)


# ---- voxel grids around the path ----
def test_voxel_grid(gpu_api):
    cases.case_voxel_grid(gpu_api)


def test_flood_fill(gpu_api):
    assert cases.case_flood_fill(gpu_api) > 90
    # a large volume with every kind of gap: a signed band around a sphere, everything else NaN
    g = 160
    GV = oracle.grid([g, g, g]) * 2 - 1
    S = np.linalg.norm(GV, axis=1) - 0.6
    S[np.abs(S) > 0.05] = np.nan
    got = gpu_api.flood_fill([g, g, g], S)
    assert np.array_equal(got, oracle.flood_fill([g, g, g], S), equal_nan=True)
    assert not np.isnan(got).any()


def test_marching_cubes(gpu_api):
    cases.case_marching_cubes(gpu_api)
    # 1.7 M cubes, 100k+ triangles: still the reference's vertex order to the bit
    g = 120
    GV = oracle.grid([g, g, g]) * 2 - 1
    S = np.sin(7 * GV[:, 0]) * np.cos(5 * GV[:, 1]) + GV[:, 2]
    Vo, Fo, Eo = oracle.marching_cubes(S, GV, g, g, g, 0.0)
    V, F, E = gpu_api.marching_cubes(S, GV, g, g, g, 0.0, return_E2V=True)
    assert Fo.shape[0] > 50_000
    assert np.array_equal(F, Fo) and np.array_equal(E, Eo) and np.array_equal(V, Vo)


def test_offset_surface(gpu_api):
    from libigl_b200.api import SIGNED_DISTANCE_TYPE_FAST_WINDING_NUMBER as FWN, SIGNED_DISTANCE_TYPE_PSEUDONORMAL as PN
    V, F = synth.icosphere(2)
    cases.case_offset_surface(gpu_api, V, F, 0.2, 20, FWN, "ico2 +0.2 fwn")
    cases.case_offset_surface(gpu_api, V, F, 0.15, 18, PN, "ico2 +0.15 pseudonormal")
    V, F = synth.torus_knot(400, 24)
    r = cases.case_offset_surface(gpu_api, V, F, 0.04, 96, FWN, "knot19k +0.04 fwn, 96 grid")
    assert r["nF"] > 20_000


def test_grid_pipeline_stays_on_device_tutorial_705(gpu_api):
    """tutorial/705_MarchingCubes/main.cpp:24-44: voxel_grid(V,0,s,1) -> signed_distance (default bounds) ->
    marching_cubes at 0 re-creates the input surface: every vertex within one cell of the mesh."""
    from libigl_b200.api import Grid, Mesh, SIGNED_DISTANCE_TYPE_PSEUDONORMAL as PN
    V, F = synth.torus_knot(600, 30)
    g = Grid(gpu_api, V.min(0), V.max(0), 128, 1)
    m = Mesh(gpu_api, V, F, 2, pseudonormals=True)
    g.signed_distance(m, PN)
    SV, SF = g.marching_cubes(0.0)
    S, GV = g.S(), g.GV()
    Vo, Fo, _ = oracle.marching_cubes(S, GV, g.side[0], g.side[1], g.side[2], 0.0)
    assert np.array_equal(SF, Fo) and np.array_equal(SV, Vo)
    h = (g.corners()[3] - g.corners()[0]) / (g.side[0] - 1)
    d = np.sqrt(oracle.PortAABB(V, F).squared_distance(SV)[0])
    assert SF.shape[0] > 10_000 and d.max() < h
    # closed surface: every edge shared by exactly two triangles
    e = np.sort(np.concatenate([SF[:, [0, 1]], SF[:, [1, 2]], SF[:, [2, 0]]]), axis=1)
    _, cnt = np.unique(e, axis=0, return_counts=True)
    assert (cnt == 2).all()
    m.close(); g.close()


# ---- exact winding number: SIGNED_DISTANCE_TYPE_WINDING_NUMBER / _DEFAULT, igl::winding_number ----
def test_exact_winding_number(gpu_api):
    cases.case_winding_number(gpu_api, 1500)
    # a 320k-face closed knot: integers inside/outside to 1e-9, and the face-by-face sum on a sample
    V, F = synth.torus_knot(*synth.knot_dims(320_000))
    m = gpu_api.Mesh(V, F, 2, winding_number=True)
    P = synth.grid_points(*synth.bbox(V), 64, 1.1)
    W = m.winding_number(P)
    assert np.abs(W - np.round(W)).max() < 1e-9 and set(np.unique(np.round(W)).tolist()) <= {0.0, 1.0}
    k = np.arange(0, P.shape[0], 997)
    assert np.abs(W[k] - oracle.winding_number(V, F, P[k])).max() < 1e-10
    # the fast winding number agrees with it to its own truncation error
    Wf = m.fast_winding_number(P)
    assert np.abs(Wf - W).max() < 2e-2
    m.close()


def test_signed_distance_winding_number_and_default(gpu_api):
    V, F = synth.icosphere(4)
    P = synth.grid_points(*synth.bbox(V), 24, 1.4)
    cases.case_signed_distance_exact_winding(gpu_api, V, F, P, "ico4")
    cases.case_signed_distance_exact_winding(gpu_api, V, F[V[F].mean(1)[:, 2] > 0.0], P, "open hemisphere")
    cases.case_signed_distance_exact_winding(gpu_api, V, F, P, "ico4 bounded", 0.1, 0.4)
    V, F = synth.torus_knot(300, 16)
    cases.case_signed_distance_exact_winding(gpu_api, V, F, synth.grid_points(*synth.bbox(V), 20, 1.1), "knot")


def test_marching_cubes_sparse(gpu_api):
    cases.case_marching_cubes_sparse(gpu_api)


def test_sharded_device_path_equals_the_single_call(gpu_api):
    """iglb200_shard_begin/run/finish (what bench.py's `value` times, here with one shard) == the single-call device entry
    points, bit for bit, for signed distance (FWN, UNSIGNED, bounded), squared distance and the winding number alone."""
    import ctypes as C
    import torch
    api, L = gpu_api, gpu_api.lib
    V, F = synth.torus_knot(600, 24)
    P = synth.grid_points(*synth.bbox(V), 64, 1.1)
    n = P.shape[0]
    dev = torch.device("cuda", 0)
    dP = torch.from_numpy(np.ascontiguousarray(P.T)).to(dev)
    m = api.Mesh(V, F, 2)
    scratch = torch.empty(int(L.iglb200_shard_scratch_bytes(n)), dtype=torch.uint8, device=dev)
    qs = torch.empty(int(L.iglb200_query_scratch_bytes(n)), dtype=torch.uint8, device=dev)
    tiny, huge = float(np.finfo(float).tiny), float(np.finfo(float).max)
    sup = int(L.iglb200_shard_super_rows(n, 1))
    vp = C.c_void_p
    for mode, sign, lo, hi in ((0, 4, tiny, huge), (0, 4, 0.02, 0.3), (0, 3, tiny, huge), (1, 0, 0.0, 0.0), (2, 0, 0.0, 0.0)):
        out = [torch.full((n,), -5.0, dtype=torch.float64, device=dev), torch.full((n,), -5, dtype=torch.int32, device=dev),
               torch.full((3, n), -5.0, dtype=torch.float64, device=dev), torch.full((n,), -5.0, dtype=torch.float64, device=dev)]
        ref = [torch.empty_like(t) for t in out]
        po, pr = [vp(t.data_ptr()) for t in out], [vp(t.data_ptr()) for t in ref]
        api.check(L.iglb200_shard_begin_dev(m.h, vp(dP.data_ptr()), n, vp(scratch.data_ptr()), 1, sup, None))
        api.check(L.iglb200_shard_run_dev(m.h, vp(dP.data_ptr()), n, vp(scratch.data_ptr()), mode, sign, lo, hi, 2.0, 1, *po, sup, 0, 0, 1, None))
        api.check(L.iglb200_shard_finish_dev(m.h, n, vp(scratch.data_ptr()), mode, sign, lo, hi, 1, *po, None))
        if mode == 0:
            api.check(L.iglb200_signed_distance_dev(m.h, vp(dP.data_ptr()), n, sign, lo, hi, *pr, None, vp(qs.data_ptr()), 1, None))
        elif mode == 1:
            api.check(L.iglb200_squared_distance_dev(m.h, vp(dP.data_ptr()), n, pr[0], pr[1], pr[2], vp(qs.data_ptr()), 1, None))
        else:
            api.check(L.iglb200_fwn_eval_dev(m.h, vp(dP.data_ptr()), n, 2.0, pr[3], vp(qs.data_ptr()), 1, None))
        torch.cuda.synchronize()
        api.check(L.iglb200_poll_device_error())
        eq = lambda a, b: bool(((a == b) | (torch.isnan(a) & torch.isnan(b))).all())
        if mode != 2:
            assert eq(out[0], ref[0]) and torch.equal(out[1], ref[1]) and eq(out[2], ref[2]), (mode, sign)
        if mode != 1:
            assert eq(out[3], ref[3]), (mode, sign)
    m.close()


def test_device_group_on_gpu(gpu_api):
    """iglb200_init(n) with every GPU of the box (one on the driver's test box, eight under `gpurun --gpus 8`): meshes are
    replicated, host queries dealt to all devices, results identical to single-device operation."""
    api = gpu_api
    V, F = synth.torus_knot(600, 24)
    P = synth.grid_points(*synth.bbox(V), 72, 1.1)
    api.init(1)
    m1 = api.Mesh(V, F, 2)
    ref = m1.signed_distance(P, return_W=True)
    m1.close()
    n = min(api.device_count(), 8)
    try:
        api.init(n)
        g = api.group_info()
        assert g["devices"] == n
        m = api.Mesh(V, F, 2)
        assert m.info()["n_devices"] == n
        out = m.signed_distance(np.asfortranarray(P), return_W=True)
        for a, b in zip(out, ref):
            assert np.array_equal(a, b, equal_nan=True)
        D, I, Cc = m.squared_distance(P.astype(np.float32))
        assert np.array_equal(I, m.squared_distance(P.astype(np.float32).astype(np.float64))[1])
        m.close()
    finally:
        api.init(1)
