"""CPU tests of the KERNEL LOGIC: the .cu sources compiled for the host
(tests/emul.py) against the oracle.  Not the product path — see test_gpu_parity.py."""
import numpy as np

import cases
from libigl_b200 import synth


def test_known_answers(emul_api):
    cases.case_known_answers(emul_api)


def test_tree_topology_and_moments_match_reference(emul_api):
    ms = cases.meshes_small()
    ms["axis"] = cases.axis_aligned_box_mesh(8)
    ms["ico5"] = synth.icosphere(5)
    ms["knot40k"] = synth.torus_knot(1000, 20)
    cases.case_tree_exact(emul_api, ms)


def test_tiny_meshes_every_size(emul_api):
    """1..40 faces: every small-range regime of the builder (n<=4, <=6 exhaustive, <=32 sweep, >32 binned)."""
    V, F = synth.icosphere(2)
    rng = np.random.default_rng(0)
    ms = {f"n{n}": (V, F[rng.permutation(F.shape[0])[:n]]) for n in list(range(1, 41)) + [64, 65, 100]}
    cases.case_tree_exact(emul_api, ms)


def test_signed_distance_c1_small(emul_api):
    V, F = synth.icosphere(3)
    P = synth.grid_points(*synth.bbox(V), 24, 1.5)
    r = cases.case_signed_distance(emul_api, V, F, P, "ico3 grid")
    assert r["nan"] == 0


def test_signed_distance_knot(emul_api):
    V, F = synth.torus_knot(300, 16)
    P = synth.grid_points(*synth.bbox(V), 20, 1.1)
    cases.case_signed_distance(emul_api, V, F, P, "knot grid")


def test_closest_vs_brute(emul_api):
    for name, (V, F) in cases.meshes_small().items():
        P = synth.random_points(1500, -2.0, 2.0, 1)
        cases.case_closest_vs_brute(emul_api, V, F, P, name)


def test_fwn(emul_api):
    V, F = synth.noisy_sphere(3, 0.02, 7)
    cases.case_fwn(emul_api, V, F, synth.random_points(4000, -1.5, 1.5, 2), "noisy")


def test_bounds_and_sentinels(emul_api):
    cases.case_bounds_and_sentinels(emul_api)


def test_layouts(emul_api):
    cases.case_layouts(emul_api)


def test_degenerate(emul_api):
    cases.case_degenerate(emul_api)


def test_arena_roundtrip(emul_api):
    cases.case_arena_roundtrip(emul_api)


def test_adversarial_triangles(emul_api):
    cases.case_adversarial_triangles(emul_api)


def test_cloud_golden(emul_api):
    cases.case_cloud_golden(emul_api)


def test_cloud_vs_oracle(emul_api):
    cases.case_cloud(emul_api, 3000, 2000)


def test_golden_fixtures(emul_api):
    cases.case_golden_fixtures(emul_api)


def test_builder_corner_paths(emul_api):
    cases.case_builder_corner_paths(emul_api)


def test_pseudonormal_sign(emul_api):
    cases.case_pseudonormal(emul_api, cases.meshes_small())


def test_segments_points_hausdorff(emul_api):
    cases.case_segments_points_hausdorff(emul_api)


def test_voxel_grid(emul_api):
    cases.case_voxel_grid(emul_api)


def test_flood_fill(emul_api):
    assert cases.case_flood_fill(emul_api) > 90


def test_marching_cubes(emul_api):
    cases.case_marching_cubes(emul_api)


def test_offset_surface(emul_api):
    from libigl_b200.api import SIGNED_DISTANCE_TYPE_FAST_WINDING_NUMBER as FWN, SIGNED_DISTANCE_TYPE_PSEUDONORMAL as PN
    V, F = synth.icosphere(2)
    r = cases.case_offset_surface(emul_api, V, F, 0.2, 20, FWN, "ico2 +0.2 fwn")
    assert r["nF"] > 500
    cases.case_offset_surface(emul_api, V, F, 0.15, 18, PN, "ico2 +0.15 pseudonormal")
    V, F = synth.torus_knot(120, 10)
    cases.case_offset_surface(emul_api, V, F, 0.05, 28, FWN, "knot +0.05 fwn")


def test_exact_winding_number(emul_api):
    cases.case_winding_number(emul_api, 400)


def test_signed_distance_winding_number_and_default(emul_api):
    V, F = synth.icosphere(3)
    P = synth.grid_points(*synth.bbox(V), 12, 1.4)
    cases.case_signed_distance_exact_winding(emul_api, V, F, P, "ico3")
    cases.case_signed_distance_exact_winding(emul_api, V, F[V[F].mean(1)[:, 2] > 0.0], P, "open hemisphere")
    cases.case_signed_distance_exact_winding(emul_api, V, F, P, "ico3 bounded", 0.1, 0.4)


def test_marching_cubes_sparse(emul_api):
    cases.case_marching_cubes_sparse(emul_api)


def test_closest_extreme_scales(emul_api):
    cases.case_closest_extreme_scales(emul_api)


def test_iterative_closest_point(emul_api):
    cases.case_icp(emul_api)
