"""libigl_b200 — B200-native drop-in for libigl's batched point-vs-mesh query
path (igl::signed_distance with SIGNED_DISTANCE_TYPE_FAST_WINDING_NUMBER =
AABB::squared_distance + fast_winding_number), CUDA sm_100a only.

The compute lives in libigl_b200/libiglb200.so (C ABI: include/iglb200.h), built
in-tree by `python -m libigl_b200.build`.  There is no CPU path: importing works
on any machine (so the ABI can be inspected), but every compute call fails with
IglB200Error when the library or a CUDA device is missing.
"""
from __future__ import annotations

import ctypes
import os

from .api import (  # noqa: F401
    SIGNED_DISTANCE_TYPE_DEFAULT,
    SIGNED_DISTANCE_TYPE_FAST_WINDING_NUMBER,
    SIGNED_DISTANCE_TYPE_PSEUDONORMAL,
    SIGNED_DISTANCE_TYPE_UNSIGNED,
    SIGNED_DISTANCE_TYPE_WINDING_NUMBER,
    AABB as _AABB,
    Api,
    Cloud,
    IglB200Error,
    Mesh,
)

# IGLB200_LIBRARY lets a tuning run point at another build of the SAME CUDA library
LIB_PATH = os.environ.get("IGLB200_LIBRARY") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libiglb200.so")
_api = None


def lib() -> Api:
    """The bound CUDA library.  Raises (never falls back) if it has not been built."""
    global _api
    if _api is None:
        if not os.path.exists(LIB_PATH):
            raise IglB200Error(-2, f"{LIB_PATH} is missing: run `python -m libigl_b200.build` (needs nvcc); "
                                   "there is no CPU fallback")
        _api = Api(ctypes.CDLL(LIB_PATH))
    return _api


def signed_distance(P, V, F, sign_type=SIGNED_DISTANCE_TYPE_FAST_WINDING_NUMBER, lower_bound=None,
                    upper_bound=None, return_W=False):
    return lib().signed_distance(P, V, F, sign_type, lower_bound, upper_bound, return_W)


def point_mesh_squared_distance(P, V, Ele):
    return lib().point_mesh_squared_distance(P, V, Ele)


def hausdorff(VA, FA, VB, FB):
    return lib().hausdorff(VA, FA, VB, FB)


def fast_winding_number(*args, **kw):
    return lib().fast_winding_number(*args, **kw)


def mesh(V, F, order=2) -> Mesh:
    return lib().Mesh(V, F, order)


def cloud(P, N, A, order=2) -> Cloud:
    return lib().Cloud(P, N, A, order)


def winding_number(V, F, O):
    return lib().winding_number(V, F, O)


def voxel_grid(*args):
    return lib().voxel_grid(*args)


def flood_fill(res, S):
    return lib().flood_fill(res, S)


def marching_cubes(S, GV, nx, ny, nz, isovalue=0.0, return_E2V=False):
    return lib().marching_cubes(S, GV, nx, ny, nz, isovalue, return_E2V)


def offset_surface(V, F, isolevel, s, sign_type=SIGNED_DISTANCE_TYPE_FAST_WINDING_NUMBER):
    return lib().offset_surface(V, F, isolevel, s, sign_type)


def AABB():
    return _AABB(lib())
