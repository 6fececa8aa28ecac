"""Host-side mirror of the reference interface for the point-vs-mesh query path,
on top of the C ABI in include/iglb200.h (ctypes; no torch, no CPU compute).

Names and argument meaning follow libigl (paths under /root/reference/include/igl):
  signed_distance                signed_distance.h:61-89
  signed_distance_fast_winding_number  signed_distance.h:283-312
  AABB.init / AABB.squared_distance    AABB.h:494-496, 722-728
  point_mesh_squared_distance    point_mesh_squared_distance.h:54-60
  fast_winding_number (soup)     fast_winding_number.h:160-206
  fast_winding_number (cloud)    fast_winding_number.h:44-53, 88-131
Outputs are returned instead of being written through reference arguments, as
the libigl python bindings do.  numpy arrays are passed by pointer with their
own strides (row- or column-major, float32 or float64) — nothing is copied on
the host unless the dtype is unsupported.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

SIGNED_DISTANCE_TYPE_PSEUDONORMAL = 0
SIGNED_DISTANCE_TYPE_WINDING_NUMBER = 1
SIGNED_DISTANCE_TYPE_DEFAULT = 2
SIGNED_DISTANCE_TYPE_UNSIGNED = 3
SIGNED_DISTANCE_TYPE_FAST_WINDING_NUMBER = 4

_vp = C.c_void_p
_i64 = C.c_int64
_int = C.c_int
_dbl = C.c_double


class IglB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"iglb200 error {code}: {msg}")
        self.code = code


def _mat(a, name):
    """(array kept alive, pointer, is_f64, rows, row_stride, col_stride) of an n x 3 matrix."""
    a = np.asarray(a)
    if a.ndim != 2 or a.shape[1] != 3:
        raise ValueError(f"{name} must be #x3, got {a.shape}")
    if a.dtype not in (np.float64, np.float32):
        a = a.astype(np.float64)
    es = a.dtype.itemsize
    if a.shape[0] > 0 and (a.strides[0] <= 0 or a.strides[1] <= 0 or a.strides[0] % es or a.strides[1] % es):
        a = np.ascontiguousarray(a)
    return a, a.ctypes.data_as(_vp), int(a.dtype == np.float64), a.shape[0], a.strides[0] // es, a.strides[1] // es


class Api:
    """Bound C ABI.  The product instance is created by libigl_b200/__init__.py from
    libigl_b200/libiglb200.so; nothing here knows about any other library."""

    SYMBOLS = [
        "iglb200_last_error", "iglb200_version", "iglb200_device_count", "iglb200_set_device", "iglb200_init", "iglb200_shutdown",
        "iglb200_group_info", "iglb200_mesh_replicate", "iglb200_signed_distance_coop", "iglb200_host_chunk_rows",
        "iglb200_shard_scratch_bytes", "iglb200_shard_super_rows", "iglb200_shard_begin_dev", "iglb200_shard_run_dev",
        "iglb200_shard_finish_dev", "iglb200_ipc_alloc", "iglb200_ipc_open", "iglb200_ipc_close", "iglb200_ipc_free",
        "iglb200_host_alloc", "iglb200_host_free", "iglb200_host_register", "iglb200_host_unregister",
        "iglb200_mesh_create", "iglb200_mesh_create_ex", "iglb200_mesh_destroy", "iglb200_mesh_info", "iglb200_mesh_build_ms",
        "iglb200_mesh_arena", "iglb200_mesh_arena_download", "iglb200_mesh_from_arena",
        "iglb200_signed_distance", "iglb200_squared_distance", "iglb200_squared_distance_bounded", "iglb200_fwn_eval",
        "iglb200_query_scratch_bytes", "iglb200_signed_distance_dev", "iglb200_squared_distance_dev",
        "iglb200_fwn_eval_dev", "iglb200_last_launch_count", "iglb200_profile_enable", "iglb200_profile_collect",
        "iglb200_debug_visit_counts", "iglb200_poll_device_error", "iglb200_debug_fp32_peak_tflops",
        "iglb200_cloud_create", "iglb200_cloud_destroy", "iglb200_cloud_info", "iglb200_cloud_eval",
        "iglb200_cloud_eval_dev",
        "iglb200_winding_number", "iglb200_winding_number_dev", "iglb200_iterative_closest_point",
        "iglb200_grid_create", "iglb200_grid_create_from", "iglb200_grid_destroy", "iglb200_grid_info", "iglb200_grid_device",
        "iglb200_grid_get_gv", "iglb200_grid_set_s", "iglb200_grid_get_s", "iglb200_grid_signed_distance",
        "iglb200_grid_flood_fill", "iglb200_grid_marching_cubes", "iglb200_surface_info", "iglb200_surface_get",
        "iglb200_surface_device", "iglb200_surface_destroy", "iglb200_flood_fill", "iglb200_marching_cubes",
        "iglb200_marching_cubes_sparse",
    ]

    def __init__(self, lib: C.CDLL):
        self.lib = L = lib
        for s in self.SYMBOLS:
            getattr(L, s)  # AttributeError if the library does not export it
        L.iglb200_last_error.restype = C.c_char_p
        L.iglb200_version.restype = C.c_char_p
        L.iglb200_host_alloc.argtypes = [C.POINTER(_vp), C.c_uint64]
        L.iglb200_host_free.argtypes = [_vp]
        L.iglb200_host_register.argtypes = [_vp, C.c_uint64]
        L.iglb200_host_unregister.argtypes = [_vp]
        L.iglb200_mesh_create.argtypes = [_vp, _int, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _int, C.POINTER(_vp)]
        L.iglb200_mesh_create_ex.argtypes = [_vp, _int, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _int, _int, C.POINTER(_vp)]
        L.iglb200_mesh_destroy.argtypes = [_vp]
        L.iglb200_mesh_info.argtypes = [_vp, C.POINTER(_i64)]
        L.iglb200_mesh_build_ms.argtypes = [_vp]
        L.iglb200_mesh_build_ms.restype = _dbl
        L.iglb200_mesh_arena.argtypes = [_vp, C.POINTER(_vp), C.POINTER(C.c_uint64)]
        L.iglb200_mesh_arena_download.argtypes = [_vp, _vp, C.c_uint64]
        L.iglb200_mesh_from_arena.argtypes = [_vp, C.c_uint64, _int, C.POINTER(_vp)]
        L.iglb200_signed_distance.argtypes = [_vp, _vp, _int, _i64, _i64, _i64, _int, _dbl, _dbl,
                                              _vp, _int, _vp, _vp, _int, _i64, _i64, _vp, _int, _vp, _int, _i64, _i64]
        L.iglb200_signed_distance_coop.argtypes = [_vp, _vp, _int, _i64, _i64, _i64, _int, _dbl, _dbl,
                                                   _vp, _int, _vp, _vp, _int, _i64, _i64, _vp, _int, _vp, _int, _i64, _i64, _vp, _i64]
        L.iglb200_shard_scratch_bytes.argtypes = [_i64]
        L.iglb200_shard_scratch_bytes.restype = C.c_uint64
        L.iglb200_shard_super_rows.argtypes = [_i64, _int]
        L.iglb200_shard_super_rows.restype = _i64
        L.iglb200_shard_begin_dev.argtypes = [_vp, _vp, _i64, _vp, _int, _i64, _vp]
        L.iglb200_shard_run_dev.argtypes = [_vp, _vp, _i64, _vp, _int, _int, _dbl, _dbl, C.c_float, _int, _vp, _vp, _vp, _vp, _i64, _i64, _int, _int, _vp]
        L.iglb200_shard_finish_dev.argtypes = [_vp, _i64, _vp, _int, _int, _dbl, _dbl, _int, _vp, _vp, _vp, _vp, _vp]
        L.iglb200_ipc_alloc.argtypes = [C.c_uint64, C.POINTER(_vp), C.c_char_p]
        L.iglb200_ipc_open.argtypes = [C.c_char_p, C.POINTER(_vp)]
        L.iglb200_ipc_close.argtypes = [_vp]
        L.iglb200_ipc_free.argtypes = [_vp]
        L.iglb200_host_chunk_rows.argtypes = [_i64, _int]
        L.iglb200_host_chunk_rows.restype = _i64
        L.iglb200_group_info.argtypes = [C.POINTER(_i64), C.c_char_p, C.c_uint64]
        L.iglb200_mesh_replicate.argtypes = [_vp]
        L.iglb200_squared_distance.argtypes = [_vp, _vp, _int, _i64, _i64, _i64, _vp, _int, _vp, _vp, _int, _i64, _i64]
        L.iglb200_squared_distance_bounded.argtypes = [_vp, _vp, _int, _i64, _i64, _i64, _dbl, _dbl, _vp, _int, _vp, _vp, _int, _i64, _i64]
        L.iglb200_fwn_eval.argtypes = [_vp, _vp, _int, _i64, _i64, _i64, C.c_float, _vp, _int]
        L.iglb200_query_scratch_bytes.argtypes = [_i64]
        L.iglb200_query_scratch_bytes.restype = C.c_uint64
        L.iglb200_signed_distance_dev.argtypes = [_vp, _vp, _i64, _int, _dbl, _dbl, _vp, _vp, _vp, _vp, _vp, _vp, _int, _vp]
        L.iglb200_squared_distance_dev.argtypes = [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _int, _vp]
        L.iglb200_fwn_eval_dev.argtypes = [_vp, _vp, _i64, C.c_float, _vp, _vp, _int, _vp]
        L.iglb200_debug_visit_counts.argtypes = [_vp, _vp, _i64, C.POINTER(_dbl)]
        L.iglb200_debug_fp32_peak_tflops.restype = _dbl
        L.iglb200_profile_enable.argtypes = [_int]
        L.iglb200_profile_enable.restype = None
        L.iglb200_profile_collect.argtypes = [C.c_char_p, C.c_uint64]
        L.iglb200_cloud_create.argtypes = [_vp, _vp, _vp, _int, _i64, _int, C.POINTER(_vp)]
        L.iglb200_cloud_destroy.argtypes = [_vp]
        L.iglb200_cloud_info.argtypes = [_vp, C.POINTER(_i64)]
        L.iglb200_cloud_eval.argtypes = [_vp, _vp, _int, _i64, _i64, _i64, _dbl, _vp, _int]
        L.iglb200_cloud_eval_dev.argtypes = [_vp, _vp, _i64, _dbl, _vp, _vp]
        L.iglb200_winding_number.argtypes = [_vp, _vp, _int, _i64, _i64, _i64, _vp, _int]
        L.iglb200_winding_number_dev.argtypes = [_vp, _vp, _i64, _vp, _vp, _int, _vp]
        L.iglb200_iterative_closest_point.argtypes = [_vp, _vp, _int, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _int, _int, C.c_uint64, _vp, _vp]
        L.iglb200_grid_create.argtypes = [_vp, _vp, _int, _int, C.POINTER(_vp)]
        L.iglb200_grid_create_from.argtypes = [_vp, _i64, _i64, _int, _int, _int, C.POINTER(_vp)]
        L.iglb200_grid_destroy.argtypes = [_vp]
        L.iglb200_grid_info.argtypes = [_vp, _vp, C.POINTER(_i64), _vp]
        L.iglb200_grid_device.argtypes = [_vp, C.POINTER(_vp), C.POINTER(_vp)]
        L.iglb200_grid_get_gv.argtypes = [_vp, _vp, _i64, _i64]
        L.iglb200_grid_set_s.argtypes = [_vp, _vp]
        L.iglb200_grid_get_s.argtypes = [_vp, _vp]
        L.iglb200_grid_signed_distance.argtypes = [_vp, _vp, _int, _dbl, _dbl]
        L.iglb200_grid_flood_fill.argtypes = [_vp]
        L.iglb200_grid_marching_cubes.argtypes = [_vp, _dbl, _dbl, C.POINTER(_vp)]
        L.iglb200_surface_info.argtypes = [_vp, C.POINTER(_i64), C.POINTER(_i64)]
        L.iglb200_surface_get.argtypes = [_vp, _vp, _vp, _vp]
        L.iglb200_surface_device.argtypes = [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]
        L.iglb200_surface_destroy.argtypes = [_vp]
        L.iglb200_flood_fill.argtypes = [_vp, _vp]
        L.iglb200_marching_cubes.argtypes = [_vp, _vp, _i64, _i64, _int, _int, _int, _dbl, C.POINTER(_vp)]
        L.iglb200_marching_cubes_sparse.argtypes = [_vp, _vp, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _dbl, C.POINTER(_vp)]

    # -- plumbing ---------------------------------------------------------
    def check(self, rc):
        if rc != 0:
            raise IglB200Error(rc, self.lib.iglb200_last_error().decode())

    def version(self):
        return self.lib.iglb200_version().decode()

    def device_count(self):
        return int(self.lib.iglb200_device_count())

    def set_device(self, d):
        self.check(self.lib.iglb200_set_device(int(d)))

    def init(self, n_gpus):
        """iglb200_init: devices 0..n_gpus-1 become the process's group (meshes are replicated, host queries dealt)."""
        self.check(self.lib.iglb200_init(int(n_gpus)))

    def group_info(self):
        info = (_i64 * 4)()
        buf = C.create_string_buffer(128)
        self.check(self.lib.iglb200_group_info(info, buf, 128))
        return {"devices": int(info[0]), "nccl": bool(info[1]), "peer": bool(info[2]), "replicate_ms": info[3] / 1000.0,
                "transport": buf.value.decode()}

    def last_launch_count(self):
        return int(self.lib.iglb200_last_launch_count())

    def profile_enable(self, on=True):
        self.lib.iglb200_profile_enable(int(on))

    def profile_collect(self):
        """{kernel group: milliseconds} accumulated since the previous collect."""
        buf = C.create_string_buffer(4096)
        self.lib.iglb200_profile_collect(buf, 4096)
        out = {}
        for kv in buf.value.decode().split(";"):
            if "=" in kv:
                k, v = kv.split("=")
                out[k] = float(v)
        return out

    # -- handles ----------------------------------------------------------
    def Mesh(self, V, F, order=2, pseudonormals=False, winding_number=False):
        return Mesh(self, V, F, order, pseudonormals=pseudonormals, winding_number=winding_number)

    def Cloud(self, P, N, A, order=2):
        return Cloud(self, P, N, A, order)

    def Grid(self, box_min=None, box_max=None, s=None, pad_count=0, GV=None, side=None):
        return Grid(self, box_min, box_max, s, pad_count, GV=GV, side=side)

    # -- reference-named free functions -------------------------------------
    def voxel_grid(self, V_or_box, offset_or_s, s_or_pad, pad_count=None):
        """igl::voxel_grid: (V, offset, s, pad_count) or ((box_min, box_max), s, pad_count) -> GV, side
        (voxel_grid.cpp:68-88 / :14-66)."""
        if pad_count is None:
            (bmin, bmax), s, pad = V_or_box, offset_or_s, s_or_pad
        else:
            V = np.asarray(V_or_box, np.float64)
            bmin, bmax, s, pad = V.min(0) - offset_or_s, V.max(0) + offset_or_s, s_or_pad, pad_count
        g = Grid(self, bmin, bmax, s, pad)
        try:
            return g.GV(), g.side
        finally:
            g.close()

    def flood_fill(self, res, S):
        """igl::flood_fill(res, S): returns the filled copy (flood_fill.cpp:11-98)."""
        res = np.ascontiguousarray(res, dtype=np.int32)
        S = np.array(S, dtype=np.float64, copy=True).reshape(-1)
        if S.size != int(np.prod(res.astype(np.int64))):
            raise ValueError("S must hold prod(res) values")
        self.check(self.lib.iglb200_flood_fill(res.ctypes.data_as(_vp), S.ctypes.data_as(_vp)))
        return S

    def marching_cubes(self, S, GV, nx, ny, nz, isovalue=0.0, return_E2V=False):
        """igl::marching_cubes(S, GV, nx, ny, nz, isovalue, V, F[, E2V]) (marching_cubes.cpp:19-112)."""
        S = np.ascontiguousarray(S, dtype=np.float64).reshape(-1)
        GV = np.asarray(GV, np.float64)
        if GV.ndim != 2 or GV.shape[1] != 3 or GV.shape[0] != S.size or S.size != int(nx) * int(ny) * int(nz):
            raise ValueError("GV must be (nx*ny*nz) x 3 and S nx*ny*nz")
        if not (GV.flags.c_contiguous or GV.flags.f_contiguous):
            GV = np.ascontiguousarray(GV)
        h = _vp()
        self.check(self.lib.iglb200_marching_cubes(S.ctypes.data_as(_vp), GV.ctypes.data_as(_vp), GV.strides[0] // 8,
                                                   GV.strides[1] // 8, int(nx), int(ny), int(nz), float(isovalue), C.byref(h)))
        sf = Surface(self, h)
        try:
            return sf.get(return_E2V)
        finally:
            sf.close()

    def marching_cubes_sparse(self, S, GV, GI, isovalue=0.0, return_E2V=False):
        """igl::marching_cubes(S, GV, GI, isovalue, V, F[, E2V]): cubes listed by their 8 corner indices
        (marching_cubes.cpp:114-176)."""
        S = np.ascontiguousarray(S, dtype=np.float64).reshape(-1)
        GV = np.asarray(GV, np.float64)
        GI = np.ascontiguousarray(GI, dtype=np.int32)
        if GV.ndim != 2 or GV.shape[1] != 3 or GV.shape[0] != S.size or GI.ndim != 2 or (GI.shape[0] and GI.shape[1] != 8):
            raise ValueError("GV must be #GV x 3, S #GV, GI #cubes x 8")
        if not (GV.flags.c_contiguous or GV.flags.f_contiguous):
            GV = np.ascontiguousarray(GV)
        h = _vp()
        self.check(self.lib.iglb200_marching_cubes_sparse(S.ctypes.data_as(_vp), GV.ctypes.data_as(_vp), S.size,
                                                          GV.strides[0] // 8 if S.size else 3, GV.strides[1] // 8 if S.size else 1,
                                                          GI.ctypes.data_as(_vp), GI.shape[0], 8, 1, float(isovalue), C.byref(h)))
        sf = Surface(self, h)
        try:
            return sf.get(return_E2V)
        finally:
            sf.close()

    def offset_surface(self, V, F, isolevel, s, sign_type=SIGNED_DISTANCE_TYPE_FAST_WINDING_NUMBER):
        """igl::offset_surface(V, F, isolevel, s, signed_distance_type, SV, SF, GV, side, S)
        (offset_surface.cpp:18-48); the grid never leaves the device between the stages."""
        V = np.asarray(V, np.float64)
        pseudo = int(sign_type) == SIGNED_DISTANCE_TYPE_PSEUDONORMAL
        exact = int(sign_type) in (SIGNED_DISTANCE_TYPE_WINDING_NUMBER, SIGNED_DISTANCE_TYPE_DEFAULT)
        g = Grid(self, V.min(0) - isolevel, V.max(0) + isolevel, s, 1)     # voxel_grid(V, isolevel, s, 1, GV, side)
        m = Mesh(self, V, F, 2, pseudonormals=pseudo, winding_number=exact)
        try:
            c = g.corners()
            h = (c[3] - c[0]) / float(g.side[0] - 1)                        # :31-32
            g.signed_distance(m, sign_type, isolevel - np.sqrt(3.0) * h, isolevel + np.sqrt(3.0) * h)
            g.flood_fill()
            SV, SF = g.marching_cubes(0.0, shift=isolevel)
            return SV, SF, g.GV(), g.side, g.S()
        finally:
            m.close()
            g.close()

    def signed_distance(self, P, V, F, sign_type=SIGNED_DISTANCE_TYPE_FAST_WINDING_NUMBER,
                        lower_bound=None, upper_bound=None, return_W=False):
        """igl::signed_distance: returns S, I, C (+ N for PSEUDONORMAL, the only sign type that
        writes it in the reference; + W if asked)."""
        F = np.asarray(F)
        pseudo = int(sign_type) == SIGNED_DISTANCE_TYPE_PSEUDONORMAL
        if F.shape[0] == 0:  # signed_distance.cpp:398-405
            n = np.asarray(P).shape[0]
            out = (np.full(n, np.nan), np.full(n, -1, np.int32), np.full((n, 3), np.nan))
            if pseudo:
                out = out + (np.full((n, 3), np.nan),)
            return out + (np.full(n, np.nan),) if return_W else out
        exact = int(sign_type) in (SIGNED_DISTANCE_TYPE_WINDING_NUMBER, SIGNED_DISTANCE_TYPE_DEFAULT)
        m = Mesh(self, V, F, 2, pseudonormals=pseudo, winding_number=exact)
        try:
            return m.signed_distance(P, sign_type, lower_bound, upper_bound, return_W)
        finally:
            m.close()

    def point_mesh_squared_distance(self, P, V, Ele):
        m = Mesh(self, V, Ele, 0)
        try:
            return m.squared_distance(P)
        finally:
            m.close()

    def winding_number(self, V, F, O):
        """igl::winding_number(V,F,O,W) for triangle meshes (winding_number.cpp:17-52): exact, double."""
        m = Mesh(self, V, F, 2, winding_number=True)
        try:
            return m.winding_number(O)
        finally:
            m.close()

    def hausdorff(self, VA, FA, VB, FB):
        """igl::hausdorff (hausdorff.cpp:20-39): vertex-to-surface distances both ways, on the device."""
        dBA = self.point_mesh_squared_distance(VA, VB, FB)[0]
        dAB = self.point_mesh_squared_distance(VB, VA, FA)[0]
        return float(np.sqrt(max(dBA.max(), dAB.max())))

    def fast_winding_number(self, *args, **kw):
        """Overloads of igl::fast_winding_number:
        (V, F, Q) -> W                         soup one-shot (fast_winding_number.cpp:367-378)
        (P, N, A, Q[, order, beta]) -> WN      point cloud (fast_winding_number.cpp:320-365)"""
        if len(args) == 3:
            V, F, Q = args
            m = Mesh(self, V, F, 2)
            try:
                return m.fast_winding_number(Q, 2.0)
            finally:
                m.close()
        if len(args) in (4, 6):
            P, N, A, Q = args[:4]
            order, beta = (args[4], args[5]) if len(args) == 6 else (kw.get("expansion_order", 2), kw.get("beta", 2.0))
            c = Cloud(self, P, N, A, order)
            try:
                return c.eval(Q, beta)
            finally:
                c.close()
        raise TypeError("fast_winding_number(V,F,Q) or fast_winding_number(P,N,A,Q[,order,beta])")


class Mesh:
    """AABB<MatrixXd,3> + FastWindingNumberBVH built on the device (one handle)."""

    def __init__(self, api: Api, V, F, order=2, _handle=None, pseudonormals=False, winding_number=False):
        self.api = api
        self.h = _vp()
        if _handle is not None:
            self.h = _handle
            return
        Vk, vp, vf64, nV, vrs, vcs = _mat(V, "V")
        F = np.asarray(F)
        if F.ndim == 2 and F.shape[1] in (1, 2):
            # segments / points: the reference reads corner k as Ele(i, k % cols)
            # (point_simplex_squared_distance.cpp:126-130), i.e. a degenerate triangle
            F = F[:, [0, 1 % F.shape[1], 2 % F.shape[1]]]
        if F.ndim != 2 or F.shape[1] != 3:
            raise ValueError(f"F must be #Fx3 (or #Fx2 / #Fx1 element lists), got {F.shape}")
        if F.dtype != np.int32:
            F = F.astype(np.int32)
        if F.shape[0] and (F.strides[0] <= 0 or F.strides[1] <= 0):
            F = np.ascontiguousarray(F)
        api.check(api.lib.iglb200_mesh_create_ex(vp, vf64, nV, vrs, vcs, F.ctypes.data_as(_vp), F.shape[0],
                                                 F.strides[0] // 4 if F.shape[0] else 3, F.strides[1] // 4 if F.shape[0] else 1,
                                                 int(order), (1 if pseudonormals else 0) | (2 if winding_number else 0),
                                                 C.byref(self.h)))

    def close(self):
        if self.h:
            self.api.lib.iglb200_mesh_destroy(self.h)
            self.h = _vp()

    __del__ = close

    def info(self):
        a = (_i64 * 8)()
        self.api.check(self.api.lib.iglb200_mesh_info(self.h, a))
        return dict(n_faces=a[0], n_nodes=a[1], n_levels=a[2], device_bytes=a[3], n_excluded=a[4], order=a[5],
                    pseudonormals=bool(a[6]), n_devices=int(a[7]),
                    build_ms=float(self.api.lib.iglb200_mesh_build_ms(self.h)))

    def arena(self):
        p = _vp()
        b = C.c_uint64()
        self.api.check(self.api.lib.iglb200_mesh_arena(self.h, C.byref(p), C.byref(b)))
        return p.value, b.value

    def download(self):
        """Host copy of the flat tree image, parsed into numpy views (inspection / caching)."""
        _, nbytes = self.arena()
        buf = np.empty(nbytes, np.uint8)
        self.api.check(self.api.lib.iglb200_mesh_arena_download(self.h, buf.ctypes.data_as(_vp), nbytes))
        return parse_arena(buf)

    @staticmethod
    def from_arena(api: Api, device_ptr: int, nbytes: int, copy: bool = True):
        h = _vp()
        api.check(api.lib.iglb200_mesh_from_arena(_vp(device_ptr), nbytes, int(copy), C.byref(h)))
        return Mesh(api, None, None, _handle=h)

    def visit_counts(self, P):
        """Mean nodes / faces touched: closest-point walk per warp of 32 queries, winding per query."""
        P = np.asfortranarray(np.asarray(P, np.float64))
        out = (_dbl * 4)()
        self.api.check(self.api.lib.iglb200_debug_visit_counts(self.h, P.ctypes.data_as(_vp), P.shape[0], out))
        return dict(cp_nodes_per_warp=out[0], cp_tris_per_warp=out[1], fwn_nodes=out[2], fwn_tris=out[3])

    # igl::signed_distance with prebuilt structures
    def signed_distance(self, P, sign_type=SIGNED_DISTANCE_TYPE_FAST_WINDING_NUMBER, lower_bound=None,
                        upper_bound=None, return_W=False):
        Pk, pp, pf64, n, prs, pcs = _mat(P, "P")
        if lower_bound is None:  # signed_distance.cpp:436-439
            lower_bound = np.finfo(np.float64).tiny
        if upper_bound is None:
            upper_bound = np.finfo(np.float64).max
        S = np.empty(n, np.float64)
        I = np.empty(n, np.int32)
        Cc = np.empty((n, 3), np.float64, order="F")
        W = np.empty(n, np.float64) if return_W else None
        pseudo = int(sign_type) == SIGNED_DISTANCE_TYPE_PSEUDONORMAL
        N = np.empty((n, 3), np.float64, order="F") if pseudo else None
        self.api.check(self.api.lib.iglb200_signed_distance(
            self.h, pp, pf64, n, prs, pcs, int(sign_type), float(lower_bound), float(upper_bound),
            S.ctypes.data_as(_vp), 1, I.ctypes.data_as(_vp), Cc.ctypes.data_as(_vp), 1, 1, max(n, 1),
            W.ctypes.data_as(_vp) if return_W else None, 1,
            N.ctypes.data_as(_vp) if pseudo else None, 1, 1, max(n, 1)))
        out = (S, I, Cc, N) if pseudo else (S, I, Cc)
        return out + (W,) if return_W else out

    def signed_distance_coop(self, sh, chunk_rows, sign_type=SIGNED_DISTANCE_TYPE_FAST_WINDING_NUMBER, lower_bound=None,
                             upper_bound=None, counter_index=0, want_W=True):
        """This process's share of ONE query set held in memory shared between processes (dist.SharedArrays):
        chunks of `chunk_rows` rows are taken off sh.counter until the rows run out."""
        if lower_bound is None:
            lower_bound = np.finfo(np.float64).tiny
        if upper_bound is None:
            upper_bound = np.finfo(np.float64).max
        n = sh.n
        self.api.check(self.api.lib.iglb200_signed_distance_coop(
            self.h, sh.P.ctypes.data_as(_vp), 1, n, 1, n, int(sign_type), float(lower_bound), float(upper_bound),
            sh.S.ctypes.data_as(_vp), 1, sh.I.ctypes.data_as(_vp), sh.C.ctypes.data_as(_vp), 1, 1, n,
            sh.W.ctypes.data_as(_vp) if want_W else None, 1, None, 1, 1, n,
            _vp(sh.counter.ctypes.data + 8 * int(counter_index)), int(chunk_rows)))

    # igl::signed_distance_fast_winding_number(P,V,F,tree,fwn_bvh,S): S only
    def signed_distance_fast_winding_number(self, P):
        return self.signed_distance(P)[0]

    # AABB::squared_distance(V,Ele,P,sqrD,I,C); with bounds: the batch form of AABB.h:575-614
    def squared_distance(self, P, low_sqr_d=0.0, up_sqr_d=np.inf):
        Pk, pp, pf64, n, prs, pcs = _mat(P, "P")
        D = np.empty(n, np.float64)
        I = np.empty(n, np.int32)
        Cc = np.empty((n, 3), np.float64, order="F")
        self.api.check(self.api.lib.iglb200_squared_distance_bounded(
            self.h, pp, pf64, n, prs, pcs, float(low_sqr_d), float(up_sqr_d), D.ctypes.data_as(_vp), 1, I.ctypes.data_as(_vp),
            Cc.ctypes.data_as(_vp), 1, 1, max(n, 1)))
        return D, I, Cc

    # fast_winding_number(fwn_bvh, accuracy_scale, Q, W)
    def winding_number(self, O):
        """Exact generalised winding number at the rows of O (mesh created with winding_number=True)."""
        Ok, op, of64, n, ors, ocs = _mat(O, "O")
        W = np.empty(n, np.float64)
        self.api.check(self.api.lib.iglb200_winding_number(self.h, op, of64, n, ors, ocs, W.ctypes.data_as(_vp), 1))
        return W

    def iterative_closest_point(self, VX, FX, num_samples, max_iters, seed=0):
        """igl::iterative_closest_point(VX,FX,VY,FY,num_samples,max_iters,R,t) with this mesh as Y: returns R (3x3), t (3)
        such that VX @ R + t ~ Y."""
        Vk, vp, vf64, nV, vrs, vcs = _mat(VX, "VX")
        FX = np.ascontiguousarray(FX, dtype=np.int32)
        R = np.empty((3, 3))
        t = np.empty(3)
        self.api.check(self.api.lib.iglb200_iterative_closest_point(self.h, vp, vf64, nV, vrs, vcs, FX.ctypes.data_as(_vp), FX.shape[0], 3, 1,
                                                                    int(num_samples), int(max_iters), int(seed),
                                                                    R.ctypes.data_as(_vp), t.ctypes.data_as(_vp)))
        return R, t

    def fast_winding_number(self, Q, accuracy_scale=2.0):
        Qk, qp, qf64, n, qrs, qcs = _mat(Q, "Q")
        W = np.empty(n, np.float64)
        self.api.check(self.api.lib.iglb200_fwn_eval(self.h, qp, qf64, n, qrs, qcs, float(accuracy_scale),
                                                     W.ctypes.data_as(_vp), 1))
        return W


class AABB:
    """igl::AABB<DerivedV,3> subset: init + squared_distance."""

    def __init__(self, api: Api):
        self.api = api
        self.mesh = None

    def init(self, V, Ele):
        self.mesh = Mesh(self.api, V, Ele, 0)

    def squared_distance(self, V, Ele, P):
        if self.mesh is None:
            raise IglB200Error(-1, "AABB.init was not called")
        return self.mesh.squared_distance(P)


class Cloud:
    """octree + CM/R/EC of fast_winding_number(P,N,A,...) kept on the device."""

    def __init__(self, api: Api, P, N, A, order=2):
        self.api = api
        self.h = _vp()
        P = np.ascontiguousarray(P, dtype=np.float64)
        N = np.ascontiguousarray(N, dtype=np.float64)
        A = np.ascontiguousarray(A, dtype=np.float64).reshape(-1)
        if P.ndim != 2 or P.shape[1] != 3 or N.shape != P.shape or A.shape[0] != P.shape[0]:
            raise ValueError("P, N must be #Px3 and A #P")
        api.check(api.lib.iglb200_cloud_create(P.ctypes.data_as(_vp), N.ctypes.data_as(_vp), A.ctypes.data_as(_vp), 1,
                                               P.shape[0], int(order), C.byref(self.h)))

    def close(self):
        if self.h:
            self.api.lib.iglb200_cloud_destroy(self.h)
            self.h = _vp()

    __del__ = close

    def info(self):
        a = (_i64 * 8)()
        self.api.check(self.api.lib.iglb200_cloud_info(self.h, a))
        return dict(n_points=a[0], n_cells=a[1], n_levels=a[2], device_bytes=a[3], order=a[4])

    def eval(self, Q, beta=2.0):
        Qk, qp, qf64, n, qrs, qcs = _mat(Q, "Q")
        W = np.empty(n, np.float64)
        self.api.check(self.api.lib.iglb200_cloud_eval(self.h, qp, qf64, n, qrs, qcs, float(beta),
                                                       W.ctypes.data_as(_vp), 1))
        return W


class Surface:
    """Result of marching cubes, resident on the device until fetched."""

    def __init__(self, api: Api, handle):
        self.api = api
        self.h = handle

    def close(self):
        if self.h:
            self.api.lib.iglb200_surface_destroy(self.h)
            self.h = _vp()

    __del__ = close

    def sizes(self):
        nv, nf = _i64(), _i64()
        self.api.check(self.api.lib.iglb200_surface_info(self.h, C.byref(nv), C.byref(nf)))
        return int(nv.value), int(nf.value)

    def get(self, with_E2V=False):
        nv, nf = self.sizes()
        V = np.empty((nv, 3), np.float64)
        F = np.empty((nf, 3), np.int32)
        E = np.empty((nv, 2), np.int32) if with_E2V else None
        self.api.check(self.api.lib.iglb200_surface_get(self.h, V.ctypes.data_as(_vp), F.ctypes.data_as(_vp),
                                                        E.ctypes.data_as(_vp) if with_E2V else None))
        return (V, F, E) if with_E2V else (V, F)


class Grid:
    """A voxel grid (positions GV + one value S per vertex) kept on the device:
    voxel_grid -> signed_distance -> flood_fill -> marching_cubes without a host round trip."""

    def __init__(self, api: Api, box_min=None, box_max=None, s=None, pad_count=0, GV=None, side=None):
        self.api = api
        self.h = _vp()
        if GV is not None:
            GV = np.asarray(GV, np.float64)
            if not (GV.flags.c_contiguous or GV.flags.f_contiguous):
                GV = np.ascontiguousarray(GV)
            api.check(api.lib.iglb200_grid_create_from(GV.ctypes.data_as(_vp), GV.strides[0] // 8, GV.strides[1] // 8,
                                                       int(side[0]), int(side[1]), int(side[2]), C.byref(self.h)))
        else:
            bmin = np.ascontiguousarray(box_min, dtype=np.float64)
            bmax = np.ascontiguousarray(box_max, dtype=np.float64)
            api.check(api.lib.iglb200_grid_create(bmin.ctypes.data_as(_vp), bmax.ctypes.data_as(_vp), int(s), int(pad_count),
                                                  C.byref(self.h)))
        sd = (C.c_int * 3)()
        n = _i64()
        api.check(api.lib.iglb200_grid_info(self.h, sd, C.byref(n), None))
        self.side = np.array(list(sd), np.int32)
        self.n = int(n.value)

    def close(self):
        if self.h:
            self.api.lib.iglb200_grid_destroy(self.h)
            self.h = _vp()

    __del__ = close

    def corners(self):
        c = (_dbl * 6)()
        self.api.check(self.api.lib.iglb200_grid_info(self.h, None, None, c))
        return np.array(list(c))

    def device_ptrs(self):
        gv, s = _vp(), _vp()
        self.api.check(self.api.lib.iglb200_grid_device(self.h, C.byref(gv), C.byref(s)))
        return gv.value, s.value

    def GV(self):
        out = np.empty((self.n, 3), np.float64)
        self.api.check(self.api.lib.iglb200_grid_get_gv(self.h, out.ctypes.data_as(_vp), 3, 1))
        return out

    def S(self):
        out = np.empty(self.n, np.float64)
        self.api.check(self.api.lib.iglb200_grid_get_s(self.h, out.ctypes.data_as(_vp)))
        return out

    def set_S(self, S):
        S = np.ascontiguousarray(S, dtype=np.float64).reshape(-1)
        if S.size != self.n:
            raise ValueError("S must hold one value per grid vertex")
        self.api.check(self.api.lib.iglb200_grid_set_s(self.h, S.ctypes.data_as(_vp)))

    def signed_distance(self, mesh, sign_type=SIGNED_DISTANCE_TYPE_FAST_WINDING_NUMBER, lower_bound=None, upper_bound=None):
        if lower_bound is None:
            lower_bound = np.finfo(np.float64).tiny
        if upper_bound is None:
            upper_bound = np.finfo(np.float64).max
        self.api.check(self.api.lib.iglb200_grid_signed_distance(self.h, mesh.h, int(sign_type), float(lower_bound),
                                                                 float(upper_bound)))

    def flood_fill(self):
        self.api.check(self.api.lib.iglb200_grid_flood_fill(self.h))

    def marching_cubes(self, isovalue=0.0, shift=0.0, return_E2V=False, keep=False):
        h = _vp()
        self.api.check(self.api.lib.iglb200_grid_marching_cubes(self.h, float(shift), float(isovalue), C.byref(h)))
        sf = Surface(self.api, h)
        if keep:
            return sf
        try:
            return sf.get(return_E2V)
        finally:
            sf.close()


def parse_arena(buf: np.ndarray):
    """Decode the arena layout of libigl_b200/csrc/mesh.h (ArenaHeader + four arrays)."""
    hdr = np.dtype([("magic", "<u8"), ("bytes", "<u8"), ("nF", "<i8"), ("nV", "<i8"), ("nnodes", "<i8"),
                    ("nlevels", "<i8"), ("nexcluded", "<i8"), ("order", "<i4"), ("pad", "<i4"),
                    ("bbox_lo", "<f8", 3), ("bbox_hi", "<f8", 3), ("absmax", "<f4"), ("padf", "<f4"),
                    ("off_tri64", "<u8"), ("off_tri32", "<u8"), ("off_fnodes", "<u8"), ("off_cnodes", "<u8"),
                    ("off_triq", "<u8"), ("off_pn", "<u8"), ("reserved", "<u8", 6), ("sigma", "<f8"), ("off_pnodes", "<u8")])
    h = buf[:hdr.itemsize].view(hdr)[0]
    nF, nn = int(h["nF"]), int(h["nnodes"])
    tri64 = buf[int(h["off_tri64"]):int(h["off_tri64"]) + 72 * nF].view(np.float64).reshape(nF, 3, 3)
    tri32 = buf[int(h["off_tri32"]):int(h["off_tri32"]) + 48 * nF].view(np.float32).reshape(nF, 3, 4)
    fn = buf[int(h["off_fnodes"]):int(h["off_fnodes"]) + 384 * nn]
    fchild = fn.view(np.uint32).reshape(nn, 96)[:, :4].copy()
    fdata = fn.view(np.float32).reshape(nn, 96)[:, 4:].reshape(nn, 23, 4).copy()
    pnb = buf[int(h["off_pnodes"]):int(h["off_pnodes"]) + 160 * nn]
    pill = pnb.view(np.float32).reshape(nn, 40)[:, :32].reshape(nn, 8, 4).copy()     # rows cx cy cz nx ny nz R t, 4 lanes
    cchild = pnb.view(np.uint32).reshape(nn, 40)[:, 32:36].copy()
    cbox = None
    if int(h["off_cnodes"]):
        cn = buf[int(h["off_cnodes"]):int(h["off_cnodes"]) + 128 * nn]
        cbox = cn.view(np.float32).reshape(nn, 32)[:, :24].reshape(nn, 6, 4).copy()
    pn = None
    if int(h["off_pn"]):
        pn = buf[int(h["off_pn"]):int(h["off_pn"]) + 168 * nF].view(np.float64).reshape(nF, 7, 3).copy()
    return dict(header=h, tri64=tri64, tri32=tri32, fwn_child=fchild, fwn_data=fdata, cp_box=cbox, cp_pill=pill, cp_child=cchild,
                pseudonormals=pn)
