"""TEST INFRASTRUCTURE ONLY — the checker, never the product.

ctypes front-end to
  oracle/_ref/libref_fwn.so        genuine reference soup FWN (oracle/ref_fwn.cpp)
  oracle/_build/liboracle_port.so  CPU restatement of AABB / Ericson / octree /
                                   cloud FWN (oracle/oracle_port.cpp)
and the gluing of signed_distance.cpp:79-139 on top of them.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  libigl_b200/ never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PORT = os.path.join(_HERE, "_build", "liboracle_port.so")
_REF = os.path.join(_HERE, "_ref", "libref_fwn.so")

_c_dp = C.POINTER(C.c_double)
_c_fp = C.POINTER(C.c_float)
_c_ip = C.POINTER(C.c_int)
_c_lp = C.POINTER(C.c_int64)
_c_up = C.POINTER(C.c_uint32)


def build(force: bool = False) -> None:
    """Compile the checker (port always; _ref only where /root/reference exists)."""
    need = force or not os.path.exists(_PORT) or (
        os.path.getmtime(_PORT) < os.path.getmtime(os.path.join(_HERE, "oracle_port.cpp")))
    need_ref = os.path.isdir("/root/reference/include/igl") and (
        force or not os.path.exists(_REF)
        or os.path.getmtime(_REF) < os.path.getmtime(os.path.join(_HERE, "ref_fwn.cpp")))
    if need or need_ref:
        subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)


def _dp(a):
    return a.ctypes.data_as(_c_dp)


def _fp(a):
    return a.ctypes.data_as(_c_fp)


def _ip(a):
    return a.ctypes.data_as(_c_ip)


_port = None
_ref = None


def port():
    global _port
    if _port is None:
        build()
        L = C.CDLL(_PORT)
        L.port_cp_create.restype = C.c_void_p
        L.port_cp_create.argtypes = [_c_dp, C.c_int64, _c_ip, C.c_int64]
        L.port_cp_destroy.argtypes = [C.c_void_p]
        L.port_cp_num_nodes.restype = C.c_int64
        L.port_cp_num_nodes.argtypes = [C.c_void_p]
        L.port_cp_query.argtypes = [C.c_void_p, _c_dp, C.c_int64, C.c_double, C.c_double,
                                    _c_dp, _c_ip, _c_dp, _c_lp]
        L.port_cp_brute.argtypes = [_c_dp, _c_ip, C.c_int64, _c_dp, C.c_int64, _c_dp, _c_ip, _c_dp]
        L.port_cp_face_distance.argtypes = [_c_dp, _c_ip, _c_dp, _c_ip, C.c_int64, _c_dp, _c_dp]
        L.port_pseudo_create.restype = C.c_void_p
        L.port_pseudo_create.argtypes = [_c_dp, C.c_int64, _c_ip, C.c_int64]
        L.port_pseudo_destroy.argtypes = [C.c_void_p]
        L.port_pseudo_normals.argtypes = [C.c_void_p, _c_dp, _c_dp, _c_dp]
        L.port_pseudo_test.argtypes = [C.c_void_p, _c_dp, _c_ip, _c_dp, C.c_int64, _c_dp, _c_dp]
        L.port_cloud_create.restype = C.c_void_p
        L.port_cloud_create.argtypes = [_c_dp, _c_dp, _c_dp, C.c_int64, C.c_int]
        L.port_cloud_destroy.argtypes = [C.c_void_p]
        L.port_cloud_num_cells.restype = C.c_int64
        L.port_cloud_num_cells.argtypes = [C.c_void_p]
        L.port_cloud_cells.argtypes = [C.c_void_p, _c_ip, _c_dp, _c_dp, _c_dp, _c_dp, _c_dp, _c_ip]
        L.port_cloud_eval.argtypes = [C.c_void_p, _c_dp, C.c_int64, C.c_double, _c_dp]
        _port = L
    return _port


def ref():
    global _ref
    if _ref is None:
        build()
        if not os.path.exists(_REF):
            raise RuntimeError("oracle/_ref/libref_fwn.so missing (built only where /root/reference exists)")
        L = C.CDLL(_REF)
        L.ref_fwn_create.restype = C.c_void_p
        L.ref_fwn_create.argtypes = [C.c_int, _c_ip, C.c_int, _c_fp, C.c_int]
        L.ref_fwn_destroy.argtypes = [C.c_void_p]
        L.ref_fwn_solid_angle.argtypes = [C.c_void_p, _c_fp, C.c_int64, C.c_float, _c_fp]
        L.ref_fwn_winding.argtypes = [C.c_void_p, _c_fp, C.c_int64, C.c_float, _c_dp]
        L.ref_fwn_num_nodes.argtypes = [C.c_void_p]
        L.ref_fwn_nodes.argtypes = [C.c_void_p, _c_up]
        L.ref_fwn_boxdata.argtypes = [C.c_void_p, _c_fp]
        L.ref_bvh4_build.argtypes = [_c_fp, C.c_int, _c_up, C.c_int]
        _ref = L
    return _ref


def num_threads() -> int:
    return int(port().port_threads())


# ---------------------------------------------------------------------------
class RefFwn:
    """Genuine UT_SolidAngle<float,float> (fast_winding_number.cpp:383-416)."""

    def __init__(self, V, F, order: int = 2):
        self.U = np.ascontiguousarray(V, dtype=np.float32)  # V.cast<float>(), signed_distance.cpp:66
        self.F = np.ascontiguousarray(F, dtype=np.int32)
        self.h = ref().ref_fwn_create(self.F.shape[0], _ip(self.F), self.U.shape[0], _fp(self.U), order)

    def __del__(self):
        if getattr(self, "h", None):
            ref().ref_fwn_destroy(self.h)
            self.h = None

    def winding(self, Q, accuracy_scale: float = 2.0):
        """float64 W = omega/(4 pi) per query; Q is cast to float as the reference does."""
        Qf = np.ascontiguousarray(Q, dtype=np.float32)
        W = np.empty(Qf.shape[0], dtype=np.float64)
        ref().ref_fwn_winding(self.h, _fp(Qf), Qf.shape[0], accuracy_scale, _dp(W))
        return W

    def solid_angle(self, Q, accuracy_scale: float = 2.0):
        Qf = np.ascontiguousarray(Q, dtype=np.float32)
        W = np.empty(Qf.shape[0], dtype=np.float32)
        ref().ref_fwn_solid_angle(self.h, _fp(Qf), Qf.shape[0], accuracy_scale, _fp(W))
        return W

    def nodes(self):
        n = ref().ref_fwn_num_nodes(self.h)
        out = np.empty((n, 4), dtype=np.uint32)
        ref().ref_fwn_nodes(self.h, out.ctypes.data_as(_c_up))
        return out

    def boxdata(self):
        n = ref().ref_fwn_num_nodes(self.h)
        out = np.empty((n, 23, 4), dtype=np.float32)
        ref().ref_fwn_boxdata(self.h, _fp(out))
        return out


def ref_bvh4(boxes):
    """Reference UT_BVH<4> topology for explicit boxes [n,3,2] (min,max per axis)."""
    b = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 6)
    cap = b.shape[0] + 8
    out = np.empty((cap, 4), dtype=np.uint32)
    n = ref().ref_bvh4_build(_fp(b), b.shape[0], out.ctypes.data_as(_c_up), cap)
    return out[:n].copy()


class PortAABB:
    """Restated igl::AABB<MatrixXd,3> (AABB.cpp:845-978, 1135-1218)."""

    def __init__(self, V, F):
        self.V = np.ascontiguousarray(V, dtype=np.float64)
        self.F = np.ascontiguousarray(F, dtype=np.int32)
        self.h = port().port_cp_create(_dp(self.V), self.V.shape[0], _ip(self.F), self.F.shape[0])

    def __del__(self):
        if getattr(self, "h", None):
            port().port_cp_destroy(self.h)
            self.h = None

    def num_nodes(self):
        return int(port().port_cp_num_nodes(self.h))

    def squared_distance(self, P, low=0.0, up=np.inf, count=False):
        P = np.ascontiguousarray(P, dtype=np.float64)
        n = P.shape[0]
        sqrD = np.empty(n)
        I = np.empty(n, dtype=np.int32)
        Cc = np.empty((n, 3))
        cnt = np.zeros(2, dtype=np.int64)
        port().port_cp_query(self.h, _dp(P), n, low, up, _dp(sqrD), _ip(I), _dp(Cc),
                             cnt.ctypes.data_as(_c_lp) if count else None)
        if count:
            return sqrD, I, Cc, cnt
        return sqrD, I, Cc


def brute_closest(V, F, P):
    V = np.ascontiguousarray(V, dtype=np.float64)
    F = np.ascontiguousarray(F, dtype=np.int32)
    P = np.ascontiguousarray(P, dtype=np.float64)
    n = P.shape[0]
    sqrD = np.empty(n)
    I = np.empty(n, dtype=np.int32)
    Cc = np.empty((n, 3))
    port().port_cp_brute(_dp(V), _ip(F), F.shape[0], _dp(P), n, _dp(sqrD), _ip(I), _dp(Cc))
    return sqrD, I, Cc


def face_distance(V, F, P, faces):
    """|P[k] - closest point on face faces[k]|^2 (tie analysis in parity checks)."""
    V = np.ascontiguousarray(V, dtype=np.float64)
    F = np.ascontiguousarray(F, dtype=np.int32)
    P = np.ascontiguousarray(P, dtype=np.float64)
    faces = np.ascontiguousarray(faces, dtype=np.int32)
    out = np.empty(P.shape[0])
    port().port_cp_face_distance(_dp(V), _ip(F), _dp(P), _ip(faces), P.shape[0], _dp(out), None)
    return out


def signed_distance_bounds(lower, upper):
    """signed_distance.cpp:81-84."""
    max_abs = max(abs(lower), abs(upper))
    with np.errstate(over="ignore"):
        up = float(np.float64(max_abs) ** 2)
        low = float(max(np.float64(max_abs) - (np.float64(upper) - np.float64(lower)), 0.0) ** 2)
    return low, up


def signed_distance_fwn(P, V, F, lower=None, upper=None, tree=None, fwn=None):
    """igl::signed_distance(..., SIGNED_DISTANCE_TYPE_FAST_WINDING_NUMBER, ...)
    signed_distance.cpp:28-139 (+ :398-405 empty F, :436-439 default bounds).
    Returns S, I, C, W (W float64 holding the float winding number)."""
    P = np.ascontiguousarray(P, dtype=np.float64)
    n = P.shape[0]
    F = np.ascontiguousarray(F, dtype=np.int32)
    if F.shape[0] == 0:
        return (np.full(n, np.nan), np.full(n, -1, np.int32), np.full((n, 3), np.nan), np.full(n, np.nan))
    if lower is None:
        lower = np.finfo(np.float64).tiny
    if upper is None:
        upper = np.finfo(np.float64).max
    low, up = signed_distance_bounds(lower, upper)
    tree = tree or PortAABB(V, F)
    fwn = fwn or RefFwn(V, F, 2)
    sqrD, I, Cc = tree.squared_distance(P, low, up)
    # w is float (computeSolidAngle/(4 pi) narrowed to Derivedp::Scalar=float,
    # fast_winding_number.cpp:452) then widened to Scalar (signed_distance.cpp:121)
    w = fwn.winding(P).astype(np.float32).astype(np.float64)
    s = 1.0 - 2.0 * np.abs(w)
    S = s * np.sqrt(sqrD)
    oob = (sqrD >= up) | (sqrD < low)
    S[oob] = np.nan
    I = I.copy()
    I[oob] = F.shape[0] + 1
    Cc[oob] = 0.0
    w[oob] = np.nan
    return S, I, Cc, w


class PortPseudonormal:
    """Restated normals + igl::pseudonormal_test (SIGNED_DISTANCE_TYPE_PSEUDONORMAL)."""

    def __init__(self, V, F):
        self.V = np.ascontiguousarray(V, dtype=np.float64)
        self.F = np.ascontiguousarray(F, dtype=np.int32)
        self.h = port().port_pseudo_create(_dp(self.V), self.V.shape[0], _ip(self.F), self.F.shape[0])

    def __del__(self):
        if getattr(self, "h", None):
            port().port_pseudo_destroy(self.h)
            self.h = None

    def normals(self):
        FN = np.empty((self.F.shape[0], 3)); VN = np.empty((self.V.shape[0], 3)); EN = np.empty((self.F.shape[0], 3, 3))
        port().port_pseudo_normals(self.h, _dp(FN), _dp(VN), _dp(EN))
        return FN, VN, EN

    def test(self, Q, I, Cc):
        Q = np.ascontiguousarray(Q, dtype=np.float64); I = np.ascontiguousarray(I, dtype=np.int32)
        Cc = np.ascontiguousarray(Cc, dtype=np.float64)
        S = np.empty(Q.shape[0]); N = np.empty((Q.shape[0], 3))
        port().port_pseudo_test(self.h, _dp(Q), _ip(I), _dp(Cc), Q.shape[0], _dp(S), _dp(N))
        return S, N


def signed_distance_pseudonormal(P, V, F, lower=None, upper=None, tree=None, pseudo=None):
    """igl::signed_distance(..., SIGNED_DISTANCE_TYPE_PSEUDONORMAL, ...): S, I, C, N (signed_distance.cpp:28-139)."""
    P = np.ascontiguousarray(P, dtype=np.float64)
    F = np.ascontiguousarray(F, dtype=np.int32)
    if lower is None:
        lower = np.finfo(np.float64).tiny
    if upper is None:
        upper = np.finfo(np.float64).max
    low, up = signed_distance_bounds(lower, upper)
    tree = tree or PortAABB(V, F)
    pseudo = pseudo or PortPseudonormal(V, F)
    sqrD, I, Cc = tree.squared_distance(P, low, up)
    oob = (sqrD >= up) | (sqrD < low)
    s, N = pseudo.test(P, np.where(oob, -1, I), Cc)
    S = s * np.sqrt(sqrD)
    S[oob] = np.nan
    I = I.copy()
    I[oob] = F.shape[0] + 1
    Cc[oob] = 0.0
    N[oob] = 0.0   # the reference leaves these rows of N untouched (N.resize at :75)
    return S, I, Cc, N


class PortCloud:
    """Restated igl::octree + point-cloud fast_winding_number (double)."""

    def __init__(self, P, N, A, order=2):
        self.P = np.ascontiguousarray(P, dtype=np.float64)
        self.N = np.ascontiguousarray(N, dtype=np.float64)
        self.A = np.ascontiguousarray(A, dtype=np.float64)
        self.order = order
        self.h = port().port_cloud_create(_dp(self.P), _dp(self.N), _dp(self.A), self.P.shape[0], order)

    def __del__(self):
        if getattr(self, "h", None):
            port().port_cloud_destroy(self.h)
            self.h = None

    def num_cells(self):
        return int(port().port_cloud_num_cells(self.h))

    def cells(self):
        n = self.num_cells()
        nt = {0: 3, 1: 12, 2: 39}[self.order]
        CH = np.empty((n, 8), np.int32)
        CN = np.empty((n, 3))
        W = np.empty(n)
        CM = np.empty((n, 3))
        R = np.empty(n)
        EC = np.empty((n, nt))
        cnt = np.empty(n, np.int32)
        port().port_cloud_cells(self.h, _ip(CH), _dp(CN), _dp(W), _dp(CM), _dp(R), _dp(EC), _ip(cnt))
        return dict(CH=CH, CN=CN, W=W, CM=CM, R=R, EC=EC, counts=cnt)

    def eval(self, Q, beta=2.0):
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        WN = np.empty(Q.shape[0])
        port().port_cloud_eval(self.h, _dp(Q), Q.shape[0], float(beta), _dp(WN))
        return WN


# ---------------------------------------------------------------------------
# voxel grids around the path: grid / voxel_grid / flood_fill / marching_cubes / offset_surface
def mc_tables():
    """(edge_flags[256], edge_conn[12,2], tri[256,16]) as the genuine reference header defines them."""
    L = ref()
    ef = np.empty(256, np.int32); cn = np.empty((12, 2), np.int32); tr = np.empty((256, 16), np.int32)
    L.ref_mc_tables.argtypes = [_c_ip, _c_ip, _c_ip]
    L.ref_mc_tables(_ip(ef), _ip(cn), _ip(tr))
    return ef, cn, tr


def grid(res):
    """igl::grid (grid.cpp:12-45): (prod(res), 3) positions in [0,1]^3, x fastest."""
    res = np.ascontiguousarray(res, dtype=np.int32)
    GV = np.empty((int(np.prod(res.astype(np.int64))), 3))
    L = port()
    L.port_grid.argtypes = [_c_ip, _c_dp]
    L.port_grid(_ip(res), _dp(GV))
    return GV


def voxel_grid(bmin, bmax, s, pad_count):
    """igl::voxel_grid(box, s, pad_count, GV, side) (voxel_grid.cpp:14-66) -> GV (n,3), side (3,)."""
    bmin = np.ascontiguousarray(bmin, dtype=np.float64); bmax = np.ascontiguousarray(bmax, dtype=np.float64)
    side = np.zeros(3, np.int32)
    L = port()
    L.port_voxel_grid.argtypes = [_c_dp, _c_dp, C.c_int, C.c_int, _c_ip, _c_dp]
    L.port_voxel_grid(_dp(bmin), _dp(bmax), int(s), int(pad_count), _ip(side), None)
    GV = np.empty((int(np.prod(side.astype(np.int64))), 3))
    L.port_voxel_grid(_dp(bmin), _dp(bmax), int(s), int(pad_count), _ip(side), _dp(GV))
    return GV, side


def voxel_grid_mesh(V, offset, s, pad_count):
    """The (V, offset) overload (voxel_grid.cpp:68-88)."""
    V = np.asarray(V, np.float64)
    return voxel_grid(V.min(0) - offset, V.max(0) + offset, s, pad_count)


def flood_fill(res, S):
    """igl::flood_fill (flood_fill.cpp:11-98); returns the filled copy."""
    res = np.ascontiguousarray(res, dtype=np.int32)
    S = np.array(S, dtype=np.float64, copy=True).ravel()
    assert S.size == int(np.prod(res.astype(np.int64)))
    L = port()
    L.port_flood_fill.argtypes = [_c_ip, _c_dp]
    L.port_flood_fill(_ip(res), _dp(S))
    return S


def marching_cubes(S, GV, nx, ny, nz, isovalue=0.0, tables=None):
    """igl::marching_cubes(S,GV,nx,ny,nz,isovalue,V,F[,E2V]) -> V (n,3), F (m,3), E (n,2) = E2V keys by vertex."""
    S = np.ascontiguousarray(S, dtype=np.float64).ravel(); GV = np.ascontiguousarray(GV, dtype=np.float64)
    ef, cn, tr = tables or mc_tables()
    ef = np.ascontiguousarray(ef, np.int32); cn = np.ascontiguousarray(cn, np.int32); tr = np.ascontiguousarray(tr, np.int32)
    L = port()
    L.port_marching_cubes.restype = C.c_void_p
    L.port_marching_cubes.argtypes = [_c_dp, _c_dp, C.c_int, C.c_int, C.c_int, C.c_double, _c_ip, _c_ip, _c_ip, _c_lp, _c_lp]
    L.port_mc_get.argtypes = [C.c_void_p, _c_dp, _c_ip, _c_ip]
    L.port_mc_destroy.argtypes = [C.c_void_p]
    nV = C.c_int64(); nF = C.c_int64()
    h = L.port_marching_cubes(_dp(S), _dp(GV), int(nx), int(ny), int(nz), float(isovalue), _ip(ef), _ip(cn), _ip(tr),
                              C.byref(nV), C.byref(nF))
    V = np.empty((nV.value, 3)); F = np.empty((nF.value, 3), np.int32); E = np.empty((nV.value, 2), np.int32)
    L.port_mc_get(h, _dp(V), _ip(F), _ip(E))
    L.port_mc_destroy(h)
    return V, F, E


def marching_cubes_sparse(S, GV, GI, isovalue=0.0, tables=None):
    """igl::marching_cubes(S,GV,GI,isovalue,V,F[,E2V]) (marching_cubes.cpp:114-176) -> V, F, E."""
    S = np.ascontiguousarray(S, dtype=np.float64).ravel(); GV = np.ascontiguousarray(GV, dtype=np.float64)
    GI = np.ascontiguousarray(GI, dtype=np.int32)
    ef, cn, tr = tables or mc_tables()
    ef = np.ascontiguousarray(ef, np.int32); cn = np.ascontiguousarray(cn, np.int32); tr = np.ascontiguousarray(tr, np.int32)
    L = port()
    L.port_marching_cubes_sparse.restype = C.c_void_p
    L.port_marching_cubes_sparse.argtypes = [_c_dp, _c_dp, _c_ip, C.c_int64, C.c_double, _c_ip, _c_ip, _c_ip, _c_lp, _c_lp]
    L.port_mc_get.argtypes = [C.c_void_p, _c_dp, _c_ip, _c_ip]
    L.port_mc_destroy.argtypes = [C.c_void_p]
    nV = C.c_int64(); nF = C.c_int64()
    h = L.port_marching_cubes_sparse(_dp(S), _dp(GV), _ip(GI), GI.shape[0], float(isovalue), _ip(ef), _ip(cn), _ip(tr),
                                     C.byref(nV), C.byref(nF))
    V = np.empty((nV.value, 3)); F = np.empty((nF.value, 3), np.int32); E = np.empty((nV.value, 2), np.int32)
    L.port_mc_get(h, _dp(V), _ip(F), _ip(E))
    L.port_mc_destroy(h)
    return V, F, E


def offset_surface(V, F, isolevel, s, sign_type="fwn"):
    """igl::offset_surface (offset_surface.cpp:18-48) -> SV, SF, GV, side, S (S after flood_fill)."""
    GV, side = voxel_grid_mesh(V, isolevel, s, 1)
    h = (GV[:, 0].max() - GV[:, 0].min()) / float(side[0] - 1)
    lower, upper = isolevel - np.sqrt(3.0) * h, isolevel + np.sqrt(3.0) * h
    if sign_type == "fwn":
        S = signed_distance_fwn(GV, V, F, lower, upper)[0]
    elif sign_type == "pseudonormal":
        S = signed_distance_pseudonormal(GV, V, F, lower, upper)[0]
    else:
        raise ValueError(sign_type)
    S = flood_fill(side, S)
    SV, SF, _ = marching_cubes(S - isolevel, GV, side[0], side[1], side[2], 0.0)
    return SV, SF, GV, side, S


# ---------------------------------------------------------------------------
def winding_number(V, F, O):
    """igl::winding_number(V,F,O,W) for triangles: the exact generalised winding number, summed face by face
    (winding_number.cpp:62-89, solid_angle.cpp:10-55).  O(#O * #F): keep the product below ~1e10."""
    V = np.ascontiguousarray(V, dtype=np.float64); F = np.ascontiguousarray(F, dtype=np.int32)
    O = np.ascontiguousarray(O, dtype=np.float64)
    W = np.empty(O.shape[0])
    L = port()
    L.port_winding_number.argtypes = [_c_dp, _c_ip, C.c_int64, _c_dp, C.c_int64, _c_dp]
    L.port_winding_number(_dp(V), _ip(F), F.shape[0], _dp(O), O.shape[0], _dp(W))
    return W


def signed_distance_winding_number(P, V, F, lower=None, upper=None, tree=None):
    """igl::signed_distance(..., SIGNED_DISTANCE_TYPE_WINDING_NUMBER / _DEFAULT, ...): S, I, C, W
    (signed_distance.cpp:60-64, 107-113, 611-629: s = 1 - 2 w, S = s * sqrt(sqrd))."""
    P = np.ascontiguousarray(P, dtype=np.float64)
    F = np.ascontiguousarray(F, dtype=np.int32)
    if lower is None:
        lower = np.finfo(np.float64).tiny
    if upper is None:
        upper = np.finfo(np.float64).max
    low, up = signed_distance_bounds(lower, upper)
    tree = tree or PortAABB(V, F)
    sqrD, I, Cc = tree.squared_distance(P, low, up)
    oob = (sqrD >= up) | (sqrD < low)
    w = np.full(P.shape[0], np.nan)
    if (~oob).any():
        w[~oob] = winding_number(V, F, P[~oob])
    S = (1.0 - 2.0 * w) * np.sqrt(sqrD)
    S[oob] = np.nan
    I = I.copy()
    I[oob] = F.shape[0] + 1
    Cc[oob] = 0.0
    return S, I, Cc, w
