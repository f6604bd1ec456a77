"""ctypes binding of the CPU oracle (oracle/oracle.cpp).

ORACLE = TEST INFRASTRUCTURE ONLY. This module may be imported by tests/, by
``__graft_entry__.smoke()`` and by ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs —
never by anything under ``mp2p_icp_b200/``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

PAIR_PT2PT = np.dtype(
    [("globalIdx", "<u4"), ("localIdx", "<u4"), ("global", "<f4", 3), ("local", "<f4", 3), ("errSq", "<f4")]
)
PAIR_PT2PL = np.dtype(
    [("coefs", "<f8", 4), ("centroid", "<f8", 3), ("local", "<f4", 3), ("_pad", "<f4")]
)
PAIR_PT2LN = np.dtype([("pBase", "<f8", 3), ("director", "<f8", 3), ("local", "<f8", 3)])  # point_line_pair_t
assert PAIR_PT2PT.itemsize == 36 and PAIR_PT2PL.itemsize == 72 and PAIR_PT2LN.itemsize == 72


class _MatchPt2PtParams(C.Structure):
    _fields_ = [
        ("threshold", C.c_double),
        ("thresholdAngularDeg", C.c_double),
        ("pairingsPerPoint", C.c_uint32),
        ("allowMatchAlreadyMatchedPoints", C.c_int32),
        ("allowMatchAlreadyMatchedGlobalPoints", C.c_int32),
        ("bounding_box_intersection_check_epsilon", C.c_double),
    ]


class _MatchPt2PlParams(C.Structure):
    _fields_ = [
        ("distanceThreshold", C.c_double),
        ("searchRadius", C.c_double),
        ("knn", C.c_uint32),
        ("minimumPlanePoints", C.c_uint32),
        ("planeEigenThreshold", C.c_double),
        ("allowMatchAlreadyMatchedPoints", C.c_int32),
        ("bounding_box_intersection_check_epsilon", C.c_double),
    ]


class _HornParams(C.Structure):
    _fields_ = [
        ("use_scale_outlier_detector", C.c_int32),
        ("scale_outlier_threshold", C.c_double),
        ("w_pt2pt", C.c_double),
        ("robust_kernel", C.c_int32),
        ("robust_kernel_param", C.c_double),
        ("currentEstimateForRobust", C.c_double * 12),
    ]


class _GNParams(C.Structure):
    _fields_ = [
        ("maxInnerLoopIterations", C.c_uint32),
        ("minDelta", C.c_double),
        ("maxCost", C.c_double),
        ("w_pt2pt", C.c_double),
        ("w_pt2pl", C.c_double),
        ("kernel", C.c_int32),
        ("kernelParam", C.c_double),
    ]


KERNELS = {"None": 0, "GemanMcClure": 1, "Cauchy": 2}


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("oracle.cpp", "kdtree.hpp", "se3.hpp", "Makefile")]
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs
    )
    if force or stale:
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_kdtree_build.restype = C.c_void_p
        L.orc_kdtree_build.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        L.orc_kdtree_free.argtypes = [C.c_void_p]
        L.orc_match_pt2pt.restype = C.c_size_t
        L.orc_match_pt2pl.restype = C.c_size_t
        L.orc_pt2pl_to_pt2pt.restype = C.c_size_t
        L.orc_optimal_tf_horn.restype = C.c_int
        L.orc_optimal_tf_gauss_newton.restype = C.c_int
        L.orc_optimal_tf_gauss_newton_ex.restype = C.c_int
        L.orc_knn.restype = C.c_int
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _T(a):
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
    assert a.size == 12
    return a


# ------------------------------------------------------------------ SE(3)
def pose_from_xyzypr(x, y, z, yaw=0.0, pitch=0.0, roll=0.0):
    v = np.array([x, y, z, yaw, pitch, roll], dtype=np.float64)
    T = np.zeros(12)
    lib().orc_pose_from_xyzypr(_p(v), _p(T))
    return T.reshape(3, 4)


def compose(A, B):
    out = np.zeros(12)
    lib().orc_pose_compose(_p(_T(A)), _p(_T(B)), _p(out))
    return out.reshape(3, 4)


def inverse(A):
    out = np.zeros(12)
    lib().orc_pose_inverse(_p(_T(A)), _p(out))
    return out.reshape(3, 4)


def inverse_compose(A, B):
    """A - B = inverse(B) o A   (ICP.cpp:166,203)"""
    return compose(inverse(B), A)


def se3_exp(xi):
    xi = np.ascontiguousarray(xi, dtype=np.float64)
    out = np.zeros(12)
    lib().orc_se3_exp(_p(xi), _p(out))
    return out.reshape(3, 4)


def se3_log(T):
    out = np.zeros(6)
    lib().orc_se3_log(_p(_T(T)), _p(out))
    return out


def eig_sym(A):
    A = np.ascontiguousarray(A, dtype=np.float64)
    n = A.shape[0]
    V = np.zeros((n, n))
    vals = np.zeros(n)
    (lib().orc_eig_sym3 if n == 3 else lib().orc_eig_sym4)(_p(A), _p(V), _p(vals))
    return vals, V


def ldlt_solve6(H, b):
    x = np.zeros(6)
    lib().orc_ldlt_solve6(_p(np.ascontiguousarray(H, dtype=np.float64)), _p(np.ascontiguousarray(b, dtype=np.float64)), _p(x))
    return x


def transform_local_to_global(lx, ly, lz, T):
    lx, ly, lz = _f32(lx), _f32(ly), _f32(lz)
    n = lx.size
    gx, gy, gz = (np.empty(n, np.float32) for _ in range(3))
    bbmin, bbmax = np.empty(3, np.float32), np.empty(3, np.float32)
    lib().orc_transform_local_to_global(_p(lx), _p(ly), _p(lz), C.c_size_t(n), _p(_T(T)), _p(gx), _p(gy), _p(gz), _p(bbmin), _p(bbmax))
    return gx, gy, gz, bbmin, bbmax


# ------------------------------------------------------------------ NN index
class KDTree:
    """Global map layer + its nanoflann-style KD-tree (nn_prepare_for_3d_queries)."""

    def __init__(self, x, y, z, leaf_max: int = 10):
        self.x, self.y, self.z = _f32(x), _f32(y), _f32(z)
        self.n = self.x.size
        self._h = lib().orc_kdtree_build(_p(self.x), _p(self.y), _p(self.z), C.c_size_t(self.n), leaf_max)

    def __del__(self):
        try:
            if self._h:
                lib().orc_kdtree_free(C.c_void_p(self._h))
                self._h = None
        except Exception:
            pass

    def knn(self, qx, qy, qz, k: int, radius2: float = np.inf, bruteforce=False, nthreads=1):
        qx, qy, qz = _f32(qx), _f32(qy), _f32(qz)
        nq = qx.size
        idx = np.zeros((nq, k), np.uint32)
        d2 = np.full((nq, k), np.inf, np.float32)
        found = np.zeros(nq, np.int32)
        lib().orc_knn_batch(C.c_void_p(self._h), _p(qx), _p(qy), _p(qz), C.c_size_t(nq), k, C.c_float(radius2), _p(idx), _p(d2), _p(found), int(bruteforce), nthreads)
        return idx, d2, found


# ------------------------------------------------------------------ matchers
@dataclass
class MatchPt2PtParams:
    threshold: float
    thresholdAngularDeg: float = 0.0
    pairingsPerPoint: int = 1
    allowMatchAlreadyMatchedPoints: bool = False
    allowMatchAlreadyMatchedGlobalPoints: bool = False
    bounding_box_intersection_check_epsilon: float = 0.20


@dataclass
class MatchPt2PlParams:
    distanceThreshold: float
    searchRadius: float
    knn: int = 5
    minimumPlanePoints: int = 5
    planeEigenThreshold: float = 0.01
    allowMatchAlreadyMatchedPoints: bool = False
    bounding_box_intersection_check_epsilon: float = 0.20


def match_pt2pt(tree: KDTree, lx, ly, lz, T, prm: MatchPt2PtParams, local_paired=None, global_paired=None, nthreads=1):
    lx, ly, lz = _f32(lx), _f32(ly), _f32(lz)
    n = lx.size
    cp = _MatchPt2PtParams(prm.threshold, prm.thresholdAngularDeg, prm.pairingsPerPoint, int(prm.allowMatchAlreadyMatchedPoints), int(prm.allowMatchAlreadyMatchedGlobalPoints), prm.bounding_box_intersection_check_epsilon)
    if local_paired is None:
        local_paired = np.zeros(n, np.uint8)
    if global_paired is None:
        global_paired = np.zeros(tree.n, np.uint8)
    cap = n * prm.pairingsPerPoint
    out = np.zeros(cap, PAIR_PT2PT)
    pot = C.c_uint64(0)
    cnt = lib().orc_match_pt2pt(C.c_void_p(tree._h), _p(lx), _p(ly), _p(lz), C.c_size_t(n), _p(_T(T)), C.byref(cp), _p(local_paired), _p(global_paired), _p(out), C.c_size_t(cap), C.byref(pot), nthreads)
    return out[:cnt], pot.value


class _MatchInlierParams(C.Structure):
    _fields_ = [
        ("inliersRatio", C.c_double),
        ("allowMatchAlreadyMatchedPoints", C.c_int32),
        ("allowMatchAlreadyMatchedGlobalPoints", C.c_int32),
        ("bounding_box_intersection_check_epsilon", C.c_double),
    ]


@dataclass
class MatchInlierRatioParams:
    inliersRatio: float
    allowMatchAlreadyMatchedPoints: bool = False
    allowMatchAlreadyMatchedGlobalPoints: bool = False
    bounding_box_intersection_check_epsilon: float = 0.20


def match_inlier_ratio(tree: KDTree, lx, ly, lz, T, prm: MatchInlierRatioParams, local_paired=None, global_paired=None, nthreads=1):
    """Matcher_Points_InlierRatio (Matcher_Points_InlierRatio.cpp:41-143). The bitfields (one byte per
    point) are updated in place like MatchState. Raises where the reference throws."""
    lx, ly, lz = _f32(lx), _f32(ly), _f32(lz)
    n = lx.size
    cp = _MatchInlierParams(prm.inliersRatio, int(prm.allowMatchAlreadyMatchedPoints), int(prm.allowMatchAlreadyMatchedGlobalPoints), prm.bounding_box_intersection_check_epsilon)
    if local_paired is None:
        local_paired = np.zeros(n, np.uint8)
    if global_paired is None:
        global_paired = np.zeros(tree.n, np.uint8)
    out = np.zeros(max(n, 1), PAIR_PT2PT)
    pot = C.c_uint64(0)
    fn = lib().orc_match_inlier_ratio
    fn.restype = C.c_long
    cnt = fn(C.c_void_p(tree._h), _p(lx), _p(ly), _p(lz), C.c_size_t(n), _p(_T(T)), C.byref(cp), _p(local_paired), _p(global_paired), _p(out), C.c_size_t(n), C.byref(pot), nthreads)
    if cnt < 0:
        raise RuntimeError("Matcher_Points_InlierRatio: reference assertion (inliersRatio outside (0,1), or no tentative pairing)")
    return out[:cnt], pot.value


def match_pt2pl(tree: KDTree, lx, ly, lz, T, prm: MatchPt2PlParams, local_paired=None, nthreads=1):
    lx, ly, lz = _f32(lx), _f32(ly), _f32(lz)
    n = lx.size
    cp = _MatchPt2PlParams(prm.distanceThreshold, prm.searchRadius, prm.knn, prm.minimumPlanePoints, prm.planeEigenThreshold, int(prm.allowMatchAlreadyMatchedPoints), prm.bounding_box_intersection_check_epsilon)
    if local_paired is None:
        local_paired = np.zeros(n, np.uint8)
    out = np.zeros(n, PAIR_PT2PL)
    pot = C.c_uint64(0)
    cnt = lib().orc_match_pt2pl(C.c_void_p(tree._h), _p(lx), _p(ly), _p(lz), C.c_size_t(n), _p(_T(T)), C.byref(cp), _p(local_paired), _p(out), C.c_size_t(n), C.byref(pot), nthreads)
    return out[:cnt], pot.value


# ------------------------------------------------------------------ solvers
@dataclass
class HornParams:
    use_scale_outlier_detector: bool = False
    scale_outlier_threshold: float = 1.20
    w_pt2pt: float = 1.0
    robust_kernel: str = "None"
    robust_kernel_param: float = 1.0
    currentEstimateForRobust: np.ndarray = field(default_factory=lambda: np.eye(3, 4))


def optimal_tf_horn(pairs, wp: HornParams = None, point_weights=None):
    wp = wp or HornParams()
    pairs = np.ascontiguousarray(pairs, dtype=PAIR_PT2PT)
    cp = _HornParams(int(wp.use_scale_outlier_detector), wp.scale_outlier_threshold, wp.w_pt2pt, KERNELS[wp.robust_kernel], wp.robust_kernel_param, (C.c_double * 12)(*_T(wp.currentEstimateForRobust)))
    wc = wv = None
    nb = 0
    if point_weights:
        wc = np.array([c for c, _ in point_weights], np.uint64)
        wv = np.array([w for _, w in point_weights], np.float64)
        nb = len(point_weights)
    T = np.zeros(12)
    nout = C.c_uint64(0)
    ok = lib().orc_optimal_tf_horn(_p(pairs), C.c_size_t(pairs.size), C.byref(cp), _p(wc), _p(wv), C.c_size_t(nb), _p(T), C.byref(nout))
    return bool(ok), T.reshape(3, 4)


@dataclass
class GNParams:
    maxInnerLoopIterations: int = 6
    minDelta: float = 1e-7
    maxCost: float = 0.0
    w_pt2pt: float = 1.0
    w_pt2pl: float = 1.0
    kernel: str = "None"
    kernelParam: float = 1.0

    def c(self):
        return _GNParams(self.maxInnerLoopIterations, self.minDelta, self.maxCost, self.w_pt2pt, self.w_pt2pl, KERNELS[self.kernel], self.kernelParam)


def _pairs(p2p, p2l):
    p2p = np.ascontiguousarray(p2p if p2p is not None else np.zeros(0, PAIR_PT2PT), dtype=PAIR_PT2PT)
    p2l = np.ascontiguousarray(p2l if p2l is not None else np.zeros(0, PAIR_PT2PL), dtype=PAIR_PT2PL)
    return p2p, p2l


def gn_accumulate(p2p, p2l, T, prm: GNParams, nthreads=1):
    p2p, p2l = _pairs(p2p, p2l)
    H, g, e = np.zeros((6, 6)), np.zeros(6), C.c_double(0)
    cp = prm.c()
    lib().orc_gn_accumulate(_p(p2p), C.c_size_t(p2p.size), _p(p2l), C.c_size_t(p2l.size), _p(_T(T)), C.byref(cp), _p(H), _p(g), C.byref(e), nthreads)
    return H, g, e.value


def optimal_tf_gauss_newton(p2p, p2l, prm: GNParams, T_init, nthreads=1):
    p2p, p2l = _pairs(p2p, p2l)
    T = np.zeros(12)
    it = C.c_uint32(0)
    cp = prm.c()
    ok = lib().orc_optimal_tf_gauss_newton(_p(p2p), C.c_size_t(p2p.size), _p(p2l), C.c_size_t(p2l.size), C.byref(cp), _p(_T(T_init)), _p(T), C.byref(it), nthreads)
    return bool(ok), T.reshape(3, 4), it.value


def optimal_tf_gauss_newton_ex(p2p, p2l, p2ln, prm: GNParams, T_init, w_pt2ln=1.0, nthreads=1):
    """optimal_tf_gauss_newton over pt2pt + pt2pl + pt2ln pairings (optimal_tf_gauss_newton.cpp:36-372)."""
    p2p, p2l = _pairs(p2p, p2l)
    p2ln = np.zeros(0, PAIR_PT2LN) if p2ln is None else np.ascontiguousarray(p2ln, dtype=PAIR_PT2LN)
    T = np.zeros(12)
    it = C.c_uint32(0)
    cp = prm.c()
    ok = lib().orc_optimal_tf_gauss_newton_ex(_p(p2p), C.c_size_t(p2p.size), _p(p2l), C.c_size_t(p2l.size), _p(p2ln), C.c_size_t(p2ln.size), C.c_double(w_pt2ln), C.byref(cp), _p(_T(T_init)), _p(T), C.byref(it), nthreads)
    return bool(ok), T.reshape(3, 4), it.value


def covariance(p2p, p2l, p2ln, x6, finDif_xyz=1e-7, finDif_angles=1e-7):
    """mp2p_icp::covariance (covariance.cpp:28-141) at x6 = (x, y, z, yaw, pitch, roll). Returns (cov 6x6, hessian 6x6).
    The reference never assigns the z slot of its xInitial (:41-47) — pass x6[2] = 0 for its number."""
    p2p, p2l = _pairs(p2p, p2l)
    p2ln = np.zeros(0, PAIR_PT2LN) if p2ln is None else np.ascontiguousarray(p2ln, dtype=PAIR_PT2LN)
    cov, hes = np.zeros(36), np.zeros(36)
    x6 = np.ascontiguousarray(x6, dtype=np.float64)
    lib().orc_covariance(_p(p2p), C.c_size_t(p2p.size), _p(p2l), C.c_size_t(p2l.size), _p(p2ln), C.c_size_t(p2ln.size), _p(x6), C.c_double(finDif_xyz), C.c_double(finDif_angles), _p(cov), _p(hes))
    return cov.reshape(6, 6), hes.reshape(6, 6)


def gn_accumulate_ex(p2p, p2l, p2ln, T, prm: GNParams, w_pt2ln=1.0, nthreads=1):
    p2p, p2l = _pairs(p2p, p2l)
    p2ln = np.zeros(0, PAIR_PT2LN) if p2ln is None else np.ascontiguousarray(p2ln, dtype=PAIR_PT2LN)
    H, g, err = np.zeros(36), np.zeros(6), C.c_double(0)
    cp = prm.c()
    lib().orc_gn_accumulate_ex(_p(p2p), C.c_size_t(p2p.size), _p(p2l), C.c_size_t(p2l.size), _p(p2ln), C.c_size_t(p2ln.size), C.c_double(w_pt2ln), _p(_T(T)), C.byref(cp), _p(H), _p(g), C.byref(err), nthreads)
    return H.reshape(6, 6), g, err.value


def error_and_jacobian_pt2ln(pair, T):
    e, J = np.zeros(3), np.zeros((3, 6))
    pp = np.ascontiguousarray(pair, dtype=PAIR_PT2LN)
    lib().orc_error_and_jacobian_pt2ln(_p(pp), _p(_T(T)), _p(e), _p(J))
    return e, J


class _MatchPt2LnParams(C.Structure):
    _fields_ = [
        ("distanceThreshold", C.c_double),
        ("knn", C.c_uint32),
        ("minimumLinePoints", C.c_uint32),
        ("lineEigenThreshold", C.c_double),
        ("allowMatchAlreadyMatchedPoints", C.c_int32),
        ("bounding_box_intersection_check_epsilon", C.c_double),
    ]


@dataclass
class MatchPt2LnParams:
    distanceThreshold: float
    knn: int = 4
    minimumLinePoints: int = 4
    lineEigenThreshold: float = 0.01
    allowMatchAlreadyMatchedPoints: bool = False
    bounding_box_intersection_check_epsilon: float = 0.20


def match_pt2ln(tree: KDTree, lx, ly, lz, T, prm: MatchPt2LnParams, local_paired=None, nthreads=1):
    """Matcher_Point2Line (Matcher_Point2Line.cpp:46-163)."""
    lx, ly, lz = _f32(lx), _f32(ly), _f32(lz)
    n = lx.size
    cp = _MatchPt2LnParams(prm.distanceThreshold, prm.knn, prm.minimumLinePoints, prm.lineEigenThreshold, int(prm.allowMatchAlreadyMatchedPoints), prm.bounding_box_intersection_check_epsilon)
    if local_paired is None:
        local_paired = np.zeros(n, np.uint8)
    out = np.zeros(max(n, 1), PAIR_PT2LN)
    pot = C.c_uint64(0)
    fn = lib().orc_match_pt2ln
    fn.restype = C.c_size_t
    cnt = fn(C.c_void_p(tree._h), _p(lx), _p(ly), _p(lz), C.c_size_t(n), _p(_T(T)), C.byref(cp), _p(local_paired), _p(out), C.c_size_t(n), C.byref(pot), nthreads)
    return out[:cnt], pot.value


class _MatchAdaptiveParams(C.Structure):
    _fields_ = [
        ("confidenceInterval", C.c_double),
        ("firstToSecondDistanceMax", C.c_double),
        ("absoluteMaxSearchDistance", C.c_double),
        ("minimumCorrDist", C.c_double),
        ("enableDetectPlanes", C.c_int32),
        ("planeSearchPoints", C.c_uint32),
        ("planeMinimumFoundPoints", C.c_uint32),
        ("maxPt2PtCorrespondences", C.c_uint32),
        ("planeEigenThreshold", C.c_double),
        ("planeMinimumDistance", C.c_double),
        ("allowMatchAlreadyMatchedPoints", C.c_int32),
        ("allowMatchAlreadyMatchedGlobalPoints", C.c_int32),
        ("bounding_box_intersection_check_epsilon", C.c_double),
    ]


@dataclass
class MatchAdaptiveParams:  # Matcher_Adaptive.h:66-75
    confidenceInterval: float = 0.80
    firstToSecondDistanceMax: float = 1.2
    absoluteMaxSearchDistance: float = 5.0
    minimumCorrDist: float = 0.1
    enableDetectPlanes: bool = False
    planeSearchPoints: int = 8
    planeMinimumFoundPoints: int = 4
    maxPt2PtCorrespondences: int = 1
    planeEigenThreshold: float = 0.01
    planeMinimumDistance: float = 0.10
    allowMatchAlreadyMatchedPoints: bool = False
    allowMatchAlreadyMatchedGlobalPoints: bool = False
    bounding_box_intersection_check_epsilon: float = 0.20

    def c(self):
        return _MatchAdaptiveParams(self.confidenceInterval, self.firstToSecondDistanceMax, self.absoluteMaxSearchDistance, self.minimumCorrDist, int(self.enableDetectPlanes), self.planeSearchPoints, self.planeMinimumFoundPoints, self.maxPt2PtCorrespondences, self.planeEigenThreshold, self.planeMinimumDistance, int(self.allowMatchAlreadyMatchedPoints), int(self.allowMatchAlreadyMatchedGlobalPoints), self.bounding_box_intersection_check_epsilon)


def match_adaptive(tree: KDTree, lx, ly, lz, T, prm: MatchAdaptiveParams, local_paired=None, global_paired=None, nthreads=1):
    """Matcher_Adaptive (Matcher_Adaptive.cpp:59-314). Returns (pt2pt pairs, pt2pl pairs, potential_pairings, ci_high)."""
    lx, ly, lz = _f32(lx), _f32(ly), _f32(lz)
    n = lx.size
    cp = prm.c()
    if local_paired is None:
        local_paired = np.zeros(n, np.uint8)
    if global_paired is None:
        global_paired = np.zeros(tree.n, np.uint8)
    cap2p = max(n * prm.maxPt2PtCorrespondences, 1)
    out2p, out2l = np.zeros(cap2p, PAIR_PT2PT), np.zeros(max(n, 1), PAIR_PT2PL)
    n2p, n2l, pot, ci = C.c_size_t(0), C.c_size_t(0), C.c_uint64(0), C.c_double(0)
    fn = lib().orc_match_adaptive
    fn.restype = C.c_long
    rc = fn(C.c_void_p(tree._h), _p(lx), _p(ly), _p(lz), C.c_size_t(n), _p(_T(T)), C.byref(cp), _p(local_paired), _p(global_paired), _p(out2p), C.c_size_t(cap2p), C.byref(n2p), _p(out2l), C.c_size_t(n), C.byref(n2l), C.byref(pot), C.byref(ci), nthreads)
    if rc < 0:
        raise RuntimeError("Matcher_Adaptive: reference assertion (no neighbour within absoluteMaxSearchDistance, or all first/second errors equal)")
    return out2p[: n2p.value], out2l[: n2l.value], pot.value, ci.value


def pt2pl_to_pt2pt(p2l, T_guess):
    p2l = np.ascontiguousarray(p2l, dtype=PAIR_PT2PL)
    out = np.zeros(max(p2l.size, 1), PAIR_PT2PT)
    n = lib().orc_pt2pl_to_pt2pt(_p(p2l), C.c_size_t(p2l.size), _p(_T(T_guess)), _p(out), C.c_size_t(out.size))
    return out[:n]


def error_and_jacobian(kind: int, pair, T):
    e, J = np.zeros(3), np.zeros((3, 6))
    pp = np.ascontiguousarray(pair, dtype=PAIR_PT2PT if kind == 0 else PAIR_PT2PL)
    lib().orc_error_and_jacobian(kind, _p(pp) if kind == 0 else None, _p(pp) if kind == 1 else None, _p(_T(T)), _p(e), _p(J))
    return e, J


def decimate_voxels(x, y, z, resolution: float, method: str = "FirstPoint", flatten_to=None):
    """FilterDecimateVoxels over one layer (FilterDecimateVoxels.cpp:109-378). Returns (xyz (m, 3) float32, src (m,)
    int64: source index or -1 for averaged points), in ascending (cx, cy, cz) voxel order."""
    x, y, z = _f32(x), _f32(y), _f32(z)
    n = x.size
    ox, oy, oz = (np.zeros(max(n, 1), np.float32) for _ in range(3))
    src = np.zeros(max(n, 1), np.int64)
    fn = lib().orc_decimate_voxels
    fn.restype = C.c_size_t
    m = fn(_p(x), _p(y), _p(z), C.c_size_t(n), C.c_float(resolution), {"FirstPoint": 0, "ClosestToAverage": 1, "VoxelAverage": 2}[method], int(flatten_to is not None), C.c_float(0.0 if flatten_to is None else flatten_to), _p(ox), _p(oy), _p(oz), _p(src), C.c_size_t(n))
    return np.stack([ox[:m], oy[:m], oz[:m]], 1), src[:m]


def max_threads() -> int:
    return lib().orc_max_threads()
