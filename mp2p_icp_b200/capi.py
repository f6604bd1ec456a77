"""ctypes binding of include/mp2p_b200.h (the drop-in C ABI).

Mirrors the C structs one to one. Host numpy arrays are passed as plain pointers; device-resident
buffers (torch tensors) are passed as integer addresses with ``on_device=True``.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.environ.get("MP2P_B200_LIB") or os.path.join(_HERE, "libmp2p_b200.so")  # env override: A/B builds only

PACKET_DOUBLES = 32
MAX_KNN = 32
PAIRS_LAST_MATCH = 2  # MP2P_B200_PAIRS_LAST_MATCH

PAIR_PT2PT = np.dtype(
    [("globalIdx", "<u4"), ("localIdx", "<u4"), ("global", "<f4", 3), ("local", "<f4", 3), ("errSq", "<f4")]
)
PAIR_PT2PL = np.dtype([("coefs", "<f8", 4), ("centroid", "<f8", 3), ("local", "<f4", 3), ("_pad", "<f4")])
PAIR_PT2LN = np.dtype([("pBase", "<f8", 3), ("director", "<f8", 3), ("local", "<f8", 3)])  # point_line_pair_t
assert PAIR_PT2PT.itemsize == 36 and PAIR_PT2PL.itemsize == 72 and PAIR_PT2LN.itemsize == 72

KERNELS = {"None": 0, "GemanMcClure": 1, "Cauchy": 2}


class Mp2pError(RuntimeError):
    pass


class _Pt2PtParams(C.Structure):
    _fields_ = [
        ("threshold", C.c_double),
        ("thresholdAngularDeg", C.c_double),
        ("pairingsPerPoint", C.c_uint32),
        ("allowMatchAlreadyMatchedPoints", C.c_int32),
        ("allowMatchAlreadyMatchedGlobalPoints", C.c_int32),
        ("bounding_box_intersection_check_epsilon", C.c_double),
    ]


class _Pt2PlParams(C.Structure):
    _fields_ = [
        ("distanceThreshold", C.c_double),
        ("searchRadius", C.c_double),
        ("knn", C.c_uint32),
        ("minimumPlanePoints", C.c_uint32),
        ("planeEigenThreshold", C.c_double),
        ("allowMatchAlreadyMatchedPoints", C.c_int32),
        ("bounding_box_intersection_check_epsilon", C.c_double),
    ]


class _Pt2LnParams(C.Structure):
    _fields_ = [
        ("distanceThreshold", C.c_double),
        ("knn", C.c_uint32),
        ("minimumLinePoints", C.c_uint32),
        ("lineEigenThreshold", C.c_double),
        ("allowMatchAlreadyMatchedPoints", C.c_int32),
        ("bounding_box_intersection_check_epsilon", C.c_double),
    ]


class _AdaptiveParams(C.Structure):
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


class _InlierRatioParams(C.Structure):
    _fields_ = [
        ("inliersRatio", C.c_double),
        ("allowMatchAlreadyMatchedPoints", C.c_int32),
        ("allowMatchAlreadyMatchedGlobalPoints", C.c_int32),
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


class _DecimateParams(C.Structure):
    _fields_ = [("voxel_filter_resolution", C.c_float), ("decimate_method", C.c_int32), ("has_flatten_to", C.c_int32), ("flatten_to", C.c_float)]


DECIMATE_METHODS = {"FirstPoint": 0, "ClosestToAverage": 1, "VoxelAverage": 2, "RandomPoint": 3}


def _decimate_params(resolution, method, flatten_to):
    return _DecimateParams(float(resolution), DECIMATE_METHODS[method] if isinstance(method, str) else int(method), int(flatten_to is not None), float(flatten_to or 0.0))


class _MapInfo(C.Structure):
    _fields_ = [
        ("n_points", C.c_uint64),
        ("bbox_min", C.c_float * 3),
        ("bbox_max", C.c_float * 3),
        ("finest_cell_size", C.c_float),
        ("n_levels", C.c_uint32),
        ("n_finest_cells", C.c_uint64),
        ("index_bytes", C.c_uint64),
        ("build_ms", C.c_float),
    ]


@dataclass
class Pt2PtParams:
    """Parameters of Matcher_Points_DistanceThreshold (same names as the reference's YAML keys)."""

    threshold: float
    thresholdAngularDeg: float = 0.0
    pairingsPerPoint: int = 1
    allowMatchAlreadyMatchedPoints: bool = False
    allowMatchAlreadyMatchedGlobalPoints: bool = False
    bounding_box_intersection_check_epsilon: float = 0.20

    def c(self):
        return _Pt2PtParams(self.threshold, self.thresholdAngularDeg, self.pairingsPerPoint, int(self.allowMatchAlreadyMatchedPoints), int(self.allowMatchAlreadyMatchedGlobalPoints), self.bounding_box_intersection_check_epsilon)


@dataclass
class Pt2PlParams:
    distanceThreshold: float
    searchRadius: float
    knn: int = 5
    minimumPlanePoints: int = 5
    planeEigenThreshold: float = 0.01
    allowMatchAlreadyMatchedPoints: bool = False
    bounding_box_intersection_check_epsilon: float = 0.20

    def c(self):
        return _Pt2PlParams(self.distanceThreshold, self.searchRadius, self.knn, self.minimumPlanePoints, self.planeEigenThreshold, int(self.allowMatchAlreadyMatchedPoints), self.bounding_box_intersection_check_epsilon)


@dataclass
class Pt2LnParams:
    """Parameters of Matcher_Point2Line (same names and defaults as the reference)."""

    distanceThreshold: float = 0.50
    knn: int = 4
    minimumLinePoints: int = 4
    lineEigenThreshold: float = 0.01
    allowMatchAlreadyMatchedPoints: bool = False
    bounding_box_intersection_check_epsilon: float = 0.20

    def c(self):
        return _Pt2LnParams(self.distanceThreshold, self.knn, self.minimumLinePoints, self.lineEigenThreshold, int(self.allowMatchAlreadyMatchedPoints), self.bounding_box_intersection_check_epsilon)


@dataclass
class AdaptiveParams:
    """Parameters of Matcher_Adaptive (same names and defaults as the reference, Matcher_Adaptive.h:66-75)."""

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
        return _AdaptiveParams(self.confidenceInterval, self.firstToSecondDistanceMax, self.absoluteMaxSearchDistance, self.minimumCorrDist, int(self.enableDetectPlanes), self.planeSearchPoints, self.planeMinimumFoundPoints, self.maxPt2PtCorrespondences, self.planeEigenThreshold, self.planeMinimumDistance, int(self.allowMatchAlreadyMatchedPoints), int(self.allowMatchAlreadyMatchedGlobalPoints), self.bounding_box_intersection_check_epsilon)


@dataclass
class InlierRatioParams:
    """Parameters of Matcher_Points_InlierRatio (same names as the reference's YAML keys)."""

    inliersRatio: float = 0.80
    allowMatchAlreadyMatchedPoints: bool = False
    allowMatchAlreadyMatchedGlobalPoints: bool = False
    bounding_box_intersection_check_epsilon: float = 0.20

    def c(self):
        return _InlierRatioParams(self.inliersRatio, int(self.allowMatchAlreadyMatchedPoints), int(self.allowMatchAlreadyMatchedGlobalPoints), self.bounding_box_intersection_check_epsilon)


@dataclass
class HornParams:
    use_scale_outlier_detector: bool = False
    scale_outlier_threshold: float = 1.20
    w_pt2pt: float = 1.0
    robust_kernel: str = "None"
    robust_kernel_param: float = 1.0
    currentEstimateForRobust: np.ndarray = field(default_factory=lambda: np.eye(3, 4))

    def c(self):
        T = np.ascontiguousarray(self.currentEstimateForRobust, dtype=np.float64).reshape(-1)
        return _HornParams(int(self.use_scale_outlier_detector), self.scale_outlier_threshold, self.w_pt2pt, KERNELS[self.robust_kernel], self.robust_kernel_param, (C.c_double * 12)(*T))


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


EXPORTS = [
    "mp2p_b200_ctx_last_count",
    "mp2p_b200_last_error", "mp2p_b200_device_count", "mp2p_b200_ctx_create", "mp2p_b200_ctx_destroy",
    "mp2p_b200_ctx_synchronize", "mp2p_b200_ctx_launch_count", "mp2p_b200_map_create", "mp2p_b200_map_destroy",
    "mp2p_b200_map_get_info", "mp2p_b200_knn", "mp2p_b200_match_pt2pt", "mp2p_b200_match_pt2pl", "mp2p_b200_match_inlier_ratio", "mp2p_b200_match_pt2ln", "mp2p_b200_solve_gauss_newton_ex", "mp2p_b200_adaptive_search", "mp2p_b200_adaptive_threshold", "mp2p_b200_adaptive_emit", "mp2p_b200_match_adaptive", "mp2p_b200_read_kitti_bin", "mp2p_b200_map_create_xyzi", "mp2p_b200_cloud_create_xyzi",
    "mp2p_b200_solve_horn", "mp2p_b200_solve_gauss_newton", "mp2p_b200_gn_accumulate",
    "mp2p_b200_gn_step_from_packet", "mp2p_b200_horn_sums", "mp2p_b200_horn_moments", "mp2p_b200_horn_finish",
    "mp2p_b200_host_alloc", "mp2p_b200_host_free", "mp2p_b200_ctx_set_profiling",
    "mp2p_b200_ctx_get_timings", "mp2p_b200_ctx_get_search_stats",
    "mp2p_b200_match_pt2pt_shard_search", "mp2p_b200_match_pt2pt_shard_resolve",
    "mp2p_b200_iterate_pt2pt_horn", "mp2p_b200_iterate_pt2pl_gn",
    "mp2p_b200_cloud_create", "mp2p_b200_cloud_destroy", "mp2p_b200_cloud_get_info",
    "mp2p_b200_shard_record_words",
    "mp2p_b200_gn_device_begin", "mp2p_b200_gn_device_accumulate", "mp2p_b200_gn_device_step",
    "mp2p_b200_pt2pl_to_pt2pt", "mp2p_b200_solve_horn_pt2pl",
    "mp2p_b200_peer_create", "mp2p_b200_peer_connect", "mp2p_b200_peer_destroy", "mp2p_b200_peer_record_slot",
    "mp2p_b200_peer_allgather_records", "mp2p_b200_peer_allreduce_packet",
    "mp2p_b200_peer_iterate_pt2pt", "mp2p_b200_peer_iterate_pt2pl_gn",
    "mp2p_b200_filter_decimate_voxels", "mp2p_b200_cloud_create_decimated", "mp2p_b200_covariance", "mp2p_b200_ctx_get_tile_trace",
    "mp2p_b200_cloud_cached",
    "mp2p_b200_layer_fingerprint",
    "mp2p_b200_layer_invalidate",
    "mp2p_b200_map_cached",
    "mp2p_b200_peer_claims_connect",
    "mp2p_b200_peer_claims_create",
]
PEER_HANDLE_BYTES = 64
GN_STATE_DOUBLES = 16
COUNT_ON_DEVICE = (1 << 64) - 1  # MP2P_B200_COUNT_ON_DEVICE

_lib = None


def library_path() -> str:
    return _SO


def load_library():
    """Load libmp2p_b200.so. Fails loudly if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise Mp2pError(f"{_SO} not built: run `python -c 'import __graft_entry__ as g; g.build()'` (or make -C mp2p_icp_b200/csrc)")
    L = C.CDLL(_SO)
    L.mp2p_b200_last_error.restype = C.c_char_p
    L.mp2p_b200_ctx_launch_count.restype = C.c_uint64
    L.mp2p_b200_ctx_launch_count.argtypes = [C.c_void_p]
    L.mp2p_b200_ctx_create.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
    L.mp2p_b200_ctx_destroy.argtypes = [C.c_void_p]
    L.mp2p_b200_ctx_synchronize.argtypes = [C.c_void_p]
    L.mp2p_b200_map_destroy.argtypes = [C.c_void_p]
    L.mp2p_b200_cloud_destroy.argtypes = [C.c_void_p]
    L.mp2p_b200_shard_record_words.restype = C.c_uint64
    L.mp2p_b200_shard_record_words.argtypes = [C.c_uint64, C.c_uint32]
    L.mp2p_b200_host_free.argtypes = [C.c_void_p]
    L.mp2p_b200_peer_destroy.argtypes = [C.c_void_p]
    _lib = L
    return L


def shard_record_words(per_shard: int, k: int) -> int:
    return int(load_library().mp2p_b200_shard_record_words(per_shard, k))


def _check(rc: int):
    if rc != 0:
        raise Mp2pError(f"mp2p_b200 error {rc}: {load_library().mp2p_b200_last_error().decode()}")


def _ptr(a):
    """numpy array -> void*, int (device address) -> void*, None -> NULL."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return a if isinstance(a, (int, np.integer)) else np.ascontiguousarray(a, dtype=np.float32)


def _pose(T):
    T = np.ascontiguousarray(T, dtype=np.float64).reshape(-1)
    assert T.size == 12
    return T


def pack_bits(flags) -> np.ndarray:
    """bool/uint8 per point -> uint32 words, bit i of word i//32 (MatchState bitfield layout)."""
    flags = np.asarray(flags).astype(bool)
    pad = (-len(flags)) % 32
    b = np.concatenate([flags, np.zeros(pad, bool)]).reshape(-1, 32)
    return (b.astype(np.uint32) << np.arange(32, dtype=np.uint32)).sum(1).astype(np.uint32)


class Context:
    """One per (process, GPU). `stream` = a cudaStream_t address (e.g. torch's) or None."""

    def __init__(self, device: int = 0, stream: int | None = None):
        L = load_library()
        h = C.c_void_p()
        if stream is not None and int(stream) == 0:
            # handle 0 is the legacy default stream; the C ABI reads NULL as "create a private stream"
            raise Mp2pError("cannot share the legacy default stream (handle 0): make a torch.cuda.Stream() current and pass its .cuda_stream")
        _check(L.mp2p_b200_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(h)))
        self._h = h
        self.device = device
        self.stream = int(stream) if stream else None  # None: the context owns a private stream
        self._maps = weakref.WeakSet()

    def close(self):
        if getattr(self, "_h", None):
            for m in list(self._maps):  # maps hold device memory of this context: free them first
                m.close()
            load_library().mp2p_b200_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def last_count(self) -> int:
        """Pairings the last matcher call left on the device (fused / sharded iterations)."""
        n = C.c_uint64(0)
        _check(load_library().mp2p_b200_ctx_last_count(self._h, C.byref(n)))
        return int(n.value)

    def synchronize(self):
        _check(load_library().mp2p_b200_ctx_synchronize(self._h))

    def set_profiling(self, timings: bool, search_stats: bool = False):
        _check(load_library().mp2p_b200_ctx_set_profiling(self._h, int(timings), int(search_stats)))

    def timings(self) -> dict:
        ms = (C.c_float * 8)()
        _check(load_library().mp2p_b200_ctx_get_timings(self._h, ms))
        names = ["nn_search", "compact", "horn_sums", "horn_moments", "gn_accumulate", "call_total", "plane_fit"]
        return {n: float(ms[i]) for i, n in enumerate(names)}

    def search_stats(self) -> dict:
        st = (C.c_uint64 * 8)()
        _check(load_library().mp2p_b200_ctx_get_search_stats(self._h, st))
        return {"probes": int(st[0]), "candidates": int(st[1]), "valid": int(st[2]), "climbed": int(st[3]),
                "max_candidates_per_query": int(st[4]), "max_probes_per_query": int(st[5]), "max_levels": int(st[6]), "heavy_warps": int(st[7])}

    @property
    def launch_count(self) -> int:
        return int(load_library().mp2p_b200_ctx_launch_count(self._h))

    # ---------------------------------------------------------------- solvers
    def cached_map(self, x, y, z):
        """(Map view of the library-owned cached index of a HOST layer, rebuilt flag) — mp2p_b200_map_cached."""
        h, rb = C.c_void_p(), C.c_int32(0)
        _check(load_library().mp2p_b200_map_cached(self._h, _ptr(x), _ptr(y), _ptr(z), C.c_uint64(x.size), C.byref(h), C.byref(rb)))
        m = Map.__new__(Map)
        m.ctx, m._h, m._borrowed = self, h, True
        return m, bool(rb.value)

    def cached_cloud(self, x, y, z):
        h, rb = C.c_void_p(), C.c_int32(0)
        _check(load_library().mp2p_b200_cloud_cached(self._h, _ptr(x), _ptr(y), _ptr(z), C.c_uint64(x.size), C.byref(h), C.byref(rb)))
        c = Cloud.__new__(Cloud)
        c.ctx, c._h, c.n, c._borrowed = self, h, int(x.size), True
        return c, bool(rb.value)

    def invalidate_layer(self, x):
        load_library().mp2p_b200_layer_invalidate(self._h, _ptr(x))

    def tile_trace(self):
        """(n, 8) uint32 {SM, start ns, end ns, tile, rounds, steps, inserts, probes | levels << 16} of the last k > 1 search (needs $MP2P_KNN_TRACE)."""
        n = C.c_uint64(0)
        _check(load_library().mp2p_b200_ctx_get_tile_trace(self._h, None, C.c_uint64(0), C.byref(n)))
        out = np.zeros((n.value, 8), np.uint32)
        if n.value:
            _check(load_library().mp2p_b200_ctx_get_tile_trace(self._h, _ptr(out), C.c_uint64(n.value), C.byref(n)))
        return out

    def covariance(self, p2p, p2l, p2ln, x6, finDif_xyz=1e-7, finDif_angles=1e-7):
        """mp2p_icp::covariance on the device. Returns (cov 6x6, hessian 6x6, positive_definite)."""
        def arr(a, dt):
            return np.zeros(0, dt) if a is None else np.ascontiguousarray(a, dtype=dt)
        p2p, p2l, p2ln = arr(p2p, PAIR_PT2PT), arr(p2l, PAIR_PT2PL), arr(p2ln, PAIR_PT2LN)
        cov, hes, pd = np.zeros(36), np.zeros(36), C.c_int32(0)
        x6 = np.ascontiguousarray(x6, dtype=np.float64)
        _check(load_library().mp2p_b200_covariance(self._h, _ptr(p2p) if p2p.size else None, C.c_uint64(p2p.size), _ptr(p2l) if p2l.size else None, C.c_uint64(p2l.size), _ptr(p2ln) if p2ln.size else None, C.c_uint64(p2ln.size), 0, _ptr(x6), C.c_double(finDif_xyz), C.c_double(finDif_angles), _ptr(cov), _ptr(hes), C.byref(pd)))
        return cov.reshape(6, 6), hes.reshape(6, 6), bool(pd.value)

    def decimate_voxels(self, x, y, z, resolution: float, method="FirstPoint", flatten_to=None, n=None, on_device=False):
        """FilterDecimateVoxels over one layer on the device. Returns (xyz (m, 3) float32, src (m,) int64: source index,
        -1 for an averaged point), ascending (cx, cy, cz) voxel order."""
        if not on_device:
            x, y, z = _f32(x), _f32(y), _f32(z)
            n = x.size
        ox, oy, oz = (np.zeros(max(n, 1), np.float32) for _ in range(3))
        src = np.zeros(max(n, 1), np.int64)
        cnt = C.c_uint64(0)
        prm = _decimate_params(resolution, method, flatten_to)
        px, py, pz = (C.c_void_p(int(a)) for a in (x, y, z)) if on_device else (_ptr(x), _ptr(y), _ptr(z))
        _check(load_library().mp2p_b200_filter_decimate_voxels(self._h, px, py, pz, C.c_uint64(n), int(on_device), C.byref(prm), _ptr(ox), _ptr(oy), _ptr(oz), _ptr(src), C.c_uint64(n), 0, C.byref(cnt)))
        m = cnt.value
        return np.stack([ox[:m], oy[:m], oz[:m]], 1), src[:m]

    def solve_horn(self, pairs, n=None, prm: HornParams = None, point_weights=None, on_device=False, last_match=False):
        """last_match=True: `pairs` is the unmodified host output of the last matcher call on this
        context; the solver reads the copy that call left on the device (MP2P_B200_PAIRS_LAST_MATCH)."""
        prm = prm or HornParams()
        if last_match:
            on_device, n = PAIRS_LAST_MATCH, len(pairs)
        elif not on_device:
            pairs = np.ascontiguousarray(pairs, dtype=PAIR_PT2PT)
            n = pairs.size
        cp = prm.c()
        wc = wv = None
        nb = 0
        if point_weights:
            wc = np.array([c for c, _ in point_weights], np.uint64)
            wv = np.array([w for _, w in point_weights], np.float64)
            nb = len(point_weights)
        T = np.zeros(12)
        solved = C.c_int32(0)
        _check(load_library().mp2p_b200_solve_horn(self._h, _ptr(pairs), C.c_uint64(n), int(on_device), C.byref(cp), _ptr(wc), _ptr(wv), C.c_uint64(nb), _ptr(T), C.byref(solved)))
        return bool(solved.value), T.reshape(3, 4)

    def pt2pl_to_pt2pt(self, p2l, T_guess):
        """pt2ln_pl_to_pt2pt (plane part): host pt2pl records -> host pt2pt records (input order)."""
        p2l = np.ascontiguousarray(p2l, dtype=PAIR_PT2PL)
        out = np.zeros(max(p2l.size, 1), PAIR_PT2PT)
        cnt = C.c_uint64(0)
        _check(load_library().mp2p_b200_pt2pl_to_pt2pt(self._h, _ptr(p2l) if p2l.size else None, C.c_uint64(p2l.size), 0, _ptr(_pose(T_guess)), _ptr(out), C.c_uint64(out.size), 0, C.byref(cnt)))
        return out[: cnt.value]

    def solve_horn_pt2pl(self, p2l, T_guess, prm: HornParams = None, last_match=False):
        """Solver_Horn over pt2pl pairings: conversion + optimal_tf_horn, all on the device."""
        prm = prm or HornParams()
        p2l = np.ascontiguousarray(p2l, dtype=PAIR_PT2PL)
        cp = prm.c()
        T = np.zeros(12)
        solved = C.c_int32(0)
        _check(load_library().mp2p_b200_solve_horn_pt2pl(self._h, _ptr(p2l) if p2l.size else None, C.c_uint64(p2l.size), PAIRS_LAST_MATCH if last_match else 0, _ptr(_pose(T_guess)), C.byref(cp), _ptr(T), C.byref(solved)))
        return bool(solved.value), T.reshape(3, 4)

    def solve_gauss_newton(self, p2p, p2l, prm: GNParams, T_init, n2p=None, n2l=None, on_device=False, last_match=False):
        if last_match:  # see solve_horn
            on_device, n2p, n2l = PAIRS_LAST_MATCH, (len(p2p) if p2p is not None else 0), (len(p2l) if p2l is not None else 0)
        elif not on_device:
            p2p = np.ascontiguousarray(p2p if p2p is not None else np.zeros(0, PAIR_PT2PT), dtype=PAIR_PT2PT)
            p2l = np.ascontiguousarray(p2l if p2l is not None else np.zeros(0, PAIR_PT2PL), dtype=PAIR_PT2PL)
            n2p, n2l = p2p.size, p2l.size
        cp = prm.c()
        T = np.zeros(12)
        it, solved = C.c_uint32(0), C.c_int32(0)
        _check(load_library().mp2p_b200_solve_gauss_newton(self._h, _ptr(p2p) if n2p else None, C.c_uint64(n2p or 0), _ptr(p2l) if n2l else None, C.c_uint64(n2l or 0), int(on_device), C.byref(cp), _ptr(_pose(T_init)), _ptr(T), C.byref(it), C.byref(solved)))
        return bool(solved.value), T.reshape(3, 4), it.value

    def solve_gauss_newton_ex(self, p2p, p2l, p2ln, prm: GNParams, T_init, w_pt2ln=1.0):
        """optimal_tf_gauss_newton over host lists of pt2pt, pt2pl and pt2ln pairings."""
        p2p = np.ascontiguousarray(p2p if p2p is not None else np.zeros(0, PAIR_PT2PT), dtype=PAIR_PT2PT)
        p2l = np.ascontiguousarray(p2l if p2l is not None else np.zeros(0, PAIR_PT2PL), dtype=PAIR_PT2PL)
        p2ln = np.ascontiguousarray(p2ln if p2ln is not None else np.zeros(0, PAIR_PT2LN), dtype=PAIR_PT2LN)
        cp = prm.c()
        T = np.zeros(12)
        it, solved = C.c_uint32(0), C.c_int32(0)
        _check(load_library().mp2p_b200_solve_gauss_newton_ex(self._h, _ptr(p2p) if p2p.size else None, C.c_uint64(p2p.size), _ptr(p2l) if p2l.size else None, C.c_uint64(p2l.size), _ptr(p2ln) if p2ln.size else None, C.c_uint64(p2ln.size), 0, C.byref(cp), C.c_double(w_pt2ln), _ptr(_pose(T_init)), _ptr(T), C.byref(it), C.byref(solved)))
        return bool(solved.value), T.reshape(3, 4), it.value

    def gn_accumulate(self, p2p, p2l, prm: GNParams, T, n2p=None, n2l=None, on_device=False, packet=None, packet_on_device=False):
        if not on_device:
            p2p = np.ascontiguousarray(p2p if p2p is not None else np.zeros(0, PAIR_PT2PT), dtype=PAIR_PT2PT)
            p2l = np.ascontiguousarray(p2l if p2l is not None else np.zeros(0, PAIR_PT2PL), dtype=PAIR_PT2PL)
            n2p, n2l = p2p.size, p2l.size
        cp = prm.c()
        if packet is None:
            packet = np.zeros(PACKET_DOUBLES)
        _check(load_library().mp2p_b200_gn_accumulate(self._h, _ptr(p2p) if n2p else None, C.c_uint64(n2p or 0), _ptr(p2l) if n2l else None, C.c_uint64(n2l or 0), int(on_device), C.byref(cp), _ptr(_pose(T)), _ptr(packet), int(packet_on_device)))
        return packet

    # device-resident Gauss-Newton loop (multi-GPU building blocks): all three only enqueue work
    def gn_device_begin(self, T, state: int):
        _check(load_library().mp2p_b200_gn_device_begin(self._h, _ptr(_pose(T)), C.c_void_p(int(state))))

    def gn_device_accumulate(self, d_p2p, n2p, d_p2l, n2l, prm_c, state: int, packet: int):
        _check(load_library().mp2p_b200_gn_device_accumulate(self._h, C.c_void_p(int(d_p2p)) if d_p2p else None, C.c_uint64(n2p or 0), C.c_void_p(int(d_p2l)) if d_p2l else None, C.c_uint64(n2l or 0), C.byref(prm_c), C.c_void_p(int(state)), C.c_void_p(int(packet))))

    def gn_device_step(self, packet: int, prm_c, state: int):
        _check(load_library().mp2p_b200_gn_device_step(self._h, C.c_void_p(int(packet)), C.byref(prm_c), C.c_void_p(int(state))))

    def horn_sums(self, pairs, n=None, on_device=False, packet=None, packet_on_device=False):
        if not on_device:
            pairs = np.ascontiguousarray(pairs, dtype=PAIR_PT2PT)
            n = pairs.size
        if packet is None:
            packet = np.zeros(PACKET_DOUBLES)
        _check(load_library().mp2p_b200_horn_sums(self._h, _ptr(pairs) if n else None, C.c_uint64(n), int(on_device), _ptr(packet), int(packet_on_device)))
        return packet

    def horn_moments(self, pairs, sums_packet, n_total, n=None, prm: HornParams = None, on_device=False, sums_on_device=False, packet=None, packet_on_device=False):
        prm = prm or HornParams()
        if not on_device:
            pairs = np.ascontiguousarray(pairs, dtype=PAIR_PT2PT)
            n = pairs.size
        if packet is None:
            packet = np.zeros(PACKET_DOUBLES)
        cp = prm.c()
        _check(load_library().mp2p_b200_horn_moments(self._h, _ptr(pairs) if n else None, C.c_uint64(n), int(on_device), C.byref(cp), _ptr(sums_packet), int(sums_on_device), C.c_uint64(n_total), _ptr(packet), int(packet_on_device)))
        return packet


class Peer:
    """NVLink mailbox exchange of one rank (csrc/peer.cu). `exchange_handles(bytes) -> list of bytes`
    all-gathers the 64-byte IPC handles across ranks (rank order) by whatever means the caller has."""

    def __init__(self, ctx: Context, rank: int, world: int, record_words: int, exchange_handles):
        L = load_library()
        h = (C.c_uint8 * PEER_HANDLE_BYTES)()
        p = C.c_void_p()
        _check(L.mp2p_b200_peer_create(ctx._h, C.c_uint32(rank), C.c_uint32(world), C.c_uint64(record_words), h, C.byref(p)))
        self._h, self.ctx, self.rank, self.world = p, ctx, rank, world
        allh = exchange_handles(bytes(h))
        if len(allh) != world or any(len(x) != PEER_HANDLE_BYTES for x in allh):
            raise Mp2pError("exchange_handles must return one 64-byte handle per rank")
        blob = (C.c_uint8 * (PEER_HANDLE_BYTES * world)).from_buffer_copy(b"".join(allh))
        _check(L.mp2p_b200_peer_connect(self._h, blob))

    def enable_owner_claims(self, n_map_points: int, exchange_handles):
        """Owner-partitioned first claims over NVLink (mp2p_b200_peer_claims_*): every rank calls this with the same
        map size; `exchange_handles` as in the constructor."""
        L = load_library()
        h = (C.c_uint8 * PEER_HANDLE_BYTES)()
        _check(L.mp2p_b200_peer_claims_create(self._h, C.c_uint64(n_map_points), h))
        allh = exchange_handles(bytes(h))
        blob = (C.c_uint8 * (PEER_HANDLE_BYTES * self.world)).from_buffer_copy(b"".join(allh))
        _check(L.mp2p_b200_peer_claims_connect(self._h, blob))

    def record_slot(self) -> int:
        p = C.c_void_p()
        _check(load_library().mp2p_b200_peer_record_slot(self._h, C.byref(p)))
        return int(p.value)

    def allgather_records(self) -> int:
        p = C.c_void_p()
        _check(load_library().mp2p_b200_peer_allgather_records(self._h, C.byref(p)))
        return int(p.value)

    def allreduce_packet(self, packet_device: int):
        _check(load_library().mp2p_b200_peer_allreduce_packet(self._h, C.c_void_p(int(packet_device))))

    def make_iterator(self, gmap, lx, ly, lz, n_local, matcher_prm, solver_prm, per_shard: int, pairs_device: int, capacity: int):
        """Pre-binds one query-sharded ICP iteration of this rank (mp2p_b200_peer_iterate_*: every
        kernel and exchange enqueued natively, one synchronisation). Returns
        pose(3x4) -> (solved, pose_out 3x4, pairings of the whole cloud or -1, GN updates)."""
        L = load_library()
        mp, sp = matcher_prm.c(), solver_prm.c()
        pose_in, pose_out = (C.c_double * 12)(), (C.c_double * 12)()
        solved, n_all, iters = C.c_int32(0), C.c_uint64(0), C.c_uint32(0)
        plx, ply, plz, n_local, kind = _local(lx, ly, lz, n_local, True)
        head = [self._h, gmap._h, plx, ply, plz, C.c_uint64(n_local), kind, pose_in, C.byref(mp)]
        if isinstance(matcher_prm, Pt2PtParams):
            horn = isinstance(solver_prm, HornParams)
            fn = L.mp2p_b200_peer_iterate_pt2pt
            args = head + [C.byref(sp) if horn else None, None if horn else C.byref(sp), C.c_uint64(per_shard), C.c_void_p(int(pairs_device)), C.c_uint64(capacity), pose_out, C.byref(solved), C.byref(n_all), C.byref(iters)]
        else:
            horn = False
            fn = L.mp2p_b200_peer_iterate_pt2pl_gn
            args = head + [C.byref(sp), C.c_void_p(int(pairs_device)), C.c_uint64(capacity), pose_out, C.byref(solved), C.byref(iters)]
        pose_in_np = np.frombuffer(pose_in, dtype=np.float64)
        pose_out_np = np.frombuffer(pose_out, dtype=np.float64).reshape(3, 4)
        keep = (lx, ly, lz, mp, sp)

        def step(T, _keep=keep):
            pose_in_np[:] = np.asarray(T, dtype=np.float64).reshape(-1)
            rc = fn(*args)
            if rc != 0:
                _check(rc)
            return bool(solved.value), pose_out_np.copy(), (int(n_all.value) if horn else -1), int(iters.value)

        return step

    def close(self):
        if getattr(self, "_h", None):
            load_library().mp2p_b200_peer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def gn_step_from_packet(packet, prm: GNParams, T):
    out = np.zeros(12)
    conv = C.c_int32(0)
    cp = prm.c()
    _check(load_library().mp2p_b200_gn_step_from_packet(_ptr(np.ascontiguousarray(packet, dtype=np.float64)), C.byref(cp), _ptr(_pose(T)), _ptr(out), C.byref(conv)))
    return out.reshape(3, 4), bool(conv.value)


def horn_finish(sums, moments):
    out = np.zeros(12)
    solved = C.c_int32(0)
    _check(load_library().mp2p_b200_horn_finish(_ptr(np.ascontiguousarray(sums, dtype=np.float64)), _ptr(np.ascontiguousarray(moments, dtype=np.float64)), _ptr(out), C.byref(solved)))
    return bool(solved.value), out.reshape(3, 4)


class _CloudInfo(C.Structure):
    _fields_ = [("n_points", C.c_uint64), ("build_ms", C.c_float), ("x_device", C.c_void_p), ("y_device", C.c_void_p), ("z_device", C.c_void_p)]


class Cloud:
    """A LOCAL cloud resident on the GPU for a whole align(): caller-order copy + Morton-sorted copy
    (mp2p_b200_cloud_create). Pass it as `lx` (ly = lz = None) to the matchers / iterators."""

    def __init__(self, ctx: Context, x, y, z, n=None, on_device=False):
        self.ctx = ctx
        if not on_device:
            x, y, z = _f32(x), _f32(y), _f32(z)
            n = x.size
        h = C.c_void_p()
        _check(load_library().mp2p_b200_cloud_create(ctx._h, _ptr(x), _ptr(y), _ptr(z), C.c_uint64(n), int(on_device), C.byref(h)))
        self._h = h
        self.n = int(n)
        ctx._maps.add(self)

    @classmethod
    def from_xyzi(cls, ctx: Context, xyzi, n=None, on_device=False):
        """From interleaved KITTI (x, y, z, intensity) float32 records: an (n, 4) host array, or a
        device address + n (mp2p_b200_cloud_create_xyzi)."""
        self = cls.__new__(cls)
        self.ctx = ctx
        if not on_device:
            xyzi = np.ascontiguousarray(xyzi, dtype=np.float32).reshape(-1, 4)
            n = xyzi.shape[0]
        h = C.c_void_p()
        _check(load_library().mp2p_b200_cloud_create_xyzi(ctx._h, _ptr(xyzi), C.c_uint64(n), int(on_device), C.byref(h)))
        self._h, self.n = h, int(n)
        ctx._maps.add(self)
        return self

    @classmethod
    def decimated(cls, ctx: "Context", x, y, z, resolution: float, method="FirstPoint", flatten_to=None, n=None, on_device=False):
        """FilterDecimateVoxels followed by cloud_create without leaving the device (the `decimated` local layer of
        demos/icp-settings-kitti.yaml:76-82 goes from the filter to the matchers in HBM)."""
        if not on_device:
            x, y, z = _f32(x), _f32(y), _f32(z)
            n = x.size
        self = cls.__new__(cls)
        p, cnt = C.c_void_p(), C.c_uint64(0)
        prm = _decimate_params(resolution, method, flatten_to)
        px, py, pz = (C.c_void_p(int(a)) for a in (x, y, z)) if on_device else (_ptr(x), _ptr(y), _ptr(z))
        _check(load_library().mp2p_b200_cloud_create_decimated(ctx._h, px, py, pz, C.c_uint64(n), int(on_device), C.byref(prm), C.byref(p), C.byref(cnt)))
        self._h, self.ctx, self.n = p, ctx, int(cnt.value)
        ctx._maps.add(self)
        return self

    def close(self):
        if getattr(self, "_h", None):
            if getattr(self.ctx, "_h", None) and not getattr(self, "_borrowed", False):  # cached layers belong to the library
                load_library().mp2p_b200_cloud_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def info(self) -> dict:
        inf = _CloudInfo()
        _check(load_library().mp2p_b200_cloud_get_info(self._h, C.byref(inf)))
        return {"n_points": inf.n_points, "build_ms": inf.build_ms, "x_device": inf.x_device, "y_device": inf.y_device, "z_device": inf.z_device}


def _local(lx, ly, lz, n_local, local_on_device):
    """-> (lx, ly, lz as void*, n_local, kind) for the `lx, ly, lz, n_local, local_on_device` arguments."""
    if isinstance(lx, Cloud):
        return lx._h, None, None, lx.n, 2
    if not local_on_device:
        lx, ly, lz = _f32(lx), _f32(ly), _f32(lz)
        return _ptr(lx), _ptr(ly), _ptr(lz), lx.size, 0
    return _ptr(lx), _ptr(ly), _ptr(lz), n_local, 1


def read_kitti_bin(path: str) -> np.ndarray:
    """A KITTI velodyne .bin file as an (n, 4) float32 array living in PINNED memory owned by the
    returned array's base object (mp2p_b200_read_kitti_bin)."""
    L = load_library()
    p, n = C.c_void_p(), C.c_uint64(0)
    _check(L.mp2p_b200_read_kitti_bin(path.encode(), C.byref(p), C.byref(n)))

    class _Owner:
        def __init__(self, addr):
            self.addr = addr

        def __del__(self):
            try:
                load_library().mp2p_b200_host_free(C.c_void_p(self.addr))
            except Exception:
                pass

    buf = (C.c_float * (4 * n.value)).from_address(p.value) if n.value else (C.c_float * 0)()
    buf._owner = _Owner(p.value)  # the array's base is `buf`: the pinned block lives as long as any view of it
    return np.frombuffer(buf, dtype=np.float32).reshape(-1, 4)


class Map:
    """A global map layer resident on the GPU with its NN index (nn_prepare_for_3d_queries)."""

    @classmethod
    def from_xyzi(cls, ctx: Context, xyzi, n=None, on_device=False):
        """From interleaved KITTI (x, y, z, intensity) float32 records (mp2p_b200_map_create_xyzi)."""
        self = cls.__new__(cls)
        self.ctx = ctx
        if not on_device:
            xyzi = np.ascontiguousarray(xyzi, dtype=np.float32).reshape(-1, 4)
            n = xyzi.shape[0]
        h = C.c_void_p()
        _check(load_library().mp2p_b200_map_create_xyzi(ctx._h, _ptr(xyzi), C.c_uint64(n), int(on_device), C.byref(h)))
        self._h, self.n = h, n
        ctx._maps.add(self)
        return self

    def __init__(self, ctx: Context, x, y, z, n=None, on_device=False):
        self.ctx = ctx
        if not on_device:
            x, y, z = _f32(x), _f32(y), _f32(z)
            n = x.size
        h = C.c_void_p()
        _check(load_library().mp2p_b200_map_create(ctx._h, _ptr(x), _ptr(y), _ptr(z), C.c_uint64(n), int(on_device), C.byref(h)))
        self._h = h
        self.n = n
        ctx._maps.add(self)

    def close(self):
        if getattr(self, "_h", None):
            if getattr(self.ctx, "_h", None) and not getattr(self, "_borrowed", False):  # a map never outlives its context; cached layers belong to the library
                load_library().mp2p_b200_map_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def info(self) -> dict:
        inf = _MapInfo()
        _check(load_library().mp2p_b200_map_get_info(self._h, C.byref(inf)))
        return {
            "n_points": inf.n_points, "bbox_min": list(inf.bbox_min), "bbox_max": list(inf.bbox_max),
            "finest_cell_size": inf.finest_cell_size, "n_levels": inf.n_levels,
            "n_finest_cells": inf.n_finest_cells, "index_bytes": inf.index_bytes, "build_ms": inf.build_ms,
        }

    def knn(self, qx, qy, qz, k: int, radius2: float = np.inf):
        qx, qy, qz = _f32(qx), _f32(qy), _f32(qz)
        nq = qx.size
        idx = np.zeros((nq, k), np.uint32)
        d2 = np.full((nq, k), np.inf, np.float32)
        found = np.zeros(nq, np.int32)
        r2 = np.float32(min(radius2, 3.0e38))
        _check(load_library().mp2p_b200_knn(self.ctx._h, self._h, _ptr(qx), _ptr(qy), _ptr(qz), C.c_uint64(nq), C.c_uint32(k), C.c_float(r2), _ptr(idx), _ptr(d2), _ptr(found)))
        return idx, d2, found

    def match_pt2pt(self, lx, ly, lz, T, prm: Pt2PtParams, local_paired=None, global_paired=None, n_local=None, local_on_device=False, out=None, out_on_device=False, capacity=None, sync=True):
        """Returns (pairs, potential_pairings); `pairs` is a numpy view of `out[:count]` for host
        output, or the count for device output."""
        plx, ply, plz, n_local, kind = _local(lx, ly, lz, n_local, local_on_device)
        cap = capacity if capacity is not None else n_local * prm.pairingsPerPoint
        if out is None and not out_on_device:
            out = np.empty(max(cap, 1), PAIR_PT2PT)
        lb = pack_bits(local_paired) if local_paired is not None else None
        gb = pack_bits(global_paired) if global_paired is not None else None
        cp = prm.c()
        cnt, pot = C.c_uint64(0), C.c_uint64(0)
        _check(load_library().mp2p_b200_match_pt2pt(self.ctx._h, self._h, plx, ply, plz, C.c_uint64(n_local), kind, _ptr(_pose(T)), C.byref(cp), _ptr(lb), _ptr(gb), _ptr(out), C.c_uint64(cap), int(out_on_device), C.byref(cnt) if sync else None, C.byref(pot)))
        if not sync:
            return None, pot.value
        if out_on_device:
            return cnt.value, pot.value
        return out[: cnt.value], pot.value

    def match_inlier_ratio(self, lx, ly, lz, T, prm: InlierRatioParams, local_paired=None, global_paired=None, n_local=None, local_on_device=False, out=None, out_on_device=False, capacity=None):
        """Matcher_Points_InlierRatio. Returns (pairs in ascending-distance order, potential_pairings)."""
        plx, ply, plz, n_local, kind = _local(lx, ly, lz, n_local, local_on_device)
        cap = capacity if capacity is not None else n_local
        if out is None and not out_on_device:
            out = np.empty(max(cap, 1), PAIR_PT2PT)
        lb = pack_bits(local_paired) if local_paired is not None else None
        gb = pack_bits(global_paired) if global_paired is not None else None
        cp = prm.c()
        cnt, pot = C.c_uint64(0), C.c_uint64(0)
        _check(load_library().mp2p_b200_match_inlier_ratio(self.ctx._h, self._h, plx, ply, plz, C.c_uint64(n_local), kind, _ptr(_pose(T)), C.byref(cp), _ptr(lb), _ptr(gb), _ptr(out), C.c_uint64(cap), int(out_on_device), C.byref(cnt), C.byref(pot)))
        if out_on_device:
            return cnt.value, pot.value
        return out[: cnt.value], pot.value

    def shard_search_pt2pt(self, lx, ly, lz, T, prm: Pt2PtParams, per_shard: int, record_out: int, n_local=None, local_on_device=False, local_paired=None):
        """Phase A of the query-sharded matcher; record_out is a DEVICE address with room for
        shard_record_words(per_shard, K) 64-bit words. Asynchronous."""
        plx, ply, plz, n_local, kind = _local(lx, ly, lz, n_local, local_on_device)
        lb = pack_bits(local_paired) if local_paired is not None else None
        cp = prm.c()
        _check(load_library().mp2p_b200_match_pt2pt_shard_search(self.ctx._h, self._h, plx, ply, plz, C.c_uint64(n_local), kind, _ptr(_pose(T)), C.byref(cp), _ptr(lb), C.c_uint64(per_shard), C.c_void_p(int(record_out))))

    def shard_resolve_pt2pt(self, n_local, shard_rank, n_shards, per_shard, records: int, prm: Pt2PtParams, global_paired=None, out=None, out_on_device=False, capacity=None, sync=True, horn_sums: int = 0):
        """Phase B. sync=False (device output only): nothing is read back, returns None; the count
        stays on the device for solver calls given n=COUNT_ON_DEVICE."""
        cap = capacity if capacity is not None else n_local * prm.pairingsPerPoint
        if out is None and not out_on_device:
            out = np.empty(max(cap, 1), PAIR_PT2PT)
        gb = pack_bits(global_paired) if global_paired is not None else None
        cp = prm.c()
        cnt = C.c_uint64(0)
        _check(load_library().mp2p_b200_match_pt2pt_shard_resolve(self.ctx._h, self._h, C.c_uint64(n_local), C.c_uint32(shard_rank), C.c_uint32(n_shards), C.c_uint64(per_shard), C.c_void_p(int(records)), C.byref(cp), _ptr(gb), _ptr(out), C.c_uint64(cap), int(out_on_device), C.byref(cnt) if sync else None, C.c_void_p(int(horn_sums)) if horn_sums else None))
        if not sync:
            return None
        if out_on_device:
            return cnt.value
        return out[: cnt.value]

    def make_iterator(self, lx, ly, lz, n_local, matcher_prm, solver_prm, pairs_device: int = 0, capacity: int = 0, local_on_device=True):
        """Pre-binds a fused ICP iteration (device-resident local cloud, or host arrays uploaded by
        every call with local_on_device=False) and returns a function
        pose(3x4) -> (solved, pose_out 3x4, n_pairs). All ctypes marshalling happens once here."""
        L = load_library()
        is_pt2pt = isinstance(matcher_prm, Pt2PtParams)
        mp, sp = matcher_prm.c(), solver_prm.c()
        pose_in, pose_out = (C.c_double * 12)(), (C.c_double * 12)()
        solved, n_pairs, iters, pot = C.c_int32(0), C.c_uint64(0), C.c_uint32(0), C.c_uint64(0)
        if not local_on_device and not isinstance(lx, Cloud):
            lx, ly, lz = _f32(lx), _f32(ly), _f32(lz)
        plx, ply, plz, n_local, kind = _local(lx, ly, lz, n_local, local_on_device)
        keep_local = (lx, ly, lz)  # noqa: F841
        args = [self.ctx._h, self._h, plx, ply, plz, C.c_uint64(n_local), kind, pose_in, C.byref(mp), C.byref(sp), C.c_void_p(int(pairs_device)) if pairs_device else None, C.c_uint64(capacity), pose_out, C.byref(solved), C.byref(n_pairs)]
        if is_pt2pt:
            fn, args = L.mp2p_b200_iterate_pt2pt_horn, args + [C.byref(pot)]
        else:
            fn, args = L.mp2p_b200_iterate_pt2pl_gn, args + [C.byref(iters), C.byref(pot)]
        keep = (mp, sp)  # noqa: F841  (keeps the structs alive as long as the closure)

        pose_in_np = np.frombuffer(pose_in, dtype=np.float64)  # views of the ctypes arrays: no per-call marshalling
        pose_out_np = np.frombuffer(pose_out, dtype=np.float64).reshape(3, 4)

        def step(T):
            pose_in_np[:] = np.asarray(T, dtype=np.float64).reshape(-1)
            rc = fn(*args)
            if rc != 0:
                _check(rc)
            return bool(solved.value), pose_out_np.copy(), int(n_pairs.value)

        step._keep = (keep_local, keep, self)  # the closure owns its inputs (a temporary Cloud must outlive it)
        return step

    def make_plugin_step(self, hx, hy, hz, matcher_prm, solver_prm, out_pairs, reuse_device_pairs=False, cache_local_cloud=True):
        """Pre-binds what the reference's ICP loop does per iteration through the two plugin classes
        (run_matchers then run_solvers, ICP.cpp:143,170) over HOST buffers: a matcher call that
        uploads the local cloud `hx, hy, hz` and returns the pairings into `out_pairs` (host), then a
        solver call over those host pairings. Defaults = the plugin classes' defaults: the local layer
        goes through the library's fingerprinted cache (mp2p_b200_cloud_cached, YAML cacheLocalCloud),
        the solver uploads the pairings it is given (the library compares them with the matcher's
        device copy and may hand out the result computed ahead of time). reuse_device_pairs = the YAML
        opt-in assumeUnmodifiedPairings: the solver names the pairings as the unmodified output of the
        last matcher call (MP2P_B200_PAIRS_LAST_MATCH) instead of uploading them. Returns
        pose(3x4) -> (solved, pose_out, n_pairs); all ctypes marshalling happens once here."""
        L = load_library()
        is_pt2pt = isinstance(matcher_prm, Pt2PtParams)
        mp, sp = matcher_prm.c(), solver_prm.c()
        hx, hy, hz = _f32(hx), _f32(hy), _f32(hz)
        n_local = hx.size
        pose_in, pose_out = (C.c_double * 12)(), (C.c_double * 12)()
        solved, cnt, pot, iters = C.c_int32(0), C.c_uint64(0), C.c_uint64(0), C.c_uint32(0)
        cap = out_pairs.size
        origin = PAIRS_LAST_MATCH if reuse_device_pairs else 0
        n_arg = C.c_uint64(0)
        if is_pt2pt:
            m_fn = L.mp2p_b200_match_pt2pt
            m_args = [self.ctx._h, self._h, _ptr(hx), _ptr(hy), _ptr(hz), C.c_uint64(n_local), 0, pose_in, C.byref(mp), None, None, _ptr(out_pairs), C.c_uint64(cap), 0, C.byref(cnt), C.byref(pot)]
        else:
            m_fn = L.mp2p_b200_match_pt2pl
            m_args = [self.ctx._h, self._h, _ptr(hx), _ptr(hy), _ptr(hz), C.c_uint64(n_local), 0, pose_in, C.byref(mp), None, _ptr(out_pairs), C.c_uint64(cap), 0, C.byref(cnt), C.byref(pot)]
        if isinstance(solver_prm, HornParams):
            s_fn = L.mp2p_b200_solve_horn
            s_args = [self.ctx._h, _ptr(out_pairs), n_arg, origin, C.byref(sp), None, None, C.c_uint64(0), pose_out, C.byref(solved)]
            n_idx = 2
        elif is_pt2pt:
            s_fn = L.mp2p_b200_solve_gauss_newton
            s_args = [self.ctx._h, _ptr(out_pairs), n_arg, None, C.c_uint64(0), origin, C.byref(sp), pose_in, pose_out, C.byref(iters), C.byref(solved)]
            n_idx = 2
        else:
            s_fn = L.mp2p_b200_solve_gauss_newton
            s_args = [self.ctx._h, None, C.c_uint64(0), _ptr(out_pairs), n_arg, origin, C.byref(sp), pose_in, pose_out, C.byref(iters), C.byref(solved)]
            n_idx = 4
        keep = (mp, sp, hx, hy, hz, out_pairs)  # noqa: F841
        pose_in_np = np.frombuffer(pose_in, dtype=np.float64)
        pose_out_np = np.frombuffer(pose_out, dtype=np.float64).reshape(3, 4)

        cloud_h = C.c_void_p()
        c_args = [self.ctx._h, _ptr(hx), _ptr(hy), _ptr(hz), C.c_uint64(n_local), C.byref(cloud_h), None]
        host_lx = m_args[2]

        def step(T):
            pose_in_np[:] = np.asarray(T, dtype=np.float64).reshape(-1)
            if cache_local_cloud:
                rc = L.mp2p_b200_cloud_cached(*c_args)
                if rc != 0:
                    _check(rc)
                m_args[2], m_args[6] = cloud_h, 2
            else:
                m_args[2], m_args[6] = host_lx, 0
            rc = m_fn(*m_args)
            if rc != 0:
                _check(rc)
            s_args[n_idx] = C.c_uint64(cnt.value)
            rc = s_fn(*s_args)
            if rc != 0:
                _check(rc)
            return bool(solved.value), pose_out_np.copy(), int(cnt.value)

        step._keep = (keep, self)
        return step

    def match_adaptive(self, lx, ly, lz, T, prm: AdaptiveParams, local_paired=None, global_paired=None, n_local=None, local_on_device=False, threshold_fn=None):
        """Matcher_Adaptive. Returns (pt2pt pairs, pt2pl pairs, potential_pairings, ci_high). With
        `threshold_fn(hist[50], err_min, err_max, n_samples) -> maxCorrDistSqr` the two-phase form is used
        (what a plugin built against MRPT does with mrpt::math::confidenceIntervalsFromHistogram)."""
        L = load_library()
        plx, ply, plz, n_local, kind = _local(lx, ly, lz, n_local, local_on_device)
        cap2p = max(n_local * prm.maxPt2PtCorrespondences, 1)
        out2p, out2l = np.empty(cap2p, PAIR_PT2PT), np.empty(max(n_local, 1), PAIR_PT2PL)
        lb = pack_bits(local_paired) if local_paired is not None else None
        gb = pack_bits(global_paired) if global_paired is not None else None
        cp = prm.c()
        n2p, n2l, pot, ci = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0), C.c_double(0)
        if threshold_fn is None:
            _check(L.mp2p_b200_match_adaptive(self.ctx._h, self._h, plx, ply, plz, C.c_uint64(n_local), kind, _ptr(_pose(T)), C.byref(cp), _ptr(lb), _ptr(gb), _ptr(out2p), C.c_uint64(cap2p), _ptr(out2l), C.c_uint64(n_local), 0, C.byref(n2p), C.byref(n2l), C.byref(ci), C.byref(pot)))
            return out2p[: n2p.value], out2l[: n2l.value], pot.value, ci.value
        hist = np.zeros(50, np.uint64)
        emin, emax, ns, gate = C.c_double(0), C.c_double(0), C.c_uint64(0), C.c_int32(0)
        _check(L.mp2p_b200_adaptive_search(self.ctx._h, self._h, plx, ply, plz, C.c_uint64(n_local), kind, _ptr(_pose(T)), C.byref(cp), _ptr(lb), _ptr(hist), C.byref(emin), C.byref(emax), C.byref(ns), C.byref(gate), C.byref(pot)))
        if not gate.value:
            return out2p[:0], out2l[:0], pot.value, 0.0
        thr = float(threshold_fn(hist, emin.value, emax.value, ns.value))
        _check(L.mp2p_b200_adaptive_emit(self.ctx._h, self._h, C.byref(cp), C.c_double(thr), _ptr(gb), _ptr(out2p), C.c_uint64(cap2p), _ptr(out2l), C.c_uint64(n_local), 0, C.byref(n2p), C.byref(n2l)))
        return out2p[: n2p.value], out2l[: n2l.value], pot.value, thr

    def match_pt2ln(self, lx, ly, lz, T, prm: Pt2LnParams, local_paired=None, n_local=None, local_on_device=False, out=None, out_on_device=False, capacity=None):
        """Matcher_Point2Line. Returns (point-to-line pairings in ascending local index, potential_pairings)."""
        plx, ply, plz, n_local, kind = _local(lx, ly, lz, n_local, local_on_device)
        cap = capacity if capacity is not None else n_local
        if out is None and not out_on_device:
            out = np.empty(max(cap, 1), PAIR_PT2LN)
        lb = pack_bits(local_paired) if local_paired is not None else None
        cp = prm.c()
        cnt, pot = C.c_uint64(0), C.c_uint64(0)
        _check(load_library().mp2p_b200_match_pt2ln(self.ctx._h, self._h, plx, ply, plz, C.c_uint64(n_local), kind, _ptr(_pose(T)), C.byref(cp), _ptr(lb), _ptr(out), C.c_uint64(cap), int(out_on_device), C.byref(cnt), C.byref(pot)))
        if out_on_device:
            return cnt.value, pot.value
        return out[: cnt.value], pot.value

    def match_pt2pl(self, lx, ly, lz, T, prm: Pt2PlParams, local_paired=None, n_local=None, local_on_device=False, out=None, out_on_device=False, capacity=None, sync=True):
        plx, ply, plz, n_local, kind = _local(lx, ly, lz, n_local, local_on_device)
        cap = capacity if capacity is not None else n_local
        if out is None and not out_on_device:
            out = np.empty(max(cap, 1), PAIR_PT2PL)
        lb = pack_bits(local_paired) if local_paired is not None else None
        cp = prm.c()
        cnt, pot = C.c_uint64(0), C.c_uint64(0)
        _check(load_library().mp2p_b200_match_pt2pl(self.ctx._h, self._h, plx, ply, plz, C.c_uint64(n_local), kind, _ptr(_pose(T)), C.byref(cp), _ptr(lb), _ptr(out), C.c_uint64(cap), int(out_on_device), C.byref(cnt) if sync else None, C.byref(pot)))
        if not sync:
            return None, pot.value
        if out_on_device:
            return cnt.value, pot.value
        return out[: cnt.value], pot.value
