"""GPU parity tests: the CUDA path, called through the C ABI (include/mp2p_b200.h), against the CPU
oracle on the same seeded inputs. Integer/index work must be BIT-EXACT (whole pairing records are
compared with array_equal); SE(3) poses must agree within 1e-5 m / 1e-5 rad (north_star).
"""
import os

import numpy as np
import pytest

import mp2p_icp_b200 as b200
from oracle import oracle_py as orc
from tests import fixtures as fx
from tests import icp_harness

pytestmark = pytest.mark.gpu
DEG = np.pi / 180.0
GOLD = os.path.join(os.path.dirname(__file__), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
POSE_TOL = 1e-5


@pytest.fixture(scope="module")
def ctx():
    c = b200.Context(0)
    yield c
    c.close()


def pose_err(A, B):
    d = orc.se3_log(orc.inverse_compose(A, B))
    return np.linalg.norm(d[:3]), np.linalg.norm(d[3:])


def assert_pose_close(A, B, tol=POSE_TOL):
    dt, dr = pose_err(A, B)
    assert dt < tol and dr < tol, (dt, dr)


def xyz(a):
    return np.ascontiguousarray(a[:, 0]), np.ascontiguousarray(a[:, 1]), np.ascontiguousarray(a[:, 2])


# --------------------------------------------------------------------------- raw k-NN (a5)
@pytest.mark.parametrize("k,r2", [(1, np.inf), (1, 0.3), (4, np.inf), (5, 2.0), (8, 0.5), (16, np.inf), (20, 4.0)])
def test_knn_bit_exact_uniform(ctx, k, r2):
    rng = np.random.default_rng(11 + k)
    M = rng.uniform(0, 30, (60000, 3)).astype(np.float32)
    Q = rng.uniform(-3, 33, (5000, 3)).astype(np.float32)
    tree = orc.KDTree(*xyz(M))
    gmap = b200.Map(ctx, *xyz(M))
    i0, d0, f0 = tree.knn(*xyz(Q), k, r2, nthreads=8)
    i1, d1, f1 = gmap.knn(*xyz(Q), k, r2)
    assert np.array_equal(f0, f1)
    mask = np.arange(k)[None, :] < f0[:, None]
    assert np.array_equal(i0[mask], i1[mask])
    assert np.array_equal(d0[mask], d1[mask])


def test_knn_bit_exact_surfaces_and_duplicates(ctx):
    """Planar (2-D manifold) data with many exact ties and duplicated points: lowest index wins."""
    rng = np.random.default_rng(5)
    g = np.stack(np.meshgrid(np.arange(200) * 0.05, np.arange(200) * 0.05, indexing="ij"), -1).reshape(-1, 2)
    M = np.concatenate([np.c_[g, np.zeros(len(g))], np.c_[g[:5000], np.zeros(5000)], np.c_[g[:, 0], np.full(len(g), 10.0), g[:, 1]]]).astype(np.float32)
    Q = np.c_[rng.uniform(0, 10, 4000), rng.uniform(0, 10, 4000), rng.normal(0, 0.02, 4000)].astype(np.float32)
    Q[:500] = M[rng.integers(0, len(M), 500)]  # queries exactly on map points
    tree, gmap = orc.KDTree(*xyz(M)), b200.Map(ctx, *xyz(M))
    for k, r2 in [(1, np.inf), (8, 1.0), (5, 0.0026)]:
        i0, d0, f0 = tree.knn(*xyz(Q), k, r2, nthreads=8)
        i1, d1, f1 = gmap.knn(*xyz(Q), k, r2)
        assert np.array_equal(f0, f1)
        mask = np.arange(k)[None, :] < f0[:, None]
        assert np.array_equal(i0[mask], i1[mask]) and np.array_equal(d0[mask], d1[mask])


def test_knn_edge_cases(ctx):
    one = b200.Map(ctx, np.array([1.0], np.float32), np.array([2.0], np.float32), np.array([3.0], np.float32))
    i, d, f = one.knn(np.array([1.0, 50.0], np.float32), np.array([2.0, 0], np.float32), np.array([3.5, 0], np.float32), 3)
    assert list(f) == [1, 1] and i[0, 0] == 0 and d[0, 0] == np.float32(0.25)
    empty = b200.Map(ctx, np.zeros(0, np.float32), np.zeros(0, np.float32), np.zeros(0, np.float32))
    i, d, f = empty.knn(np.zeros(3, np.float32), np.zeros(3, np.float32), np.zeros(3, np.float32), 2)
    assert not f.any()
    # far-away and huge-radius queries
    rng = np.random.default_rng(1)
    M = rng.normal(0, 1, (3000, 3)).astype(np.float32)
    Q = np.array([[1e4, 0, 0], [-500, 300, 2], [0, 0, 0]], np.float32)
    tree, gmap = orc.KDTree(*xyz(M)), b200.Map(ctx, *xyz(M))
    i0, d0, f0 = tree.knn(*xyz(Q), 4)
    i1, d1, f1 = gmap.knn(*xyz(Q), 4)
    assert np.array_equal(i0, i1) and np.array_equal(d0, d1) and np.array_equal(f0, f1)


_VARIANT_CHILD = r"""
import sys, hashlib, numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
import fixtures as fx, mp2p_icp_b200 as b200
M = fx.make_street_scene(n_map=400_000, length=60.0)
S = fx.make_lidar_scan((30.0, 0.4, 0.0), seed=3, length=60.0, max_range=28.0)
pose = fx.pose_xyzypr(30.1, 0.3, 0.05, 0.02, 0.003, -0.002)
ctx = b200.Context(0); gmap = b200.Map(ctx, M[:, 0].copy(), M[:, 1].copy(), M[:, 2].copy())
l = [S[:, 0].copy(), S[:, 1].copy(), S[:, 2].copy()]
h = hashlib.blake2b(digest_size=8)
p2l, _ = gmap.match_pt2pl(*l, pose, b200.Pt2PlParams(distanceThreshold=0.5, searchRadius=1.0, knn=8, minimumPlanePoints=5, planeEigenThreshold=0.01))
h.update(p2l.tobytes())
p2p, _ = gmap.match_pt2pt(*l, pose, b200.Pt2PtParams(threshold=0.6, thresholdAngularDeg=0.0, pairingsPerPoint=3))
h.update(p2p.tobytes())
G = (S.astype(np.float64) @ pose[:, :3].T + pose[:, 3]).astype(np.float32)
i, d, f = gmap.knn(G[:, 0].copy(), G[:, 1].copy(), G[:, 2].copy(), 20, 4.0)
mask = np.arange(20)[None, :] < f[:, None]
h.update(f.tobytes()); h.update(i[mask].tobytes()); h.update(d[mask].tobytes())
print("HASH", h.hexdigest(), len(S), len(p2l), len(p2p))
"""


@pytest.mark.parametrize("env", [{"MP2P_KNN_THREAD": "1"}, {"MP2P_KNN_THREAD": "1", "MP2P_KNN_DEFER_PROBES": "6", "MP2P_KNN_DEFER_CANDS": "40"}, {"MP2P_INDEX_BOX": "1"}, {"MP2P_INDEX_BOX": "1", "MP2P_KNN_THREAD": "1"}, {"MP2P_KNN_V1": "1"},
                                 {"MP2P_HOST_COPY_THREADS": "0"}, {"MP2P_HOST_COPY_THREADS": "3"}])
def test_search_variants_return_the_same_keys(env):
    """The A/B variants of the k > 1 search (one thread per query, with and without most queries handed over to
    the warp-per-query pass; tight voxel boxes; the round-1 search) are all exact searches over the same (d2, index)
    order: on a street scene with off-surface queries the pt2pl records (k = 8), the pt2pt records with three
    pairings per point (first claims) and a raw 20-NN must be byte-identical to the default's. The last two
    settings move the (pageable) record arrays with plain cudaMemcpyAsync / with three helper threads instead of
    the default bounce-buffer path (hostcopy.hpp). The knobs are read once per process, hence the child processes."""
    import subprocess, sys

    def run(extra):
        e = dict(os.environ, **extra)
        out = subprocess.run([sys.executable, "-c", _VARIANT_CHILD, ROOT], env=e, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        return [l for l in out.stdout.splitlines() if l.startswith("HASH")][-1]

    base = run({"MP2P_KNN_THREAD": "0"})
    n_scan, n_2l, n_2p = (int(v) for v in base.split()[-3:])
    assert n_scan > 100_000 and n_2l > 20_000 and n_2p > 50_000
    assert run(env) == base


# --------------------------------------------------------------------------- pt2pt matcher (a3,a4)
@pytest.mark.parametrize(
    "pose,expected",
    [((0, 0, 0, 0, 0, 0), []), ((0, 5, 0, 0, 0, 0), [(0, 0)]), ((-2, 5, 0, 0, 0, 0), [(1, 0)]), ((8.5, -1.0, 1, 45 * DEG, 0, 0), [(1, 19)])],
)
def test_matcher_pt2pt_known_answers(ctx, pose, expected):
    """tests/test-mp2p_matcher_pt2pt.cpp:56-107 through the C ABI."""
    gmap = b200.Map(ctx, *fx.pt2pt_fixture_global())
    pairs, pot = gmap.match_pt2pt(*fx.two_local_points(), fx.pose_xyzypr(*pose), b200.Pt2PtParams(threshold=1.05, thresholdAngularDeg=0.001))
    assert [(int(p["localIdx"]), int(p["globalIdx"])) for p in pairs] == expected
    assert pot == 2


def _c2(n_map, decim):
    M, L, gt = fx.make_c2(n_map=n_map, decim=decim)
    return M, L, gt


@pytest.mark.parametrize(
    "kw",
    [
        dict(threshold=1.0),
        dict(threshold=1.0, thresholdAngularDeg=0.5),
        dict(threshold=2.5, pairingsPerPoint=3),
        dict(threshold=1.5, pairingsPerPoint=8, thresholdAngularDeg=0.2),
        dict(threshold=1.0, allowMatchAlreadyMatchedGlobalPoints=True),
        dict(threshold=0.05),
    ],
)
def test_match_pt2pt_bit_exact_c2_small(ctx, kw):
    M, L, gt = _c2(200_000, 10)
    # add duplicated queries so that first-claim dedup really triggers
    L = np.concatenate([L, L[:3000] + np.float32(1e-3)])
    tree, gmap = orc.KDTree(*xyz(M)), b200.Map(ctx, *xyz(M))
    for pose in (np.eye(3, 4), gt):
        p0, pot0 = orc.match_pt2pt(tree, *xyz(L), pose, orc.MatchPt2PtParams(**kw), nthreads=8)
        p1, pot1 = gmap.match_pt2pt(*xyz(L), pose, b200.Pt2PtParams(**kw))
        assert pot0 == pot1
        assert len(p0) == len(p1)
        assert p0.tobytes() == p1.tobytes()  # every field of every record, bit for bit
    assert len(p1) > 1000  # at the GT pose the clouds overlap


def test_match_pt2pt_matchstate_bitfields(ctx):
    """Incoming MatchState bits (Matcher.cpp:46-88): paired locals are skipped, paired globals rejected."""
    M, L, gt = _c2(100_000, 10)
    rng = np.random.default_rng(2)
    lp = (rng.random(len(L)) < 0.3).astype(np.uint8)
    gp = (rng.random(len(M)) < 0.3).astype(np.uint8)
    tree, gmap = orc.KDTree(*xyz(M)), b200.Map(ctx, *xyz(M))
    for allow_local in (False, True):
        kw = dict(threshold=1.0, allowMatchAlreadyMatchedPoints=allow_local)
        p0, _ = orc.match_pt2pt(tree, *xyz(L), gt, orc.MatchPt2PtParams(**kw), lp.copy(), gp.copy(), nthreads=8)
        p1, _ = gmap.match_pt2pt(*xyz(L), gt, b200.Pt2PtParams(**kw), local_paired=lp, global_paired=gp)
        assert len(p0) > 100 and p0.tobytes() == p1.tobytes()


def test_match_pt2pt_bbox_gate_and_empty(ctx):
    M, L, gt = _c2(50_000, 10)
    gmap = b200.Map(ctx, *xyz(M))
    far = fx.pose_xyzypr(500, 0, 0)
    p1, pot = gmap.match_pt2pt(*xyz(L), far, b200.Pt2PtParams(threshold=1.0))
    assert len(p1) == 0 and pot == len(L)
    z = np.zeros(0, np.float32)
    p1, pot = gmap.match_pt2pt(z, z, z, np.eye(3, 4), b200.Pt2PtParams(threshold=1.0))
    assert len(p1) == 0 and pot == 0
    empty = b200.Map(ctx, z, z, z)
    p1, pot = empty.match_pt2pt(*xyz(L), np.eye(3, 4), b200.Pt2PtParams(threshold=1.0))
    assert len(p1) == 0 and pot == len(L)
    with pytest.raises(b200.Mp2pError):
        gmap.match_pt2pt(*xyz(L), np.eye(3, 4), b200.Pt2PtParams(threshold=-1.0))


def test_match_pt2pt_repeated_calls_are_identical(ctx):
    """The claim array is never cleared between calls (epoch tags): results must not drift."""
    M, L, gt = _c2(100_000, 5)
    gmap = b200.Map(ctx, *xyz(M))
    first, _ = gmap.match_pt2pt(*xyz(L), gt, b200.Pt2PtParams(threshold=1.0))
    first = first.copy()
    for _ in range(5):
        again, _ = gmap.match_pt2pt(*xyz(L), gt, b200.Pt2PtParams(threshold=1.0))
        assert first.tobytes() == again.tobytes()


# --------------------------------------------------------------------------- pt2pl matcher (a7,a8)
PT2PL_PRM = dict(distanceThreshold=0.1, searchRadius=0.1, minimumPlanePoints=5, knn=5, planeEigenThreshold=0.1)


def test_matcher_pt2pl_known_answers(ctx):
    """tests/test-mp2p_matcher_pt2pl.cpp:71-131 (disabled upstream) through the C ABI."""
    gmap = b200.Map(ctx, *fx.pt2pl_fixture_global())
    l = fx.two_local_points()
    prm = b200.Pt2PlParams(**PT2PL_PRM)
    assert len(gmap.match_pt2pl(*l, fx.pose_xyzypr(0, 0, 0), prm)[0]) == 0
    assert len(gmap.match_pt2pl(*l, fx.pose_xyzypr(0, 5, 0), prm)[0]) == 1
    pairs, _ = gmap.match_pt2pl(*l, fx.pose_xyzypr(8.04, 0, 0), prm)
    assert len(pairs) == 1
    np.testing.assert_allclose(pairs[0]["local"], [2, 0, 0], atol=1e-3)
    np.testing.assert_allclose(pairs[0]["centroid"], [10, 0, 0], atol=0.01)
    np.testing.assert_allclose(pairs[0]["coefs"], [1, 0, 0, -10], atol=1e-3)
    assert len(gmap.match_pt2pl(*l, fx.pose_xyzypr(18.053, 0.05, 0.03), prm)[0]) == 0


def test_match_pt2pl_parity_street_scene(ctx):
    """KITTI-shaped synthetic (C3, reduced): plane records must agree with the oracle; the plane fit
    is float/double arithmetic -> tolerance 1e-9 on coefficients (in practice bit-equal)."""
    M = fx.make_street_scene(n_map=400_000, length=60.0)
    S = fx.make_lidar_scan((30.0, 0.5, 0.0), n_rings=32, n_az=600, length=60.0)
    gt = fx.pose_xyzypr(30.0, 0.5, 0.0, 0.02, 0.0, 0.0)
    guess = fx.pose_xyzypr(30.1, 0.45, 0.02, 0.025, 0.001, -0.001)
    tree, gmap = orc.KDTree(*xyz(M)), b200.Map(ctx, *xyz(M))
    kw = dict(distanceThreshold=0.5, searchRadius=1.0, knn=8, minimumPlanePoints=5, planeEigenThreshold=0.01)
    p0, pot0 = orc.match_pt2pl(tree, *xyz(S), guess, orc.MatchPt2PlParams(**kw), nthreads=8)
    p1, pot1 = gmap.match_pt2pl(*xyz(S), guess, b200.Pt2PlParams(**kw))
    assert pot0 == pot1 and len(p0) == len(p1) and len(p0) > 2000
    assert np.array_equal(p0["local"], p1["local"])
    np.testing.assert_allclose(p1["coefs"], p0["coefs"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(p1["centroid"], p0["centroid"], rtol=0, atol=1e-9)
    del gt


# --------------------------------------------------------------------------- Matcher_Points_InlierRatio (§8f N1)
@pytest.mark.parametrize("ratio,allow_global", [(0.80, False), (0.5, False), (0.33, True), (0.999, False)])
def test_match_inlier_ratio_bit_exact(ctx, ratio, allow_global):
    """Records, their ORDER (ascending distance, ties in reverse local order) and the count must be the
    oracle's (Matcher_Points_InlierRatio.cpp:41-143), byte for byte."""
    M, L, gt = _c2(200_000, 10)
    L = np.concatenate([L, L[:3000], L[100:1100] + np.float32(1e-3)])  # exact duplicates: distance ties + contested globals
    tree, gmap = orc.KDTree(*xyz(M)), b200.Map(ctx, *xyz(M))
    guess = fx.pose_xyzypr(0.25, -0.15, 0.08, 1.7 * DEG, -0.8 * DEG, 1.2 * DEG)
    for T in (guess, gt):
        p0, pot0 = orc.match_inlier_ratio(tree, *xyz(L), T, orc.MatchInlierRatioParams(ratio, allowMatchAlreadyMatchedGlobalPoints=allow_global), nthreads=8)
        p1, pot1 = gmap.match_inlier_ratio(*xyz(L), T, b200.InlierRatioParams(ratio, allowMatchAlreadyMatchedGlobalPoints=allow_global))
        assert pot0 == pot1 == len(L) and len(p0) == len(p1) > 1000
        assert p0.tobytes() == p1.tobytes()
        assert np.all(np.diff(p1["errSq"]) >= 0)
    # resident (Morton-sorted) local cloud: same bytes
    p2, _ = gmap.match_inlier_ratio(b200.Cloud(ctx, *xyz(L)), None, None, gt, b200.InlierRatioParams(ratio, allowMatchAlreadyMatchedGlobalPoints=allow_global))
    assert p2.tobytes() == p0.tobytes()


def test_match_inlier_ratio_matchstate_and_edge_cases(ctx):
    M, L, gt = _c2(100_000, 10)
    tree, gmap = orc.KDTree(*xyz(M)), b200.Map(ctx, *xyz(M))
    rng = np.random.default_rng(3)
    lp = (rng.random(len(L)) < 0.3).astype(np.uint8)
    gp = (rng.random(len(M)) < 0.3).astype(np.uint8)
    for kw in (dict(), dict(allowMatchAlreadyMatchedPoints=True), dict(allowMatchAlreadyMatchedGlobalPoints=True)):
        lp0, gp0 = lp.copy(), gp.copy()
        p0, _ = orc.match_inlier_ratio(tree, *xyz(L), gt, orc.MatchInlierRatioParams(0.7, **kw), lp0, gp0, nthreads=8)
        p1, _ = gmap.match_inlier_ratio(*xyz(L), gt, b200.InlierRatioParams(0.7, **kw), local_paired=lp, global_paired=gp)
        assert len(p0) > 100 and p0.tobytes() == p1.tobytes()
    # the known-answer cases of tests/test_oracle_golden.py::test_inlier_ratio_semantics
    g = np.array([[0, 0, 0], [10, 0, 0], [20, 0, 0], [30, 0, 0]], np.float32)
    Ls = np.array([[0.5, 0, 0], [10.25, 0, 0], [9.75, 0, 0], [21, 0, 0], [32, 0, 0]], np.float32)
    small = b200.Map(ctx, *xyz(g))
    I = np.eye(3, 4)
    p, pot = small.match_inlier_ratio(*xyz(Ls), I, b200.InlierRatioParams(0.5, allowMatchAlreadyMatchedGlobalPoints=True))
    assert pot == 5 and list(p["localIdx"]) == [2, 1] and list(p["globalIdx"]) == [1, 1]
    p, _ = small.match_inlier_ratio(*xyz(Ls), I, b200.InlierRatioParams(0.5))
    assert list(p["localIdx"]) == [2]
    for ratio, keep in [(0.7, 4), (0.9, 4), (0.95, 5), (0.1, 0)]:
        p, _ = small.match_inlier_ratio(*xyz(Ls), I, b200.InlierRatioParams(ratio, allowMatchAlreadyMatchedGlobalPoints=True))
        assert len(p) == keep
    p, _ = small.match_inlier_ratio(*xyz(Ls), I, b200.InlierRatioParams(0.95, allowMatchAlreadyMatchedGlobalPoints=True))
    assert list(p["localIdx"]) == [2, 1, 0, 3, 4] and np.array_equal(p["global"], g[[1, 1, 0, 2, 3]]) and np.array_equal(p["local"], Ls[[2, 1, 0, 3, 4]])
    p, _ = small.match_inlier_ratio(*xyz(Ls), I, b200.InlierRatioParams(0.95), global_paired=np.array([1, 0, 0, 0], np.uint8))
    assert list(p["localIdx"]) == [2, 3, 4]
    with pytest.raises(b200.Mp2pError):  # the reference asserts 0 < ratio < 1
        small.match_inlier_ratio(*xyz(Ls), I, b200.InlierRatioParams(1.0))
    with pytest.raises(b200.Mp2pError):  # ... and nTotal > 0
        small.match_inlier_ratio(*xyz(Ls), I, b200.InlierRatioParams(0.5), local_paired=np.ones(5, np.uint8))
    p, _ = small.match_inlier_ratio(*xyz(Ls), fx.pose_xyzypr(1000, 0, 0, 0, 0, 0), b200.InlierRatioParams(0.5))
    assert len(p) == 0  # no bounding-box overlap: nothing, no error
    p, pot = small.match_inlier_ratio(*xyz(Ls[:0]), I, b200.InlierRatioParams(0.5))
    assert len(p) == 0 and pot == 0


@pytest.mark.parametrize("solver", ["horn", "gn"])
def test_icp_align_bunny_inlier_ratio_matches_oracle(ctx, solver):
    """tests/test-mp2p_icp_algos.cpp:250-262 protocol with Matcher_Points_InlierRatio: the GPU loop
    must follow the oracle loop iteration by iteration and end within 0.1 of the ground truth."""
    x, y, z = icp_harness.load_xyz_gz(os.path.join(GOLD, "bunny_decim.xyz.gz"))
    x, y, z = x[::10], y[::10], z[::10]
    P = np.stack([x, y, z], 1).astype(np.float64)
    size = P.max(0) - P.min(0)
    rng = np.random.default_rng(77)
    gt = orc.pose_from_xyzypr(*(rng.uniform(-0.15, 0.15, 3) * size), *(rng.uniform(-10, 10, 3) * DEG))
    L = ((P - gt[:, 3]) @ gt[:, :3]).astype(np.float32)
    tree, gmap = orc.KDTree(x, y, z), b200.Map(ctx, x, y, z)
    gn = dict(maxInnerLoopIterations=6)

    def run(match, horn, gauss):
        def solve(pairs, guess, it):
            if solver == "horn":
                return horn(pairs)
            ok, T, _ = gauss(pairs, guess)
            return ok, T

        return icp_harness.align(match, solve, np.eye(3, 4), icp_harness.IcpParams(maxIterations=100))

    r0 = run(lambda T, it: orc.match_inlier_ratio(tree, *xyz(L), T, orc.MatchInlierRatioParams(0.80), nthreads=4)[0], orc.optimal_tf_horn,
             lambda p, T: orc.optimal_tf_gauss_newton(p, None, orc.GNParams(**gn), T))
    r1 = run(lambda T, it: gmap.match_inlier_ratio(*xyz(L), T, b200.InlierRatioParams(0.80))[0], ctx.solve_horn,
             lambda p, T: ctx.solve_gauss_newton(p, None, b200.GNParams(**gn), T))
    assert r0.nIterations == r1.nIterations and r0.terminationReason == r1.terminationReason
    assert_pose_close(r0.pose, r1.pose)
    assert np.linalg.norm(orc.se3_log(orc.inverse_compose(r1.pose, gt))) < 0.1


# --------------------------------------------------------------------------- Matcher_Adaptive (§8f N1)
@pytest.mark.parametrize("kw", [dict(confidenceInterval=0.8, absoluteMaxSearchDistance=0.5, minimumCorrDist=0.01),
                                dict(confidenceInterval=0.75, absoluteMaxSearchDistance=2.0, minimumCorrDist=0.1),
                                dict(confidenceInterval=0.6, absoluteMaxSearchDistance=0.7, minimumCorrDist=0.02, maxPt2PtCorrespondences=3, firstToSecondDistanceMax=1.5),
                                dict(confidenceInterval=0.9, absoluteMaxSearchDistance=1.0, minimumCorrDist=0.05, maxPt2PtCorrespondences=2, allowMatchAlreadyMatchedGlobalPoints=True)])
def test_match_adaptive_pt2pt_bit_exact(ctx, kw):
    """No planes: pt2pt records, their count, the histogram-derived threshold — all the oracle's."""
    M, L, gt = _c2(200_000, 10)
    rng = np.random.default_rng(8)
    L = L.copy()
    L[:1500] += rng.normal(0, 0.6, (1500, 3)).astype(np.float32)  # a tail of poor matches for the histogram
    tree, gmap = orc.KDTree(*xyz(M)), b200.Map(ctx, *xyz(M))
    lp = (rng.random(len(L)) < 0.1).astype(np.uint8)
    gp = (rng.random(len(M)) < 0.1).astype(np.uint8)
    for T in (fx.pose_xyzypr(0.25, -0.15, 0.08, 1.7 * DEG, -0.8 * DEG, 1.2 * DEG), gt):
        for paired in (False, True):
            a0, l0, pot0, ci0 = orc.match_adaptive(tree, *xyz(L), T, orc.MatchAdaptiveParams(**kw), lp.copy() if paired else None, gp if paired else None, nthreads=8)
            a1, l1, pot1, ci1 = gmap.match_adaptive(*xyz(L), T, b200.AdaptiveParams(**kw), local_paired=lp if paired else None, global_paired=gp if paired else None)
            assert pot0 == pot1 and ci0 == ci1 and len(l0) == len(l1) == 0
            assert len(a0) == len(a1) > 100 and a0.tobytes() == a1.tobytes()
    # two-phase form (the histogram crosses to the host: what an MRPT-linked plugin uses) with a caller's threshold
    seen = {}

    def thr(hist, emin, emax, ns):
        seen.update(hist=hist.copy(), emin=emin, emax=emax, ns=ns)
        return 0.04

    a2, _, _, _ = gmap.match_adaptive(*xyz(L), gt, b200.AdaptiveParams(**kw), threshold_fn=thr)
    k = kw.get("maxPt2PtCorrespondences", 1)
    idx, d2, found = tree.knn(*orc.transform_local_to_global(*xyz(L), gt)[:3], min(k, 10), np.nextafter(np.float32(kw["absoluteMaxSearchDistance"] ** 2), np.float32(np.inf)) if k == 1 else np.float32(kw["absoluteMaxSearchDistance"] ** 2), nthreads=8)
    first_two = np.concatenate([d2[found > r, r] for r in range(min(k, 2))])
    assert seen["ns"] == len(first_two) == seen["hist"].sum() and seen["emin"] == first_two.min() and seen["emax"] == first_two.max()
    inv = 49.0 / (float(first_two.max()) - float(first_two.min()))
    assert np.array_equal(seen["hist"], np.bincount((inv * (first_two.astype(np.float64) - float(first_two.min()))).astype(np.int64), minlength=50)[:50].astype(np.uint64))
    assert np.all(a2["errSq"] < 0.04) and len(a2) > 1000


def test_match_adaptive_planes_and_edge_cases(ctx):
    Ms = fx.make_street_scene(n_map=300_000, length=50.0)
    S = fx.make_lidar_scan((25.0, 0.4, 0.0), n_rings=32, n_az=500, length=50.0)
    tree, smap = orc.KDTree(*xyz(Ms)), b200.Map(ctx, *xyz(Ms))
    T = fx.pose_xyzypr(25.03, 0.38, 0.01, 0.005, 0.0, 0.0)
    for kw in (dict(enableDetectPlanes=True, absoluteMaxSearchDistance=1.0, confidenceInterval=0.8, planeMinimumDistance=50.0),
               dict(enableDetectPlanes=True, absoluteMaxSearchDistance=0.6, confidenceInterval=0.7, planeSearchPoints=10, planeMinimumFoundPoints=5, planeEigenThreshold=0.02, planeMinimumDistance=30.0, maxPt2PtCorrespondences=2),
               dict(enableDetectPlanes=True, absoluteMaxSearchDistance=1.0, planeSearchPoints=16, planeMinimumFoundPoints=8, planeMinimumDistance=0.10)):
        lp0 = np.zeros(len(S), np.uint8)
        a0, l0, pot0, ci0 = orc.match_adaptive(tree, *xyz(S), T, orc.MatchAdaptiveParams(**kw), lp0, nthreads=8)
        a1, l1, pot1, ci1 = smap.match_adaptive(*xyz(S), T, b200.AdaptiveParams(**kw))
        assert pot0 == pot1 and ci0 == ci1 and len(a0) == len(a1) and len(l0) == len(l1)
        assert a0.tobytes() == a1.tobytes() and np.array_equal(l0["local"], l1["local"])
        np.testing.assert_allclose(l1["coefs"], l0["coefs"], rtol=0, atol=1e-9)
        np.testing.assert_allclose(l1["centroid"], l0["centroid"], rtol=0, atol=1e-9)
    assert len(l1) + len(a1) > 1000
    # the known answers of tests/test_oracle_golden.py::test_matcher_adaptive_planes_and_edge_cases
    gx, gy = np.meshgrid(np.arange(40) * 0.05, np.arange(40) * 0.05)
    G = np.stack([gx.ravel(), gy.ravel(), np.zeros(1600)], 1).astype(np.float32)
    G[:, 2] += np.random.default_rng(4).normal(0, 1e-4, len(G)).astype(np.float32)
    Ls = np.array([[1.0, 1.0, 0.03], [0.52, 1.31, 0.05], [1.0, 1.0, 1.5], [30, 30, 30]], np.float32)
    small = b200.Map(ctx, *xyz(G))
    I = np.eye(3, 4)
    prm = b200.AdaptiveParams(confidenceInterval=0.8, absoluteMaxSearchDistance=2.0, enableDetectPlanes=True, planeSearchPoints=8, planeMinimumFoundPoints=4, planeMinimumDistance=0.10, minimumCorrDist=0.1)
    p2p, p2l, pot, _ = small.match_adaptive(*xyz(Ls), I, prm)
    assert pot == 4 and len(p2l) == 2 and len(p2p) == 0 and np.array_equal(p2l["local"], Ls[:2])
    prm2 = b200.AdaptiveParams(confidenceInterval=0.8, absoluteMaxSearchDistance=2.0, minimumCorrDist=0.1)
    p2p, p2l, _, _ = small.match_adaptive(*xyz(Ls), I, prm2)
    assert len(p2l) == 0 and list(p2p["localIdx"]) == [0, 1]
    with pytest.raises(b200.Mp2pError):  # no neighbour at all
        small.match_adaptive(*xyz(Ls[:1]), I, b200.AdaptiveParams(absoluteMaxSearchDistance=0.01))
    with pytest.raises(b200.Mp2pError):  # one error value: CHistogram asserts max > min
        small.match_adaptive(*xyz(Ls[:1]), I, prm2)
    assert len(small.match_adaptive(*xyz(Ls), fx.pose_xyzypr(500, 0, 0, 0, 0, 0), prm2)[0]) == 0  # no bbox overlap
    assert len(small.match_adaptive(*xyz(Ls[:0]), I, prm2)[0]) == 0
    with pytest.raises(b200.Mp2pError):  # Matcher_Adaptive.cpp:50-51
        small.match_adaptive(*xyz(Ls), I, b200.AdaptiveParams(confidenceInterval=1.0))


# --------------------------------------------------------------------------- Matcher_Point2Line + pt2ln in GN (§8f N1)
def _pole_scene(seed=3, n_poles=400, n_ground=150_000):
    """Vertical and slanted poles (line-like neighbourhoods) over a ground plane."""
    rng = np.random.default_rng(seed)
    base = rng.uniform(0, 60, (n_poles, 2))
    tilt = rng.normal(0, 0.15, (n_poles, 2))
    h = np.arange(0, 3.0, 0.04)
    poles = np.stack([(base[:, None, 0] + tilt[:, None, 0] * h[None, :]).ravel(), (base[:, None, 1] + tilt[:, None, 1] * h[None, :]).ravel(),
                      np.broadcast_to(h, (n_poles, len(h))).ravel()], 1)
    poles += rng.normal(0, 0.002, poles.shape)
    ground = np.stack([rng.uniform(0, 60, n_ground), rng.uniform(0, 60, n_ground), rng.normal(-0.3, 0.005, n_ground)], 1)
    M = np.concatenate([poles, ground]).astype(np.float32)
    # local points: near poles (most), on the ground, and far away
    pick = rng.integers(0, len(poles), 20_000)
    near = poles[pick] + rng.normal(0, 0.05, (len(pick), 3))
    L = np.concatenate([near, ground[:4000] + [0, 0, 0.05], rng.uniform(100, 120, (500, 3))]).astype(np.float32)
    return M, L


@pytest.mark.parametrize("kw", [dict(), dict(distanceThreshold=0.12, knn=8, minimumLinePoints=3, lineEigenThreshold=0.02), dict(distanceThreshold=1.0, knn=16, minimumLinePoints=6, lineEigenThreshold=0.05)])
def test_match_pt2ln_parity(ctx, kw):
    """Same queries accepted, same order, local points bit-identical; line base / director within 1e-9
    (float mean and fp64 moments as upstream; the 3x3 eigen-solve differs in the last bits)."""
    M, L = _pole_scene()
    tree, gmap = orc.KDTree(*xyz(M)), b200.Map(ctx, *xyz(M))
    T = fx.pose_xyzypr(0.03, -0.02, 0.01, 0.004, -0.002, 0.003)
    rng = np.random.default_rng(1)
    lp = (rng.random(len(L)) < 0.2).astype(np.uint8)
    for paired in (None, lp):
        p0, pot0 = orc.match_pt2ln(tree, *xyz(L), T, orc.MatchPt2LnParams(**{**dict(distanceThreshold=0.5), **kw}), None if paired is None else paired.copy(), nthreads=8)
        p1, pot1 = gmap.match_pt2ln(*xyz(L), T, b200.Pt2LnParams(**kw), local_paired=paired)
        assert pot0 == pot1 == len(L) and len(p0) == len(p1) > 2000
        assert np.array_equal(p0["local"], p1["local"])
        np.testing.assert_allclose(p1["pBase"], p0["pBase"], rtol=0, atol=1e-9)
        np.testing.assert_allclose(p1["director"], p0["director"], rtol=0, atol=1e-9)
        np.testing.assert_allclose(np.linalg.norm(p1["director"], axis=1), 1.0, atol=1e-12)
    p2, _ = gmap.match_pt2ln(b200.Cloud(ctx, *xyz(L)), None, None, T, b200.Pt2LnParams(**kw), local_paired=lp)
    assert p2.tobytes() == p1.tobytes()  # resident (Morton-sorted) cloud: invisible
    # the oracle's known answers (tests/test_oracle_golden.py::test_matcher_pt2ln_known_answers)
    zs = np.arange(41) * 0.05
    pole = np.stack([np.full(41, 5.0), np.full(41, 5.0), zs], 1)
    gx, gy = np.meshgrid(np.arange(20) * 0.1, np.arange(20) * 0.1)
    G = np.concatenate([pole, np.stack([gx.ravel(), gy.ravel(), np.zeros(400)], 1)]).astype(np.float32)
    Ls = np.array([[5.1, 5.0, 1.0], [5.0, 5.2, 0.52], [1.0, 1.0, 0.05], [5.0, 9.0, 1.0]], np.float32)
    small = b200.Map(ctx, *xyz(G))
    p, pot = small.match_pt2ln(*xyz(Ls), np.eye(3, 4), b200.Pt2LnParams(distanceThreshold=0.5))
    assert pot == 4 and len(p) == 2 and np.allclose(np.abs(p["director"]), [[0, 0, 1]] * 2, atol=1e-6)
    assert np.allclose(p["pBase"][:, 2], [0.975, 0.525], atol=1e-6) and np.array_equal(p["local"], Ls[:2].astype(np.float64))
    assert len(small.match_pt2ln(*xyz(Ls), np.eye(3, 4), b200.Pt2LnParams(distanceThreshold=0.12))[0]) == 0
    p3, _ = small.match_pt2ln(*xyz(Ls), np.eye(3, 4), b200.Pt2LnParams(distanceThreshold=0.12, minimumLinePoints=2))
    assert len(p3) == 1 and np.allclose(p3["pBase"][0, 2], 0.975, atol=1e-6)
    with pytest.raises(b200.Mp2pError):
        small.match_pt2ln(*xyz(Ls), np.eye(3, 4), b200.Pt2LnParams(minimumLinePoints=1))  # Matcher_Point2Line.cpp:44
    assert len(small.match_pt2ln(*xyz(Ls[:0]), np.eye(3, 4), b200.Pt2LnParams())[0]) == 0


def test_gn_pt2ln_known_answers_and_mixed_lists(ctx):
    """tests/test-mp2p_optimize_pt2ln.cpp (three axis lines, 15 ground-truth poses, 1e-3), then random
    mixed pt2pt + pt2pl + pt2ln lists with a robust kernel against the oracle."""
    from tests.test_oracle_golden import PT2LN_GT, pt2ln_fixture

    for gt6 in PT2LN_GT:
        gt = orc.pose_from_xyzypr(*gt6[:3], *(np.array(gt6[3:]) * DEG))
        pairs = pt2ln_fixture(gt)
        ok1, T1, it1 = ctx.solve_gauss_newton_ex(None, None, pairs, b200.GNParams(maxInnerLoopIterations=25), np.eye(3, 4))
        ok0, T0, it0 = orc.optimal_tf_gauss_newton_ex(None, None, pairs, orc.GNParams(maxInnerLoopIterations=25), np.eye(3, 4))
        assert ok0 and ok1 and np.linalg.norm(orc.se3_log(orc.inverse_compose(T1, gt))) < 1e-3
        assert_pose_close(T0, T1, 1e-7)
    rng = np.random.default_rng(12)
    p2p, gt = _random_pairs(5000, 77, sigma=0.02)
    n = 3000
    l = rng.uniform(0, 50, (n, 3))
    g = l @ gt[:, :3].T + gt[:, 3]
    u = rng.normal(0, 1, (n, 3))
    u /= np.linalg.norm(u, axis=1)[:, None]
    p2ln = np.zeros(n, orc.PAIR_PT2LN)
    p2ln["director"], p2ln["local"] = u, l + rng.normal(0, 0.02, (n, 3))
    p2ln["pBase"] = g + u * rng.uniform(-5, 5, (n, 1))  # any point of the line through g
    nrm = rng.normal(0, 1, (n, 3))
    nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    p2l = np.zeros(n, orc.PAIR_PT2PL)
    p2l["coefs"][:, :3], p2l["coefs"][:, 3] = nrm, -(nrm * g).sum(1)
    p2l["centroid"], p2l["local"] = g, (l + rng.normal(0, 0.02, (n, 3))).astype(np.float32)
    guess = fx.pose_xyzypr(*(gt[:, 3] + 0.05), 0.0, 0.0, 0.0)
    for kernel, w in (("None", 1.0), ("GemanMcClure", 2.5), ("Cauchy", 0.4)):
        prm = dict(maxInnerLoopIterations=5, kernel=kernel, kernelParam=0.3)
        for lists in ((None, None, p2ln), (p2p, None, p2ln), (p2p, p2l, p2ln)):
            ok0, T0, it0 = orc.optimal_tf_gauss_newton_ex(*lists, orc.GNParams(**prm), guess, w_pt2ln=w, nthreads=8)
            ok1, T1, it1 = ctx.solve_gauss_newton_ex(*lists, b200.GNParams(**prm), guess, w_pt2ln=w)
            assert ok0 and ok1 and it0 == it1
            assert_pose_close(T0, T1, 1e-9)
    # without lines the extended entry point is the plain solver
    ok0, T0, _ = ctx.solve_gauss_newton(p2p, p2l, b200.GNParams(maxInnerLoopIterations=4), guess)
    ok1, T1, _ = ctx.solve_gauss_newton_ex(p2p, p2l, None, b200.GNParams(maxInnerLoopIterations=4), guess)
    assert_pose_close(T0, T1, 1e-14)


# --------------------------------------------------------------------------- solvers (a11-a14)
def _random_pairs(n, seed, sigma=0.02):
    rng = np.random.default_rng(seed)
    A = rng.uniform(0, 50, (n, 3))
    gt = fx.pose_xyzypr(*rng.uniform(-1, 1, 3), *(rng.uniform(-5, 5, 3) * DEG))
    B = (A - gt[:, 3]) @ gt[:, :3] + rng.normal(0, sigma, (n, 3))
    pairs = np.zeros(n, orc.PAIR_PT2PT)
    pairs["globalIdx"] = pairs["localIdx"] = np.arange(n)
    pairs["global"], pairs["local"] = A, B
    return pairs, gt


@pytest.mark.parametrize("n", [3, 4, 33, 1000, 100_003])
def test_horn_parity(ctx, n):
    pairs, gt = _random_pairs(n, 100 + n)
    ok0, T0 = orc.optimal_tf_horn(pairs)
    ok1, T1 = ctx.solve_horn(pairs)
    assert ok0 and ok1
    assert_pose_close(T0, T1)


def test_horn_variants(ctx):
    pairs, gt = _random_pairs(20_000, 9)
    pairs["local"][::50] += 30.0  # gross outliers for the scale detector
    for kw in (
        dict(use_scale_outlier_detector=True),
        dict(robust_kernel="GemanMcClure", robust_kernel_param=0.5, currentEstimateForRobust=gt),
        dict(robust_kernel="Cauchy", robust_kernel_param=0.5, currentEstimateForRobust=gt),
    ):
        ok0, T0 = orc.optimal_tf_horn(pairs, orc.HornParams(**kw))
        ok1, T1 = ctx.solve_horn(pairs, prm=b200.HornParams(**kw))
        assert ok0 and ok1
        assert_pose_close(T0, T1)
    w = [(5000, 1.0), (10000, 0.25), (5000, 2.0)]
    ok0, T0 = orc.optimal_tf_horn(pairs, point_weights=w)
    ok1, T1 = ctx.solve_horn(pairs, point_weights=w)
    assert_pose_close(T0, T1)
    ok, _ = ctx.solve_horn(pairs[:2])
    assert not ok  # optimal_tf_horn.cpp:96


def _random_planes(n, seed, gt):
    rng = np.random.default_rng(seed)
    nrm = rng.normal(0, 1, (n, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    c = rng.uniform(0, 50, (n, 3))
    g = c + np.cross(nrm, rng.normal(0, 1, (n, 3))) * 2.0  # a point on each plane
    p = np.zeros(n, orc.PAIR_PT2PL)
    p["coefs"][:, :3], p["coefs"][:, 3] = nrm, -(nrm * c).sum(1)
    p["centroid"] = c
    p["local"] = (g - gt[:, 3]) @ gt[:, :3] + rng.normal(0, 0.01, (n, 3))
    return p


@pytest.mark.parametrize("kernel", ["None", "GemanMcClure", "Cauchy"])
def test_gn_accumulate_and_solve_parity(ctx, kernel):
    p2p, gt = _random_pairs(50_001, 21)
    p2l = _random_planes(30_007, 22, gt)
    guess = orc.compose(gt, orc.se3_exp(np.array([0.05, -0.03, 0.02, 0.01, -0.008, 0.012])))
    prm = dict(maxInnerLoopIterations=5, kernel=kernel, kernelParam=0.15, w_pt2pt=1.0, w_pt2pl=0.7)
    H0, g0, e0 = orc.gn_accumulate(p2p, p2l, guess, orc.GNParams(**prm), nthreads=8)
    pk = ctx.gn_accumulate(p2p, p2l, b200.GNParams(**prm), guess)
    H1 = np.zeros((6, 6))
    H1[np.triu_indices(6)] = pk[:21]
    H1 = H1 + np.triu(H1, 1).T
    scale = np.abs(H0).max()
    assert np.abs(H1 - H0).max() < 1e-10 * scale
    assert np.abs(pk[21:27] - g0).max() < 1e-10 * np.abs(g0).max() + 1e-9
    assert abs(pk[27] - e0) < 1e-10 * e0 and pk[28] == len(p2p) + len(p2l)
    ok0, T0, it0 = orc.optimal_tf_gauss_newton(p2p, p2l, orc.GNParams(**prm), guess, nthreads=8)
    ok1, T1, it1 = ctx.solve_gauss_newton(p2p, p2l, b200.GNParams(**prm), guess)
    assert ok0 and ok1 and it0 == it1
    assert_pose_close(T0, T1)


def test_gn_known_answers(ctx):
    """tests/test-mp2p_optimize_pt2pl.cpp:107-128 through the C ABI."""
    from tests.test_oracle_golden import GT_POSES_PT2PL, make_pt2pl_case

    for gt in GT_POSES_PT2PL:
        GT = orc.pose_from_xyzypr(*gt)
        p2p, p2l = make_pt2pl_case(GT)
        ok, T, _ = ctx.solve_gauss_newton(p2p, p2l, b200.GNParams(maxInnerLoopIterations=25), np.eye(3, 4))
        assert ok and np.linalg.norm(orc.se3_log(orc.inverse_compose(T, GT))) < 1e-3


# --------------------------------------------------------------------------- whole iterations / align()
@pytest.mark.parametrize("solver", ["horn", "gn"])
def test_icp_align_bunny_matches_oracle(ctx, solver):
    """C1: full align() loop (tests/test-mp2p_icp_algos.cpp protocol) — the GPU run must follow the
    oracle iteration by iteration: identical pairings, pose within 1e-5."""
    x, y, z = icp_harness.load_xyz_gz(os.path.join(GOLD, "bunny_decim.xyz.gz"))
    P = np.stack([x, y, z], 1).astype(np.float64)
    gt = fx.pose_xyzypr(0.015, -0.010, 0.008, 4 * DEG, -3 * DEG, 2 * DEG)
    L = fx.to_local_frame(P, gt)
    thr = 0.40 * (P.max(0) - P.min(0)).max()
    tree, gmap = orc.KDTree(x, y, z), b200.Map(ctx, x, y, z)
    log = {"cpu": [], "gpu": []}

    def mk(which):
        def match(pose, it):
            if which == "cpu":
                p = orc.match_pt2pt(tree, *xyz(L), pose, orc.MatchPt2PtParams(threshold=thr), nthreads=8)[0]
            else:
                p = gmap.match_pt2pt(*xyz(L), pose, b200.Pt2PtParams(threshold=thr))[0].copy()
            log[which].append(p)
            return p

        def solve(pairs, guess, it):
            if solver == "horn":
                return orc.optimal_tf_horn(pairs) if which == "cpu" else ctx.solve_horn(pairs)
            if which == "cpu":
                ok, T, _ = orc.optimal_tf_gauss_newton(pairs, None, orc.GNParams(maxInnerLoopIterations=3), guess)
            else:
                ok, T, _ = ctx.solve_gauss_newton(pairs, None, b200.GNParams(maxInnerLoopIterations=3), guess)
            return ok, T

        return match, solve

    prm = icp_harness.IcpParams(maxIterations=100, minAbsStep_trans=1e-4, minAbsStep_rot=1e-4)
    r0 = icp_harness.align(*mk("cpu"), np.eye(3, 4), prm)
    r1 = icp_harness.align(*mk("gpu"), np.eye(3, 4), prm)
    assert r0.nIterations == r1.nIterations and r0.terminationReason == r1.terminationReason
    assert_pose_close(r0.pose, r1.pose)
    assert_pose_close(r1.pose, gt, tol=5e-3)
    same = sum(a.tobytes() == b.tobytes() for a, b in zip(log["cpu"], log["gpu"]))
    # poses differ by ~1e-12 between the two runs (reduction order), which can flip a float rounding
    # of a transformed query once in a while; the first iteration (identical pose) must be identical.
    assert log["cpu"][0].tobytes() == log["gpu"][0].tobytes()
    assert same >= len(log["cpu"]) - 2


# --------------------------------------------------------------------------- query sharding (§8e)
@pytest.mark.parametrize("kw", [dict(threshold=1.0), dict(threshold=2.0, pairingsPerPoint=3), dict(threshold=1.0, allowMatchAlreadyMatchedGlobalPoints=True)])
def test_sharded_match_equals_unsharded(ctx, kw):
    """Phase A (search per shard) + gathered candidates + phase B (global first-claim replay) must
    reproduce the single-call result bit for bit. The shards are run one after another on the one
    GPU here; across processes the only difference is who owns which slice (tests/test_multi_gloo.py
    covers the host-side exchange)."""
    import torch

    M, L, gt = _c2(150_000, 5)
    L = np.concatenate([L, L[:4000] + np.float32(2e-3)])  # cross-shard duplicate claims
    gmap = b200.Map(ctx, *xyz(M))
    prm = b200.Pt2PtParams(**kw)
    K = prm.pairingsPerPoint
    ref, _ = gmap.match_pt2pt(*xyz(L), gt, prm)
    ref = ref.copy()
    ok_ref, T_ref = ctx.solve_horn(ref)
    n_total, n_sh = len(L), 3
    per = -(-n_total // n_sh)  # the last shard is shorter: its padding slots must stay inert
    words = b200.capi.shard_record_words(per, K)
    assert words == per * K + 4
    records = torch.empty(n_sh * words, dtype=torch.int64, device="cuda")
    sums = torch.zeros(32 * n_sh, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    bounds = [min(s * per, n_total) for s in range(n_sh + 1)]
    for use_cloud in (False, True):
        records.fill_(0x5A5A5A5A)  # stale garbage: every word of a record must be rewritten by phase A
        shard = []
        for s in range(n_sh):
            a, b = bounds[s], bounds[s + 1]
            shard.append((b200.Cloud(ctx, *xyz(L[a:b])), None, None) if use_cloud else xyz(L[a:b]))
            gmap.shard_search_pt2pt(*shard[s], gt, prm, per, records.data_ptr() + s * words * 8)
        ctx.synchronize()
        parts = []
        for s in range(n_sh):
            a, b = bounds[s], bounds[s + 1]
            # phase B reuses the shard staged by phase A on this context: restage it
            gmap.shard_search_pt2pt(*shard[s], gt, prm, per, records.data_ptr() + s * words * 8)
            parts.append(gmap.shard_resolve_pt2pt(b - a, s, n_sh, per, records.data_ptr(), prm, horn_sums=sums.data_ptr() + s * 256).copy())
        got = np.concatenate(parts)
        assert len(got) == len(ref) and got.tobytes() == ref.tobytes()
        # HORN1 sums produced by the compaction pass: summed over shards = sums of the whole cloud
        h = sums.cpu().numpy().reshape(n_sh, 32).sum(axis=0)
        assert h[6] == len(ref) and h[7] == len(ref)
        np.testing.assert_allclose(h[0:3], ref["local"].astype(np.float64).sum(axis=0), rtol=1e-12)
        np.testing.assert_allclose(h[3:6], ref["global"].astype(np.float64).sum(axis=0), rtol=1e-12)
    del ok_ref, T_ref


def test_sharded_iteration_one_sync_world1():
    """ShardedMatcherSolver.iterate_* (asynchronous phases, counts stay on the device) at world = 1
    must equal the fused single-GPU iteration and the two-call path."""
    import torch

    from mp2p_icp_b200.sharded import ShardedMatcherSolver

    with pytest.raises(ValueError):  # a context on a private stream cannot be ordered with NCCL
        ShardedMatcherSolver(b200.Context(0), None, 0, 1, 10)
    prev = torch.cuda.current_stream()
    torch.cuda.set_stream(torch.cuda.Stream())
    try:
        _sharded_iteration_world1(b200.Context(0, stream=torch.cuda.current_stream().cuda_stream))
    finally:
        torch.cuda.set_stream(prev)


def _sharded_iteration_world1(ctx):
    import torch

    from mp2p_icp_b200.sharded import ShardedMatcherSolver

    M, L, gt = _c2(150_000, 5)
    gmap = b200.Map(ctx, *xyz(M))
    guess = fx.pose_xyzypr(0.25, -0.15, 0.08, 1.7 * DEG, -0.8 * DEG, 1.2 * DEG)
    mprm = b200.Pt2PtParams(threshold=1.0)
    pairs, _ = gmap.match_pt2pt(*xyz(L), guess, mprm)
    pairs = pairs.copy()
    ok_ref, T_ref = ctx.solve_horn(pairs)
    cloud = b200.Cloud(ctx, *xyz(L))
    d_pairs = torch.zeros(len(L) * 36, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    sh = ShardedMatcherSolver(ctx, gmap, 0, 1, len(L), k_max=1)
    for _ in range(3):
        ok, T, n = sh.iterate_pt2pt_horn((cloud, None, None), guess, mprm, b200.HornParams(), d_pairs.data_ptr(), len(L))
        assert ok and ok_ref and n == len(pairs)
        assert_pose_close(T, T_ref, 1e-9)
    assert d_pairs.cpu().numpy().view(b200.PAIR_PT2PT)[:n].tobytes() == pairs.tobytes()
    gprm = b200.GNParams(maxInnerLoopIterations=4, kernel="Cauchy", kernelParam=0.3)
    ok_g, T_g, it_g = sh.iterate_pt2pt_gn((cloud, None, None), guess, mprm, gprm, d_pairs.data_ptr(), len(L))
    ok_r, T_r, it_r = ctx.solve_gauss_newton(pairs, None, gprm, guess)
    assert ok_g and ok_r and it_g == it_r
    assert_pose_close(T_g, T_r, 1e-9)
    # pt2pl + GN through the asynchronous matcher form and the device-resident GN loop
    Ms = fx.make_street_scene(n_map=200_000, length=40.0)
    S = fx.make_lidar_scan((20.0, 0.5, 0.0), n_rings=32, n_az=400, length=40.0)
    g2 = fx.pose_xyzypr(20.08, 0.46, 0.02, 0.02, 0.001, -0.001)
    smap, scloud = b200.Map(ctx, *xyz(Ms)), b200.Cloud(ctx, *xyz(S))
    mkw = b200.Pt2PlParams(distanceThreshold=0.5, searchRadius=1.0, knn=8, minimumPlanePoints=5, planeEigenThreshold=0.01)
    skw = b200.GNParams(maxInnerLoopIterations=3, kernel="GemanMcClure", kernelParam=0.15)
    ref_l, _ = smap.match_pt2pl(*xyz(S), g2, mkw)
    ok_r, T_r, it_r = ctx.solve_gauss_newton(None, ref_l, skw, g2)
    d_pl = torch.zeros(len(S) * 72, dtype=torch.uint8, device="cuda")
    sh2 = ShardedMatcherSolver(ctx, smap, 0, 1, len(S), k_max=1)
    ok_g, T_g, it_g = sh2.iterate_pt2pl_gn((scloud, None, None), g2, mkw, skw, d_pl.data_ptr(), len(S))
    assert ok_g and ok_r and it_g == it_r and len(ref_l) > 1000
    assert_pose_close(T_g, T_r, 1e-9)
    assert d_pl.cpu().numpy().view(b200.PAIR_PT2PL)[: len(ref_l)].tobytes() == ref_l.tobytes()
    # nothing matches: the on-device count is zero, the iteration reports "not solved"
    far = fx.pose_xyzypr(500, 0, 0)
    ok, T, n = sh.iterate_pt2pt_horn((cloud, None, None), far, mprm, b200.HornParams(), d_pairs.data_ptr(), len(L))
    assert not ok and n == 0


# --------------------------------------------------------------------------- fused iterations
def test_fused_iteration_pt2pt_horn_equals_two_calls(ctx):
    """mp2p_b200_iterate_pt2pt_horn (pairings never leave HBM, one sync) == match + solve_horn."""
    import torch

    M, L, gt = _c2(200_000, 10)
    gmap = b200.Map(ctx, *xyz(M))
    guess = fx.pose_xyzypr(0.25, -0.15, 0.08, 1.7 * DEG, -0.8 * DEG, 1.2 * DEG)
    mprm, sprm = b200.Pt2PtParams(threshold=1.0), b200.HornParams()
    pairs, _ = gmap.match_pt2pt(*xyz(L), guess, mprm)
    ok_ref, T_ref = ctx.solve_horn(pairs)
    d = [torch.from_numpy(a).cuda() for a in xyz(L)]
    d_pairs = torch.zeros(len(L) * 36, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    step = gmap.make_iterator(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), len(L), mprm, sprm, d_pairs.data_ptr(), len(L))
    for _ in range(3):
        ok, T, n = step(guess)
        assert ok and ok_ref and n == len(pairs)
        assert_pose_close(T, T_ref, 1e-9)  # same kernels; the reduction grid differs (capacity vs count)
    got = d_pairs.cpu().numpy().view(b200.PAIR_PT2PT)[:n]
    assert got.tobytes() == pairs.tobytes()
    ok0, T0 = orc.optimal_tf_horn(pairs)
    assert_pose_close(T, T0)


def test_fused_iteration_pt2pl_gn_equals_two_calls(ctx):
    import torch

    M = fx.make_street_scene(n_map=300_000, length=50.0)
    S = fx.make_lidar_scan((25.0, 0.5, 0.0), n_rings=32, n_az=500, length=50.0)
    guess = fx.pose_xyzypr(25.08, 0.46, 0.02, 0.02, 0.001, -0.001)
    gmap = b200.Map(ctx, *xyz(M))
    mkw = dict(distanceThreshold=0.5, searchRadius=1.0, knn=8, minimumPlanePoints=5, planeEigenThreshold=0.01)
    skw = dict(maxInnerLoopIterations=3, kernel="GemanMcClure", kernelParam=0.15)
    pairs, _ = gmap.match_pt2pl(*xyz(S), guess, b200.Pt2PlParams(**mkw))
    ok_ref, T_ref, it_ref = ctx.solve_gauss_newton(None, pairs, b200.GNParams(**skw), guess)
    d = [torch.from_numpy(a).cuda() for a in xyz(S)]
    torch.cuda.synchronize()
    step = gmap.make_iterator(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), len(S), b200.Pt2PlParams(**mkw), b200.GNParams(**skw))
    ok, T, n = step(guess)
    assert ok and n == len(pairs)
    assert_pose_close(T, T_ref, 1e-9)
    tree = orc.KDTree(*xyz(M))
    p0, _ = orc.match_pt2pl(tree, *xyz(S), guess, orc.MatchPt2PlParams(**mkw), nthreads=8)
    ok0, T0, it0 = orc.optimal_tf_gauss_newton(None, p0, orc.GNParams(**skw), guess, nthreads=8)
    assert it0 == it_ref
    assert_pose_close(T, T0)


# --------------------------------------------------------------------------- Solver_Horn over pt2pl pairings (a14)
def test_horn_over_pt2pl_pairings_matches_oracle(ctx):
    """pt2ln_pl_to_pt2pt (plane part) + optimal_tf_horn on the device: the projected records are the
    oracle's as a SET (the reference orders them by descending |distance|, the device keeps the input
    order; Horn's sums do not depend on it), pose within 1e-5."""
    S = fx.make_street_scene(n_map=200_000, length=40.0)
    scan = fx.make_lidar_scan((20.0, 0.3, 0.0), n_rings=16, n_az=400, length=40.0)
    guess = fx.pose_xyzypr(20.05, 0.28, 0.01, 0.01, 0.0, 0.0)
    smap = b200.Map(ctx, *xyz(S))
    mprm = b200.Pt2PlParams(distanceThreshold=0.5, searchRadius=1.0, knn=8, minimumPlanePoints=5, planeEigenThreshold=0.01)
    q, _ = smap.match_pt2pl(*xyz(scan), guess, mprm)
    assert len(q) > 1000
    ref = orc.pt2pl_to_pt2pt(q, guess)
    got = ctx.pt2pl_to_pt2pt(q, guess)
    assert 3 <= len(ref) < len(q) and len(got) == len(ref)
    key = lambda a: np.sort(np.frombuffer(a.tobytes(), dtype="S36"))
    assert np.array_equal(key(ref), key(got))
    ok_c, T_c = orc.optimal_tf_horn(ref)
    ok_g, T_g = ctx.solve_horn_pt2pl(q, guess)
    ok_l, T_l = ctx.solve_horn_pt2pl(q, guess, last_match=True)
    assert ok_c and ok_g and ok_l
    assert_pose_close(T_g, T_c)
    assert_pose_close(T_l, T_g, 1e-12)
    # "at least 3": one dominant error, everything else far below 25 % of it
    few = q[:50].copy()
    few["coefs"][0, 3] += 100.0
    ref = orc.pt2pl_to_pt2pt(few, guess)
    got = ctx.pt2pl_to_pt2pt(few, guess)
    assert len(ref) == 3 and np.array_equal(key(ref), key(got))
    assert len(ctx.pt2pl_to_pt2pt(q[:0], guess)) == 0
    assert ctx.solve_horn_pt2pl(q[:2], guess)[0] is False  # fewer than 3 pairings (optimal_tf_horn.cpp:96)


# --------------------------------------------------------------------------- solver over the matcher's device copy
def test_solver_reads_last_match_device_copy(ctx):
    """MP2P_B200_PAIRS_LAST_MATCH: a solver handed the unmodified host output of the last matcher
    call reads the copy that call left on the device — same pose as uploading the records again
    (and as the oracle); a count that does not belong to the last matcher call is refused."""
    M, L, gt = _c2(200_000, 10)
    gmap = b200.Map(ctx, *xyz(M))
    guess = fx.pose_xyzypr(0.25, -0.15, 0.08, 1.7 * DEG, -0.8 * DEG, 1.2 * DEG)
    pairs, _ = gmap.match_pt2pt(*xyz(L), guess, b200.Pt2PtParams(threshold=1.0))
    ok_a, T_a = ctx.solve_horn(pairs, last_match=True)
    ok_b, T_b = ctx.solve_horn(pairs)
    ok_c, T_c = orc.optimal_tf_horn(pairs)
    assert ok_a and ok_b and ok_c
    assert_pose_close(T_a, T_b, 1e-12)
    assert_pose_close(T_a, T_c)
    gn = b200.GNParams(maxInnerLoopIterations=4)
    ok_a, T_a, it_a = ctx.solve_gauss_newton(pairs, None, gn, guess, last_match=True)
    ok_b, T_b, it_b = ctx.solve_gauss_newton(pairs, None, gn, guess)
    assert ok_a and ok_b and it_a == it_b
    assert_pose_close(T_a, T_b, 1e-12)
    with pytest.raises(b200.Mp2pError):
        ctx.solve_horn(pairs[:-1], last_match=True)
    # a later matcher call replaces the copy: the old list no longer qualifies unless the count matches
    pairs2, _ = gmap.match_pt2pt(*xyz(L[: len(L) // 2]), guess, b200.Pt2PtParams(threshold=1.0))
    assert len(pairs2) != len(pairs)
    with pytest.raises(b200.Mp2pError):
        ctx.solve_horn(pairs, last_match=True)
    ok_a, T_a = ctx.solve_horn(pairs2, last_match=True)
    assert_pose_close(T_a, orc.optimal_tf_horn(pairs2)[1])
    # pt2pl list + Gauss-Newton, and the pre-bound plugin step (matcher call + solver call over host buffers)
    S = fx.make_street_scene(n_map=200_000, length=40.0)
    scan = fx.make_lidar_scan((20.0, 0.3, 0.0), n_rings=16, n_az=400, length=40.0)
    g2 = fx.pose_xyzypr(20.05, 0.28, 0.01, 0.01, 0.0, 0.0)
    smap = b200.Map(ctx, *xyz(S))
    mprm = b200.Pt2PlParams(distanceThreshold=0.5, searchRadius=1.0, knn=8, minimumPlanePoints=5, planeEigenThreshold=0.01)
    sprm = b200.GNParams(maxInnerLoopIterations=3, kernel="GemanMcClure", kernelParam=0.15)
    q, _ = smap.match_pt2pl(*xyz(scan), g2, mprm)
    ok_a, T_a, _ = ctx.solve_gauss_newton(None, q, sprm, g2, last_match=True)
    ok_b, T_b, _ = ctx.solve_gauss_newton(None, q, sprm, g2)
    assert ok_a and ok_b
    assert_pose_close(T_a, T_b, 1e-12)
    out = np.empty(len(scan), b200.PAIR_PT2PL)
    for reuse in (True, False):
        step = smap.make_plugin_step(*xyz(scan), mprm, sprm, out, reuse_device_pairs=reuse)
        ok, T, n = step(g2)
        assert ok and n == len(q) and out[:n].tobytes() == q.tobytes()
        assert_pose_close(T, T_b, 1e-12)


def test_speculative_solver_matches_plain_calls(ctx):
    """After a solver call over the last matcher output, the next matcher call runs that solver
    speculatively while the records travel to the host; the solver call that follows must return what
    the plain two-call sequence returns (and a solver call with other parameters / another start
    pose must not be served from the speculation)."""
    M, L, gt = _c2(200_000, 10)
    gmap = b200.Map(ctx, *xyz(M))
    poses = [fx.pose_xyzypr(0.25, -0.15, 0.08, 1.7 * DEG, -0.8 * DEG, 1.2 * DEG), fx.pose_xyzypr(0.28, -0.18, 0.09, 1.9 * DEG, -0.9 * DEG, 1.4 * DEG), gt]
    mprm = b200.Pt2PtParams(threshold=1.0)
    for T in poses * 2:  # Horn
        pairs, _ = gmap.match_pt2pt(*xyz(L), T, mprm)
        ok_s, T_s = ctx.solve_horn(pairs, last_match=True)  # from the 2nd round on: speculated
        ok_p, T_p = ctx.solve_horn(pairs.copy())            # plain upload
        ok_s2, T_s2 = ctx.solve_horn(pairs, last_match=True)
        assert ok_s and ok_p and ok_s2
        assert_pose_close(T_s, T_p, 1e-12)
        assert_pose_close(T_s2, T_p, 1e-12)
    gn = b200.GNParams(maxInnerLoopIterations=4, kernel="Cauchy", kernelParam=0.3)
    for k, T in enumerate(poses * 2):  # Gauss-Newton over the pt2pt list
        pairs, _ = gmap.match_pt2pt(*xyz(L), T, mprm)
        ok_s, T_s, it_s = ctx.solve_gauss_newton(pairs, None, gn, T, last_match=True)
        ok_p, T_p, it_p = ctx.solve_gauss_newton(pairs.copy(), None, gn, T)
        assert ok_s and ok_p and it_s == it_p
        assert_pose_close(T_s, T_p, 1e-12)
        if k:  # another start pose than the matcher's: must be solved for real
            T2 = poses[(k + 1) % 3]
            pairs, _ = gmap.match_pt2pt(*xyz(L), T, mprm)
            ok_a, T_a, _ = ctx.solve_gauss_newton(pairs, None, gn, T2, last_match=True)
            ok_b, T_b, _ = ctx.solve_gauss_newton(pairs.copy(), None, gn, T2)
            assert_pose_close(T_a, T_b, 1e-12)
    S = fx.make_street_scene(n_map=200_000, length=40.0)
    scan = fx.make_lidar_scan((20.0, 0.3, 0.0), n_rings=16, n_az=400, length=40.0)
    smap = b200.Map(ctx, *xyz(S))
    pl = b200.Pt2PlParams(distanceThreshold=0.5, searchRadius=1.0, knn=8, minimumPlanePoints=5, planeEigenThreshold=0.01)
    gm = b200.GNParams(maxInnerLoopIterations=3, kernel="GemanMcClure", kernelParam=0.15)
    for T in [fx.pose_xyzypr(20.05, 0.28, 0.01, 0.01, 0.0, 0.0), fx.pose_xyzypr(20.02, 0.29, 0.0, 0.01, 0.0, 0.0)] * 2:
        q, _ = smap.match_pt2pl(*xyz(scan), T, pl)
        ok_s, T_s, it_s = ctx.solve_gauss_newton(None, q, gm, T, last_match=True)
        ok_p, T_p, it_p = ctx.solve_gauss_newton(None, q.copy(), gm, T)
        assert ok_s and ok_p and it_s == it_p
        assert_pose_close(T_s, T_p, 1e-12)


def test_pinned_host_buffers_take_the_zero_copy_path_and_change_nothing(ctx):
    """Pinned (page-locked) local arrays are read by the k = 1 search kernel in place and a pinned
    pairings buffer is written by the compaction kernel itself (no copies around the kernels): the
    records, the count and the solver results must be those of the pageable-buffer calls."""
    import torch

    M, L, gt = _c2(200_000, 10)
    gmap = b200.Map(ctx, *xyz(M))
    pinned = [torch.from_numpy(a).pin_memory() for a in xyz(L)]
    px, py, pz = (t.numpy() for t in pinned)
    out_t = torch.zeros(len(L) * 36, dtype=torch.uint8).pin_memory()
    out = out_t.numpy().view(b200.PAIR_PT2PT)
    poses = [fx.pose_xyzypr(0.25, -0.15, 0.08, 1.7 * DEG, -0.8 * DEG, 1.2 * DEG), gt, fx.pose_xyzypr(5.0, 0.0, 0.0, 0.0, 0.0, 0.0)]
    for kw in (dict(threshold=1.0), dict(threshold=1.0, allowMatchAlreadyMatchedGlobalPoints=True), dict(threshold=0.05)):
        prm = b200.Pt2PtParams(**kw)
        for T in poses * 2:
            ref, _ = gmap.match_pt2pt(*xyz(L), T, prm)
            out_t.fill_(0xAB)
            got, _ = gmap.match_pt2pt(px, py, pz, T, prm, out=out)
            assert len(got) == len(ref) and got.tobytes() == ref.tobytes()
            if len(ref) >= 3:
                ok_a, T_a = ctx.solve_horn(got, last_match=True)  # second round on: speculated behind the kernel
                ok_b, T_b = ctx.solve_horn(ref.copy())
                assert ok_a and ok_b
                assert_pose_close(T_a, T_b, 1e-12)
    # a pinned output that is too small: same error as the copy path, nothing written past its end
    small_t = torch.zeros(1000 * 36 + 36, dtype=torch.uint8).pin_memory()
    small_t[-36:] = 0xCD
    with pytest.raises(b200.Mp2pError):
        gmap.match_pt2pt(px, py, pz, gt, b200.Pt2PtParams(threshold=1.0), out=small_t.numpy()[: 1000 * 36].view(b200.PAIR_PT2PT), capacity=1000)
    assert bool((small_t[-36:] == 0xCD).all())
    # empty local cloud / k > 1 (copy path for the local arrays, zero-copy output only for k = 1)
    got, _ = gmap.match_pt2pt(px[:0], py[:0], pz[:0], gt, b200.Pt2PtParams(threshold=1.0), out=out)
    assert len(got) == 0
    prm3 = b200.Pt2PtParams(threshold=2.5, pairingsPerPoint=3)
    ref, _ = gmap.match_pt2pt(*xyz(L[:20000]), gt, prm3)
    out3 = torch.zeros(20000 * 3 * 36, dtype=torch.uint8).pin_memory()
    got, _ = gmap.match_pt2pt(px[:20000], py[:20000], pz[:20000], gt, prm3, out=out3.numpy().view(b200.PAIR_PT2PT))
    assert got.tobytes() == ref.tobytes()


# --------------------------------------------------------------------------- resident (Morton-sorted) local cloud
@pytest.mark.parametrize("kw", [dict(threshold=1.0), dict(threshold=2.5, pairingsPerPoint=3), dict(threshold=1.0, thresholdAngularDeg=0.5, allowMatchAlreadyMatchedGlobalPoints=True)])
def test_resident_cloud_pt2pt_is_invisible(ctx, kw):
    """mp2p_b200_cloud: the search walks a Morton-sorted copy, the records must not change by a bit
    (same order, same indices) — against the array path AND the oracle."""
    M, L, gt = _c2(200_000, 10)
    L = np.concatenate([L, L[:3000] + np.float32(1e-3)])  # duplicates: first-claim dedup must keep caller order
    rng = np.random.default_rng(5)
    lp = (rng.random(len(L)) < 0.2).astype(np.uint8)
    gp = (rng.random(len(M)) < 0.2).astype(np.uint8)
    tree, gmap = orc.KDTree(*xyz(M)), b200.Map(ctx, *xyz(M))
    cloud = b200.Cloud(ctx, *xyz(L))
    assert cloud.info["n_points"] == len(L)
    for pose in (np.eye(3, 4), gt):
        for bits in (dict(), dict(local_paired=lp, global_paired=gp)):
            ref, pot0 = gmap.match_pt2pt(*xyz(L), pose, b200.Pt2PtParams(**kw), **bits)
            ref = ref.copy()
            got, pot1 = gmap.match_pt2pt(cloud, None, None, pose, b200.Pt2PtParams(**kw), **bits)
            assert pot0 == pot1 and ref.tobytes() == got.tobytes()
    p0, _ = orc.match_pt2pt(tree, *xyz(L), gt, orc.MatchPt2PtParams(**kw), lp.copy(), gp.copy(), nthreads=8)
    assert len(p0) > 1000 and p0.tobytes() == got.tobytes()
    with pytest.raises(b200.Mp2pError):
        gmap.match_pt2pt(cloud, None, None, gt, b200.Pt2PtParams(threshold=-1.0))


def test_resident_cloud_pt2pl_and_fused_iterations(ctx):
    M = fx.make_street_scene(n_map=300_000, length=50.0)
    S = fx.make_lidar_scan((25.0, 0.5, 0.0), n_rings=32, n_az=500, length=50.0)
    guess = fx.pose_xyzypr(25.08, 0.46, 0.02, 0.02, 0.001, -0.001)
    gmap, cloud = b200.Map(ctx, *xyz(M)), b200.Cloud(ctx, *xyz(S))
    mkw = dict(distanceThreshold=0.5, searchRadius=1.0, knn=8, minimumPlanePoints=5, planeEigenThreshold=0.01)
    skw = dict(maxInnerLoopIterations=3, kernel="GemanMcClure", kernelParam=0.15)
    lp = (np.random.default_rng(6).random(len(S)) < 0.2).astype(np.uint8)
    for bits in (dict(), dict(local_paired=lp)):
        ref, _ = gmap.match_pt2pl(*xyz(S), guess, b200.Pt2PlParams(**mkw), **bits)
        ref = ref.copy()
        got, _ = gmap.match_pt2pl(cloud, None, None, guess, b200.Pt2PlParams(**mkw), **bits)
        assert len(ref) > 2000 and ref.tobytes() == got.tobytes()
    ref, _ = gmap.match_pt2pl(*xyz(S), guess, b200.Pt2PlParams(**mkw))
    ok_ref, T_ref, _ = ctx.solve_gauss_newton(None, ref, b200.GNParams(**skw), guess)
    ok, T, n = gmap.make_iterator(cloud, None, None, None, b200.Pt2PlParams(**mkw), b200.GNParams(**skw))(guess)
    assert ok and ok_ref and n == len(ref)
    assert_pose_close(T, T_ref, 1e-9)
    # pt2pt + Horn on the same data
    mprm = b200.Pt2PtParams(threshold=0.5)
    pairs, _ = gmap.match_pt2pt(*xyz(S), guess, mprm)
    ok_ref, T_ref = ctx.solve_horn(pairs)
    ok, T, n = gmap.make_iterator(cloud, None, None, None, mprm, b200.HornParams())(guess)
    assert ok and ok_ref and n == len(pairs)
    assert_pose_close(T, T_ref, 1e-9)


def test_resident_cloud_edge_cases(ctx):
    z = np.zeros(0, np.float32)
    M, L, gt = _c2(20_000, 10)
    gmap = b200.Map(ctx, *xyz(M))
    empty = b200.Cloud(ctx, z, z, z)
    p, pot = gmap.match_pt2pt(empty, None, None, gt, b200.Pt2PtParams(threshold=1.0))
    assert len(p) == 0 and pot == 0
    for n in (1, 31, 257):  # ragged tiles
        c = b200.Cloud(ctx, *xyz(L[:n]))
        ref, _ = gmap.match_pt2pt(*xyz(L[:n]), gt, b200.Pt2PtParams(threshold=1.0))
        got, _ = gmap.match_pt2pt(c, None, None, gt, b200.Pt2PtParams(threshold=1.0))
        assert ref.tobytes() == got.tobytes()
    same = np.repeat(L[:1], 1000, axis=0)  # zero-extent cloud
    c = b200.Cloud(ctx, *xyz(same))
    ref, _ = gmap.match_pt2pt(*xyz(same), gt, b200.Pt2PtParams(threshold=1.0))
    got, _ = gmap.match_pt2pt(c, None, None, gt, b200.Pt2PtParams(threshold=1.0))
    assert len(ref) == 1 and ref.tobytes() == got.tobytes()


def test_stale_and_foreign_handles_are_refused(ctx):
    """A destroyed cloud / map handle, or one created on another context, is an argument error — not a
    dereference of freed memory."""
    import ctypes as C

    M, L, gt = _c2(20_000, 10)
    gmap = b200.Map(ctx, *xyz(M))
    cloud = b200.Cloud(ctx, *xyz(L))
    prm = b200.Pt2PtParams(threshold=1.0)
    ok, _ = gmap.match_pt2pt(cloud, None, None, gt, prm)
    assert len(ok) > 100
    stale = b200.Cloud.__new__(b200.Cloud)  # a wrapper around the raw handle value that outlives the cloud
    stale.ctx, stale._h, stale.n = ctx, C.c_void_p(cloud._h.value), cloud.n
    cloud.close()
    with pytest.raises(b200.Mp2pError):
        gmap.match_pt2pt(stale, None, None, gt, prm)
    stale._h = None
    other = b200.Context(0)
    try:
        omap, ocloud = b200.Map(other, *xyz(M)), b200.Cloud(other, *xyz(L))
        with pytest.raises(b200.Mp2pError):  # a cloud of another context
            gmap.match_pt2pt(ocloud, None, None, gt, prm)
        foreign = b200.Map.__new__(b200.Map)  # `omap`'s handle presented as a map of `ctx`
        foreign.ctx, foreign._h, foreign.n = ctx, C.c_void_p(omap._h.value), omap.n
        with pytest.raises(b200.Mp2pError):
            foreign.match_pt2pt(*xyz(L), gt, prm)
        foreign._h = None
    finally:
        other.close()
    again, _ = gmap.match_pt2pt(*xyz(L), gt, prm)  # the context is still healthy
    assert again.tobytes() == ok.tobytes()


# --------------------------------------------------------------------------- KITTI .bin straight to the device (§8f N4)
def test_kitti_bin_loader_and_xyzi_entry_points(ctx, tmp_path):
    """A KITTI velodyne file (float32 x, y, z, intensity records, apps/kitti2mm/main.cpp:55-69) read into
    pinned memory and split on the device: map and cloud built from the interleaved records must behave
    exactly like the ones built from SoA arrays."""
    import torch

    Ms = fx.make_street_scene(n_map=150_000, length=40.0)
    S = fx.make_lidar_scan((20.0, 0.4, 0.0), n_rings=32, n_az=400, length=40.0)
    rng = np.random.default_rng(0)
    scan4 = np.concatenate([S, rng.random((len(S), 1), dtype=np.float32)], axis=1).astype(np.float32)
    path = str(tmp_path / "000000.bin")
    scan4.tofile(path)
    got = b200.capi.read_kitti_bin(path)
    assert got.shape == scan4.shape and got.dtype == np.float32 and np.array_equal(got, scan4)
    T = fx.pose_xyzypr(20.05, 0.38, 0.01, 0.01, 0.0, 0.0)
    smap = b200.Map(ctx, *xyz(Ms))
    kw = b200.Pt2PlParams(distanceThreshold=0.5, searchRadius=1.0, knn=8, minimumPlanePoints=5, planeEigenThreshold=0.01)
    ref, _ = smap.match_pt2pl(b200.Cloud(ctx, *xyz(S)), None, None, T, kw)
    c_host = b200.Cloud.from_xyzi(ctx, got)  # pinned host records
    d4 = torch.from_numpy(scan4).cuda()
    c_dev = b200.Cloud.from_xyzi(ctx, d4.data_ptr(), n=len(scan4), on_device=True)
    for c in (c_host, c_dev):
        assert c.n == len(S)
        p, _ = smap.match_pt2pl(c, None, None, T, kw)
        assert len(ref) > 1000 and p.tobytes() == ref.tobytes()
    map4 = np.concatenate([Ms, np.zeros((len(Ms), 1), np.float32)], axis=1)
    m2 = b200.Map.from_xyzi(ctx, map4)
    q, _ = m2.match_pt2pl(c_host, None, None, T, kw)
    assert q.tobytes() == ref.tobytes()
    a, _ = m2.match_pt2pt(c_dev, None, None, T, b200.Pt2PtParams(threshold=0.3))
    b, _ = smap.match_pt2pt(*xyz(S), T, b200.Pt2PtParams(threshold=0.3))
    assert len(a) > 1000 and a.tobytes() == b.tobytes()
    assert b200.Cloud.from_xyzi(ctx, scan4[:0]).n == 0
    with pytest.raises(b200.Mp2pError):
        b200.capi.read_kitti_bin(str(tmp_path / "missing.bin"))
    (tmp_path / "bad.bin").write_bytes(b"\0" * 20)
    with pytest.raises(b200.Mp2pError):  # not a whole number of 16-byte records
        b200.capi.read_kitti_bin(str(tmp_path / "bad.bin"))


# --------------------------------------------------------------------------- multi-GPU (needs >= 2 GPUs)
@pytest.mark.parametrize("transport", ["peer", "peer_replay", "nccl"])
def test_multi_gpu_sharded_iteration_equals_single_gpu(transport):
    import subprocess
    import sys

    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run under `gpurun --gpus 2`)")
    world = 2 if n < 4 else 4
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", "29571", os.path.join(here, "multi_gpu_parity_worker.py")]
    # "peer" = NVLink mailboxes + owner-partitioned claims (the default), "peer_replay" = mailboxes with the record
    # all-gather and replicated claim replay of round 1, "nccl" = torch.distributed collectives
    env = dict(os.environ, MP2P_B200_TRANSPORT="peer" if transport.startswith("peer") else transport,
               MP2P_B200_OWNER_CLAIMS="0" if transport == "peer_replay" else "1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "identical=True" in r.stdout and f"transport={env['MP2P_B200_TRANSPORT']}" in r.stdout


# ---------------------------------------------------------------------------------------------
# FilterDecimateVoxels on the device: points AND order equal to the oracle
# (FilterDecimateVoxels.cpp:109-378; order = the reference's std::map walk, ascending (cx, cy, cz))
# ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("method", ["FirstPoint", "ClosestToAverage", "VoxelAverage"])
@pytest.mark.parametrize("flatten_to", [None, -1.5])
def test_filter_decimate_voxels_bit_exact(ctx, method, flatten_to):
    rng = np.random.default_rng(11)
    # a scan-shaped cloud around the origin: negative coordinates exercise the truncation toward zero
    P = np.concatenate([fx.make_lidar_scan((20.0, 0.3, 0.0), n_rings=16, n_az=600, length=40.0), rng.uniform(-3, 3, (5000, 3)).astype(np.float32)])
    x, y, z = (np.ascontiguousarray(P[:, k]) for k in range(3))
    for res in (2.0, 0.35):
        want_xyz, want_src = orc.decimate_voxels(x, y, z, res, method, flatten_to)
        got_xyz, got_src = ctx.decimate_voxels(x, y, z, res, method, flatten_to)
        assert len(got_xyz) == len(want_xyz) > 50
        assert got_xyz.tobytes() == want_xyz.tobytes(), (method, flatten_to, res)
        assert np.array_equal(got_src, want_src)


@pytest.mark.gpu
def test_filter_decimate_voxels_edge_cases_and_resident_cloud(ctx):
    x, y, z = (np.ascontiguousarray(a) for a in (np.array([0.1, 0.2, -0.1, 5.0, 5.1], np.float32), np.zeros(5, np.float32), np.zeros(5, np.float32)))
    got, src = ctx.decimate_voxels(x, y, z, 1.0)  # -0.1 and 0.1 truncate to the SAME voxel 0 (PointCloudToVoxelGridSingle.h:105)
    assert src.tolist() == [0, 3] and got.tobytes() == np.array([[0.1, 0, 0], [5.0, 0, 0]], np.float32).tobytes()
    e, es = ctx.decimate_voxels(x[:0], y[:0], z[:0], 1.0)
    assert len(e) == 0 and len(es) == 0
    one, osrc = ctx.decimate_voxels(x[:1], y[:1], z[:1], 0.5, "VoxelAverage")
    assert osrc.tolist() == [-1] and one.tobytes() == np.array([[0.1, 0, 0]], np.float32).tobytes()
    with pytest.raises(b200.Mp2pError):
        ctx.decimate_voxels(x, y, z, 1.0, "RandomPoint")  # unseeded generator upstream: refused
    with pytest.raises(b200.Mp2pError):
        ctx.decimate_voxels(x, y, z, 0.0)
    with pytest.raises(b200.Mp2pError):  # voxel indices spanning more than 64 bits
        ctx.decimate_voxels(np.array([-3e9, 3e9], np.float32), np.array([-3e9, 3e9], np.float32), np.array([-3e9, 3e9], np.float32), 1.0)
    # filter -> resident cloud -> matcher without leaving the device == matcher over the oracle's decimated cloud
    M, L, gt = fx.make_c2(n_map=50_000, decim=5)
    gmap = b200.Map(ctx, *xyz(M))
    want_xyz, _ = orc.decimate_voxels(*xyz(L), 4.0)
    cloud = b200.Cloud.decimated(ctx, *xyz(L), 4.0)
    assert cloud.n == len(want_xyz)
    prm = b200.Pt2PtParams(threshold=1.0)
    a, _ = gmap.match_pt2pt(cloud, None, None, np.eye(3, 4), prm)
    b, _ = gmap.match_pt2pt(*xyz(want_xyz), np.eye(3, 4), prm)
    assert len(a) > 10 and a.tobytes() == b.tobytes()


# ---------------------------------------------------------------------------------------------
# covariance() on the device (covariance.cpp:28-141) vs the oracle's restatement
# ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_covariance_matches_oracle(ctx):
    rng = np.random.default_rng(3)
    x6 = np.array([0.3, -0.2, 0.0, 0.2, -0.1, 0.05])  # z slot = 0: what the reference evaluates at (covariance.cpp:41-47)
    T = orc.pose_from_xyzypr(*x6)
    n, m, k = 5000, 3000, 700
    L = rng.uniform(-20, 20, (n, 3))
    p2p = np.zeros(n, b200.PAIR_PT2PT)
    p2p["global"], p2p["local"] = (L @ T[:, :3].T + T[:, 3] + rng.normal(0, 0.02, (n, 3))).astype(np.float32), L.astype(np.float32)
    Lp = rng.uniform(-20, 20, (m, 3))
    nrm = rng.normal(size=(m, 3))
    nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    p2l = np.zeros(m, b200.PAIR_PT2PL)
    p2l["coefs"][:, :3] = nrm
    p2l["coefs"][:, 3] = -(nrm * (Lp @ T[:, :3].T + T[:, 3])).sum(1) + rng.normal(0, 0.02, m)
    p2l["local"] = Lp.astype(np.float32)
    Ln = rng.uniform(-20, 20, (k, 3))
    u = rng.normal(size=(k, 3))
    u /= np.linalg.norm(u, axis=1)[:, None]
    p2ln = np.zeros(k, b200.PAIR_PT2LN)
    p2ln["pBase"], p2ln["director"], p2ln["local"] = Ln @ T[:, :3].T + T[:, 3] + rng.normal(0, 0.02, (k, 3)), u, Ln
    for lists in ((p2p, None, None), (None, p2l, None), (p2p, p2l, p2ln)):
        cov_c, hes_c = orc.covariance(*lists, x6)
        cov_g, hes_g, pd = ctx.covariance(*lists, x6)
        assert pd
        assert np.allclose(hes_g, hes_c, rtol=1e-9, atol=1e-9 * np.abs(hes_c).max())
        assert np.allclose(cov_g, cov_c, rtol=1e-6, atol=1e-9 * np.abs(cov_c).max())
    c0, _, _ = ctx.covariance(None, None, None, x6)
    assert np.array_equal(c0, np.eye(6) * 1e6)  # no pairings (covariance.cpp:33-38)
    # a single pt2pl pairing: the hessian is singular -> reported, no exception
    _, _, pd1 = ctx.covariance(None, p2l[:1], None, x6)
    assert not pd1


@pytest.mark.gpu
def test_solver_over_host_pairings_is_checked_against_the_device_copy(ctx):
    """The safe default of the plugin (no assumeUnmodifiedPairings): the solver uploads the host pairings and the
    library compares them byte for byte with the copy the matcher call left on the device. Unmodified -> the
    result the matcher call computed ahead of time is handed out; ANY edit in between -> the solve runs over
    the edited records (here: the oracle's result over the edited list)."""
    S = fx.make_street_scene(n_map=300_000, length=60.0)
    scan = fx.make_lidar_scan((30.0, 0.3, 0.0), n_rings=32, n_az=600, length=60.0)
    guess = fx.pose_xyzypr(30.05, 0.28, 0.01, 0.01, 0.0, 0.0)
    smap = b200.Map(ctx, *xyz(S))
    mprm = b200.Pt2PlParams(distanceThreshold=0.5, searchRadius=1.0, knn=8, minimumPlanePoints=5, planeEigenThreshold=0.01)
    gprm = b200.GNParams(maxInnerLoopIterations=3, kernel="GemanMcClure", kernelParam=0.15)
    ogp = orc.GNParams(maxInnerLoopIterations=3, kernel="GemanMcClure", kernelParam=0.15)
    for rep in range(4):  # the matcher starts computing ahead once it has seen the solver's request
        pairs, _ = smap.match_pt2pl(*xyz(scan), guess, mprm)
        assert len(pairs) > 500
        if rep == 3:  # an edit of ONE record in the middle of the list (what a sampled witness would miss)
            pairs = pairs.copy()
            pairs["coefs"][len(pairs) // 2 + 1, 3] += 0.05
        ok, T, _ = ctx.solve_gauss_newton(None, pairs, gprm, guess)
        _, T_ref, _ = orc.optimal_tf_gauss_newton(None, pairs, ogp, guess, nthreads=4)
        assert ok
        assert_pose_close(T, T_ref, 1e-9)
    # the same for pt2pt + Horn
    M, L, gt = fx.make_c2(n_map=100_000, decim=10)
    gmap = b200.Map(ctx, *xyz(M))
    for rep in range(4):
        p2p, _ = gmap.match_pt2pt(*xyz(L), np.eye(3, 4), b200.Pt2PtParams(threshold=1.0))
        if rep == 3:
            p2p = p2p.copy()
            p2p["global"][len(p2p) // 3, 0] += 0.5
        ok, T = ctx.solve_horn(p2p)
        _, T_ref = orc.optimal_tf_horn(p2p)
        assert ok
        assert_pose_close(T, T_ref, 1e-9)
