"""Pins the CPU oracle against every known-answer fixture the reference's own tests hold for the
Matcher+Solver path. Each test names the reference test it ports.

CPU-only (`-m "not gpu"`).
"""
import os

import numpy as np
import pytest

from oracle import oracle_py as orc
from tests import icp_harness
from tests.fixtures import (
    pt2pl_fixture_global,
    pt2pt_fixture_global,
    two_local_points,
)

GOLD = os.path.join(os.path.dirname(__file__), "golden")
DEG = np.pi / 180.0


# ---------------------------------------------------------------------------------------------
# tests/test-mp2p_matcher_pt2pt.cpp:26-107 — exact (localIdx, globalIdx) for 3 poses + identity
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize(
    "pose,expected",
    [
        ((0, 0, 0, 0, 0, 0), []),  # :68-74
        ((0, 5, 0, 0, 0, 0), [(0, 0)]),  # :78-86  (localIdx, globalIdx)
        ((-2, 5, 0, 0, 0, 0), [(1, 0)]),  # :88-96
        ((8.5, -1.0, 1, 45 * DEG, 0, 0), [(1, 19)]),  # :98-107
    ],
)
def test_matcher_pt2pt_known_answers(pose, expected):
    gx, gy, gz = pt2pt_fixture_global()
    lx, ly, lz = two_local_points()
    tree = orc.KDTree(gx, gy, gz)
    prm = orc.MatchPt2PtParams(threshold=1.05, thresholdAngularDeg=0.001)
    pairs, pot = orc.match_pt2pt(tree, lx, ly, lz, orc.pose_from_xyzypr(*pose), prm)
    assert [(int(p["localIdx"]), int(p["globalIdx"])) for p in pairs] == expected
    assert pot == 2


# ---------------------------------------------------------------------------------------------
# tests/test-mp2p_matcher_pt2pl.cpp:29-168 (DISABLED upstream, tests/CMakeLists.txt:37): the only
# known answers that exist for pt2pl matching.
# ---------------------------------------------------------------------------------------------
PT2PL_PRM = dict(distanceThreshold=0.1, searchRadius=0.1, minimumPlanePoints=5, knn=5, planeEigenThreshold=0.1)


def test_matcher_pt2pl_known_answers():
    gx, gy, gz = pt2pl_fixture_global()
    lx, ly, lz = two_local_points()
    tree = orc.KDTree(gx, gy, gz)
    prm = orc.MatchPt2PlParams(**PT2PL_PRM)

    pairs, _ = orc.match_pt2pl(tree, lx, ly, lz, orc.pose_from_xyzypr(0, 0, 0), prm)
    assert len(pairs) == 0  # :85-91

    pairs, _ = orc.match_pt2pl(tree, lx, ly, lz, orc.pose_from_xyzypr(0, 5, 0), prm)
    assert len(pairs) == 1  # :93-100

    pairs, _ = orc.match_pt2pl(tree, lx, ly, lz, orc.pose_from_xyzypr(8.04, 0, 0), prm)
    assert len(pairs) == 1  # :102-123
    p0 = pairs[0]
    np.testing.assert_allclose(p0["local"], [2.0, 0.0, 0.0], atol=1e-3)
    np.testing.assert_allclose(p0["centroid"], [10.0, 0.0, 0.0], atol=0.01)
    np.testing.assert_allclose(p0["coefs"], [1.0, 0.0, 0.0, -10.0], atol=1e-3)

    pairs, _ = orc.match_pt2pl(tree, lx, ly, lz, orc.pose_from_xyzypr(18.053, 0.05, 0.03), prm)
    assert len(pairs) == 0  # :125-131


@pytest.mark.parametrize("allow,expected_total", [(True, 2), (False, 1)])
def test_matcher_pt2pl_then_pt2pt_pipeline(allow, expected_total):
    """:133-168 — run_matchers({pt2pl, pt2pt}) sharing one MatchState (Matcher.cpp:46-88)."""
    gx, gy, gz = pt2pl_fixture_global()
    lx, ly, lz = two_local_points()
    tree = orc.KDTree(gx, gy, gz)
    T = orc.pose_from_xyzypr(8.04, 0, 0)
    local_paired = np.zeros(2, np.uint8)
    global_paired = np.zeros(tree.n, np.uint8)
    p2l, _ = orc.match_pt2pl(tree, lx, ly, lz, T, orc.MatchPt2PlParams(**PT2PL_PRM), local_paired)
    p2p, _ = orc.match_pt2pt(
        tree, lx, ly, lz, T,
        orc.MatchPt2PtParams(threshold=0.1, thresholdAngularDeg=0.0, allowMatchAlreadyMatchedPoints=allow),
        local_paired, global_paired,
    )
    assert len(p2l) == 1
    assert len(p2l) + len(p2p) == expected_total


# ---------------------------------------------------------------------------------------------
# tests/test-mp2p_optimize_pt2pl.cpp:36-128 — GN on 3 pt2pl + 1 pt2pt recovers 15 GT poses (1e-3)
# ---------------------------------------------------------------------------------------------
GT_POSES_PT2PL = [
    (0, 0, 0, 0, 0, 0), (1, 0, 0, 0, 0, 0), (0, 1, 0, 0, 0, 0), (0, 0, 1, 0, 0, 0),
    (-2, 0, 0, 0, 0, 0), (0, -3, 0, 0, 0, 0), (0, 0, -4, 0, 0, 0),
    (0, 0, 0, 20 * DEG, 0, 0), (0, 0, 0, -20 * DEG, 0, 0),
    (0, 0, 0, 0, 10 * DEG, 0), (0, 0, 0, 0, -10 * DEG, 0),
    (0, 0, 0, 0, 0, 15 * DEG), (0, 0, 0, 0, 0, -15 * DEG),
    (1, 2, 3, 0, 0, 0), (1, 2, 3, -10 * DEG, 5 * DEG, 30 * DEG),
]


def _inv_compose_point(T, g):
    T = np.asarray(T).reshape(3, 4)
    return T[:, :3].T @ (np.asarray(g, float) - T[:, 3])


def make_pt2pl_case(gt):
    p2l = np.zeros(3, orc.PAIR_PT2PL)
    for k, (normal, g) in enumerate([((0, 0, 1), (0.5, 0, 0)), ((1, 0, 0), (0, 0.8, 0)), ((0, 1, 0), (0, 0, 0.3))]):
        p2l[k]["coefs"] = [*normal, 0.0]  # TPlane::FromPointAndNormal({0,0,0}, n)
        p2l[k]["centroid"] = 0
        p2l[k]["local"] = _inv_compose_point(gt, g).astype(np.float32)
    p2p = np.zeros(1, orc.PAIR_PT2PT)
    p2p[0]["global"] = 0
    p2p[0]["local"] = _inv_compose_point(gt, (0, 0, 0)).astype(np.float32)
    return p2p, p2l


@pytest.mark.parametrize("gt", GT_POSES_PT2PL)
def test_gn_pt2pl_known_answers(gt):
    GT = orc.pose_from_xyzypr(*gt)
    p2p, p2l = make_pt2pl_case(GT)
    ok, T, _ = orc.optimal_tf_gauss_newton(p2p, p2l, orc.GNParams(maxInnerLoopIterations=25), np.eye(3, 4))
    assert ok
    assert np.linalg.norm(orc.se3_log(orc.inverse_compose(T, GT))) < 1e-3


# tests/test-mp2p_optimize_with_prior.cpp:38-52,81-89 — case 0 (no prior): 3 pt2pt pairings
def test_gn_pt2pt_no_prior_known_answer():
    GT = orc.pose_from_xyzypr(1.0, 2.0, 3.0, -10 * DEG, 5 * DEG, 30 * DEG)
    p2p = np.zeros(3, orc.PAIR_PT2PT)
    for k, g in enumerate(np.eye(3)):
        p2p[k]["global"] = g
        p2p[k]["local"] = _inv_compose_point(GT, g).astype(np.float32)
    ok, T, _ = orc.optimal_tf_gauss_newton(p2p, None, orc.GNParams(maxInnerLoopIterations=25), np.eye(3, 4))
    assert ok
    assert np.linalg.norm(orc.se3_log(orc.inverse_compose(T, GT))) < 1e-3


# ---------------------------------------------------------------------------------------------
# tests/test-mp2p_error_terms_jacobians.cpp:45-102,186-252 — analytic J vs finite differences
# on D*exp(eps), step 1e-6, tolerance 1e-5, 1000 random draws, seed 1234.
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", [0, 1])
def test_jacobians_vs_finite_differences(kind):
    rng = np.random.default_rng(1234)
    for _ in range(1000):
        D = orc.pose_from_xyzypr(*rng.normal(0, 10, 3), rng.uniform(-np.pi, np.pi), *rng.uniform(-np.pi / 2, np.pi / 2, 2))
        if kind == 0:
            pair = np.zeros(1, orc.PAIR_PT2PT)
            pair["global"] = rng.normal(0, 20, 3)
            pair["local"] = rng.normal(0, 20, 3)
        else:
            pair = np.zeros(1, orc.PAIR_PT2PL)
            n = rng.normal(0, 1, 3)
            n /= np.linalg.norm(n)
            c = rng.normal(0, 20, 3)
            pair["coefs"] = [*n, -n @ c]
            pair["centroid"] = c
            pair["local"] = rng.normal(0, 20, 3)
        _, J = orc.error_and_jacobian(kind, pair, D)
        num = np.zeros((3, 6))
        for a in range(6):
            ep, em = np.zeros(6), np.zeros(6)
            ep[a], em[a] = 1e-6, -1e-6
            e_p, _ = orc.error_and_jacobian(kind, pair, orc.compose(D, orc.se3_exp(ep)))
            e_m, _ = orc.error_and_jacobian(kind, pair, orc.compose(D, orc.se3_exp(em)))
            num[:, a] = (e_p - e_m) / 2e-6
        assert np.abs(num - J).max() < 1e-5


# ---------------------------------------------------------------------------------------------
# tests/test-mp2p_optimal_tf_algos.cpp:49-63,112-157,285,369-377 — Horn on random pairings with
# sigma_xyz = 1 mm noise: SO(3) error < min(1, 0.2 + 10*sigma_xyz + 50*sigma_n) (outlier-free).
# We also check the far stronger property that the noiseless case is exact.
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [3, 4, 10, 100, 1000])
def test_horn_random_pairings(n):
    rng = np.random.default_rng(1234 + n)
    for sigma in (0.0, 1e-3):
        A = rng.uniform(0, 50, (n, 3))
        gt = orc.pose_from_xyzypr(*rng.uniform(-1, 1, 3), *(rng.uniform(-5, 5, 3) * DEG))
        B = (A - gt[:, 3]) @ gt[:, :3] + rng.normal(0, sigma, (n, 3)) if sigma else (A - gt[:, 3]) @ gt[:, :3]
        pairs = np.zeros(n, orc.PAIR_PT2PT)
        pairs["globalIdx"] = pairs["localIdx"] = np.arange(n)
        pairs["global"], pairs["local"] = A, B
        ok, T = orc.optimal_tf_horn(pairs)
        assert ok
        err = orc.se3_log(orc.inverse_compose(T, gt))
        if sigma == 0.0:
            assert np.linalg.norm(err) < 2e-4  # float32 storage of the pairings
        else:
            assert np.linalg.norm(err[3:]) < min(1.0, 0.2 + 10 * sigma)


def test_horn_needs_three_pairings():
    pairs = np.zeros(2, orc.PAIR_PT2PT)  # optimal_tf_horn.cpp:96
    ok, _ = orc.optimal_tf_horn(pairs)
    assert not ok


# ---------------------------------------------------------------------------------------------
# tests/test-mp2p_icp_algos.cpp:85-108,161-169,187-223 — full ICP::align on bunny / buddha:
# GT pose within +-15 % bbox and +-10 deg, threshold = 0.40*max_dim, thresholdAngularDeg = 0,
# maxIterations 100, identity guess; assert |log(GT - est)| < 0.1.
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("model", ["bunny_decim.xyz.gz", "happy_buddha_decim.xyz.gz"])
@pytest.mark.parametrize("solver", ["horn", "gn"])
def test_icp_align_protocol(model, solver):
    x, y, z = icp_harness.load_xyz_gz(os.path.join(GOLD, model))
    x, y, z = x[::10], y[::10], z[::10]  # decimation 10 (:250-262)
    P = np.stack([x, y, z], 1).astype(np.float64)
    size = P.max(0) - P.min(0)
    rng = np.random.default_rng(1234)
    tree = orc.KDTree(x, y, z)
    for _ in range(3):
        gt = orc.pose_from_xyzypr(*(rng.uniform(-0.15, 0.15, 3) * size), *(rng.uniform(-10, 10, 3) * DEG))
        L = ((P - gt[:, 3]) @ gt[:, :3]).astype(np.float32)  # changeCoordinatesReference(pts, -gt)
        lx, ly, lz = (np.ascontiguousarray(L[:, k]) for k in range(3))
        prm = orc.MatchPt2PtParams(threshold=0.40 * size.max(), thresholdAngularDeg=0.0)

        def match(pose, it):
            return orc.match_pt2pt(tree, lx, ly, lz, pose, prm, nthreads=4)[0]

        def solve(pairs, guess, it):
            if solver == "horn":
                return orc.optimal_tf_horn(pairs)
            ok, T, _ = orc.optimal_tf_gauss_newton(pairs, None, orc.GNParams(maxInnerLoopIterations=6), guess)
            return ok, T

        res = icp_harness.align(match, solve, np.eye(3, 4), icp_harness.IcpParams(maxIterations=100))
        assert np.linalg.norm(orc.se3_log(orc.inverse_compose(res.pose, gt))) < 0.1


# ---------------------------------------------------------------------------------------------
# Oracle-internal cross checks (SURVEY §4 plan item 2): KD-tree == brute force == scipy cKDTree
# ---------------------------------------------------------------------------------------------
def test_kdtree_matches_bruteforce_and_scipy():
    from scipy.spatial import cKDTree

    rng = np.random.default_rng(7)
    M = rng.uniform(0, 20, (20000, 3)).astype(np.float32)
    Q = rng.uniform(-1, 21, (2000, 3)).astype(np.float32)
    tree = orc.KDTree(M[:, 0], M[:, 1], M[:, 2])
    for k, r2 in [(1, np.inf), (5, np.inf), (8, 1.5), (3, 0.05)]:
        i1, d1, f1 = tree.knn(Q[:, 0], Q[:, 1], Q[:, 2], k, r2)
        i2, d2, f2 = tree.knn(Q[:, 0], Q[:, 1], Q[:, 2], k, r2, bruteforce=True)
        assert np.array_equal(f1, f2)
        for q in range(len(Q)):
            assert np.array_equal(i1[q, : f1[q]], i2[q, : f1[q]])
            assert np.array_equal(d1[q, : f1[q]], d2[q, : f1[q]])
    # scipy (float64 metric): identical 1-NN except float32-rounding near-ties
    i1, _, _ = tree.knn(Q[:, 0], Q[:, 1], Q[:, 2], 1)
    _, isp = cKDTree(M.astype(np.float64)).query(Q.astype(np.float64), k=1)
    assert (i1[:, 0] != isp).sum() <= 2


def test_kdtree_ties_lowest_index_wins():
    # duplicate points: nanoflann's winner is traversal dependent (unpinned); ours is the lowest index
    x = np.array([1, 1, 1, 5, 1], np.float32)
    tree = orc.KDTree(x, np.zeros(5, np.float32), np.zeros(5, np.float32), leaf_max=1)
    i, d, f = tree.knn(np.array([1.0]), np.array([0.0]), np.array([0.0]), 3)
    assert list(i[0]) == [0, 1, 2] and f[0] == 3


def test_transform_is_double_then_float():
    """Matcher_Points_Base.cpp:216 — composePoint(float...) computes in double, rounds once."""
    rng = np.random.default_rng(3)
    l = rng.normal(0, 50, (1000, 3)).astype(np.float32)
    T = orc.pose_from_xyzypr(3.3, -7.1, 0.4, 0.7, -0.2, 0.1)
    gx, gy, gz, bmin, bmax = orc.transform_local_to_global(l[:, 0], l[:, 1], l[:, 2], T)
    ref = (l.astype(np.float64) @ T[:, :3].T + T[:, 3])
    # numpy's matmul may reassociate; allow 1 ulp but require exact equality for the explicit formula
    exp = np.empty_like(ref)
    for r in range(3):
        exp[:, r] = ((T[r, 0] * l[:, 0].astype(np.float64) + T[r, 1] * l[:, 1].astype(np.float64)) + T[r, 2] * l[:, 2].astype(np.float64)) + T[r, 3]
    assert np.array_equal(np.stack([gx, gy, gz], 1), exp.astype(np.float32))
    assert np.allclose(ref, exp)
    assert np.array_equal(bmin, exp.astype(np.float32).min(0)) and np.array_equal(bmax, exp.astype(np.float32).max(0))


# ---------------------------------------------------------------------------------------------
# Matcher_Points_InlierRatio. The reference pins it only through the ICP protocol of
# tests/test-mp2p_icp_algos.cpp:250-262 (class default inliersRatio = 0.80,
# Matcher_Points_InlierRatio.h:56); the hand-made cases below pin the restatement's reading of
# Matcher_Points_InlierRatio.cpp:41-143 (sorted emission, reverse-insertion tie order of
# multimap::emplace_hint(begin()), mrpt::round, first-in-sorted-order dedup, bitfields).
# ---------------------------------------------------------------------------------------------
def test_inlier_ratio_semantics():
    g = np.array([[0, 0, 0], [10, 0, 0], [20, 0, 0], [30, 0, 0]], np.float32)
    tree = orc.KDTree(*(np.ascontiguousarray(g[:, k]) for k in range(3)))
    # local points: distances to their NN 0.5 (g0), 0.25 (g1), 0.25 (g1, tie with the previous), 1.0 (g2), 2.0 (g3)
    L = np.array([[0.5, 0, 0], [10.25, 0, 0], [9.75, 0, 0], [21, 0, 0], [32, 0, 0]], np.float32)
    lx, ly, lz = (np.ascontiguousarray(L[:, k]) for k in range(3))
    I = np.eye(3, 4)
    # nTotal = 5, ratio 0.5 -> mrpt::round(2.5) = 2 (lrint: ties to even); sorted: d2 = 0.0625 twice
    # (local 2 BEFORE local 1: reverse insertion order), then 0.25, 1, 4
    p, pot = orc.match_inlier_ratio(tree, lx, ly, lz, I, orc.MatchInlierRatioParams(0.5, allowMatchAlreadyMatchedGlobalPoints=True))
    assert pot == 5 and list(p["localIdx"]) == [2, 1] and list(p["globalIdx"]) == [1, 1]
    # same with dedup: local 2 takes g1, local 1 is dropped (its global point is taken), NOT replaced
    lp, gp = np.zeros(5, np.uint8), np.zeros(4, np.uint8)
    p, _ = orc.match_inlier_ratio(tree, lx, ly, lz, I, orc.MatchInlierRatioParams(0.5), lp, gp)
    assert list(p["localIdx"]) == [2] and list(lp) == [0, 0, 1, 0, 0] and list(gp) == [0, 1, 0, 0]
    # ratio 0.7 -> round(3.5) = 4 (even); ratio 0.9 -> round(4.5) = 4; ratio 0.95 -> round(4.75) = 5
    for ratio, keep in [(0.7, 4), (0.9, 4), (0.95, 5), (0.1, 0)]:
        p, _ = orc.match_inlier_ratio(tree, lx, ly, lz, I, orc.MatchInlierRatioParams(ratio, allowMatchAlreadyMatchedGlobalPoints=True))
        assert len(p) == keep
    p, _ = orc.match_inlier_ratio(tree, lx, ly, lz, I, orc.MatchInlierRatioParams(0.95, allowMatchAlreadyMatchedGlobalPoints=True))
    assert list(p["localIdx"]) == [2, 1, 0, 3, 4] and np.allclose(p["errSq"], [0.0625, 0.0625, 0.25, 1.0, 4.0])
    assert np.array_equal(p["local"], L[[2, 1, 0, 3, 4]]) and np.array_equal(p["global"], g[[1, 1, 0, 2, 3]])
    # locals already paired are not searched (:81-83) and do not count in nTotal
    lp = np.array([0, 1, 1, 0, 0], np.uint8)
    p, _ = orc.match_inlier_ratio(tree, lx, ly, lz, I, orc.MatchInlierRatioParams(0.7), lp, np.zeros(4, np.uint8))
    assert list(p["localIdx"]) == [0, 3]  # nTotal 3 -> round(2.1) = 2
    # global points paired on entry are skipped at emission (:126-128)
    p, _ = orc.match_inlier_ratio(tree, lx, ly, lz, I, orc.MatchInlierRatioParams(0.95), None, np.array([1, 0, 0, 0], np.uint8))
    assert list(p["localIdx"]) == [2, 3, 4]
    # the reference throws: ratio outside (0,1) (:49-50); every local already paired (:117)
    with pytest.raises(RuntimeError):
        orc.match_inlier_ratio(tree, lx, ly, lz, I, orc.MatchInlierRatioParams(1.0))
    with pytest.raises(RuntimeError):
        orc.match_inlier_ratio(tree, lx, ly, lz, I, orc.MatchInlierRatioParams(0.5), np.ones(5, np.uint8), None)
    # no bounding-box overlap (:64-67): nothing, no throw
    far = orc.pose_from_xyzypr(1000, 0, 0)
    p, _ = orc.match_inlier_ratio(tree, lx, ly, lz, far, orc.MatchInlierRatioParams(0.5))
    assert len(p) == 0


@pytest.mark.parametrize("solver", ["horn", "gn"])
def test_icp_align_protocol_inlier_ratio(solver):
    """tests/test-mp2p_icp_algos.cpp:250-262: Solver_Horn / Solver_GaussNewton with
    Matcher_Points_InlierRatio (defaults) on the bunny, decimation 10, |log(GT - est)| < 0.1."""
    x, y, z = icp_harness.load_xyz_gz(os.path.join(GOLD, "bunny_decim.xyz.gz"))
    x, y, z = x[::10], y[::10], z[::10]
    P = np.stack([x, y, z], 1).astype(np.float64)
    size = P.max(0) - P.min(0)
    rng = np.random.default_rng(4321)
    tree = orc.KDTree(x, y, z)
    for _ in range(3):
        gt = orc.pose_from_xyzypr(*(rng.uniform(-0.15, 0.15, 3) * size), *(rng.uniform(-10, 10, 3) * DEG))
        L = ((P - gt[:, 3]) @ gt[:, :3]).astype(np.float32)
        lx, ly, lz = (np.ascontiguousarray(L[:, k]) for k in range(3))

        def match(pose, it):
            return orc.match_inlier_ratio(tree, lx, ly, lz, pose, orc.MatchInlierRatioParams(0.80), nthreads=4)[0]

        def solve(pairs, guess, it):
            if solver == "horn":
                return orc.optimal_tf_horn(pairs)
            ok, T, _ = orc.optimal_tf_gauss_newton(pairs, None, orc.GNParams(maxInnerLoopIterations=6), guess)
            return ok, T

        res = icp_harness.align(match, solve, np.eye(3, 4), icp_harness.IcpParams(maxIterations=100))
        assert np.linalg.norm(orc.se3_log(orc.inverse_compose(res.pose, gt))) < 0.1


# ---------------------------------------------------------------------------------------------
# pt2ln: the error term and its Jacobian are pinned by the reference's own tests —
# tests/test-mp2p_error_terms_jacobians.cpp:105-176 (finite differences on D*exp(eps), 1e-5) and
# :492-512 (known answer) — and the Gauss-Newton solver over point-to-line pairings by
# tests/test-mp2p_optimize_pt2ln.cpp:25-111 (three axis lines, 17 ground-truth poses, 1e-3).
# Matcher_Point2Line has no upstream test: hand-made known answers below.
# ---------------------------------------------------------------------------------------------
def test_pt2ln_jacobian_vs_finite_differences_and_known_answer():
    rng = np.random.default_rng(1234)
    for _ in range(1000):
        D = orc.pose_from_xyzypr(*rng.normal(0, 10, 3), rng.uniform(-np.pi, np.pi), *rng.uniform(-np.pi / 2, np.pi / 2, 2))
        pair = np.zeros(1, orc.PAIR_PT2LN)
        u = rng.normal(0, 1, 3)
        pair["pBase"], pair["director"], pair["local"] = rng.normal(0, 20, 3), u / np.linalg.norm(u), rng.normal(0, 10, 3)
        _, J = orc.error_and_jacobian_pt2ln(pair, D)
        num = np.zeros((3, 6))
        for a in range(6):
            ep, em = np.zeros(6), np.zeros(6)
            ep[a], em[a] = 1e-6, -1e-6
            e_p, _ = orc.error_and_jacobian_pt2ln(pair, orc.compose(D, orc.se3_exp(ep)))
            e_m, _ = orc.error_and_jacobian_pt2ln(pair, orc.compose(D, orc.se3_exp(em)))
            num[:, a] = (e_p - e_m) / 2e-6
        assert np.abs(num - J).max() < 1e-5
    pair = np.zeros(1, orc.PAIR_PT2LN)  # :492-512
    pair["pBase"], pair["director"], pair["local"] = [10, 11, 12], [0, 0, 1], [10, 11, -1.0]
    e, _ = orc.error_and_jacobian_pt2ln(pair, np.eye(3, 4))
    assert np.abs(e).max() < 1e-6


PT2LN_GT = [(0, 0, 0, 0, 0, 0), (1, 0, 0, 0, 0, 0), (0, 1, 0, 0, 0, 0), (0, 0, 1, 0, 0, 0), (-2, 0, 0, 0, 0, 0), (0, -3, 0, 0, 0, 0),
            (0, 0, -4, 0, 0, 0), (0, 0, 0, 20, 0, 0), (0, 0, 0, -20, 0, 0), (0, 0, 0, 0, 10, 0), (0, 0, 0, 0, -10, 0),
            (0, 0, 0, 0, 0, 15), (0, 0, 0, 0, 0, -15), (1, 2, 3, 0, 0, 0), (1, 2, 3, -10, 5, 30)]


def pt2ln_fixture(gt):
    """tests/test-mp2p_optimize_pt2ln.cpp:36-50: the three axes as lines, one point on each."""
    pairs = np.zeros(3, orc.PAIR_PT2LN)
    for k, (axis, pt) in enumerate([((1, 0, 0), (0.5, 0, 0)), ((0, 1, 0), (0, 0.4, 0)), ((0, 0, 1), (0, 0, 0.2))]):
        pairs["director"][k] = axis
        pairs["local"][k] = (np.array(pt) - gt[:, 3]) @ gt[:, :3]  # groundTruth.inverseComposePoint
    return pairs


@pytest.mark.parametrize("gt6", PT2LN_GT)
def test_gn_pt2ln_known_answers(gt6):
    gt = orc.pose_from_xyzypr(*gt6[:3], *(np.array(gt6[3:]) * DEG))
    ok, T, _ = orc.optimal_tf_gauss_newton_ex(None, None, pt2ln_fixture(gt), orc.GNParams(maxInnerLoopIterations=25), np.eye(3, 4))
    assert ok and np.linalg.norm(orc.se3_log(orc.inverse_compose(T, gt))) < 1e-3


def test_matcher_pt2ln_known_answers():
    # global: a pole along z at (5, 5), 41 points 5 cm apart, and a plane patch z = 0 (must not look like a line)
    zs = np.arange(41) * 0.05
    pole = np.stack([np.full(41, 5.0), np.full(41, 5.0), zs], 1)
    gx, gy = np.meshgrid(np.arange(20) * 0.1, np.arange(20) * 0.1)
    plane = np.stack([gx.ravel(), gy.ravel(), np.zeros(400)], 1)
    G = np.concatenate([pole, plane]).astype(np.float32)
    tree = orc.KDTree(*(np.ascontiguousarray(G[:, k]) for k in range(3)))
    L = np.array([[5.1, 5.0, 1.0], [5.0, 5.2, 0.52], [1.0, 1.0, 0.05], [5.0, 9.0, 1.0]], np.float32)
    lx, ly, lz = (np.ascontiguousarray(L[:, k]) for k in range(3))
    prm = orc.MatchPt2LnParams(distanceThreshold=0.5, knn=4, minimumLinePoints=4, lineEigenThreshold=0.01)
    lp = np.zeros(4, np.uint8)
    p, pot = orc.match_pt2ln(tree, lx, ly, lz, np.eye(3, 4), prm, lp)
    # locals 0 and 1 sit next to the pole; local 2 is over the plane (eigen test fails); local 3 is 4 m away
    assert pot == 4 and len(p) == 2 and list(lp) == [1, 1, 0, 0]
    assert np.allclose(np.abs(p["director"]), [[0, 0, 1], [0, 0, 1]], atol=1e-6)
    assert np.allclose(p["pBase"][:, :2], 5.0, atol=1e-6) and np.array_equal(p["local"], L[:2].astype(np.float64))
    assert np.allclose(p["pBase"][:, 2], [0.975, 0.525], atol=1e-6)  # mean of the 4 nearest pole points (ties: lowest index)
    # the count check is on the neighbours within the threshold, the PCA on all knn (as written upstream)
    prm2 = orc.MatchPt2LnParams(distanceThreshold=0.12, knn=4, minimumLinePoints=4, lineEigenThreshold=0.01)
    p2, _ = orc.match_pt2ln(tree, lx, ly, lz, np.eye(3, 4), prm2)
    assert len(p2) == 0  # only 2-3 of the 4 nearest are within 12 cm
    prm3 = orc.MatchPt2LnParams(distanceThreshold=0.12, knn=4, minimumLinePoints=2, lineEigenThreshold=0.01)
    p3, _ = orc.match_pt2ln(tree, lx, ly, lz, np.eye(3, 4), prm3)
    assert len(p3) == 1 and np.allclose(p3["pBase"][0, 2], 0.975, atol=1e-6)
    # already-paired locals are skipped
    p4, _ = orc.match_pt2ln(tree, lx, ly, lz, np.eye(3, 4), prm, np.array([1, 0, 0, 0], np.uint8))
    assert len(p4) == 1 and np.array_equal(p4["local"][0], L[1].astype(np.float64))


# ---------------------------------------------------------------------------------------------
# Matcher_Adaptive — no upstream test; its threshold rests on two MRPT helpers that
# are not in the reference tree (CHistogram, confidenceIntervalsFromHistogram: restated as recalled,
# PARITY UNPINNED, see oracle.cpp). The cases below pin the restatement against an independent numpy
# statement of the same steps and hand-made known answers.
# ---------------------------------------------------------------------------------------------
def _adaptive_threshold_numpy(errs_first_two, confidenceInterval, minimumCorrDist):
    e = np.asarray(errs_first_two, np.float32).astype(np.float64)
    lo, hi, nb = e.min(), e.max(), 50
    inv = (nb - 1) / (hi - lo)
    bins = np.bincount((inv * (e - lo)).astype(np.int64), minlength=nb)[:nb].astype(np.float64)
    hits = bins * (inv / len(e))
    Hc = np.cumsum(hits)
    Hc = Hc * (1.0 / Hc.max())
    xs = lo + np.arange(nb) * (hi - lo) / (nb - 1)
    k = min(nb - 1, int(np.searchsorted(Hc, 1.0 - (1.0 - confidenceInterval), side="right")))
    return max(minimumCorrDist**2, xs[k]), xs[k]


def test_matcher_adaptive_against_numpy_statement():
    rng = np.random.default_rng(2)
    M = rng.uniform(0, 20, (30_000, 3)).astype(np.float32)
    L = (M[::6] + rng.normal(0, 0.12, (5000, 3))).astype(np.float32)
    L[:300] += 3.0  # gross outliers
    tree = orc.KDTree(*(np.ascontiguousarray(M[:, k]) for k in range(3)))
    lx, ly, lz = (np.ascontiguousarray(L[:, k]) for k in range(3))
    I = np.eye(3, 4)
    for ci, k in [(0.80, 1), (0.75, 1), (0.5, 3)]:
        prm = orc.MatchAdaptiveParams(confidenceInterval=ci, absoluteMaxSearchDistance=0.5, maxPt2PtCorrespondences=k, minimumCorrDist=0.01)
        p2p, p2l, pot, ci_high = orc.match_adaptive(tree, lx, ly, lz, I, prm, nthreads=4)
        assert pot == len(L) * k and len(p2l) == 0
        # independent statement: k-NN by the oracle's own search, thresholds by numpy
        idx, d2, found = tree.knn(lx, ly, lz, k, np.nextafter(np.float32(0.25), np.float32(np.inf)) if k == 1 else 0.25, nthreads=4)
        first_two = np.concatenate([d2[found > r, r] for r in range(min(k, 2))])
        thr, xk = _adaptive_threshold_numpy(first_two, ci, 0.01)
        assert abs(ci_high - xk) <= 1e-12 * max(1.0, xk)
        exp = []
        for i in range(len(L)):
            for r in range(min(found[i], k)):
                if d2[i, r] >= thr:
                    continue
                if r and np.float32(d2[i, r]) > np.float32(d2[i, 0]) * np.float32(1.2 * 1.2):
                    break
                exp.append((i, idx[i, r]))
        assert [(int(a), int(b)) for a, b in zip(p2p["localIdx"], p2p["globalIdx"])] == exp and len(exp) > 2000
        assert np.array_equal(p2p["local"], L[p2p["localIdx"]]) and np.array_equal(p2p["global"], M[p2p["globalIdx"]])


def test_matcher_adaptive_planes_and_edge_cases():
    # a plane patch z = 0 sampled every 5 cm; locals 3 cm above it (planar neighbourhood -> pt2pl) and far away
    gx, gy = np.meshgrid(np.arange(40) * 0.05, np.arange(40) * 0.05)
    G = np.stack([gx.ravel(), gy.ravel(), np.zeros(1600)], 1).astype(np.float32)
    rng = np.random.default_rng(4)
    G[:, 2] += rng.normal(0, 1e-4, len(G)).astype(np.float32)
    tree = orc.KDTree(*(np.ascontiguousarray(G[:, k]) for k in range(3)))
    L = np.array([[1.0, 1.0, 0.03], [0.52, 1.31, 0.05], [1.0, 1.0, 1.5], [30, 30, 30]], np.float32)
    lx, ly, lz = (np.ascontiguousarray(L[:, k]) for k in range(3))
    prm = orc.MatchAdaptiveParams(confidenceInterval=0.8, absoluteMaxSearchDistance=2.0, enableDetectPlanes=True, planeSearchPoints=8, planeMinimumFoundPoints=4, planeMinimumDistance=0.10, minimumCorrDist=0.1)
    lp = np.zeros(4, np.uint8)
    p2p, p2l, pot, _ = orc.match_adaptive(tree, lx, ly, lz, np.eye(3, 4), prm, lp)
    # locals 0, 1: plane found, |distance of the LOCAL point| 0.03 / 0.05 < 0.10 -> pt2pl; local 2: plane found but
    # 1.5 m away -> falls to the pt2pt branch (error 2.25 >= maxCorrDistSqr -> nothing); local 3: no neighbour
    assert pot == 4 and len(p2l) == 2 and len(p2p) == 0 and list(lp) == [1, 1, 0, 0]
    assert np.allclose(np.abs(p2l["coefs"][:, 2]), 1.0, atol=1e-4) and np.allclose(p2l["centroid"][:, 2], 0.0, atol=1e-3)
    assert np.array_equal(p2l["local"], L[:2])
    # planes off: pt2pt with the adaptive threshold, nearest neighbour only
    prm2 = orc.MatchAdaptiveParams(confidenceInterval=0.8, absoluteMaxSearchDistance=2.0, minimumCorrDist=0.1)
    p2p, p2l, _, ci_high = orc.match_adaptive(tree, lx, ly, lz, np.eye(3, 4), prm2)
    assert len(p2l) == 0 and list(p2p["localIdx"]) == [0, 1]  # errors 0.0009, ~0.003 < max(0.01, ci_high); 2.25 is not
    # the reference throws / crashes: no neighbour at all; a single error value (CHistogram asserts max > min)
    with pytest.raises(RuntimeError):
        orc.match_adaptive(tree, lx[:1], ly[:1], lz[:1], np.eye(3, 4), orc.MatchAdaptiveParams(absoluteMaxSearchDistance=0.01))
    with pytest.raises(RuntimeError):
        orc.match_adaptive(tree, lx[:1], ly[:1], lz[:1], np.eye(3, 4), prm2)
    # no bounding-box overlap, empty cloud: nothing, no throw
    assert len(orc.match_adaptive(tree, lx, ly, lz, orc.pose_from_xyzypr(500, 0, 0), prm2)[0]) == 0
    assert len(orc.match_adaptive(tree, lx[:0], ly[:0], lz[:0], np.eye(3, 4), prm2)[0]) == 0


# ---------------------------------------------------------------------------------------------
# C4: the schedule of demos/icp-settings-kitti.yaml:10-59 —
# Matcher_Points_DistanceThreshold(threshold 2.0) + Solver_Horn for iterations 0-5, then
# Matcher_Adaptive(confidenceInterval 0.75, firstToSecondDistanceMax 1.2, absoluteMaxSearchDistance 2.0)
# + Solver_GaussNewton(maxIterations 3, GemanMcClure 0.15); maxIterations 200, minAbsStep 1e-4.
# (`enableDetectPlanes`, required by Matcher_Adaptive::initialize, is missing from that yaml — F5 — and is
# taken as false, the class default.) No upstream expectation exists for this pipeline; the test pins
# that the restated chain converges to the ground truth.
# ---------------------------------------------------------------------------------------------
def test_c4_kitti_schedule_converges():
    """Run on a well-conditioned cloud (C2-shaped): the synthetic street of C3 is a corridor — ground and
    two facades parallel to x — whose along-street translation no point matcher can observe."""
    from tests import fixtures as fx

    M, L, gt = fx.make_c2(n_map=200_000, decim=10)
    guess = np.eye(3, 4)
    tree = orc.KDTree(*(np.ascontiguousarray(M[:, k]) for k in range(3)))
    lx, ly, lz = (np.ascontiguousarray(L[:, k]) for k in range(3))
    used = []

    def match(pose, it):
        if it <= 5:
            used.append("pt2pt")
            return orc.match_pt2pt(tree, lx, ly, lz, pose, orc.MatchPt2PtParams(threshold=2.0, thresholdAngularDeg=0.0), nthreads=4)[0]
        used.append("adaptive")
        p2p, p2l, _, _ = orc.match_adaptive(tree, lx, ly, lz, pose, orc.MatchAdaptiveParams(confidenceInterval=0.75, firstToSecondDistanceMax=1.2, absoluteMaxSearchDistance=2.0), nthreads=4)
        return p2p

    def solve(pairs, cur, it):
        if it <= 5:
            return orc.optimal_tf_horn(pairs)
        ok, T, _ = orc.optimal_tf_gauss_newton(pairs, None, orc.GNParams(maxInnerLoopIterations=3, kernel="GemanMcClure", kernelParam=0.15), cur, nthreads=4)
        return ok, T

    res = icp_harness.align(match, solve, guess, icp_harness.IcpParams(maxIterations=200, minAbsStep_trans=1e-4, minAbsStep_rot=1e-4))
    d = orc.se3_log(orc.inverse_compose(res.pose, gt))
    assert res.terminationReason in ("Stalled", "MaxIterations") and "adaptive" in used
    assert np.linalg.norm(d[:3]) < 5e-3 and np.linalg.norm(d[3:]) < 5e-4, (d, res.nIterations)


# ---------------------------------------------------------------------------------------------
# FilterDecimateVoxels — oracle first; the CUDA side is the next round's. No upstream test pins its
# output; the cases below pin the restatement's reading of FilterDecimateVoxels.cpp:109-378 and of the voxel index
# (truncation toward zero, PointCloudToVoxelGridSingle.h:105) against an independent numpy statement.
# ---------------------------------------------------------------------------------------------
def test_decimate_voxels_known_answers_and_numpy_statement():
    P = np.array([[0.05, 0.05, 0.05], [0.15, 0.02, 0.01], [-0.15, 0.0, 0.0], [0.25, 0.0, 0.0], [0.26, 0.01, 0.0], [0.21, 0.19, 0.1],
                  [0.25, 0.0, 0.35]], np.float32)
    x, y, z = (np.ascontiguousarray(P[:, k]) for k in range(3))
    # resolution 0.2: int32(c / 0.2) truncates toward zero, so (-0.2, 0.2) is ONE voxel per axis: points 0, 1, 2 share (0,0,0)
    out, src = orc.decimate_voxels(x, y, z, 0.2, "FirstPoint")
    assert list(src) == [0, 3, 6] and np.array_equal(out, P[[0, 3, 6]])  # voxels (0,0,0), (1,0,0), (1,0,1) in that order
    out, src = orc.decimate_voxels(x, y, z, 0.2, "VoxelAverage")
    assert list(src) == [-1, -1, -1]
    assert np.allclose(out[0], P[:3].mean(0), atol=1e-7) and np.allclose(out[1], P[3:6].mean(0), atol=1e-7) and np.array_equal(out[2], P[6])
    out, src = orc.decimate_voxels(x, y, z, 0.2, "ClosestToAverage")
    assert list(src) == [0, 4, 6]  # mean of voxel (1,0,0) = (0.24, 0.0667, 0.0333): point 4 is the closest member
    out, src = orc.decimate_voxels(x, y, z, 0.2, "FirstPoint", flatten_to=7.0)
    assert list(src) == [0, 3] and np.array_equal(out[:, 2], [7.0, 7.0]) and np.array_equal(out[:, :2], P[[0, 3], :2])  # (1,0,1) shares a column
    assert len(orc.decimate_voxels(x[:0], y[:0], z[:0], 0.2)[0]) == 0
    # random cloud against a numpy statement of the same steps
    rng = np.random.default_rng(6)
    Q = rng.uniform(-3, 3, (20000, 3)).astype(np.float32)
    res = np.float32(0.25)
    key = (Q / res).astype(np.int32)  # float division, truncation toward zero
    order = np.lexsort((np.arange(len(Q)), key[:, 2], key[:, 1], key[:, 0]))
    ks = key[order]
    first = np.ones(len(Q), bool)
    first[1:] = np.any(ks[1:] != ks[:-1], axis=1)
    out, src = orc.decimate_voxels(*(np.ascontiguousarray(Q[:, k]) for k in range(3)), float(res), "FirstPoint")
    assert np.array_equal(src, order[first]) and np.array_equal(out, Q[order[first]])
    out, src = orc.decimate_voxels(*(np.ascontiguousarray(Q[:, k]) for k in range(3)), float(res), "VoxelAverage")
    starts = np.flatnonzero(first)
    ends = np.append(starts[1:], len(Q))
    for j in rng.choice(len(starts), 300, replace=False):
        members = Q[np.sort(order[starts[j] : ends[j]])]
        mean = np.zeros(3, np.float32)
        for m in members:
            mean = mean + m
        mean = mean * (np.float32(1.0) / np.float32(len(members)))
        assert np.array_equal(out[j], mean)
