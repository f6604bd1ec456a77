"""Independent cross-checks of the oracle's numerical building blocks (CPU only): each restated piece
is compared with a DIFFERENT algorithm for the same quantity from numpy / scipy, so that an error
shared by the restatement and the CUDA code (both written from the same reading of the reference)
would still show. These complement tests/test_oracle_golden.py, which pins the oracle on the
reference's own fixtures."""
import numpy as np
import pytest
from scipy.linalg import expm, logm
from scipy.optimize import least_squares

from oracle import oracle_py as orc

DEG = np.pi / 180.0


def _hat6(xi):
    v, w = xi[:3], xi[3:]
    M = np.zeros((4, 4))
    M[:3, :3] = [[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]
    M[:3, 3] = v
    return M


def test_se3_exp_log_against_matrix_exponential():
    """Lie::SE<3>::exp / log with tangent order (v, omega) (ICP.cpp:194-196) vs scipy's expm / logm."""
    rng = np.random.default_rng(0)
    for _ in range(200):
        xi = np.concatenate([rng.normal(0, 2, 3), rng.normal(0, 1, 3)])
        if np.linalg.norm(xi[3:]) > 3.0:
            continue
        T = orc.se3_exp(xi)
        E = expm(_hat6(xi))
        assert np.abs(T - E[:3]).max() < 1e-12
        back = orc.se3_log(T)
        assert np.abs(back - xi).max() < 1e-9
        Lm = np.real(logm(np.vstack([T, [0, 0, 0, 1]])))
        assert np.abs(np.array([*Lm[:3, 3], Lm[2, 1], Lm[0, 2], Lm[1, 0]]) - back).max() < 1e-8
    small = orc.se3_exp([1e-3, 0, 0, 1e-12, 0, 0])  # small-angle branch
    assert np.abs(small - expm(_hat6(np.array([1e-3, 0, 0, 1e-12, 0, 0])))[:3]).max() < 1e-15


def test_pose_composition_is_matrix_product():
    rng = np.random.default_rng(1)
    for _ in range(50):
        A = orc.pose_from_xyzypr(*rng.normal(0, 5, 3), *rng.uniform(-3, 3, 3))
        B = orc.pose_from_xyzypr(*rng.normal(0, 5, 3), *rng.uniform(-1.5, 1.5, 3))
        A4, B4 = np.vstack([A, [0, 0, 0, 1]]), np.vstack([B, [0, 0, 0, 1]])
        assert np.abs(orc.compose(A, B) - (A4 @ B4)[:3]).max() < 1e-12
        assert np.abs(orc.inverse(A) - np.linalg.inv(A4)[:3]).max() < 1e-12
        assert np.abs(orc.inverse_compose(A, B) - (np.linalg.inv(B4) @ A4)[:3]).max() < 1e-12
    # CPose3D(x, y, z, yaw, pitch, roll) = Rz(yaw) Ry(pitch) Rx(roll)
    y, p, r = 0.3, -0.2, 0.5
    Rz = np.array([[np.cos(y), -np.sin(y), 0], [np.sin(y), np.cos(y), 0], [0, 0, 1]])
    Ry = np.array([[np.cos(p), 0, np.sin(p)], [0, 1, 0], [-np.sin(p), 0, np.cos(p)]])
    Rx = np.array([[1, 0, 0], [0, np.cos(r), -np.sin(r)], [0, np.sin(r), np.cos(r)]])
    assert np.abs(orc.pose_from_xyzypr(1, 2, 3, y, p, r)[:, :3] - Rz @ Ry @ Rx).max() < 1e-15


def test_eig_symmetric_against_lapack():
    rng = np.random.default_rng(2)
    for n in (3, 4):
        for _ in range(100):
            B = rng.normal(size=(n, n))
            A = B + B.T
            vals, V = orc.eig_sym(A)
            w, _ = np.linalg.eigh(A)
            assert np.all(np.diff(vals) >= 0) and np.abs(vals - w).max() < 1e-10  # ascending, as eig_symmetric delivers
            assert np.abs(A @ V - V * vals).max() < 1e-9 and np.abs(V.T @ V - np.eye(n)).max() < 1e-12


def _kabsch(local, glob):
    """Least-squares rigid transform by SVD (Kabsch / Umeyama without scale): a different algorithm for
    what Horn's quaternion method computes."""
    cl, cg = local.mean(0), glob.mean(0)
    H = (local - cl).T @ (glob - cg)
    U, _, Vt = np.linalg.svd(H)
    D = np.diag([1, 1, np.sign(np.linalg.det(Vt.T @ U.T))])
    R = Vt.T @ D @ U.T
    return np.hstack([R, (cg - R @ cl)[:, None]])


@pytest.mark.parametrize("n,sigma", [(3, 0.0), (10, 0.01), (500, 0.05), (5000, 0.2)])
def test_horn_equals_svd_solution(n, sigma):
    rng = np.random.default_rng(3 + n)
    gt = orc.pose_from_xyzypr(*rng.uniform(-2, 2, 3), *(rng.uniform(-40, 40, 3) * DEG))
    L = rng.uniform(-20, 20, (n, 3)).astype(np.float32).astype(np.float64)
    G = (L @ gt[:, :3].T + gt[:, 3] + rng.normal(0, sigma, (n, 3))).astype(np.float32).astype(np.float64)
    pairs = np.zeros(n, orc.PAIR_PT2PT)
    pairs["globalIdx"] = pairs["localIdx"] = np.arange(n)
    pairs["global"], pairs["local"] = G, L
    ok, T = orc.optimal_tf_horn(pairs)
    K = _kabsch(L, G)
    assert ok and np.abs(T - K).max() < 1e-8


def test_gauss_newton_reaches_the_least_squares_minimum():
    """optimal_tf_gauss_newton over mixed pt2pt + pt2pl + pt2ln pairings vs scipy.optimize.least_squares on
    the same residuals (no robust kernel): same minimiser."""
    rng = np.random.default_rng(7)
    gt = orc.pose_from_xyzypr(0.4, -0.3, 0.2, 5 * DEG, -3 * DEG, 4 * DEG)
    n = 300
    l = rng.uniform(-10, 10, (n, 3))
    g = l @ gt[:, :3].T + gt[:, 3]
    p2p = np.zeros(n, orc.PAIR_PT2PT)
    p2p["global"], p2p["local"] = g + rng.normal(0, 0.02, (n, 3)), l
    nrm = rng.normal(size=(n, 3))
    nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    p2l = np.zeros(n, orc.PAIR_PT2PL)
    p2l["coefs"][:, :3], p2l["coefs"][:, 3] = nrm, -(nrm * g).sum(1) + rng.normal(0, 0.02, n)
    p2l["local"] = l
    u = rng.normal(size=(n, 3))
    u /= np.linalg.norm(u, axis=1)[:, None]
    p2ln = np.zeros(n, orc.PAIR_PT2LN)
    p2ln["pBase"], p2ln["director"], p2ln["local"] = g + rng.normal(0, 0.02, (n, 3)), u, l
    lf = p2p["local"].astype(np.float64), p2l["local"].astype(np.float64), p2ln["local"]
    gf, cf = p2p["global"].astype(np.float64), p2l["coefs"]

    def residuals(xi):
        T = orc.se3_exp(xi)
        R, t = T[:, :3], T[:, 3]
        r1 = (lf[0] @ R.T + t - gf).ravel()
        r2 = ((lf[1] @ R.T + t) * cf[:, :3]).sum(1) + cf[:, 3]
        q = lf[2] @ R.T + t - p2ln["pBase"]
        r3 = (q - u * (q * u).sum(1)[:, None]).ravel()
        return np.concatenate([r1, r2, r3])

    sol = least_squares(residuals, np.zeros(6), xtol=1e-14, ftol=1e-14, gtol=1e-14)
    ok, T, _ = orc.optimal_tf_gauss_newton_ex(p2p, p2l, p2ln, orc.GNParams(maxInnerLoopIterations=30), np.eye(3, 4), nthreads=2)
    assert ok and np.abs(T - orc.se3_exp(sol.x)).max() < 1e-7


def test_estimate_points_eigen_against_numpy_covariance():
    """The plane of Matcher_Point2Plane: normal = eigenvector of the smallest eigenvalue of the (population)
    covariance of the neighbours (estimate_points_eigen.cpp:27-123), checked through the pt2pl matcher on a
    tilted plane against numpy's eigh of the same neighbourhoods."""
    rng = np.random.default_rng(9)
    n_true = np.array([0.3, -0.2, 0.93])
    n_true /= np.linalg.norm(n_true)
    a, b = np.cross(n_true, [1, 0, 0]), None
    a /= np.linalg.norm(a)
    b = np.cross(n_true, a)
    uv = rng.uniform(-3, 3, (20000, 2))
    G = (uv[:, :1] * a + uv[:, 1:] * b + rng.normal(0, 1e-3, (20000, 1)) * n_true).astype(np.float32)
    tree = orc.KDTree(*(np.ascontiguousarray(G[:, k]) for k in range(3)))
    Q = (rng.uniform(-2, 2, (200, 1)) * a + rng.uniform(-2, 2, (200, 1)) * b + 0.02 * n_true).astype(np.float32)
    prm = orc.MatchPt2PlParams(distanceThreshold=0.1, searchRadius=0.5, knn=10, minimumPlanePoints=5, planeEigenThreshold=0.01)
    pl, _ = orc.match_pt2pl(tree, *(np.ascontiguousarray(Q[:, k]) for k in range(3)), np.eye(3, 4), prm)
    assert len(pl) == 200
    idx, _, found = tree.knn(*(np.ascontiguousarray(Q[:, k]) for k in range(3)), 10, 0.25)
    for k in range(200):
        P = G[idx[k, : found[k]]].astype(np.float64)
        w, V = np.linalg.eigh(np.cov(P.T, bias=True))
        nk = pl["coefs"][k, :3]
        assert abs(abs(nk @ V[:, 0]) - 1) < 1e-5 and abs(np.linalg.norm(nk) - 1) < 1e-9  # float mean upstream: 1e-5
        assert np.abs(pl["centroid"][k] - P.mean(0)).max() < 1e-5
        assert abs(nk @ pl["centroid"][k] + pl["coefs"][k, 3]) < 1e-9


def test_covariance_against_an_independent_numpy_statement():
    """mp2p_icp::covariance (covariance.cpp:28-141): numerical Jacobian of the stacked error vector w.r.t.
    (x, y, z, yaw, pitch, roll), hessian = J^T J, cov = hessian^-1. Independent statement: the ANALYTIC Jacobian
    of the same residuals through scipy's rotation derivatives (finite differences of a different step and
    scheme would share the method), numpy's inverse."""
    rng = np.random.default_rng(5)
    n = 200
    x6 = np.array([0.3, -0.2, 0.1, 0.2, -0.1, 0.05])
    T = orc.pose_from_xyzypr(*x6)
    L = rng.uniform(-5, 5, (n, 3))
    G = L @ T[:, :3].T + T[:, 3] + rng.normal(0, 0.01, (n, 3))
    p2p = np.zeros(n, orc.PAIR_PT2PT)
    p2p["global"], p2p["local"] = G.astype(np.float32), L.astype(np.float32)
    m = 150
    Lp = rng.uniform(-5, 5, (m, 3))
    nrm = rng.normal(size=(m, 3))
    nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    Gp = Lp @ T[:, :3].T + T[:, 3]
    p2l = np.zeros(m, orc.PAIR_PT2PL)
    p2l["coefs"][:, :3] = nrm
    p2l["coefs"][:, 3] = -(nrm * Gp).sum(1) + rng.normal(0, 0.01, m)
    p2l["local"] = Lp.astype(np.float32)
    cov, hes = orc.covariance(p2p, p2l, None, x6)

    def residuals(x):
        Tx = orc.pose_from_xyzypr(*x)
        l1 = p2p["local"].astype(np.float64)
        e1 = (l1 @ Tx[:, :3].T + Tx[:, 3] - p2p["global"].astype(np.float64)).reshape(-1)
        l2 = p2l["local"].astype(np.float64)
        g2 = l2 @ Tx[:, :3].T + Tx[:, 3]
        c = p2l["coefs"]
        ev = (c[:, :3] * g2).sum(1) + c[:, 3]
        e2 = (-(c[:, :3] / (c[:, :3] ** 2).sum(1)[:, None]) * ev[:, None]).reshape(-1)
        return np.concatenate([e1, e2])

    # complex-step-free independent derivative: Richardson-extrapolated central differences at two larger steps
    def jac(h):
        J = np.zeros((3 * (n + m), 6))
        for i in range(6):
            d = np.zeros(6)
            d[i] = h
            J[:, i] = (residuals(x6 + d) - residuals(x6 - d)) / (2 * h)
        return J

    J = (4 * jac(1e-4) - jac(2e-4)) / 3
    H = J.T @ J
    assert np.allclose(hes, H, rtol=1e-5, atol=1e-6 * np.abs(H).max())
    assert np.allclose(cov, np.linalg.inv(H), rtol=1e-4, atol=1e-6 * np.abs(np.linalg.inv(H)).max())
    # no pairings: diag(1e6) (covariance.cpp:33-38)
    c0, _ = orc.covariance(None, None, None, x6)
    assert np.array_equal(c0, np.eye(6) * 1e6)
