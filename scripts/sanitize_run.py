"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mp2p_icp_b200 as b200
from tests import fixtures as fx

xyz = lambda a: (np.ascontiguousarray(a[:, 0]), np.ascontiguousarray(a[:, 1]), np.ascontiguousarray(a[:, 2]))
M, L, gt = fx.make_c2(n_map=60_000, decim=10)
ctx = b200.Context(0)
gmap = b200.Map(ctx, *xyz(M))
pose = fx.pose_xyzypr(0.25, -0.15, 0.08, 0.03, -0.014, 0.02)
p1, _ = gmap.match_pt2pt(*xyz(L), pose, b200.Pt2PtParams(threshold=1.0))
p3, _ = gmap.match_pt2pt(*xyz(L), pose, b200.Pt2PtParams(threshold=2.5, pairingsPerPoint=3))
pin = [torch.from_numpy(a).pin_memory() for a in xyz(L)]
out = torch.zeros(len(L) * 36, dtype=torch.uint8).pin_memory()
for _ in range(3):  # zero copy, then the single-launch matcher once a Horn solve is expected
    pz, _ = gmap.match_pt2pt(*(t.numpy() for t in pin), pose, b200.Pt2PtParams(threshold=1.0), out=out.numpy().view(b200.PAIR_PT2PT))
    ok, T = ctx.solve_horn(pz, last_match=True)
assert pz.tobytes() == p1.tobytes()
cloud = b200.Cloud(ctx, *xyz(L))
ok, T2, n = gmap.make_iterator(cloud, None, None, len(L), b200.Pt2PtParams(threshold=1.0), b200.HornParams())(pose)
assert ok and n == len(p1)
pi, _ = gmap.match_inlier_ratio(*xyz(L), pose, b200.InlierRatioParams(0.8))
ok, T3, it = ctx.solve_gauss_newton(p1, None, b200.GNParams(maxInnerLoopIterations=3, kernel="Cauchy", kernelParam=0.3), pose)
S = fx.make_street_scene(n_map=100_000, length=30.0)
scan = fx.make_lidar_scan((15.0, 0.3, 0.0), n_rings=16, n_az=300, length=30.0)
g2 = fx.pose_xyzypr(15.05, 0.28, 0.01, 0.01, 0.0, 0.0)
smap = b200.Map(ctx, *xyz(S))
kw = dict(distanceThreshold=0.5, searchRadius=1.0, knn=8, minimumPlanePoints=5, planeEigenThreshold=0.01)
q, _ = smap.match_pt2pl(*xyz(scan), g2, b200.Pt2PlParams(**kw))
gn = b200.GNParams(maxInnerLoopIterations=3, kernel="GemanMcClure", kernelParam=0.15)
ok, T4, it = ctx.solve_gauss_newton(None, q, gn, g2)
ok, T5, n5 = smap.make_iterator(b200.Cloud(ctx, *xyz(scan)), None, None, len(scan), b200.Pt2PlParams(**kw), gn)(g2)
ok, T6 = ctx.solve_horn_pt2pl(q, g2)
i, d, f = smap.knn(*xyz(scan[:2000]), 8, 1.0)
# Matcher_Point2Line + the pt2ln term of Gauss-Newton, Matcher_Adaptive (both branches)
ln, _ = smap.match_pt2ln(*xyz(scan), g2, b200.Pt2LnParams(distanceThreshold=1.0, knn=8, minimumLinePoints=4, lineEigenThreshold=0.5))
if len(ln) >= 3:
    ok, T7, it7 = ctx.solve_gauss_newton_ex(None, q[:500], ln[:2000], gn, g2, w_pt2ln=0.5)
a1, l1, _, _ = smap.match_adaptive(*xyz(scan), g2, b200.AdaptiveParams(enableDetectPlanes=True, absoluteMaxSearchDistance=1.0, planeMinimumDistance=50.0))
a2, l2, _, _ = gmap.match_adaptive(*xyz(L), pose, b200.AdaptiveParams(absoluteMaxSearchDistance=1.5, maxPt2PtCorrespondences=3))
# round 2: voxel decimation filter, covariance(), library-side layer cache, pageable records >= 256 KB (hostcopy.hpp)
dv, src = ctx.decimate_voxels(*xyz(scan), 0.5, method="VoxelAverage")
dx = dv
cov, hes, pd = ctx.covariance(p1, None, None, np.zeros(6))
cov = np.asarray(cov)
cm, _ = ctx.cached_map(*xyz(S))
q2, _ = cm.match_pt2pl(*xyz(scan), g2, b200.Pt2PlParams(**kw))
assert q2.tobytes() == q.tobytes()
big, _ = gmap.match_pt2pt(*xyz(M[:30_000]), fx.pose_xyzypr(0, 0, 0, 0, 0, 0), b200.Pt2PtParams(threshold=0.5))  # 1 MB of records
ok, T8 = ctx.solve_horn(big)
print("extra:", len(ln), len(a1), len(l1), len(a2), len(dx), cov.shape, len(big))
print("sanitize run OK:", len(p1), len(p3), len(pi), len(q), ctx.launch_count, "launches")
