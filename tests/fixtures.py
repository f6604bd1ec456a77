"""Toy clouds of the reference's matcher unit tests, regenerated with the same float arithmetic
(`i * 0.01f`, `5.0f + iy * 0.01f`), and the synthetic clouds of BASELINE.md / SURVEY.md §8d."""
import numpy as np

f32 = np.float32


def pt2pt_fixture_global():
    """tests/test-mp2p_matcher_pt2pt.cpp:26-34 — 20 global points."""
    pts = [(f32(i) * f32(0.01), f32(5.0), f32(0.0)) for i in range(10)]
    pts += [(f32(10.0), f32(i) * f32(0.01), f32(1.0)) for i in range(10)]
    a = np.array(pts, dtype=np.float32)
    return a[:, 0].copy(), a[:, 1].copy(), a[:, 2].copy()


def two_local_points():
    """tests/test-mp2p_matcher_pt2pt.cpp:36-44 / test-mp2p_matcher_pt2pl.cpp:50-58."""
    a = np.array([(0, 0, 0), (2, 0, 0)], dtype=np.float32)
    return a[:, 0].copy(), a[:, 1].copy(), a[:, 2].copy()


def pt2pl_fixture_global():
    """tests/test-mp2p_matcher_pt2pl.cpp:29-48 — two 10x10 planes + one 10x10x10 blob (1200 pts)."""
    pts = []
    for ix in range(10):
        for iy in range(10):
            pts.append((f32(ix) * f32(0.01), f32(5.0) + f32(iy) * f32(0.01), f32(0.0)))
    for iy in range(10):
        for iz in range(10):
            pts.append((f32(10.0), f32(iy) * f32(0.01), f32(iz) * f32(0.01)))
    for ix in range(10):
        for iy in range(10):
            for iz in range(10):
                pts.append((f32(20.0) + f32(ix) * f32(0.01), f32(iy) * f32(0.01), f32(iz) * f32(0.01)))
    a = np.array(pts, dtype=np.float32)
    return a[:, 0].copy(), a[:, 1].copy(), a[:, 2].copy()


def pose_xyzypr(x, y, z, yaw=0.0, pitch=0.0, roll=0.0):
    """CPose3D(x,y,z,yaw,pitch,roll): R = Rz(yaw) Ry(pitch) Rx(roll) -> 3x4 [R|t] (numpy, fp64)."""
    cy, sy, cp, sp, cr, sr = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
    return np.array(
        [
            [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr, x],
            [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr, y],
            [-sp, cp * sr, cp * cr, z],
        ],
        dtype=np.float64,
    )


def to_local_frame(P64, gt):
    """Express global points in the local frame of pose gt: l = R^T (g - t), stored as float32."""
    return ((P64 - gt[:, 3]) @ gt[:, :3]).astype(np.float32)


def make_c2(n_map=1_000_000, decim=10, seed_map=1234, seed_noise=4321, noise=0.02):
    """C2: map uniform in [0,100)^3, query = every `decim`-th map point + N(0,noise),
    expressed in the local frame of GT pose (0.30,-0.20,0.10; 2,-1,1.5 deg)."""
    rng = np.random.default_rng(seed_map)
    M = rng.uniform(0, 100, (n_map, 3)).astype(np.float32)
    rn = np.random.default_rng(seed_noise)
    Q = M[::decim].astype(np.float64) + rn.normal(0, noise, (len(M[::decim]), 3))
    gt = pose_xyzypr(0.30, -0.20, 0.10, np.deg2rad(2.0), np.deg2rad(-1.0), np.deg2rad(1.5))
    L = to_local_frame(Q, gt)
    return M, L, gt


def make_street_scene(n_map=10_000_000, seed=7, length=1000.0, half_width=10.0, height=12.0):
    """C3 map: ground plane z=-1.73 + two facade planes along a street, sampled with
    uniform random surface samples (KITTI-shaped synthetic; no KITTI data ships with the reference)."""
    rng = np.random.default_rng(seed)
    a_ground = length * 2 * half_width
    a_wall = length * height
    n_g = int(n_map * a_ground / (a_ground + 2 * a_wall))
    n_w = (n_map - n_g) // 2
    n_w2 = n_map - n_g - n_w
    g = np.stack([rng.uniform(0, length, n_g), rng.uniform(-half_width, half_width, n_g), np.full(n_g, -1.73)], 1)
    w1 = np.stack([rng.uniform(0, length, n_w), np.full(n_w, half_width), rng.uniform(-1.73, height - 1.73, n_w)], 1)
    w2 = np.stack([rng.uniform(0, length, n_w2), np.full(n_w2, -half_width), rng.uniform(-1.73, height - 1.73, n_w2)], 1)
    M = np.concatenate([g, w1, w2]).astype(np.float32)
    M += rng.normal(0, 0.005, M.shape).astype(np.float32)  # 5 mm surface roughness
    rng.shuffle(M)
    return M


def make_lidar_scan(sensor_xyz, n_rings=64, n_az=1875, max_range=80.0, seed=8, noise=0.02,
                    half_width=10.0, height=12.0, length=1000.0):
    """C3 scan: 64 rings x 1875 azimuths ray-cast against the street scene analytically, N(0,noise)
    range noise, in the SENSOR frame (returns float32 Nx3)."""
    rng = np.random.default_rng(seed)
    el = np.deg2rad(np.linspace(-24.8, 2.0, n_rings))
    az = np.linspace(-np.pi, np.pi, n_az, endpoint=False)
    E, A = np.meshgrid(el, az, indexing="ij")
    d = np.stack([np.cos(E) * np.cos(A), np.cos(E) * np.sin(A), np.sin(E)], -1).reshape(-1, 3)
    o = np.asarray(sensor_xyz, float)
    with np.errstate(divide="ignore", invalid="ignore"):
        t_g = (-1.73 - o[2]) / d[:, 2]
        t_w1 = (half_width - o[1]) / d[:, 1]
        t_w2 = (-half_width - o[1]) / d[:, 1]
    cands = []
    for t, kind in ((t_g, 0), (t_w1, 1), (t_w2, 2)):
        p = o + d * t[:, None]
        ok = (t > 0.5) & np.isfinite(t) & (p[:, 0] > 0) & (p[:, 0] < length)
        if kind == 0:
            ok &= np.abs(p[:, 1]) <= half_width
        else:
            ok &= (p[:, 2] >= -1.73) & (p[:, 2] <= height - 1.73)
        cands.append(np.where(ok, t, np.inf))
    t = np.min(np.stack(cands, 0), 0)
    keep = t < max_range
    t = t[keep] + rng.normal(0, noise, keep.sum())
    return (d[keep] * t[:, None]).astype(np.float32)
