#!/usr/bin/env python
"""bench.py — ICP iterations/s of the mp2p_icp Matcher+Solver hot path on B200 (BASELINE.json).

One "step" = one ICP iteration at a fixed pose: run_matchers + run_solvers over one batch of synthetic
input, index already built (the reference amortises its KD-tree the same way).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C1|C2|C3|C4|C5] [--impl reference]

Workloads: the default is **C3** — the configuration BASELINE.json's metric is quoted on
(10M-point map, ~120k-point scan, Matcher_Point2Plane + Solver_GaussNewton). C2 = 1M map / 100k queries,
pt2pt + Horn. C5 = 100M-point map, 1M queries, pt2pt + GN, the query cloud cut over the GPUs (strong
scaling). C1 / C4 = full ICP::align() loops (bunny; C3 data through demos/icp-settings-kitti.yaml).
The default run prints ONE JSON line for C3 and carries, as extra objects of that line, the align()
wall times of C1 / C4 (`align`) and the C5 figures (`c5`).

Keys of the line: see DESIGN.md §5.
  value     device-resident iterations/s — CUDA events on the launching stream, per step, L2 flushed
            between steps (flush excluded from the timing).
  e2e       the same iteration through the C ABI with HOST (pinned) buffers as the reference's loop
            makes it: matcher call (local cloud in, pairings out), solver call over those pairings
            (uploaded again: the plugin's safe default), pose back — wall clock.
  roofline  the matcher's kernels as launched INSIDE the timed step function (a second pass of the same
            steps with the library's per-kernel CUDA events on): algorithmic bytes of SURVEY §8(d) /
            their duration, against the measured HBM peak of MEASURED_PEAKS.json.
  cpu_baseline  the CPU oracle (line-faithful port of the reference path; MRPT cannot be built here) on
            the same inputs, all host threads.
"""
from __future__ import annotations

import argparse
import gzip
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tests import fixtures as fx  # noqa: E402  (synthetic clouds of SURVEY.md §8d; no oracle inside)

METRIC = "ICP iterations/sec"
DTYPE = "f32 metric / f64 solve"


def xyz(a):
    return np.ascontiguousarray(a[:, 0]), np.ascontiguousarray(a[:, 1]), np.ascontiguousarray(a[:, 2])


# ------------------------------------------------------------------------------------------------
# SE(3) helpers of the align() loop (ICP.cpp:203-229: increments are measured as |log(prev^-1 * cur)|)
# ------------------------------------------------------------------------------------------------
def se3_inv_compose(A, B):
    """B^-1 * A for 3x4 [R|t] poses (mrpt: A - B)."""
    Ra, ta, Rb, tb = A[:, :3], A[:, 3], B[:, :3], B[:, 3]
    return np.concatenate([Rb.T @ Ra, (Rb.T @ (ta - tb))[:, None]], axis=1)


def se3_log(T):
    """(v, w) with T = exp([v, w]) — mrpt::poses::Lie::SE<3>::log."""
    R, t = T[:, :3], T[:, 3]
    c = np.clip((np.trace(R) - 1.0) * 0.5, -1.0, 1.0)
    th = np.arccos(c)
    W = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    if th < 1e-9:
        w = 0.5 * W
        Vinv = np.eye(3) - 0.5 * _skew(w)
    else:
        w = th / (2.0 * np.sin(th)) * W
        K = _skew(w)
        Vinv = np.eye(3) - 0.5 * K + (1.0 / th**2) * (1.0 - th * np.sin(th) / (2.0 * (1.0 - np.cos(th)))) * (K @ K)
    return np.concatenate([Vinv @ t, w])


def _skew(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], dtype=np.float64)


def align_loop(step_fn, guess, max_iterations, min_trans, min_rot):
    """The caller of the hot path (ICP::align, ICP.cpp:123-308) reduced to what a timing needs:
    step_fn(pose, iteration) -> (solved, new pose, n_pairs). Returns (pose, iterations, reason)."""
    cur = np.array(guess, dtype=np.float64).reshape(3, 4)
    prev, prev2 = cur.copy(), None
    for it in range(max_iterations):
        ok, new, n = step_fn(cur, it)
        if n == 0:
            return cur, it, "NoPairings"
        if not ok:
            return cur, it, "SolverError"
        cur = np.array(new, dtype=np.float64).reshape(3, 4)
        d = se3_log(se3_inv_compose(cur, prev))
        dt, dr = np.linalg.norm(d[:3]), np.linalg.norm(d[3:])
        if prev2 is not None:
            d2 = se3_log(se3_inv_compose(cur, prev2))
            dt, dr = min(dt, np.linalg.norm(d2[:3])), min(dr, np.linalg.norm(d2[3:]))
        if dt < min_trans and dr < min_rot:
            return cur, it + 1, "Stalled"
        prev2, prev = prev, cur.copy()
    return cur, max_iterations, "MaxIterations"


# ------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------
C5_MAP_POINTS = 100_000_000
C5_QUERIES = 1_000_000
C5_LENGTH = 10_000.0


def c5_queries():
    """1M query points: nine 64-ring scans taken every kilometre along the 10-km street, expressed in
    ONE local frame (the vehicle pose `gt` in the middle of the street)."""
    gt = fx.pose_xyzypr(5000.3, -0.2, 0.1, np.deg2rad(2.0), np.deg2rad(-1.0), np.deg2rad(1.5))
    parts = []
    for j in range(9):
        o = (1000.0 * j + 500.0, 0.4 - 0.1 * j, 0.0)
        s = fx.make_lidar_scan(o, seed=80 + j, length=C5_LENGTH).astype(np.float64)
        parts.append(s + np.asarray(o))
    G = np.concatenate(parts)[:C5_QUERIES]
    L = fx.to_local_frame(G, gt)
    # mid-ICP guess: decimetres off, rotation error kept small against the 5-km lever arm
    pose = fx.pose_xyzypr(5000.3 + 0.12, -0.2 - 0.07, 0.1 + 0.03, np.deg2rad(2.0) + 2e-5, np.deg2rad(-1.0) + 2e-6, np.deg2rad(1.5) - 2e-6)
    return L, gt, pose


def c5_map_device(dev, n=C5_MAP_POINTS, seed=9):
    """The C3 street generator at 10x the length, sampled on the DEVICE (100M points = 1.2 GB; every
    rank draws the same Philox stream). Returns three float32 tensors."""
    import torch

    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    half_width, height, length = 10.0, 12.0, C5_LENGTH
    a_ground, a_wall = length * 2 * half_width, length * height
    n_g = int(n * a_ground / (a_ground + 2 * a_wall))
    n_w = (n - n_g) // 2
    n_w2 = n - n_g - n_w
    x = torch.empty(n, dtype=torch.float32, device=dev).uniform_(0, length, generator=g)
    y = torch.empty(n, dtype=torch.float32, device=dev)
    z = torch.empty(n, dtype=torch.float32, device=dev)
    y[:n_g].uniform_(-half_width, half_width, generator=g)
    z[:n_g] = -1.73
    y[n_g : n_g + n_w] = half_width
    z[n_g : n_g + n_w].uniform_(-1.73, height - 1.73, generator=g)
    y[n_g + n_w :] = -half_width
    z[n_g + n_w :].uniform_(-1.73, height - 1.73, generator=g)
    for t in (x, y, z):  # 5 mm surface roughness
        t.add_(torch.empty(n, dtype=torch.float32, device=dev).normal_(0, 0.005, generator=g))
    return x, y, z


def load_bunny():
    with gzip.open(os.path.join(ROOT, "tests", "golden", "bunny_decim.xyz.gz"), "rt") as f:
        return np.loadtxt(f, dtype=np.float32)


KITTI_SCHEDULE = dict(  # demos/icp-settings-kitti.yaml:10-60
    maxIterations=200, minAbsStep_trans=1e-4, minAbsStep_rot=1e-4, switch_at=6,
    pt2pt=dict(threshold=2.0, thresholdAngularDeg=0.0, pairingsPerPoint=1),
    adaptive=dict(confidenceInterval=0.75, firstToSecondDistanceMax=1.2, absoluteMaxSearchDistance=2.0),
    gn=dict(maxInnerLoopIterations=3, kernel="GemanMcClure", kernelParam=0.15),
)


def make_workload(name: str, shard: int = 0, n_shards: int = 1):
    """Returns dict(map Nx3 f32 | None, local Nx3 f32, pose 3x4, matcher, solver, params...)."""
    if name == "C1":
        # demos/bunny_decim.xyz.gz (10,642 points) against itself moved by the inverse of a
        # GT pose; raw <-> raw with threshold = 0.40 x the largest bbox side (tests/test-mp2p_icp_algos.cpp:166)
        M = load_bunny()
        gt = fx.pose_xyzypr(0.015, -0.010, 0.008, np.deg2rad(4.0), np.deg2rad(-3.0), np.deg2rad(2.0))
        L = fx.to_local_frame(M.astype(np.float64), gt)
        thr = 0.40 * float((M.max(0) - M.min(0)).max())
        return dict(name="C1", map=M, local=L, gt=gt, pose=np.eye(3, 4), matcher="pt2pt", solver="horn",
                    pt2pt=dict(threshold=thr, thresholdAngularDeg=0.0, pairingsPerPoint=1),
                    align=dict(maxIterations=100, minAbsStep_trans=1e-4, minAbsStep_rot=1e-4),
                    desc="bunny_decim (10,642 pts) vs itself under a GT pose, Matcher_Points_DistanceThreshold+Solver_Horn, full align()")
    if name == "C2":
        # 1M uniform map in [0,100)^3, 100k queries (every 10th point + N(0,0.02)),
        # pt2pt threshold 1.0 + Horn. Weak scaling: every rank owns its own 100k-query shard
        # (a different decimation phase of the same map), one joint ICP problem.
        M, _, gt = fx.make_c2(n_map=1_000_000, decim=10)
        rn = np.random.default_rng(4321 + shard)
        Q = M[shard % 10 :: 10].astype(np.float64)[:100_000] + rn.normal(0, 0.02, (100_000, 3))
        L = fx.to_local_frame(Q, gt)
        pose = fx.pose_xyzypr(0.25, -0.15, 0.08, np.deg2rad(1.7), np.deg2rad(-0.8), np.deg2rad(1.2))  # mid-ICP guess
        return dict(name="C2", map=M, local=L, gt=gt, pose=pose, matcher="pt2pt", solver="horn",
                    pt2pt=dict(threshold=1.0, thresholdAngularDeg=0.0, pairingsPerPoint=1),
                    desc="1M-pt uniform map vs 100k-pt query, Matcher_Points_DistanceThreshold(thr=1.0)+Solver_Horn")
    if name in ("C3", "C4"):
        # KITTI-shaped synthetic street, 10M-pt map, 64x1875 scan, pt2pl + GN(3, GM 0.15).
        # Weak scaling: every rank owns its own scan (another sensor position on the same street).
        M = fx.make_street_scene(n_map=10_000_000)
        sensor = (500.0 + 3.0 * shard, 0.4, 0.0)
        S = fx.make_lidar_scan(sensor, seed=8 + shard)
        gt = fx.pose_xyzypr(*sensor, 0.01, 0.0, 0.0)
        pose = fx.pose_xyzypr(sensor[0] + 0.12, sensor[1] - 0.07, 0.03, 0.016, 0.002, -0.002)
        if os.environ.get("MP2P_BENCH_GT_POSE") == "1":  # experiment: no pose error (every query on its surface)
            pose = gt
        w = dict(name="C3", map=M, local=S, gt=gt, pose=pose, matcher="pt2pl", solver="gn",
                 pt2pl=dict(distanceThreshold=0.5, searchRadius=1.0, knn=8, minimumPlanePoints=5, planeEigenThreshold=0.01),
                 gn=dict(maxInnerLoopIterations=3, kernel="GemanMcClure", kernelParam=0.15),
                 desc="10M-pt synthetic street map vs 64-ring scan, Matcher_Point2Plane(knn=8)+Solver_GaussNewton(3,GM0.15)")
        if name == "C4":
            w.update(name="C4", align=KITTI_SCHEDULE,
                     desc="C3 data through demos/icp-settings-kitti.yaml: DistanceThreshold(2.0)+Horn iterations 0-5, Matcher_Adaptive+GaussNewton(3,GM0.15) after, full align()")
        return w
    if name == "C5":
        # 100M-pt map (generated on the device), 1M queries cut in contiguous shards over
        # the GPUs; pt2pt (exact cross-shard first-claim dedup) + Gauss-Newton (6x6 / 6x1 packet all-reduce)
        L, gt, pose = c5_queries()
        return dict(name="C5", map=None, local=L, gt=gt, pose=pose, matcher="pt2pt", solver="gn",
                    pt2pt=dict(threshold=1.0, thresholdAngularDeg=0.0, pairingsPerPoint=1),
                    gn=dict(maxInnerLoopIterations=3, kernel="GemanMcClure", kernelParam=0.15),
                    desc="100M-pt synthetic street map (10 km) vs 1M-pt query (nine scans), Matcher_Points_DistanceThreshold(thr=1.0)+Solver_GaussNewton(3,GM0.15), queries sharded over the GPUs")
    raise SystemExit(f"unknown workload {name}")


# ------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.p, self.index = [], None, index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def host_threads():
    """All the host cores this process may use — torchrun exports OMP_NUM_THREADS=1 to its workers, so the
    reference arm must not take OpenMP's default for the answer."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU oracle (a port: the reference itself needs MRPT)
# ------------------------------------------------------------------------------------------------
def cpu_iteration_fn(w, nthreads, map_xyz=None):
    from oracle import oracle_py as orc

    t0 = time.perf_counter()
    tree = orc.KDTree(*(map_xyz if map_xyz is not None else xyz(w["map"])))
    build_s = time.perf_counter() - t0
    L = xyz(w["local"])

    def step(pose=None, sub=None):
        pose = w["pose"] if pose is None else pose
        l = L if sub is None else tuple(a[:sub] for a in L)
        if w["matcher"] == "pt2pt":
            pairs, _ = orc.match_pt2pt(tree, *l, pose, orc.MatchPt2PtParams(**w["pt2pt"]), nthreads=nthreads)
            if w["solver"] == "horn":
                ok, T = orc.optimal_tf_horn(pairs)
            else:
                ok, T, _ = orc.optimal_tf_gauss_newton(pairs, None, orc.GNParams(**w["gn"]), pose, nthreads=nthreads)
        else:
            pairs, _ = orc.match_pt2pl(tree, *l, pose, orc.MatchPt2PlParams(**w["pt2pl"]), nthreads=nthreads)
            ok, T, _ = orc.optimal_tf_gauss_newton(None, pairs, orc.GNParams(**w["gn"]), pose, nthreads=nthreads)
        return len(pairs), T

    return step, build_s, tree


def cpu_align_fn(w, nthreads, tree=None):
    """Full align() of C1 / C4 on the CPU oracle (the same loop, the oracle's matchers and solvers)."""
    from oracle import oracle_py as orc

    if tree is None:
        tree = orc.KDTree(*xyz(w["map"]))
    L = xyz(w["local"])
    al = w["align"]

    def step(pose, it):
        if w["name"] == "C1" or it < al["switch_at"]:
            pairs, _ = orc.match_pt2pt(tree, *L, pose, orc.MatchPt2PtParams(**(w["pt2pt"] if w["name"] == "C1" else al["pt2pt"])), nthreads=nthreads)
            ok, T = orc.optimal_tf_horn(pairs)
            return ok, T, len(pairs)
        p2p, p2l, _, _ = orc.match_adaptive(tree, *L, pose, orc.MatchAdaptiveParams(**al["adaptive"]), nthreads=nthreads)
        ok, T, _ = orc.optimal_tf_gauss_newton(p2p, p2l if len(p2l) else None, orc.GNParams(**al["gn"]), pose, nthreads=nthreads)
        return ok, T, len(p2p) + len(p2l)

    def run():
        t0 = time.perf_counter()
        T, iters, reason = align_loop(step, w["pose"], al["maxIterations"], al["minAbsStep_trans"], al["minAbsStep_rot"])
        return T, iters, reason, time.perf_counter() - t0

    return run


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (the oracle port) with all the
    host threads it can use, same workload / metric / unit as our arm. Rank 0 alone works."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle_py as orc

    nthreads = host_threads()
    name = args.workload
    base = {"impl": "reference", "metric": METRIC, "unit": "iterations/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "vs_baseline": None, "dtype": DTYPE, "data": "synthetic"}
    if name in ("C1", "C4"):
        w = make_workload(name)
        run = cpu_align_fn(w, nthreads)
        run()
        t0, iters = time.perf_counter(), 0
        for _ in range(max(1, min(args.steps, 3))):
            T, it, reason, _ = run()
            iters += it
        dt = time.perf_counter() - t0
        val = iters / dt
        sample = f"{max(1, min(args.steps, 3))} full align() runs of {w['name']} ({it} iterations each, {reason}), KD-tree build excluded"
        base.update(value=val, ms_per_step=1e3 / val, scaling="weak", config={"workload": w["name"], "detail": w["desc"]})
    else:
        w = make_workload(name)
        map_xyz = None
        sub = None
        if name == "C5":
            import torch

            if not torch.cuda.is_available():
                raise SystemExit("the C5 map is generated on the device (100M points): --impl reference --workload C5 needs a GPU box too")
            mx, my, mz = c5_map_device(torch.device("cuda", 0))
            map_xyz = tuple(t.cpu().numpy() for t in (mx, my, mz))
            del mx, my, mz
            sub = 100_000  # bounded sample: the first 100k of the 1M queries, scaled below
        step, build_s, _ = cpu_iteration_fn(w, nthreads, map_xyz)
        for _ in range(min(args.warmup, 2)):
            step(sub=sub)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            npairs, _ = step(sub=sub)
        dt = (time.perf_counter() - t0) / args.steps
        scale = 1.0 if sub is None else len(w["local"]) / sub
        val = 1.0 / (dt * scale)
        sample = f"{args.steps} {w['name']} iterations over " + (f"all {len(w['local'])} queries" if sub is None else f"the first {sub} of {len(w['local'])} queries (time scaled x{scale:.0f})") + f", KD-tree build ({build_s:.2f} s) excluded"
        base.update(value=val, ms_per_step=1e3 / val, scaling="weak" if name != "C5" else "strong",
                    config={"workload": w["name"], "detail": w["desc"], "pairs": npairs})
    base["cpu_baseline"] = {"value": base["value"], "unit": "iterations/s", "cores": nthreads, "kind": "port", "sample": sample}
    base["e2e"] = {"value": base["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    print(json.dumps(base))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class Bench:
    def __init__(self, args):
        import torch
        import torch.distributed as dist

        import mp2p_icp_b200 as b200

        self.torch, self.dist, self.b200, self.args = torch, dist, b200, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if args.gpus != self.world and self.world > 1:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}")
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        # one explicit stream for everything: our kernels, torch's copies / events, NCCL collectives
        self.stream = torch.cuda.Stream(self.dev)
        torch.cuda.set_stream(self.stream)
        self.ctx = b200.Context(self.local_rank, stream=self.stream.cuda_stream)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)  # > 126 MB L2
        self._matcher_timings = None

    # ---- timing ---------------------------------------------------------------------------------
    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, step_fn, steps, warmup, wall=False, do_flush=True, after=None):
        """W untimed steps, then K steps each bracketed by CUDA events on the launching stream (and the
        wall clock), L2 flushed before each bracket; mean per step, MAX over ranks."""
        torch = self.torch
        for _ in range(warmup):
            step_fn()
        self.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        walls, out = [], None
        for s, e in ev:
            if do_flush:
                self.flush.zero_()  # L2 flush between timed iterations, outside the event bracket
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            s.record(self.stream)
            out = step_fn()
            e.record(self.stream)
            torch.cuda.synchronize()
            walls.append(time.perf_counter() - t0)
            if after is not None:
                after()
        self.barrier()
        ms = [s.elapsed_time(e) for s, e in ev]
        per = float(np.mean(walls) * 1e3) if wall else float(np.mean(ms))
        return self.max_over_ranks(per), out

    def pcie_gbs(self):
        torch = self.torch
        hb = torch.empty(16 << 20, dtype=torch.uint8).pin_memory()
        db = torch.empty(16 << 20, dtype=torch.uint8, device=self.dev)
        best = [0.0, 0.0]
        for _ in range(5):
            for k, (dst, src) in enumerate(((db, hb), (hb, db))):
                s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s0.record(self.stream)
                dst.copy_(src, non_blocking=True)
                e0.record(self.stream)
                torch.cuda.synchronize()
                best[k] = max(best[k], (16 << 20) / (s0.elapsed_time(e0) * 1e-3) / 1e9)
        return best

    # ---- one iteration workload (C2, C3, C5) ------------------------------------------------------
    def build_map(self, w):
        b200, torch = self.b200, self.torch
        if w["map"] is not None:
            return b200.Map(self.ctx, *xyz(w["map"])), len(w["map"])
        mx, my, mz = c5_map_device(self.dev)
        gmap = b200.Map(self.ctx, mx.data_ptr(), my.data_ptr(), mz.data_ptr(), n=mx.numel(), on_device=True)
        torch.cuda.synchronize()
        n = mx.numel()
        del mx, my, mz
        torch.cuda.empty_cache()
        return gmap, n

    def iteration_bench(self, w, steps, warmup, strong=False, gmap=None, n_map=None, with_e2e=True, with_cpu=True, sampler=None):
        """Device-resident value, kernel timings of the same steps, e2e through host buffers, CPU
        baseline. weak scaling: every rank owns a whole workload-sized cloud; strong (C5): the cloud is
        cut over the ranks."""
        b200, torch, ctx, dev, world, rank = self.b200, self.torch, self.ctx, self.dev, self.world, self.rank
        if gmap is None:
            gmap, n_map = self.build_map(w)
        info = gmap.info
        L_all = w["local"]
        n_total = len(L_all) if strong else world * len(L_all)
        if strong:
            per = -(-len(L_all) // world)
            L = L_all[rank * per : min((rank + 1) * per, len(L_all))]
        else:
            L = L_all
        nq = len(L)
        pose = w["pose"]
        pt2pt = w["matcher"] == "pt2pt"
        rec = 36 if pt2pt else 72
        K = w.get("pt2pt", {}).get("pairingsPerPoint", 1)
        cap = max(nq * K, 1)
        d_l = [torch.from_numpy(a).to(dev) for a in xyz(L)]
        d_pairs = torch.empty(cap * rec, dtype=torch.uint8, device=dev)
        mprm = b200.Pt2PtParams(**w["pt2pt"]) if pt2pt else b200.Pt2PlParams(**w["pt2pl"])
        sprm = b200.HornParams() if w["solver"] == "horn" else b200.GNParams(**w["gn"])

        # the local layer is constant over an align(): resident cloud, uploaded + Morton-sorted once
        # (like the map index; reported as config.local_cloud.build_ms, outside the per-iteration time)
        cloud = b200.Cloud(ctx, d_l[0].data_ptr(), d_l[1].data_ptr(), d_l[2].data_ptr(), n=nq, on_device=True) if self.args.local == "cloud" else None
        lp = (cloud, None, None) if cloud is not None else tuple(t.data_ptr() for t in d_l)

        sh = None
        if world > 1:
            from mp2p_icp_b200.sharded import ShardedMatcherSolver

            sh = ShardedMatcherSolver(ctx, gmap, rank, world, n_total if strong else world * nq, k_max=K)
            if sh.n_local != nq:
                raise SystemExit(f"shard sizes disagree: {sh.n_local} vs {nq}")
        fused = None
        if world == 1:  # both plugins ours: fused iteration, pairings stay in HBM, one synchronisation
            if pt2pt and w["solver"] == "gn":
                def fused(T):  # pt2pt + GN (C5 shape): matcher with device output, solver over the device list
                    n, _ = gmap.match_pt2pt(*lp, T, mprm, n_local=nq, local_on_device=True, out=d_pairs.data_ptr(), out_on_device=True, capacity=cap)
                    if self._matcher_timings is not None:  # per-kernel events on: the matcher call's own slots
                        self._matcher_timings = ctx.timings()
                    ok, T2, _ = ctx.solve_gauss_newton(d_pairs.data_ptr(), None, sprm, T, n2p=n, n2l=0, on_device=True)
                    return ok, T2, n
            else:
                fused = gmap.make_iterator(*lp, nq, mprm, sprm, d_pairs.data_ptr(), cap)

        def step_device():
            if fused is not None:
                ok, T, n_pairs = fused(pose)
                return n_pairs, T
            if pt2pt and w["solver"] == "horn":
                ok, T, n_all = sh.iterate_pt2pt_horn(lp, pose, mprm, sprm, d_pairs.data_ptr(), cap)
                return n_all, T
            if pt2pt:
                ok, T, _ = sh.iterate_pt2pt_gn(lp, pose, mprm, sprm, d_pairs.data_ptr(), cap)
                return -1, T
            # pt2pl never dedups global points (Matcher_Point2Plane.cpp:87-90): shards match independently,
            # GN inner loop with the pose on the device and one all-reduce per inner iteration, ONE host sync
            ok, T, _ = sh.iterate_pt2pl_gn(lp, pose, mprm, sprm, d_pairs.data_ptr(), cap)
            return -1, T

        # ---- timed region: device-resident iterations
        l0 = ctx.launch_count
        ms_dev, (n_pairs, T_dev) = self.timed(step_device, steps, warmup)
        launches = (ctx.launch_count - l0) // (steps + warmup)
        ms_warm, _ = self.timed(step_device, steps, 2, do_flush=False)  # informative: no L2 flush

        # ---- the SAME steps once more with the library's per-kernel CUDA events on (roofline)
        kt = {}

        def collect():
            tm = ctx.timings()
            if self._matcher_timings:  # two-call step: matcher slots from the matcher call, the rest from the solver call
                for k in ("nn_search", "plane_fit", "compact"):
                    tm[k] = self._matcher_timings.get(k, 0.0)
                tm["call_total"] += self._matcher_timings.get("call_total", 0.0)
            for k, v in tm.items():
                kt.setdefault(k, []).append(v)

        self._matcher_timings = {}
        ctx.set_profiling(True, False)
        ms_dev_events, _ = self.timed(step_device, max(5, min(steps, 20)), 1, after=collect)
        self._matcher_timings = None
        ctx.set_profiling(False, True)
        if world == 1:
            step_device()
        else:  # the sharded calls do not bracket a single matcher call: one plain matcher call (over this shard alone) for the counters
            d_tmp = torch.empty(cap * rec, dtype=torch.uint8, device=dev)
            (gmap.match_pt2pt if pt2pt else gmap.match_pt2pl)(*lp, pose, mprm, n_local=nq, local_on_device=True, out=d_tmp.data_ptr(), out_on_device=True, capacity=cap)
            del d_tmp
        st = ctx.search_stats()
        ctx.set_profiling(False, False)
        kms = {k: float(np.mean(v)) for k, v in kt.items()}
        self._last_kernel_ms = kms
        if world > 1:  # this rank's share of the pairings: one more sharded step, its count read from the device
            step_device()
            n_pairs_rank = ctx.last_count()
        else:
            n_pairs_rank = int(n_pairs)

        # ---- roofline of the matcher kernels (SURVEY §8d unit sizes)
        peak, peak_src = measured_peaks()
        roof = None
        if world == 1 or not pt2pt:
            npr = n_pairs_rank if n_pairs_rank is not None else 0
            match_ms = kms.get("nn_search", 0.0) + kms.get("plane_fit", 0.0) + kms.get("compact", 0.0)
            alg = nq * 12 + npr * rec + nq * 4 + st["probes"] * 8 + st["candidates"] * 12
            loaded = nq * 12 + npr * rec + nq * 4 + st["probes"] * 16 + st["candidates"] * 16
            achieved = alg / (match_ms * 1e-3) / 1e9 if match_ms > 0 else 0.0
            search_alg = nq * 12 + nq * 4 + st["probes"] * 8 + st["candidates"] * 12
            roof = {"kernel": ("k_match_pt2pt_nn1 + k_compact_pt2pt" if pt2pt and K == 1 else ("k_match_pt2pt<G> + k_compact_pt2pt" if pt2pt else "k_match_pt2pt<8> + k_plane_fit<8> + k_compact_pt2pl")) + " (timed inside the step function of the timed region, library CUDA events)",
                    "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                    "peak_source": peak_src, "kernel_ms": match_ms,
                    "kernel_ms_parts": {k: kms.get(k, 0.0) for k in ("nn_search", "plane_fit", "compact", "gn_accumulate", "horn_sums", "horn_moments", "call_total")},
                    "algorithmic_bytes": int(alg),
                    "algorithmic_bytes_formula": "N_q*12 + pairs*%d + N_q*4 + probes*8 + candidates*12 (SURVEY 8d)" % rec,
                    "counts": {"N_q": nq, "pairs": npr, "probes": st["probes"], "candidates": st["candidates"]},
                    "bytes_loaded_16B_layout": int(loaded),
                    "search_kernel_alone": {"ms": kms.get("nn_search", 0.0), "algorithmic_bytes": int(search_alg),
                                            "achieved": (search_alg / (kms["nn_search"] * 1e-3) / 1e9) if kms.get("nn_search") else 0.0},
                    "step_ms_with_events_on": ms_dev_events,
                    "candidates_per_query": st["candidates"] / max(nq, 1), "probes_per_query": st["probes"] / max(nq, 1),
                    "climbed_queries": st["climbed"],
                    "per_query_max": {"candidates": st["max_candidates_per_query"], "probes": st["max_probes_per_query"], "levels": st["max_levels"], "warps_with_gt2000_candidates": st["heavy_warps"]}}
            roof["search_kernel_alone"]["frac"] = roof["search_kernel_alone"]["achieved"] / peak
            tj = os.path.join(ROOT, "profiles", "r02_traffic.json")
            if os.path.exists(tj):  # written by the SAME visit that produced the committed line (scripts/gpu_visit.sh)
                try:
                    t = json.load(open(tj)).get(w["name"])
                    if t:
                        roof["traffic"] = int(t["dram_bytes_read"] + t["dram_bytes_write"])
                        roof["traffic_source"] = t.get("source")
                except Exception:
                    pass

        # ---- e2e through host buffers: what the reference's ICP loop does per iteration through the two
        # plugin classes (run_matchers then run_solvers, ICP.cpp:143,170)
        e2e = None
        if with_e2e:
            h_l = [torch.from_numpy(a).pin_memory() for a in xyz(L)]
            h_pairs_t = torch.empty(cap * rec, dtype=torch.uint8).pin_memory()
            h_pairs = h_pairs_t.numpy().view(b200.PAIR_PT2PT if rec == 36 else b200.PAIR_PT2PL)
            hx, hy, hz = (t.numpy() for t in h_l)
            plugin_safe = gmap.make_plugin_step(hx, hy, hz, mprm, sprm, h_pairs)  # the plugin classes' defaults
            plugin_reuse = gmap.make_plugin_step(hx, hy, hz, mprm, sprm, h_pairs, reuse_device_pairs=True)
            plugin_nocache = gmap.make_plugin_step(hx, hy, hz, mprm, sprm, h_pairs, cache_local_cloud=False)
            # pageable buffers (std::vector storage) + MatchState bit marshalling: what icp-run would see
            px, py, pz = (np.array(a, copy=True) for a in xyz(L))
            p_pairs = np.empty(cap, b200.PAIR_PT2PT if rec == 36 else b200.PAIR_PT2PL)
            plugin_pageable = gmap.make_plugin_step(px, py, pz, mprm, sprm, p_pairs)
            lbits_bool = np.zeros(nq, dtype=bool)

            def step_pageable():
                b200.pack_bits(lbits_bool)  # MatchState -> bit words, every matcher call (mrpt_plugin.cpp to_bits)
                ok, T, n = plugin_pageable(pose)
                return n, T

            def wrap(f):
                def g():
                    ok, T, n = f(pose)
                    return n, T
                return g

            ms_e2e, (n_e, T_e) = self.timed(wrap(plugin_safe), steps, max(3, warmup), wall=True)
            ms_reuse, (n_r, T_r) = self.timed(wrap(plugin_reuse), steps, 3, wall=True)
            ms_nocache, (n_c, T_c) = self.timed(wrap(plugin_nocache), max(3, steps // 2), 3, wall=True)
            ms_page, (n_p, T_p) = self.timed(step_pageable, max(3, steps // 2), 3, wall=True)
            if n_c != n_e or float(np.abs(np.asarray(T_e) - np.asarray(T_c)).max()) > 1e-9:
                raise SystemExit("e2e: cached and uncached local cloud disagree")
            if n_r != n_e or n_p != n_e or float(np.abs(np.asarray(T_e) - np.asarray(T_r)).max()) > 1e-9 or float(np.abs(np.asarray(T_e) - np.asarray(T_p)).max()) > 1e-9:
                raise SystemExit("e2e: the three host-buffer paths disagree")
            pcie = self.pcie_gbs()
            h2d = 96 + n_e * rec + 96  # pose (matcher; the local layer is resident while its fingerprint holds), pairings + pose/params (solver)
            d2h = n_e * rec + 8 + (512 if w["solver"] == "horn" else 104)
            e2e = {"value": 1e3 / ms_e2e, "unit": "iterations/s", "ms_per_step": ms_e2e,
                   "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                   "timing": "wall clock around the two C-ABI calls, pinned host buffers",
                   "path": "the plugin classes' defaults: matcher call (host local layer, kept on the device while address + size + sampled fingerprint are unchanged; host pairings out) + solver call over those host pairings, uploaded again and compared byte for byte with the matcher's device copy (equal: the result computed during the matcher call's read-back is used)",
                   "local_cloud_uploaded_every_call": {"ms_per_step": ms_nocache, "value": 1e3 / ms_nocache, "h2d_bytes_per_step": int(h2d + nq * 12), "note": "YAML cacheLocalCloud: false"},
                   "assume_unmodified_pairings": {"ms_per_step": ms_reuse, "value": 1e3 / ms_reuse, "h2d_bytes_per_step": 192,
                                                  "note": "opt-in (YAML assumeUnmodifiedPairings / MP2P_B200_PAIRS_LAST_MATCH): the solver reads the device copy the matcher left"},
                   "pageable": {"ms_per_step": ms_page, "value": 1e3 / ms_page, "note": "pageable host arrays for cloud and pairings + MatchState bit marshalling per matcher call"},
                   "pcie_h2d_gbs": pcie[0], "pcie_d2h_gbs": pcie[1],
                   "pcie_floor_ms": (h2d / (pcie[0] * 1e9) + d2h / (pcie[1] * 1e9)) * 1e3}

        # ---- CPU baseline (rank 0, N=1 only), bounded sample
        cpu = None
        if with_cpu and world == 1 and rank == 0 and not self.args.no_cpu_baseline and w["map"] is not None:
            from oracle import oracle_py as orc

            nt = host_threads()
            step, build_s, tree = cpu_iteration_fn(w, nt)
            step()
            reps, t0 = 0, time.perf_counter()
            while reps < 3 or (time.perf_counter() - t0 < 10 and reps < 50):
                n_cpu, T_cpu = step()
                reps += 1
            dt = (time.perf_counter() - t0) / reps
            err = orc.se3_log(orc.inverse_compose(T_dev, T_cpu))
            cpu = {"value": 1.0 / dt, "unit": "iterations/s", "cores": nt, "kind": "port",
                   "sample": f"{reps} full {w['name']} iterations on all host threads; KD-tree build {build_s:.2f} s excluded",
                   "pairs": int(n_cpu), "pose_diff_vs_gpu": float(np.abs(err).max())}
            self._cpu_tree = tree

        unit_scale = 1 if strong else world  # weak: a step processes `world` clouds of the base size
        return {"ms": ms_dev, "value": unit_scale * 1e3 / ms_dev, "ms_warm": ms_warm, "kernel_ms": kms, "ms_with_events": ms_dev_events, "launches": int(launches), "n_pairs": n_pairs, "n_pairs_rank": n_pairs_rank,
                "T": T_dev, "roofline": roof, "e2e": e2e, "cpu": cpu, "nq": nq, "n_total": n_total, "n_map": n_map, "info": info, "cloud": cloud, "gmap": gmap, "sh": sh,
                "d_pairs": d_pairs, "rec": rec, "cap": cap}

    # ---- full align() of C1 / C4 through the plugin-style calls ----------------------------------------
    def align_bench(self, w, gmap=None, reps=3):
        b200, ctx = self.b200, self.ctx
        own_map = gmap is None
        if own_map:
            gmap = b200.Map(ctx, *xyz(w["map"]))
        L = xyz(w["local"])
        nq = len(L[0])
        al = w["align"]
        hx, hy, hz = (self.torch.from_numpy(a).pin_memory().numpy() for a in L)
        cloud = b200.Cloud(ctx, hx, hy, hz)  # the local layer of an align(): uploaded once
        if w["name"] == "C1":
            mprm = b200.Pt2PtParams(**w["pt2pt"])
        else:
            mprm = b200.Pt2PtParams(**al["pt2pt"])
            aprm = b200.AdaptiveParams(**al["adaptive"])
            gprm = b200.GNParams(**al["gn"])
        h_pairs = self.torch.empty(nq * 36, dtype=self.torch.uint8).pin_memory().numpy().view(b200.PAIR_PT2PT)
        first = gmap.make_plugin_step(hx, hy, hz, mprm, b200.HornParams(), h_pairs, reuse_device_pairs=False)
        fused = gmap.make_iterator(cloud, None, None, nq, mprm, b200.HornParams())

        def step_plugin(pose, it):
            if w["name"] == "C1" or it < al["switch_at"]:
                return first(pose)
            p2p, p2l, _, _ = gmap.match_adaptive(hx, hy, hz, pose, aprm)
            if len(p2p) + len(p2l) == 0:
                return False, pose, 0
            ok, T, _ = ctx.solve_gauss_newton(p2p if len(p2p) else None, p2l if len(p2l) else None, gprm, pose)
            return ok, T, len(p2p) + len(p2l)

        def step_device(pose, it):
            if w["name"] == "C1" or it < al["switch_at"]:
                return fused(pose)
            return step_plugin(pose, it)  # Matcher_Adaptive has its MRPT step on the host: two calls

        out = {}
        for tag, fn in (("plugin_calls_host_buffers", step_plugin), ("fused_device_resident", step_device)):
            align_loop(fn, w["pose"], al["maxIterations"], al["minAbsStep_trans"], al["minAbsStep_rot"])  # warm-up
            walls = []
            for _ in range(reps):
                self.torch.cuda.synchronize()
                t0 = time.perf_counter()
                T, iters, reason = align_loop(fn, w["pose"], al["maxIterations"], al["minAbsStep_trans"], al["minAbsStep_rot"])
                walls.append(time.perf_counter() - t0)
            d = se3_log(se3_inv_compose(T, w["gt"]))
            out[tag] = {"wall_ms": float(np.mean(walls) * 1e3), "iterations": int(iters), "termination": reason,
                        "iterations_per_s": iters / float(np.mean(walls)),
                        "pose_error_vs_gt": {"trans_m": [float(x) for x in d[:3]], "rot_rad": [float(x) for x in d[3:]]}}
        res = {"workload": w["name"], "detail": w["desc"], "points": {"map": int(gmap.info["n_points"]), "local": nq},
               "index_build_ms": gmap.info["build_ms"], **out}
        if w["name"] == "C4":
            res["note"] = "the synthetic street is a corridor: the along-street translation (x) is not observable by any point matcher; y, z and the rotation are"
        if not self.args.no_cpu_baseline:
            nt = host_threads()
            run = cpu_align_fn(w, nt, tree=getattr(self, "_cpu_tree", None) if w["name"] == "C4" else None)
            Tc, itc, rc, dt = run()
            from oracle import oracle_py as orc

            err = orc.se3_log(orc.inverse_compose(np.asarray(T), np.asarray(Tc)))
            res["cpu_baseline"] = {"wall_ms": dt * 1e3, "iterations": int(itc), "termination": rc, "cores": nt, "kind": "port",
                                   "iterations_per_s": itc / dt, "pose_diff_vs_gpu": float(np.abs(err).max())}
        if own_map:
            gmap.close()
        return res

    # ---- C5 parity across N: pair count + 64-bit hash of the concatenated records vs a single-GPU run ------
    def c5_parity(self, r, w):
        """Rank 0 runs the WHOLE 1M-query cloud on its own GPU through the single-GPU matcher and compares
        count and hash with the rank-order concatenation of the shards' device outputs."""
        torch, dist, b200 = self.torch, self.dist, self.b200
        import hashlib

        n_rank = r["n_pairs_rank"]
        mine = r["d_pairs"][: n_rank * r["rec"]].cpu().numpy().tobytes()
        if self.world > 1:
            parts = [None] * self.world
            dist.gather_object(mine, parts if self.rank == 0 else None, dst=0)
        else:
            parts = [mine]
        if self.rank != 0:
            return None
        cat = b"".join(parts)
        if self.world == 1:
            return {"pairs": len(cat) // r["rec"], "hash64": hashlib.blake2b(cat, digest_size=8).hexdigest(), "equal": True, "note": "N = 1 is the reference of the comparison"}
        L = xyz(w["local"])
        ref, _ = r["gmap"].match_pt2pt(*L, w["pose"], b200.Pt2PtParams(**w["pt2pt"]))
        rb = ref.tobytes()
        return {"pairs": len(cat) // r["rec"], "hash64": hashlib.blake2b(cat, digest_size=8).hexdigest(),
                "single_gpu_pairs": len(ref), "single_gpu_hash64": hashlib.blake2b(rb, digest_size=8).hexdigest(), "equal": cat == rb}


def run_ours(args):
    B = Bench(args)
    world, rank = B.world, B.rank
    name = args.workload
    sampler = ClockSampler(B.local_rank)
    if rank == 0:
        sampler.start()
    out = None
    if name in ("C1", "C4"):
        if world > 1:
            raise SystemExit("C1 / C4 are single-GPU align() workloads")
        w = make_workload(name)
        res = B.align_bench(w)
        a = res["plugin_calls_host_buffers"]
        d = res["fused_device_resident"]
        out = {"metric": METRIC, "value": d["iterations_per_s"], "unit": "iterations/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": d["wall_ms"] / max(d["iterations"], 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
               "config": {"workload": name, "detail": w["desc"], "timing": "wall clock of whole align() runs (3 repetitions after one warm-up run)"},
               "e2e": {"value": a["iterations_per_s"], "unit": "iterations/s", "ms_per_step": a["wall_ms"] / max(a["iterations"], 1), "h2d_bytes_per_step": None, "d2h_bytes_per_step": None},
               "align": res, "gpu_launches": int(B.ctx.launch_count), "roofline": None, "cpu_baseline": None}
        if "cpu_baseline" in res:
            c = res["cpu_baseline"]
            out["cpu_baseline"] = {"value": c["iterations_per_s"], "unit": "iterations/s", "cores": c["cores"], "kind": "port", "sample": f"one full align() of {name} ({c['iterations']} iterations)"}
    else:
        strong = name == "C5"
        w = make_workload(name, shard=rank, n_shards=world)
        r = B.iteration_bench(w, args.steps, args.warmup, strong=strong, with_e2e=(world == 1), with_cpu=not strong)
        extras = {}
        if strong:
            extras["parity_vs_n1"] = B.c5_parity(r, w)
        # ---- the rest of SURVEY §8d on the default line: align() wall times (N = 1), C5 (every N)
        if name == "C3" and not args.no_extras:
            if world == 1:
                al = {}
                try:
                    al["C1"] = B.align_bench(make_workload("C1"))
                    w4 = dict(w)
                    w4.update(name="C4", align=KITTI_SCHEDULE, desc=make_c4_desc())
                    al["C4"] = B.align_bench(w4, gmap=r["gmap"])
                except Exception as e:  # the headline must not die of an extra
                    al["error"] = repr(e)
                extras["align"] = al
            try:
                extras["c5"] = c5_extra(B, args)
            except Exception as e:
                extras["c5"] = {"error": repr(e)}
        clocks = sampler.stop() if rank == 0 else None
        if rank == 0:
            info, cloud, sh = r["info"], r["cloud"], r["sh"]
            out = {"metric": METRIC, "value": r["value"], "unit": "iterations/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                   "ms_per_step": r["ms"], "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
                   "config": {"workload": w["name"], "detail": w["desc"], "queries_per_gpu": r["nq"], "queries_total": r["n_total"], "map_points": r["n_map"],
                              "pairs": int(r["n_pairs"]) if r["n_pairs"] is not None and r["n_pairs"] >= 0 else r["n_pairs_rank"],
                              "l2": "flushed between timed steps (256 MiB memset, outside the event bracket)", "ms_per_step_l2_warm_informative": r["ms_warm"],
                              "device_path": "fused mp2p_b200_iterate_* call (N=1) / query-sharded iteration with the exchanges inside (N>1)",
                              "collectives": (None if sh is None else ("own kernels over NVLink peer memory (csrc/peer.cu)" if sh.transport == "peer" else "NCCL via torch.distributed")),
                              "unit_note": ("strong scaling: ONE iteration over the whole cloud, cut over the GPUs" if strong else "weak scaling: at N GPUs one step is one query-sharded iteration over N x queries_per_gpu; value counts N iteration-equivalents per step"),
                              "local_cloud": ({"resident": True, "order": "Morton-sorted copy, built once per align()", "build_ms": cloud.info["build_ms"]} if cloud is not None else {"resident": False}),
                              "index": {"build_ms": info["build_ms"], "finest_cell_m": info["finest_cell_size"], "levels": info["n_levels"], "bytes": info["index_bytes"]}},
                   "e2e": r["e2e"] if r["e2e"] is not None else {"value": None, "unit": "iterations/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                                                                  "note": "N > 1: no sharded host-buffer path is offered (the plugin classes are single-GPU); null rather than N replicas"},
                   "gpu_launches": int(r["launches"] * args.steps), "roofline": r["roofline"], "cpu_baseline": r["cpu"], "clocks": clocks,
                   "kernel_ms_rank0": r["kernel_ms"]}
            out.update(extras)
    if rank == 0 and out is not None:
        if "clocks" not in out or out["clocks"] is None:
            out["clocks"] = sampler.stop()
        print(json.dumps(out))
    if world > 1:
        B.dist.destroy_process_group()


def make_c4_desc():
    return "C3 data through demos/icp-settings-kitti.yaml: DistanceThreshold(2.0)+Horn iterations 0-5, Matcher_Adaptive+GaussNewton(3,GM0.15) after, full align()"


def c5_extra(B, args):
    """C5 figures carried by the default line: ONE iteration over the 1M-query cloud against the 100M-point
    map, the queries cut over the N GPUs of this run (strong scaling), pt2pt with exact cross-shard
    first-claim dedup + Gauss-Newton with the 6x6 / 6x1 packet all-reduce. Freed before returning."""
    w = make_workload("C5", shard=B.rank, n_shards=B.world)
    steps = max(5, min(args.steps, 10))
    r = B.iteration_bench(w, steps, 3, strong=True, with_e2e=False, with_cpu=False)
    par = B.c5_parity(r, w)
    res = None
    if B.rank == 0:
        res = {"workload": "C5", "detail": w["desc"], "n_gpus": B.world, "scaling": "strong", "ms_per_step": r["ms"], "value": r["value"], "unit": "iterations/s",
               "steps": steps, "queries_total": r["n_total"], "queries_per_gpu": r["nq"], "map_points": r["n_map"], "index_build_ms": r["info"]["build_ms"],
               "index_bytes": r["info"]["index_bytes"], "gpu_launches_per_step": r["launches"], "parity_vs_n1": par, "roofline": r["roofline"], "kernel_ms_rank0": r["kernel_ms"],
               "ms_per_step_l2_warm_informative": r["ms_warm"]}
    r["gmap"].close()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C3", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="C3 only: skip the align() (C1, C4) and C5 objects of the default line")
    ap.add_argument("--local", default="cloud", choices=["cloud", "arrays"], help="device path: resident Morton-sorted local cloud (default) or plain device arrays in the caller's order")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
