#!/usr/bin/env python
"""bench.py — ICP iterations/s of the mp2p_icp Matcher+Solver hot path on B200 (BASELINE.json).

One "step" = one ICP iteration at a fixed pose: run_matchers (Matcher_Points_DistanceThreshold,
or Matcher_Point2Plane for C3) + run_solvers (Solver_Horn / Solver_GaussNewton) over one batch of
synthetic input, index already built (the reference amortises its KD-tree the same way).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2|C3] [--impl reference]

Prints ONE JSON line (rank 0). Keys: see DESIGN.md "Measurement".
  value     device-resident iterations/s (inputs and pairings stay in HBM; only the pairing count
            and the 3x4 pose cross PCIe) — CUDA events on the launching stream, per step, L2 flushed
            between steps (flush excluded from the timing).
  e2e       the same iteration through the C ABI with HOST (pinned) buffers: local cloud H2D,
            pairings D2H, pairings H2D again for the solver, pose back — wall clock.
  roofline  the NN search kernel (k_match_*): algorithmic bytes / CUDA-event duration vs the
            measured HBM peak in MEASURED_PEAKS.json.
  cpu_baseline  the CPU oracle (line-faithful port of the reference path; MRPT cannot be built
            here) on the same inputs, all host threads.
"""
from __future__ import annotations

import argparse
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


def xyz(a):
    return np.ascontiguousarray(a[:, 0]), np.ascontiguousarray(a[:, 1]), np.ascontiguousarray(a[:, 2])


# ------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------
def make_workload(name: str, shard: int = 0, n_shards: int = 1):
    """Returns dict(map Nx3 f32, local Nx3 f32, pose 3x4, kind, params...)."""
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
    if name == "C3":
        # KITTI-shaped synthetic street, 10M-pt map, 64x1875 scan, pt2pl + GN(3, GM 0.15)
        M = fx.make_street_scene(n_map=10_000_000)
        sensor = (500.0 + 3.0 * shard, 0.4, 0.0)
        S = fx.make_lidar_scan(sensor, seed=8 + shard)
        gt = fx.pose_xyzypr(*sensor, 0.01, 0.0, 0.0)
        pose = fx.pose_xyzypr(sensor[0] + 0.12, sensor[1] - 0.07, 0.03, 0.016, 0.002, -0.002)
        return dict(name="C3", map=M, local=S, gt=gt, pose=pose, matcher="pt2pl", solver="gn",
                    pt2pl=dict(distanceThreshold=0.5, searchRadius=1.0, knn=8, minimumPlanePoints=5, planeEigenThreshold=0.01),
                    gn=dict(maxInnerLoopIterations=3, kernel="GemanMcClure", kernelParam=0.15),
                    desc="10M-pt synthetic street map vs 64-ring scan, Matcher_Point2Plane(knn=8)+Solver_GaussNewton(3,GM0.15)")
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


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU oracle (a port: the reference itself needs MRPT)
# ------------------------------------------------------------------------------------------------
def cpu_iteration_fn(w, nthreads):
    from oracle import oracle_py as orc

    t0 = time.perf_counter()
    tree = orc.KDTree(*xyz(w["map"]))
    build_s = time.perf_counter() - t0
    L = xyz(w["local"])

    def step():
        if w["matcher"] == "pt2pt":
            pairs, _ = orc.match_pt2pt(tree, *L, w["pose"], orc.MatchPt2PtParams(**w["pt2pt"]), nthreads=nthreads)
            ok, T = orc.optimal_tf_horn(pairs)
        else:
            pairs, _ = orc.match_pt2pl(tree, *L, w["pose"], orc.MatchPt2PlParams(**w["pt2pl"]), nthreads=nthreads)
            ok, T, _ = orc.optimal_tf_gauss_newton(None, pairs, orc.GNParams(**w["gn"]), w["pose"], nthreads=nthreads)
        return len(pairs), T

    return step, build_s


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle_py as orc

    w = make_workload(args.workload)
    nthreads = orc.max_threads()
    step, build_s = cpu_iteration_fn(w, nthreads)
    for _ in range(min(args.warmup, 2)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        npairs, _ = step()
    dt = (time.perf_counter() - t0) / args.steps
    val = 1.0 / dt
    sample = f"{args.steps} full {w['name']} iterations (all {len(w['local'])} queries), KD-tree build ({build_s:.2f} s) excluded"
    print(json.dumps({
        "impl": "reference", "metric": "ICP iterations/sec", "value": val, "unit": "iterations/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 metric / f64 solve", "data": "synthetic",
        "config": {"workload": w["name"], "detail": w["desc"], "pairs": npairs},
        "cpu_baseline": {"value": val, "unit": "iterations/s", "cores": nthreads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import mp2p_icp_b200 as b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    w = make_workload(args.workload, shard=rank, n_shards=world)
    # one explicit stream for everything: our kernels, torch's copies / events, NCCL collectives
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    ctx = b200.Context(local_rank, stream=stream.cuda_stream)
    gmap = b200.Map(ctx, *xyz(w["map"]))
    info = gmap.info
    nq = len(w["local"])
    pose = w["pose"]

    # ---- device-resident buffers (value) and pinned host buffers (e2e)
    d_l = [torch.from_numpy(a).to(dev) for a in xyz(w["local"])]
    rec = 36 if w["matcher"] == "pt2pt" else 72
    K = w.get("pt2pt", {}).get("pairingsPerPoint", 1)
    cap = nq * K
    d_pairs = torch.empty(cap * rec, dtype=torch.uint8, device=dev)
    h_l = [torch.from_numpy(a).pin_memory() for a in xyz(w["local"])]
    h_pairs_t = torch.empty(cap * rec, dtype=torch.uint8).pin_memory()
    h_pairs = h_pairs_t.numpy().view(b200.PAIR_PT2PT if rec == 36 else b200.PAIR_PT2PL)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    if w["matcher"] == "pt2pt":
        mprm, sprm = b200.Pt2PtParams(**w["pt2pt"]), b200.HornParams()
    else:
        mprm, sprm = b200.Pt2PlParams(**w["pt2pl"]), b200.GNParams(**w["gn"])

    sh = None
    if world > 1:
        from mp2p_icp_b200.sharded import ShardedMatcherSolver

        sh = ShardedMatcherSolver(ctx, gmap, rank, world, world * nq, k_max=K)

    def solve_device(n_pairs):
        if world == 1:
            if w["solver"] == "horn":
                return ctx.solve_horn(d_pairs.data_ptr(), n=n_pairs, prm=sprm, on_device=True)[1]
            return ctx.solve_gauss_newton(None, d_pairs.data_ptr(), sprm, pose, n2p=0, n2l=n_pairs, on_device=True)[1]
        # query-sharded: all-reduce the 32-double accumulator packets
        if w["solver"] == "horn":
            return sh.solve_horn(d_pairs.data_ptr(), n_pairs, sprm)[1]
        return sh.solve_gauss_newton(None, 0, d_pairs.data_ptr(), n_pairs, sprm, pose)[1]

    # the local layer is constant over an align(): resident cloud, uploaded + Morton-sorted once
    # (like the map index; reported as config.local_cloud.build_ms, outside the per-iteration time)
    cloud = None
    if args.local == "cloud":
        cloud = b200.Cloud(ctx, d_l[0].data_ptr(), d_l[1].data_ptr(), d_l[2].data_ptr(), n=nq, on_device=True)
    lp = (cloud, None, None) if cloud is not None else (d_l[0].data_ptr(), d_l[1].data_ptr(), d_l[2].data_ptr())
    fused = None
    if world == 1:  # both plugins ours: fused iteration, pairings stay in HBM, one synchronisation
        fused = gmap.make_iterator(*lp, nq, mprm, sprm, d_pairs.data_ptr(), cap)

    def step_device():
        if fused is not None:
            ok, T, n_pairs = fused(pose)
            return n_pairs, T
        if w["matcher"] == "pt2pt":
            # exact cross-shard first-claim dedup: search -> in-place all_gather of the exchange
            # records -> resolve (+HORN1 sums) -> all_reduce -> HORN2 -> all_reduce, ONE host sync
            ok, T, n_all = sh.iterate_pt2pt_horn(lp, pose, mprm, sprm, d_pairs.data_ptr(), cap)
            return n_all, T
        # pt2pl never dedups global points (Matcher_Point2Plane.cpp:87-90): shards match independently,
        # GN inner loop with the pose on the device and one all_reduce per inner iteration, ONE host sync
        ok, T, updates = sh.iterate_pt2pl_gn(lp, pose, mprm, sprm, d_pairs.data_ptr(), cap)
        return -1, T

    # e2e = what the reference's ICP loop does per iteration through the two plugin classes over HOST
    # buffers: matcher call (local cloud H2D, pairings D2H into the caller's Pairings), then solver
    # call over those host pairings. The plugin's solver recognises the pairings as the unmodified
    # output of the matcher call just before (count + sample witness) and lets the library read the
    # copy still on the device (MP2P_B200_PAIRS_LAST_MATCH); `plugin_upload` uploads them again.
    hx, hy, hz = (t.numpy() for t in h_l)
    plugin_reuse = gmap.make_plugin_step(hx, hy, hz, mprm, sprm, h_pairs, reuse_device_pairs=True)
    plugin_upload = gmap.make_plugin_step(hx, hy, hz, mprm, sprm, h_pairs, reuse_device_pairs=False)
    # informative: the fused call fed from HOST arrays (local cloud H2D each step, pairings stay in HBM)
    fused_host = gmap.make_iterator(hx, hy, hz, nq, mprm, sprm, d_pairs.data_ptr(), cap, local_on_device=False) if world == 1 else None

    def step_e2e():
        ok, T, n = plugin_reuse(pose)
        return n, T

    def step_e2e_upload():
        ok, T, n = plugin_upload(pose)
        return n, T

    def step_e2e_fused():
        ok, T, n = fused_host(pose)
        return n, T

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps, warmup, wall=False, do_flush=True):
        for _ in range(warmup):
            step_fn()
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        walls, out = [], None
        for s, e in ev:
            if do_flush:
                flush.zero_()  # L2 flush between timed iterations, outside the event bracket
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            s.record(stream)
            out = step_fn()
            e.record(stream)
            torch.cuda.synchronize()
            walls.append(time.perf_counter() - t0)
        barrier()
        ms = [s.elapsed_time(e) for s, e in ev]
        per = float(np.mean(walls) * 1e3) if wall else float(np.mean(ms))
        if world > 1:
            t = torch.tensor([per], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            per = float(t.item())
        return per, out

    # ---- timed region: device-resident iterations
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count
    ms_dev, (n_pairs, T_dev) = timed(step_device, args.steps, args.warmup)
    launches = (ctx.launch_count - l0) // (args.steps + args.warmup)
    # informative only: the same steps WITHOUT the L2 flush (consecutive ICP iterations of a real
    # align() find the map hot in the 126 MB L2); never used for `value`
    ms_warm, _ = timed(step_device, args.steps, 2, do_flush=False)
    # ---- e2e through host buffers
    ms_e2e, (n_pairs_e, T_e2e) = timed(step_e2e, args.steps, max(3, args.warmup), wall=True)
    ms_e2e_upload, (n_pairs_u, T_e2e_u) = timed(step_e2e_upload, args.steps, 3, wall=True)
    ms_e2e_fused = timed(step_e2e_fused, args.steps, 3, wall=True)[0] if fused_host is not None else None
    if n_pairs_u != n_pairs_e or float(np.abs(np.asarray(T_e2e) - np.asarray(T_e2e_u)).max()) > 1e-9:
        raise SystemExit("e2e: solver over the device copy and over re-uploaded pairings disagree")
    clocks = sampler.stop() if rank == 0 else None

    # ---- PCIe reference for the e2e number: pinned 16 MiB H2D and D2H copies (best of 5)
    def pcie_gbs():
        hb = torch.empty(16 << 20, dtype=torch.uint8).pin_memory()
        db = torch.empty(16 << 20, dtype=torch.uint8, device=dev)
        best = [0.0, 0.0]
        for _ in range(5):
            for k, (dst, src) in enumerate(((db, hb), (hb, db))):
                s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s0.record(stream)
                dst.copy_(src, non_blocking=True)
                e0.record(stream)
                torch.cuda.synchronize()
                best[k] = max(best[k], (16 << 20) / (s0.elapsed_time(e0) * 1e-3) / 1e9)
        return best

    pcie = pcie_gbs()

    # ---- roofline pass: per-kernel CUDA events (stats OFF: the counters add same-address atomics),
    # then ONE untimed call with the search statistics on. Not part of the timed region above.
    def match_once():
        if w["matcher"] == "pt2pt":
            gmap.match_pt2pt(*lp, pose, mprm, n_local=nq, local_on_device=True, out=d_pairs.data_ptr(), out_on_device=True, capacity=cap)
        else:
            gmap.match_pt2pl(*lp, pose, mprm, n_local=nq, local_on_device=True, out=d_pairs.data_ptr(), out_on_device=True, capacity=cap)

    ctx.set_profiling(True, False)
    nn_ms, tm = [], {}
    for _ in range(max(5, args.steps)):
        flush.zero_()
        torch.cuda.synchronize()
        match_once()
        tm = ctx.timings()
        nn_ms.append(tm["nn_search"])
    ctx.set_profiling(False, True)
    match_once()
    st = ctx.search_stats()
    ctx.set_profiling(False, False)
    k_out = K if w["matcher"] == "pt2pt" else 0
    # algorithmic bytes of ONE launch of the NN search kernel (DESIGN.md "Roofline"):
    #   query read 12 B + per hash probe 16 B + per candidate point 16 B + outputs
    out_bytes = nq * k_out * 8 + st["valid"] * 8 if w["matcher"] == "pt2pt" else nq * (1 + 0) + n_pairs * 56
    alg_bytes = nq * 12 + st["probes"] * 16 + st["candidates"] * 16 + out_bytes
    nn_ms_mean = float(np.mean(nn_ms))
    peak, peak_src = measured_peaks()
    achieved = alg_bytes / (nn_ms_mean * 1e-3) / 1e9

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # DRAM traffic of the dominant kernel: from the committed `ncu --set full` capture of this command
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        traffic = tj[w["name"]]["dram_bytes_read"] + tj[w["name"]]["dram_bytes_write"]
    except Exception:
        pass

    # ---- CPU baseline (rank 0, N=1 only), bounded sample
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import oracle_py as orc

        nt = orc.max_threads()
        step, build_s = cpu_iteration_fn(w, nt)
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

    unit_scale = world  # weak scaling: a step processes `world` shards of the base query count
    h2d = nq * 12 + 96 + 96  # local cloud + pose (matcher) + pose/params (solver); pairings are NOT uploaded again
    d2h = n_pairs_e * rec + 8 + (512 if w["solver"] == "horn" else 104)
    out = {
        "metric": "ICP iterations/sec", "value": unit_scale * 1e3 / ms_dev, "unit": "iterations/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 metric / f64 solve", "data": "synthetic",
        "config": {"workload": w["name"], "detail": w["desc"], "queries_per_gpu": nq, "map_points": len(w["map"]),
                   "pairs": int(n_pairs) if n_pairs >= 0 else None, "l2": "flushed between timed steps (256 MiB memset, outside the event bracket)",
                   "ms_per_step_l2_warm_informative": ms_warm,
                   "device_path": "fused mp2p_b200_iterate_* call (N=1) / sharded search+all_gather+resolve+all_reduce (N>1)",
                   "collectives": (None if sh is None else ("own kernels over NVLink peer memory (csrc/peer.cu)" if sh.transport == "peer" else "NCCL via torch.distributed")), "unit_note": "at N GPUs one step is one query-sharded iteration over N x queries_per_gpu; value counts N iteration-equivalents per step",
                   "local_cloud": ({"resident": True, "order": "Morton-sorted copy, built once per align()", "build_ms": cloud.info["build_ms"]} if cloud is not None else {"resident": False}),
                   "index": {"build_ms": info["build_ms"], "finest_cell_m": info["finest_cell_size"], "levels": info["n_levels"], "bytes": info["index_bytes"]}},
        "e2e": {"value": unit_scale * 1e3 / ms_e2e, "unit": "iterations/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "timing": "wall clock around the C-ABI calls, pinned host buffers",
                "path": "matcher call (host local cloud in, host pairings out) + solver call over the same host pairings, solver reads the device copy the matcher left (MP2P_B200_PAIRS_LAST_MATCH)",
                "ms_per_step_pairs_uploaded_again": ms_e2e_upload, "h2d_bytes_pairs_uploaded_again": int(h2d + n_pairs_e * rec),
                "ms_per_step_fused_call_host_cloud": ms_e2e_fused,
                "pcie_h2d_gbs": pcie[0], "pcie_d2h_gbs": pcie[1],
                "pcie_floor_ms": (h2d / (pcie[0] * 1e9) + d2h / (pcie[1] * 1e9)) * 1e3},
        "gpu_launches": int(launches * args.steps),
        "roofline": {"kernel": "k_match_" + w["matcher"], "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src, "kernel_ms": nn_ms_mean,
                     "algorithmic_bytes": int(alg_bytes), "probes": st["probes"], "candidates": st["candidates"],
                     "climbed_queries": st["climbed"], "per_query_max": {"candidates": st["max_candidates_per_query"], "probes": st["max_probes_per_query"], "levels": st["max_levels"], "warps_with_gt2000_candidates": st["heavy_warps"]}, "other_kernels_ms": {k: v for k, v in tm.items() if k != "nn_search"}},
        "cpu_baseline": cpu,
        "clocks": clocks,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=["C2", "C3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--local", default="cloud", choices=["cloud", "arrays"], help="device path: resident Morton-sorted local cloud (default) or plain device arrays in the caller's order")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
