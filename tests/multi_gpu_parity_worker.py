"""torchrun worker: N-GPU query-sharded iteration == 1-GPU iteration (SURVEY.md §4 plan item 5).

Every rank builds the same map, owns a contiguous shard of the local cloud, runs the sharded
matcher (search -> NCCL all_gather -> resolve) and the all-reduced Horn / GN solvers. Rank 0 also
runs the whole cloud on its own GPU and checks: concatenated pairings identical bit for bit, poses
within 1e-9 (different reduction grouping only).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import mp2p_icp_b200 as b200  # noqa: E402
from mp2p_icp_b200.sharded import ShardedMatcherSolver  # noqa: E402
from tests import fixtures as fx  # noqa: E402


def xyz(a):
    return np.ascontiguousarray(a[:, 0]), np.ascontiguousarray(a[:, 1]), np.ascontiguousarray(a[:, 2])


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    M, L, gt = fx.make_c2(n_map=300_000, decim=5)
    L = np.concatenate([L, L[:5000] + np.float32(2e-3)])  # duplicate claims across shards
    L = L[: (len(L) // world) * world]
    pose = fx.pose_xyzypr(0.25, -0.15, 0.08, np.deg2rad(1.7), np.deg2rad(-0.8), np.deg2rad(1.2))
    torch.cuda.set_stream(torch.cuda.Stream(dev))  # one stream for our kernels, torch and NCCL
    ctx = b200.Context(lr, stream=torch.cuda.current_stream().cuda_stream)
    gmap = b200.Map(ctx, *xyz(M))
    n_total = len(L)
    sh = ShardedMatcherSolver(ctx, gmap, rank, world, n_total, k_max=1)
    mine = L[sh.lo : sh.hi]
    d_l = [torch.from_numpy(a).to(dev) for a in xyz(mine)]
    d_pairs = torch.zeros(len(mine) * 36, dtype=torch.uint8, device=dev)
    prm = b200.Pt2PtParams(threshold=1.0)
    n_pairs = sh.match_pt2pt(d_l[0].data_ptr(), d_l[1].data_ptr(), d_l[2].data_ptr(), pose, prm, d_pairs.data_ptr(), len(mine))
    ok_h, T_h = sh.solve_horn(d_pairs.data_ptr(), n_pairs, b200.HornParams())
    gprm = b200.GNParams(maxInnerLoopIterations=4, kernel="Cauchy", kernelParam=0.3)
    ok_g, T_g, it_g = sh.solve_gauss_newton(d_pairs.data_ptr(), n_pairs, None, 0, gprm, pose)
    # the same iteration through the one-synchronisation path, on a resident (Morton-sorted) shard
    cloud = b200.Cloud(ctx, *xyz(mine))
    d_pairs2 = torch.zeros(len(mine) * 36, dtype=torch.uint8, device=dev)
    ok_f, T_f, n_all = sh.iterate_pt2pt_horn((cloud, None, None), pose, prm, b200.HornParams(), d_pairs2.data_ptr(), len(mine))
    d_pairs3 = torch.zeros(len(mine) * 36, dtype=torch.uint8, device=dev)
    ok_fg, T_fg, it_fg = sh.iterate_pt2pt_gn((cloud, None, None), pose, prm, gprm, d_pairs3.data_ptr(), len(mine))
    # repeated calls (mailbox parities alternate, epochs grow) and the plain-array form
    ok_f2, T_f2, n_all2 = True, T_f, n_all
    for _ in range(3):
        ok_f2, T_f2, n_all2 = sh.iterate_pt2pt_horn((d_l[0].data_ptr(), d_l[1].data_ptr(), d_l[2].data_ptr()), pose, prm, b200.HornParams(), d_pairs2.data_ptr(), len(mine))
    torch.cuda.synchronize()
    fused_same = bool((d_pairs2[: n_pairs * 36] == d_pairs[: n_pairs * 36]).all().item()) and bool((d_pairs3[: n_pairs * 36] == d_pairs[: n_pairs * 36]).all().item())
    df = max(np.abs(T_f - T_h).max(), np.abs(T_fg - T_g).max(), np.abs(T_f2 - T_h).max())
    fused_ok = int(fused_same and ok_f and ok_fg and ok_f2 and n_all2 == n_all and df < 1e-9 and it_fg == it_g)
    # a cloud that does not divide evenly: the last shard is short (its record is padded)
    Lu = L[: len(L) - 37]
    shu = ShardedMatcherSolver(ctx, gmap, rank, world, len(Lu), k_max=1)
    mu = Lu[shu.lo : shu.hi]
    d_pu = torch.zeros(max(len(mu), 1) * 36, dtype=torch.uint8, device=dev)
    ok_u, T_u, n_all_u = shu.iterate_pt2pt_horn((b200.Cloud(ctx, *xyz(mu)), None, None), pose, prm, b200.HornParams(), d_pu.data_ptr(), len(mu))
    if rank == 0:
        ref_u, _ = gmap.match_pt2pt(*xyz(Lu), pose, prm)
        ok_ru, T_ru = ctx.solve_horn(ref_u)
        mine_u = ref_u[ref_u["localIdx"] < shu.hi]
        got_u = d_pu[: len(mine_u) * 36].cpu().numpy().view(b200.PAIR_PT2PT)
        fused_ok = int(fused_ok and ok_u and n_all_u == len(ref_u) and np.abs(T_u - T_ru).max() < 1e-9 and got_u.tobytes() == mine_u.tobytes())
    fused_ok = torch.tensor([fused_ok], device=dev)
    dist.all_reduce(fused_ok, op=dist.ReduceOp.MIN)
    # pt2pl + Gauss-Newton over a sharded scan (independent shards, device-resident GN loop)
    Ms = fx.make_street_scene(n_map=200_000, length=40.0)
    S = fx.make_lidar_scan((20.0, 0.5, 0.0), n_rings=32, n_az=400, length=40.0)
    S = S[: (len(S) // world) * world]
    g2 = fx.pose_xyzypr(20.08, 0.46, 0.02, 0.02, 0.001, -0.001)
    smap = b200.Map(ctx, *xyz(Ms))
    sh2 = ShardedMatcherSolver(ctx, smap, rank, world, len(S), k_max=1)
    mine2 = S[sh2.lo : sh2.hi]
    mkw = b200.Pt2PlParams(distanceThreshold=0.5, searchRadius=1.0, knn=8, minimumPlanePoints=5, planeEigenThreshold=0.01)
    skw = b200.GNParams(maxInnerLoopIterations=3, kernel="GemanMcClure", kernelParam=0.15)
    d_pl = torch.zeros(len(mine2) * 72, dtype=torch.uint8, device=dev)
    ok_p, T_p, it_p = sh2.iterate_pt2pl_gn((b200.Cloud(ctx, *xyz(mine2)), None, None), g2, mkw, skw, d_pl.data_ptr(), len(mine2))
    # gather the shards' pairings on rank 0
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([n_pairs], dtype=torch.int64, device=dev))
    bufs = [torch.zeros(len(mine) * 36, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(bufs, d_pairs)
    fail = 0
    if rank == 0:
        got = np.concatenate([b.cpu().numpy().view(b200.PAIR_PT2PT)[: int(c.item())] for b, c in zip(bufs, counts)])
        ref, _ = gmap.match_pt2pt(*xyz(L), pose, prm)
        ok_r, T_r = ctx.solve_horn(ref)
        ok_r2, T_r2, it_r = ctx.solve_gauss_newton(ref, None, gprm, pose)
        same = len(got) == len(ref) and got.tobytes() == ref.tobytes()
        dh, dg = np.abs(T_h - T_r).max(), np.abs(T_g - T_r2).max()
        print(f"world={world} transport={sh.transport} pairs={len(got)} identical={same} horn_diff={dh:.2e} gn_diff={dg:.2e} gn_iters={it_g}/{it_r} one_sync_path_ok={int(fused_ok.item())} n_all={n_all}")
        ref_l, _ = smap.match_pt2pl(*xyz(S), g2, mkw)
        ok_pr, T_pr, it_pr = ctx.solve_gauss_newton(None, ref_l, skw, g2)
        dp = np.abs(T_p - T_pr).max()
        print(f"pt2pl+GN sharded: pairs(ref)={len(ref_l)} pose_diff={dp:.2e} iters={it_p}/{it_pr}")
        same = same and ok_p and ok_pr and dp < 1e-9 and it_p == it_pr
        fail = int(not (same and ok_h and ok_g and dh < 1e-9 and dg < 1e-9 and it_g == it_r and int(fused_ok.item()) == 1 and n_all == len(ref)))
    t = torch.tensor([fail], device=dev)
    dist.broadcast(t, 0)
    dist.destroy_process_group()
    sys.exit(int(t.item()))


if __name__ == "__main__":
    main()
