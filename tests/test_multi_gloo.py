"""world_size-2 tests on CPU (gloo) of the multi-GPU host logic: shard bookkeeping,
packet all-reduce and the redundant per-rank finish. The per-shard accumulation — a CUDA kernel on
the GPU — is played by the CPU oracle here (tests may use it), so what is verified is exactly the
code that runs between the kernels in a sharded run: mp2p_icp_b200.sharded + the host-side C-ABI
functions mp2p_b200_gn_step_from_packet / mp2p_b200_horn_finish."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mp2p_icp_b200 import capi, sharded
from oracle import oracle_py as orc

DEG = np.pi / 180.0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_problem():
    rng = np.random.default_rng(77)
    n = 20_001
    A = rng.uniform(0, 50, (n, 3))
    gt = orc.pose_from_xyzypr(0.4, -0.3, 0.2, 3 * DEG, -2 * DEG, 1 * DEG)
    B = (A - gt[:, 3]) @ gt[:, :3] + rng.normal(0, 0.02, (n, 3))
    p2p = np.zeros(n, orc.PAIR_PT2PT)
    p2p["global"], p2p["local"] = A, B
    nrm = rng.normal(0, 1, (5000, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    c = rng.uniform(0, 50, (5000, 3))
    p2l = np.zeros(5000, orc.PAIR_PT2PL)
    p2l["coefs"][:, :3], p2l["coefs"][:, 3] = nrm, -(nrm * c).sum(1)
    p2l["local"] = (c - gt[:, 3]) @ gt[:, :3]
    return p2p, p2l, gt


def _gn_packet(p2p, p2l, T, prm):
    H, g, e = orc.gn_accumulate(p2p, p2l, T, orc.GNParams(**prm))
    pk = np.zeros(32)
    pk[:21], pk[21:27], pk[27], pk[28] = H[np.triu_indices(6)], g, e, len(p2p) + len(p2l)
    return torch.from_numpy(pk)


def _horn_packets(p2p):
    def sums():
        pk = np.zeros(32)
        pk[0:3], pk[3:6], pk[6] = p2p["local"].astype(np.float64).sum(0), p2p["global"].astype(np.float64).sum(0), len(p2p)
        return torch.from_numpy(pk)

    def moments(s, n_total):
        s = s.numpy()
        cl, cg = s[0:3] / s[6], s[3:6] / s[6]
        r, b = p2p["local"].astype(np.float64) - cl, p2p["global"].astype(np.float64) - cg
        w = 1.0 / n_total
        pk = np.zeros(32)
        pk[:9] = (w * r.T @ b).reshape(-1)
        pk[9], pk[11] = w * len(p2p), len(p2p)
        return torch.from_numpy(pk)

    return sums, moments


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p2p, p2l, gt = _make_problem()
        b2p, _ = sharded.shard_bounds(len(p2p), world)
        b2l, _ = sharded.shard_bounds(len(p2l), world)
        my2p, my2l = p2p[b2p[rank] : b2p[rank + 1]], p2l[b2l[rank] : b2l[rank + 1]]
        prm = dict(maxInnerLoopIterations=6, kernel="GemanMcClure", kernelParam=0.5, w_pt2pl=0.8)
        ok, T_gn, it = sharded.allreduce_gn_solve(lambda T: _gn_packet(my2p, my2l, T, prm), capi.GNParams(**prm), np.eye(3, 4), dist)
        ok_h, T_h = sharded.allreduce_horn_solve(*_horn_packets(my2p), dist)
        q.put((rank, T_gn, it, ok_h, T_h))
    finally:
        dist.destroy_process_group()


def test_shard_bounds():
    b, per = sharded.shard_bounds(10, 4)
    assert b == [0, 3, 6, 9, 10] and per == 3
    b, per = sharded.shard_bounds(8, 2)
    assert b == [0, 4, 8] and per == 4
    b, per = sharded.shard_bounds(2, 4)
    assert b == [0, 1, 2, 2, 2] and per == 1


@pytest.mark.timeout(300)
def test_two_rank_allreduce_solvers_match_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    p2p, p2l, gt = _make_problem()
    prm = dict(maxInnerLoopIterations=6, kernel="GemanMcClure", kernelParam=0.5, w_pt2pl=0.8)
    ok, T_ref, it_ref = orc.optimal_tf_gauss_newton(p2p, p2l, orc.GNParams(**prm), np.eye(3, 4))
    ok, T_horn = orc.optimal_tf_horn(p2p)
    for rank, T_gn, it, ok_h, T_h in res:
        assert it == it_ref
        assert np.abs(orc.se3_log(orc.inverse_compose(T_gn, T_ref))).max() < 1e-9
        assert ok_h and np.abs(orc.se3_log(orc.inverse_compose(T_h, T_horn))).max() < 1e-9
    # every rank finishes redundantly with bit-identical poses (same reduced packet, same host math)
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][4], res[1][4])
