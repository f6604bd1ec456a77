"""Query-sharded multi-GPU driver: one process per GPU, torch.distributed (NCCL over
NVLink on GPUs, gloo in the CPU tests) as plumbing around the C-ABI building blocks.

The local cloud is split into contiguous shards (rank r owns [offsets[r], offsets[r+1])), the map
and its index are replicated. Per ICP iteration:

  matcher   shard_search (phase A)  ->  ONE all_gather of [candidate words | shard bbox]
            ->  shard_resolve (phase B: global first-claim replay + compaction of the own shard)
  solver    accumulate the shard into a 32-double packet -> all_reduce(SUM) -> every rank finishes the
            4x4 / 6x6 solve redundantly (mp2p_b200_horn_finish / mp2p_b200_gn_step_from_packet).

The reduce/finish helpers take the accumulation as a callable so that the same host logic runs in
the gloo CPU tests with a stand-in accumulator.
"""
from __future__ import annotations

import numpy as np

from . import capi


def shard_bounds(n_total: int, world: int):
    """Contiguous equal shards (the last may be shorter); all_gather needs equal chunk sizes, so
    every rank's buffers are padded to `per`."""
    per = -(-n_total // world)
    return [min(r * per, n_total) for r in range(world + 1)], per


def allreduce_horn_solve(sums_fn, moments_fn, dist, group=None):
    """sums_fn() -> tensor[32] (HORN1 packet of the shard); moments_fn(sums_tensor, n_total) ->
    tensor[32] (HORN2 packet). Returns (solved, pose 3x4)."""
    sums = sums_fn()
    dist.all_reduce(sums, group=group)
    n_total = int(round(float(sums[6])))
    if n_total < 3:  # optimal_tf_horn.cpp:96
        return False, np.eye(3, 4)
    mom = moments_fn(sums, n_total)
    dist.all_reduce(mom, group=group)
    return capi.horn_finish(sums.detach().cpu().numpy(), mom.detach().cpu().numpy())


def allreduce_gn_solve(accumulate_fn, prm: capi.GNParams, pose0, dist, group=None):
    """accumulate_fn(pose 3x4) -> tensor[32] (GN packet of the shard at that pose).
    optimal_tf_gauss_newton.cpp:70-366 with the reduction spread over ranks."""
    T = np.array(pose0, dtype=np.float64).reshape(3, 4)
    updates = 0  # number of pose updates applied (what mp2p_b200_solve_gauss_newton reports)
    for _ in range(prm.maxInnerLoopIterations):
        pk = accumulate_fn(T)
        dist.all_reduce(pk, group=group)
        h = pk.detach().cpu().numpy()
        if np.sqrt(h[27]) <= prm.maxCost:
            break
        T, conv = capi.gn_step_from_packet(h, prm, T)
        updates += 1
        if conv:
            break
    return True, T, updates


class _SingleProcess:
    """Stand-in for torch.distributed when world == 1 (no process group needed)."""

    @staticmethod
    def all_reduce(t, group=None):
        return t


class ShardedMatcherSolver:
    """GPU side: owns the exchange buffers of one rank."""

    def __init__(self, ctx: capi.Context, gmap: capi.Map, rank: int, world: int, n_total: int, k_max: int = 1):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, (dist if world > 1 else _SingleProcess)
        self.ctx, self.map, self.rank, self.world, self.n_total = ctx, gmap, rank, world, n_total
        self.bounds, self.per = shard_bounds(n_total, world)
        self.lo, self.hi = self.bounds[rank], self.bounds[rank + 1]
        self.n_local = self.hi - self.lo
        dev = torch.device("cuda", ctx.device)
        self.k_max = k_max
        # one exchange record per rank: per*K candidate words followed by 6 bbox floats (+2 pad)
        self.rec_words = self.per * k_max + 4
        self.send = torch.full((self.rec_words,), -1, dtype=torch.int64, device=dev)
        self.recv = torch.empty((world * self.rec_words,), dtype=torch.int64, device=dev)
        self.cand_all = torch.empty((world * self.per * k_max,), dtype=torch.int64, device=dev)
        self.boxes = torch.empty((world * 6,), dtype=torch.float32, device=dev)
        self.packets = torch.zeros((64,), dtype=torch.float64, device=dev)

    def match_pt2pt(self, d_lx: int, d_ly: int, d_lz: int, pose, prm: capi.Pt2PtParams, d_out: int, capacity: int):
        """Device addresses of THIS rank's shard; returns the number of pairs written to d_out."""
        t, K = self.torch, prm.pairingsPerPoint
        assert K <= self.k_max
        if K != self.k_max:
            raise ValueError("exchange buffers were sized for another pairingsPerPoint")
        bbox_ptr = self.send.data_ptr() + self.per * K * 8
        self.send[: self.per * K].fill_(-1)  # padding slots of a short last shard stay invalid
        self.map.shard_search_pt2pt(d_lx, d_ly, d_lz, pose, prm, self.send.data_ptr(), bbox_ptr, n_local=self.n_local, local_on_device=True)
        if self.world > 1:
            self.dist.all_gather_into_tensor(self.recv, self.send)
            r = self.recv.view(self.world, self.rec_words)
        else:
            r = self.send.view(1, self.rec_words)
        # padded layout -> the library's dense global slot numbering uses offsets[r]*K, which equals
        # r*per*K for every rank but possibly a shorter tail: keep the padded numbering end to end.
        self.cand_all.view(self.world, self.per * K).copy_(r[:, : self.per * K])
        self.boxes.view(self.world, 6).copy_(r[:, self.per * K : self.per * K + 3].contiguous().view(t.float32).view(self.world, 6))
        n_total_padded = self.world * self.per
        return self.map.shard_resolve_pt2pt(self.n_local, self.rank * self.per, n_total_padded, self.cand_all.data_ptr(), self.boxes.data_ptr(), self.world, prm, out=d_out, out_on_device=True, capacity=capacity)

    def solve_horn(self, d_pairs: int, n_pairs: int, prm: capi.HornParams):
        p = self.packets

        def sums():
            self.ctx.horn_sums(d_pairs, n=n_pairs, on_device=True, packet=p[:32].data_ptr(), packet_on_device=True)
            return p[:32]

        def moments(s, n_total):
            self.ctx.horn_moments(d_pairs, s.data_ptr(), n_total, n=n_pairs, prm=prm, on_device=True, sums_on_device=True, packet=p[32:].data_ptr(), packet_on_device=True)
            return p[32:]

        return allreduce_horn_solve(sums, moments, self.dist)

    def solve_gauss_newton(self, d_p2p: int, n2p: int, d_p2l: int, n2l: int, prm: capi.GNParams, pose0):
        p = self.packets

        def acc(T):
            self.ctx.gn_accumulate(d_p2p, d_p2l, prm, T, n2p=n2p, n2l=n2l, on_device=True, packet=p[:32].data_ptr(), packet_on_device=True)
            return p[:32]

        return allreduce_gn_solve(acc, prm, pose0, self.dist)
