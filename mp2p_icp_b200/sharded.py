"""Query-sharded multi-GPU driver: one process per GPU, torch.distributed (NCCL over
NVLink on GPUs, gloo in the CPU tests) as plumbing around the C-ABI building blocks.

The local cloud is split into contiguous shards (rank r owns [offsets[r], offsets[r+1])), the map
and its index are replicated. Per ICP iteration:

  matcher   shard_search (phase A)  ->  ONE in-place all_gather of the exchange records
            [candidate words | shard bbox]  ->  shard_resolve (phase B: global first-claim replay +
            compaction of the own shard, HORN1 sums in the same pass)
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
    """GPU side: owns the exchange buffers of one rank.

    Exchange layout (include/mp2p_b200.h, "EXCHANGE RECORDS"): `records` holds world records of
    rec_words 64-bit words; this rank's phase A writes its record IN PLACE at records[rank], the
    all_gather is NCCL's in-place form (send = recv + rank*count), phase B reads the gathered
    records where they are — no packing or unpacking kernels on either side.
    """

    def __init__(self, ctx: capi.Context, gmap: capi.Map, rank: int, world: int, n_total: int, k_max: int = 1, transport: str = None):
        """transport: "peer" = the library's own NVLink mailbox exchange (csrc/peer.cu: kernels that
        write into the peers' HBM, no host call per collective), "nccl" = torch.distributed
        collectives, None = $MP2P_B200_TRANSPORT or "peer" (falls back to "nccl", with a warning, if
        the mailboxes cannot be mapped)."""
        import os
        import warnings

        import torch
        import torch.distributed as dist

        if ctx.stream is None or ctx.stream != torch.cuda.current_stream(ctx.device).cuda_stream:
            # NCCL collectives and the packet read-back are ordered on torch's current stream
            raise ValueError("create the Context on torch's current stream: Context(dev, stream=torch.cuda.current_stream().cuda_stream)")
        self.torch, self.dist = torch, (dist if world > 1 else _SingleProcess)
        self.ctx, self.map, self.rank, self.world, self.n_total = ctx, gmap, rank, world, n_total
        self.bounds, self.per = shard_bounds(n_total, world)
        self.lo, self.hi = self.bounds[rank], self.bounds[rank + 1]
        self.n_local = self.hi - self.lo
        dev = torch.device("cuda", ctx.device)
        self.k = k_max
        self.rec_words = capi.shard_record_words(self.per, k_max)
        self.records = torch.empty((world * self.rec_words,), dtype=torch.int64, device=dev)
        self.mine = self.records[rank * self.rec_words : (rank + 1) * self.rec_words]
        self.packets = torch.zeros((64,), dtype=torch.float64, device=dev)
        self.h_packets = torch.zeros((64,), dtype=torch.float64).pin_memory()
        self.gn_state = torch.zeros((capi.GN_STATE_DOUBLES + 32,), dtype=torch.float64, device=dev)  # state | last packet
        self.h_gn_state = torch.zeros((capi.GN_STATE_DOUBLES + 32,), dtype=torch.float64).pin_memory()
        self.peer = None
        transport = transport or os.environ.get("MP2P_B200_TRANSPORT", "peer")
        if world > 1 and transport == "peer":

            def exchange(handle: bytes):
                got = [None] * world
                dist.all_gather_object(got, handle)
                return got

            err = None
            try:
                self.peer = capi.Peer(ctx, rank, world, self.rec_words, exchange)
                if os.environ.get("MP2P_B200_OWNER_CLAIMS", "1" if world > 2 else "0") != "0":
                    # first claims partitioned by owner (global point g -> rank g % world): proposals and their
                    # acceptance go straight to the owner's HBM over NVLink, nothing is gathered or replayed.
                    # Default from 3 ranks up: at 2 the replay of ONE peer's record is cheaper than remote
                    # atomics (C5: 0.335 against 0.348 ms), at 8 it is the other way round (0.287 against 0.275)
                    self.peer.enable_owner_claims(gmap.info["n_points"], exchange)
            except capi.Mp2pError as e:
                err = str(e)
            # all ranks or none: a rank that failed to map a mailbox drags everybody to NCCL
            flags = [None] * world
            dist.all_gather_object(flags, err)
            if any(f is not None for f in flags):
                if self.peer is not None:
                    self.peer.close()
                self.peer = None
                warnings.warn(f"NVLink mailbox exchange unavailable ({[f for f in flags if f][0]}); using NCCL collectives")
        self.transport = "peer" if self.peer is not None else ("nccl" if world > 1 else "single")
        self._records_ptr = self.records.data_ptr()
        self._iters = {}

    def _native(self, local, mprm, sprm, d_pairs, capacity):
        """Pre-bound native iteration (peer transport only), cached per argument set."""
        # keyed on object identity (cheap; the entry keeps the objects alive): parameter objects are
        # bound when first seen — mutate a copy, not the instance, to change them
        key = (tuple(id(x) if not isinstance(x, int) else x for x in local), id(mprm), id(sprm), int(d_pairs), int(capacity))
        ent = self._iters.get(key)
        if ent is None:
            lx, ly, lz = local
            ent = (self.peer.make_iterator(self.map, lx, ly, lz, self.n_local, mprm, sprm, self.per, d_pairs, capacity), local, mprm, sprm)
            self._iters[key] = ent
        return ent[0]

    def _allreduce(self, t):
        """SUM over ranks of a 32-double device packet, in place, enqueued on the context stream."""
        if self.peer is not None:
            self.peer.allreduce_packet(t.data_ptr())
        else:
            self.dist.all_reduce(t)

    # ---- matcher -------------------------------------------------------------------------------
    def _search_gather(self, local, pose, prm):
        if prm.pairingsPerPoint != self.k:
            raise ValueError("exchange buffers were sized for another pairingsPerPoint")
        lx, ly, lz = local
        if self.peer is not None:
            # the search writes the record into this rank's mailbox slot; a push kernel stores it into
            # every peer's mailbox over NVLink and a wait kernel acquires the peers' flags
            self.map.shard_search_pt2pt(lx, ly, lz, pose, prm, self.per, self.peer.record_slot(), n_local=self.n_local, local_on_device=True)
            self._records_ptr = self.peer.allgather_records()
            return
        self.map.shard_search_pt2pt(lx, ly, lz, pose, prm, self.per, self.mine.data_ptr(), n_local=self.n_local, local_on_device=True)
        if self.world > 1:
            self.dist.all_gather_into_tensor(self.records, self.mine)

    def match_pt2pt(self, d_lx, d_ly, d_lz, pose, prm: capi.Pt2PtParams, d_out: int, capacity: int):
        """(d_lx, d_ly, d_lz) = device addresses of THIS rank's shard, or (Cloud, None, None).
        Returns the number of pairs written to d_out (synchronises)."""
        self._search_gather((d_lx, d_ly, d_lz), pose, prm)
        return self.map.shard_resolve_pt2pt(self.n_local, self.rank, self.world, self.per, self._records_ptr, prm, out=d_out, out_on_device=True, capacity=capacity)

    # ---- whole iterations, ONE host synchronisation ----------------------------------------------
    def iterate_pt2pt_horn(self, local, pose, mprm: capi.Pt2PtParams, sprm: capi.HornParams, d_pairs: int, capacity: int):
        """Matcher_Points_DistanceThreshold + Solver_Horn over the sharded cloud. Enqueues
        search -> all_gather -> resolve(+HORN1 sums) -> all_reduce -> HORN2 moments -> all_reduce
        and reads the two packets back once. Returns (solved, pose 3x4, pairs in the whole cloud)."""
        if sprm.use_scale_outlier_detector:
            raise ValueError("use_scale_outlier_detector needs the two-call path (match_pt2pt + solve_horn)")
        if self.peer is not None:
            ok, T, n_all, _ = self._native(local, mprm, sprm, d_pairs, capacity)(pose)
            return ok, T, n_all
        p = self.packets
        self._search_gather(local, pose, mprm)
        self.map.shard_resolve_pt2pt(self.n_local, self.rank, self.world, self.per, self._records_ptr, mprm, out=d_pairs, out_on_device=True, capacity=capacity, sync=False, horn_sums=p.data_ptr())
        self._allreduce(p[:32])
        self.ctx.horn_moments(d_pairs, p.data_ptr(), 0, n=capi.COUNT_ON_DEVICE, prm=sprm, on_device=True, sums_on_device=True, packet=p[32:].data_ptr(), packet_on_device=True)
        self._allreduce(p[32:])
        self.h_packets.copy_(p, non_blocking=True)
        self.torch.cuda.current_stream().synchronize()
        h = self.h_packets.numpy()
        n_pairs = int(h[7])
        if n_pairs < 3:  # optimal_tf_horn.cpp:96
            return False, np.eye(3, 4), n_pairs
        ok, T = capi.horn_finish(h[:32], h[32:])
        return ok, T, n_pairs

    def iterate_pt2pt_gn(self, local, pose, mprm: capi.Pt2PtParams, sprm: capi.GNParams, d_pairs: int, capacity: int):
        """Matcher_Points_DistanceThreshold + Solver_GaussNewton over the sharded cloud:
        one synchronisation per inner Gauss-Newton iteration (the reduced 6x6 system comes to the host)."""
        if self.peer is not None:
            ok, T, _, it = self._native(local, mprm, sprm, d_pairs, capacity)(pose)
            return ok, T, it
        self._search_gather(local, pose, mprm)
        self.map.shard_resolve_pt2pt(self.n_local, self.rank, self.world, self.per, self._records_ptr, mprm, out=d_pairs, out_on_device=True, capacity=capacity, sync=False)
        return self.gn_device_loop(d_pairs, capi.COUNT_ON_DEVICE, None, 0, sprm, pose)

    def iterate_pt2pl_gn(self, local, pose, mprm: capi.Pt2PlParams, sprm: capi.GNParams, d_pairs: int, capacity: int):
        """Matcher_Point2Plane + Solver_GaussNewton (C3). pt2pl never dedups global points
        (Matcher_Point2Plane.cpp:87-90), so the shards match independently; the solver all-reduces one
        packet per inner iteration with the pose on the device. ONE host synchronisation."""
        if self.peer is not None:
            ok, T, _, it = self._native(local, mprm, sprm, d_pairs, capacity)(pose)
            return ok, T, it
        lx, ly, lz = local
        self.map.match_pt2pl(lx, ly, lz, pose, mprm, n_local=self.n_local, local_on_device=True, out=d_pairs, out_on_device=True, capacity=capacity, sync=False)
        return self.gn_device_loop(None, 0, d_pairs, capi.COUNT_ON_DEVICE, sprm, pose)

    def gn_device_loop(self, d_p2p, n2p, d_p2l, n2l, prm: capi.GNParams, pose0):
        """optimal_tf_gauss_newton.cpp:70-366 with the reduction spread over ranks and the pose kept
        on the device: maxInnerLoopIterations x {accumulate, all_reduce, step} enqueued back to back
        (after convergence the remaining rounds are no-ops on every rank alike), one read-back."""
        st, p, cp = self.gn_state, self.gn_state[capi.GN_STATE_DOUBLES :], prm.c()
        self.ctx.gn_device_begin(pose0, st.data_ptr())
        for _ in range(prm.maxInnerLoopIterations):
            self.ctx.gn_device_accumulate(d_p2p, n2p, d_p2l, n2l, cp, st.data_ptr(), p.data_ptr())
            self._allreduce(p)
            self.ctx.gn_device_step(p.data_ptr(), cp, st.data_ptr())
        self.h_gn_state.copy_(st, non_blocking=True)
        self.torch.cuda.current_stream().synchronize()
        h = self.h_gn_state.numpy()
        flags = h[12:13].view(np.uint32)
        return True, h[:12].reshape(3, 4).copy(), int(flags[1])

    # ---- solvers over pairings already on the device -------------------------------------------
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
