// Peer exchange over NVLink for the query-sharded path (product code), one process per GPU.
//
// The two collectives of a sharded ICP iteration are tiny and latency-bound:
//   * all-gather of the shards' exchange records (candidate words + bbox, ~8 B per query), and
//   * all-reduce (SUM) of one 32-double accumulator packet per solver pass.
// A library collective costs a host call plus ~20 us on the device for each; here every rank owns a
// MAILBOX in its own HBM that its peers write straight into through NVLink (CUDA IPC mapping of
// cudaMalloc memory), and small kernels on the context stream do the rest:
//   k_peer_push_record    copies this rank's record into slot [parity][rank] of EVERY peer's mailbox
//                         with coalesced stores; the last CTA (ticket) releases one flag word
//                         per peer (st.release.sys)
//   k_peer_wait_records   one warp: lane r acquires flag r (ld.acquire.sys) — everything enqueued
//                         behind it on the stream may read the gathered records
//   k_peer_allreduce      one warp: lane t owns double t of the packet: store to slot [parity][rank]
//                         of every peer, release the flags, acquire the world flags of the own
//                         mailbox, sum the world slots IN RANK ORDER (bit-identical on every rank)
// Flags carry a call counter (epoch) that only grows, so nothing is ever cleared; slots are double
// buffered by the parity of the epoch: a rank can run at most one exchange ahead of a peer (it needs
// that peer's flag of exchange e to finish e), so exchange e+2 can never overwrite a slot a peer
// still reads for e. A flag that does not arrive within kPeerTimeoutNs traps (a lost rank must not
// hang the others).
#include <algorithm>
#include <cstring>
#include <new>

#include "common.cuh"
#include "peer.cuh"

namespace mp2p
{
namespace
{
__global__ void __launch_bounds__(256)
    k_peer_push_record(PeerView pv, uint32_t parity, uint32_t epoch, unsigned int* __restrict__ ticket, uint64_t first_word)
{
    // the record sits in the own mailbox already (the search kernel wrote it there); words [first_word, rec_words)
    // travel (everything, or — with owner-partitioned claims — only the shard's bounding box at its end)
    const size_t              off   = rec_offset(pv.rec_words, pv.world, parity, pv.rank);
    const unsigned long long* src   = reinterpret_cast<const unsigned long long*>(pv.box[pv.rank] + off);
    const size_t              n_vec = pv.rec_words;
    for (uint32_t p = 1; p < pv.world; p++)
    {
        const uint32_t      dst_rank = (pv.rank + p) % pv.world;  // start with the right neighbour: spreads the traffic
        unsigned long long* dst      = reinterpret_cast<unsigned long long*>(pv.box[dst_rank] + off);
        for (size_t i = first_word + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x)
            dst[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    __shared__ unsigned last;
    if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (!last) return;
    __threadfence_system();
    if (threadIdx.x < pv.world) st_release_sys(rec_flag(pv.box[threadIdx.x], parity, pv.rank), epoch);
    if (threadIdx.x == 0) *ticket = 0u;
}

__global__ void k_peer_wait_records(PeerView pv, uint32_t parity, uint32_t epoch)
{
    if (threadIdx.x < pv.world) wait_flag(rec_flag(pv.box[pv.rank], parity, threadIdx.x), epoch);
}

__global__ void k_peer_allreduce(PeerView pv, uint32_t epoch, double* __restrict__ packet)
{
    peer_allreduce_warp(pv, epoch, packet, threadIdx.x);  // one warp
}
}  // namespace
}  // namespace mp2p

using namespace mp2p;

extern "C"
{
    int mp2p_b200_peer_create(mp2p_b200_ctx* ctx, uint32_t rank, uint32_t world, uint64_t record_words,
                              uint8_t handle_out[MP2P_B200_PEER_HANDLE_BYTES], mp2p_b200_peer** out)
    {
        static_assert(sizeof(cudaIpcMemHandle_t) == MP2P_B200_PEER_HANDLE_BYTES, "IPC handle size");
        if (!ctx || !out || !handle_out || world < 1 || world > (uint32_t)kMaxPeers || rank >= world)
        {
            set_error("peer_create: need 1 <= world <= %d and rank < world", kMaxPeers);
            return MP2P_B200_ERR_ARG;
        }
        *out = nullptr;
        MP2P_CUDA_TRY(cudaSetDevice(ctx->device));
        auto* p = new (std::nothrow) mp2p_b200_peer();
        if (!p) return MP2P_B200_ERR_NOMEM;
        p->ctx            = ctx;
        p->view.rank      = rank, p->view.world = world, p->view.rec_words = record_words;
        p->bytes          = rec_offset(record_words, world, 2, 0) + 256;
        cudaError_t e     = cudaMalloc(&p->own, p->bytes);
        if (e == cudaSuccess) e = cudaMemset(p->own, 0, p->bytes);
        if (e == cudaSuccess) e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle_out), p->own);
        if (e != cudaSuccess)
        {
            set_error("peer_create: %s", cudaGetErrorString(e));
            cudaGetLastError();
            if (p->own) cudaFree(p->own);
            delete p;
            return MP2P_B200_ERR_CUDA;
        }
        p->ticket            = reinterpret_cast<unsigned int*>(static_cast<char*>(p->own) + p->bytes - 64);
        p->view.box[rank]    = static_cast<char*>(p->own);
        *out                 = p;
        return 0;
    }

    int mp2p_b200_peer_connect(mp2p_b200_peer* p, const uint8_t* handles)
    {
        if (!p || !handles) return MP2P_B200_ERR_ARG;
        MP2P_CUDA_TRY(cudaSetDevice(p->ctx->device));
        for (uint32_t r = 0; r < p->view.world; r++)
        {
            if (r == p->view.rank) continue;
            cudaIpcMemHandle_t h;
            std::memcpy(&h, handles + (size_t)r * MP2P_B200_PEER_HANDLE_BYTES, sizeof(h));
            void*             ptr = nullptr;
            const cudaError_t e   = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess)
            {
                set_error("peer_connect: cannot map the mailbox of rank %u: %s", r, cudaGetErrorString(e));
                cudaGetLastError();
                return MP2P_B200_ERR_CUDA;
            }
            p->opened[r]   = ptr;
            p->view.box[r] = static_cast<char*>(ptr);
        }
        p->connected = true;
        return 0;
    }

    void mp2p_b200_peer_destroy(mp2p_b200_peer* p)
    {
        if (!p) return;
        cudaSetDevice(p->ctx->device);
        cudaStreamSynchronize(p->ctx->stream);
        for (void* o : p->opened)
            if (o) cudaIpcCloseMemHandle(o);
        for (void* o : p->claims_opened)
            if (o) cudaIpcCloseMemHandle(o);
        if (p->d_claim_parts) cudaFree(p->d_claim_parts);
        if (p->claims_own) cudaFree(p->claims_own);
        if (p->own) cudaFree(p->own);
        delete p;
    }

    int mp2p_b200_peer_record_slot(mp2p_b200_peer* p, uint64_t** slot_device)
    {
        if (!p || !slot_device) return MP2P_B200_ERR_ARG;
        const uint32_t parity = (p->rec_epoch + 1) & 1u;
        *slot_device = reinterpret_cast<uint64_t*>(p->view.box[p->view.rank] + rec_offset(p->view.rec_words, p->view.world, parity, p->view.rank));
        return 0;
    }

    // first_word: only words [first_word, rec_words) of the records are exchanged (0 = whole records)
    static int peer_allgather_tail(mp2p_b200_peer* p, const uint64_t** records_device, uint64_t first_word)
    {
        if (!p || !records_device || !p->connected)
        {
            set_error("peer_allgather_records: not connected");
            return MP2P_B200_ERR_ARG;
        }
        MP2P_CUDA_TRY(cudaSetDevice(p->ctx->device));
        const uint32_t epoch = ++p->rec_epoch, parity = epoch & 1u;
        cudaStream_t   st    = p->ctx->stream;
        const size_t   n_vec = p->view.rec_words - first_word;
        const int      blocks = (int)std::max<size_t>(1, std::min<size_t>((n_vec + 255) / 256, 148 * 2));
        k_peer_push_record<<<blocks, 256, 0, st>>>(p->view, parity, epoch, p->ticket, first_word);
        k_peer_wait_records<<<1, 32, 0, st>>>(p->view, parity, epoch);
        count_launch(p->ctx, 2);
        MP2P_CUDA_TRY(cudaGetLastError());
        *records_device = reinterpret_cast<const uint64_t*>(p->view.box[p->view.rank] + rec_offset(p->view.rec_words, p->view.world, parity, 0));
        return 0;
    }

    int mp2p_b200_peer_claims_create(mp2p_b200_peer* p, uint64_t n_map_points, uint8_t handle_out[MP2P_B200_PEER_HANDLE_BYTES])
    {
        if (!p || !handle_out || !n_map_points)
        {
            set_error("peer_claims_create: NULL argument or empty map");
            return MP2P_B200_ERR_ARG;
        }
        MP2P_CUDA_TRY(cudaSetDevice(p->ctx->device));
        if (p->claims_own)
        {
            set_error("peer_claims_create: this peer object already owns a claim part");
            return MP2P_B200_ERR_ARG;
        }
        const uint64_t words = (n_map_points + p->view.world - 1) / p->view.world + 1;
        cudaError_t    e     = cudaMalloc(&p->claims_own, words * 8);
        if (e == cudaSuccess) e = cudaMemset(p->claims_own, 0xff, words * 8);  // "unclaimed": loses against any proposal
        if (e == cudaSuccess) e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle_out), p->claims_own);
        if (e != cudaSuccess)
        {
            set_error("peer_claims_create: %s", cudaGetErrorString(e));
            cudaGetLastError();
            if (p->claims_own) cudaFree(p->claims_own), p->claims_own = nullptr;
            return MP2P_B200_ERR_CUDA;
        }
        p->claims_n = n_map_points;
        return 0;
    }

    int mp2p_b200_peer_claims_connect(mp2p_b200_peer* p, const uint8_t* handles)
    {
        if (!p || !handles || !p->claims_own)
        {
            set_error("peer_claims_connect: call peer_claims_create first");
            return MP2P_B200_ERR_ARG;
        }
        MP2P_CUDA_TRY(cudaSetDevice(p->ctx->device));
        unsigned long long* parts[kMaxPeers] = {};
        for (uint32_t r = 0; r < p->view.world; r++)
        {
            if (r == p->view.rank)
            {
                parts[r] = static_cast<unsigned long long*>(p->claims_own);
                continue;
            }
            cudaIpcMemHandle_t h;
            std::memcpy(&h, handles + (size_t)r * MP2P_B200_PEER_HANDLE_BYTES, sizeof(h));
            void*             ptr = nullptr;
            const cudaError_t e   = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess)
            {
                set_error("peer_claims_connect: cannot map the claim part of rank %u: %s", r, cudaGetErrorString(e));
                cudaGetLastError();
                return MP2P_B200_ERR_CUDA;
            }
            p->claims_opened[r] = ptr;
            parts[r]            = static_cast<unsigned long long*>(ptr);
        }
        MP2P_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&p->d_claim_parts), sizeof(parts)));
        MP2P_CUDA_TRY(cudaMemcpy(p->d_claim_parts, parts, sizeof(parts), cudaMemcpyHostToDevice));
        p->claims_connected = true;
        return 0;
    }

    int mp2p_b200_peer_allgather_records(mp2p_b200_peer* p, const uint64_t** records_device)
    {
        return peer_allgather_tail(p, records_device, 0);
    }
    int mp2p_b200_peer_allreduce_packet(mp2p_b200_peer* p, double* packet_device)
    {
        if (!p || !packet_device || !p->connected)
        {
            set_error("peer_allreduce_packet: not connected");
            return MP2P_B200_ERR_ARG;
        }
        MP2P_CUDA_TRY(cudaSetDevice(p->ctx->device));
        const uint32_t epoch = ++p->pkt_epoch;
        k_peer_allreduce<<<1, MP2P_B200_PACKET_DOUBLES, 0, p->ctx->stream>>>(p->view, epoch, packet_device);
        count_launch(p->ctx);
        MP2P_CUDA_TRY(cudaGetLastError());
        return 0;
    }
    // ---- whole query-sharded iterations, enqueued natively, ONE host synchronisation ---------------
    // (the same sequences mp2p_icp_b200/sharded.py issues call by call; here the host cost is a few
    // kernel launches instead of a dozen interpreter round trips)
    static int peer_gn_loop(mp2p_b200_peer* p, const mp2p_b200_pair_pt2pt* d2p, uint64_t n2p, const mp2p_b200_pair_pt2pl* d2l,
                            uint64_t n2l, const mp2p_b200_gn_params* sprm, const double pose[12], double pose_out[12],
                            int32_t* solved, uint32_t* iterations_done)
    {
        mp2p_b200_ctx* ctx    = p->ctx;
        double*        state  = ctx->d_pose.as<double>();  // MP2P_B200_GN_STATE_DOUBLES doubles
        double*        packet = ctx->d_packet.as<double>() + 5 * MP2P_B200_PACKET_DOUBLES;
        MP2P_TRY(mp2p_b200_gn_device_begin(ctx, pose, state));
        // the whole inner loop, all-reduces included, as one cooperative launch (solve.cu, k_gn_loop<true>);
        // rc 1 = not possible here: one launch per accumulate / all-reduce / step instead (same protocol, same epochs)
        uint64_t                  n2p_h = n2p, n2l_h = n2l;
        const unsigned long long *dn2p = nullptr, *dn2l = nullptr;
        if (n2p == MP2P_B200_COUNT_ON_DEVICE) n2p_h = ctx->last_count ? ctx->last_capacity : 0, dn2p = ctx->last_count;
        if (n2l == MP2P_B200_COUNT_ON_DEVICE) n2l_h = ctx->last_count ? ctx->last_capacity : 0, dn2l = ctx->last_count;
        int rc_coop = 1;
        if ((n2p != MP2P_B200_COUNT_ON_DEVICE || ctx->last_count) && (n2l != MP2P_B200_COUNT_ON_DEVICE || ctx->last_count))
            rc_coop = run_gn_coop_loop(ctx, d2p, n2p_h, d2l, n2l_h, sprm, state, reinterpret_cast<uint32_t*>(state + 12), packet, dn2p,
                                       dn2l, p);
        if (rc_coop < 0) return rc_coop;
        for (uint32_t it = 0; rc_coop == 1 && it < sprm->maxInnerLoopIterations; it++)
        {
            MP2P_TRY(mp2p_b200_gn_device_accumulate(ctx, d2p, n2p, d2l, n2l, sprm, state, packet));
            MP2P_TRY(mp2p_b200_peer_allreduce_packet(p, packet));
            MP2P_TRY(mp2p_b200_gn_device_step(ctx, packet, sprm, state));
        }
        double* h = reinterpret_cast<double*>(static_cast<char*>(ctx->h_pinned) + 2048);
        MP2P_CUDA_TRY(cudaMemcpyAsync(h, state, MP2P_B200_GN_STATE_DOUBLES * 8, cudaMemcpyDeviceToHost, ctx->stream));
        MP2P_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        MP2P_CUDA_TRY(cudaGetLastError());
        std::memcpy(pose_out, h, 96);
        if (iterations_done) *iterations_done = reinterpret_cast<const uint32_t*>(h + 12)[1];
        *solved = 1;
        return 0;
    }

    int mp2p_b200_peer_iterate_pt2pt(mp2p_b200_peer* p, mp2p_b200_map* map, const float* lx, const float* ly, const float* lz,
                                     uint64_t n_local, int local_on_device, const double pose[12],
                                     const mp2p_b200_pt2pt_params* mprm, const mp2p_b200_horn_params* horn,
                                     const mp2p_b200_gn_params* gn, uint64_t per_shard, mp2p_b200_pair_pt2pt* pairs_device,
                                     uint64_t capacity, double pose_out[12], int32_t* solved, uint64_t* n_pairs_total,
                                     uint32_t* iterations_done)
    {
        if (!p || !p->connected || !map || !pose || !mprm || (!horn == !gn) || !pairs_device || !pose_out || !solved)
        {
            set_error("peer_iterate_pt2pt: NULL argument (exactly one of the Horn / Gauss-Newton parameter blocks is needed)");
            return MP2P_B200_ERR_ARG;
        }
        if (mp2p_b200_shard_record_words(per_shard, mprm->pairingsPerPoint) != p->view.rec_words)
        {
            set_error("peer_iterate_pt2pt: the mailboxes were sized for another per_shard / pairingsPerPoint");
            return MP2P_B200_ERR_ARG;
        }
        mp2p_b200_ctx* ctx = p->ctx;
        *solved            = 0;
        if (n_pairs_total) *n_pairs_total = 0;
        struct Prof  // timing slot 5 = the whole call (measurement hook, no-op unless profiling is on)
        {
            mp2p_b200_ctx* c;
            explicit Prof(mp2p_b200_ctx* x) : c(x) { prof_reset(c); }
            ~Prof() { prof_collect(c); }
        } prof(ctx);
        double* p0 = ctx->d_packet.as<double>() + 5 * MP2P_B200_PACKET_DOUBLES;
        double* p1 = p0 + MP2P_B200_PACKET_DOUBLES;
        if (horn && mprm->pairingsPerPoint == 1 && horn->robust_kernel == 0 && horn->w_pt2pt > 0.0 &&
            !horn->use_scale_outlier_detector && n_local && map->view.n_points && n_local <= per_shard &&
            capacity >= n_local && lx && (local_on_device == 2 || (ly && lz)))
        {
            // ONE cooperative launch carries search, record exchange, replay, compaction, both Horn passes
            // and both all-reduces (match.cu, k_iterate_nn1_horn<true>); rc 1 = the grid does not fit
            MP2P_CUDA_TRY(cudaSetDevice(ctx->device));
            DeviceMatch dm;
            dm.want_horn_sums = p0, dm.fuse_moments_w = horn->w_pt2pt;
            dm.peer = p, dm.per_shard = per_shard;
            uint64_t  dummy = 0;
            const int rc    = run_match_pt2pt(ctx, map, lx, ly, lz, n_local, local_on_device, pose, mprm, nullptr, nullptr,
                                              pairs_device, capacity, 1, &dummy, &dm);
            if (rc < 0) return rc;
            if (rc == 0)
            {
                ctx->last_count = dm.d_count, ctx->last_capacity = dm.capacity;
                double* h = nullptr;
                MP2P_TRY(read_iteration_packets(ctx, true, p0, &h));
                const uint64_t n_all = (uint64_t)h[7];
                if (n_pairs_total) *n_pairs_total = n_all;
                if (n_all < 3) return 0;  // optimal_tf_horn.cpp:96
                return mp2p_b200_horn_finish(h, h + MP2P_B200_PACKET_DOUBLES, pose_out, solved);
            }
        }
        uint64_t*       slot = nullptr;
        const uint64_t* recs = nullptr;
        MP2P_TRY(mp2p_b200_peer_record_slot(p, &slot));
        if (p->claims_connected && p->claims_n >= map->view.n_points && !mprm->allowMatchAlreadyMatchedGlobalPoints &&
            n_local <= per_shard && per_shard * mprm->pairingsPerPoint * p->view.world < 0xFFFFFFFFull)
        {
            // OWNER-PARTITIONED claims: the search proposes straight into the owners' HBM over NVLink (system-scope
            // atomicMin), only the 32-byte bounding boxes are gathered (the flags of that exchange are also the
            // barrier "every rank has proposed"), the compaction reads each claim word back from its owner. No
            // record all-gather, no replay: the per-GPU work is the shard's own proposals, whatever the world size.
            // (The next call's proposals cannot overtake this call's reads: every rank passes at least one packet
            // all-reduce behind its compaction before it returns.)
            MP2P_CUDA_TRY(cudaSetDevice(ctx->device));
            if (++p->claim_epoch == 0xFFFFFFFFu)
            {
                set_error("peer_iterate_pt2pt: claim epochs exhausted (2^32 sharded calls): recreate the peer object");
                return MP2P_B200_ERR_ARG;
            }
            OwnerClaims oc;
            oc.parts = p->d_claim_parts, oc.world = p->view.world;
            oc.tag   = (unsigned long long)(0xFFFFFFFFu - p->claim_epoch) << 32;
            MP2P_TRY(run_shard_search_pt2pt(ctx, map, lx, ly, lz, n_local, local_on_device, pose, mprm, nullptr, per_shard,
                                            reinterpret_cast<unsigned long long*>(slot), &oc, p->view.rank));
            MP2P_TRY(peer_allgather_tail(p, &recs, per_shard * mprm->pairingsPerPoint));
            MP2P_TRY(run_shard_resolve_pt2pt(ctx, map, n_local, p->view.rank, p->view.world, per_shard,
                                             reinterpret_cast<const unsigned long long*>(recs), mprm, nullptr, pairs_device, capacity, 1,
                                             nullptr, horn ? p0 : nullptr, &oc));
        }
        else
        {
            MP2P_TRY(mp2p_b200_match_pt2pt_shard_search(ctx, map, lx, ly, lz, n_local, local_on_device, pose, mprm, nullptr, per_shard, slot));
            MP2P_TRY(mp2p_b200_peer_allgather_records(p, &recs));
            MP2P_TRY(mp2p_b200_match_pt2pt_shard_resolve(ctx, map, n_local, p->view.rank, p->view.world, per_shard, recs, mprm, nullptr,
                                                         pairs_device, capacity, 1, nullptr, horn ? p0 : nullptr));
        }
        if (gn)
            return peer_gn_loop(p, pairs_device, MP2P_B200_COUNT_ON_DEVICE, nullptr, 0, gn, pose, pose_out, solved, iterations_done);
        // Solver_Horn: HORN1 sums came out of the compaction; reduce, moments, reduce, one read-back
        MP2P_TRY(mp2p_b200_peer_allreduce_packet(p, p0));
        MP2P_TRY(mp2p_b200_horn_moments(ctx, pairs_device, MP2P_B200_COUNT_ON_DEVICE, 1, horn, p0, 1, 0, p1, 1));
        MP2P_TRY(mp2p_b200_peer_allreduce_packet(p, p1));
        double* h = reinterpret_cast<double*>(static_cast<char*>(ctx->h_pinned) + 2048);
        MP2P_CUDA_TRY(cudaMemcpyAsync(h, p0, 2 * MP2P_B200_PACKET_DOUBLES * 8, cudaMemcpyDeviceToHost, ctx->stream));
        MP2P_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        MP2P_CUDA_TRY(cudaGetLastError());
        const uint64_t n_all = (uint64_t)h[7];
        if (n_pairs_total) *n_pairs_total = n_all;
        if (n_all < 3) return 0;  // optimal_tf_horn.cpp:96
        return mp2p_b200_horn_finish(h, h + MP2P_B200_PACKET_DOUBLES, pose_out, solved);
    }

    int mp2p_b200_peer_iterate_pt2pl_gn(mp2p_b200_peer* p, mp2p_b200_map* map, const float* lx, const float* ly,
                                        const float* lz, uint64_t n_local, int local_on_device, const double pose[12],
                                        const mp2p_b200_pt2pl_params* mprm, const mp2p_b200_gn_params* sprm,
                                        mp2p_b200_pair_pt2pl* pairs_device, uint64_t capacity, double pose_out[12],
                                        int32_t* solved, uint32_t* iterations_done)
    {
        if (!p || !p->connected || !map || !pose || !mprm || !sprm || !pairs_device || !pose_out || !solved)
        {
            set_error("peer_iterate_pt2pl_gn: NULL argument");
            return MP2P_B200_ERR_ARG;
        }
        *solved = 0;
        // pt2pl never dedups global points (Matcher_Point2Plane.cpp:87-90): shards match independently
        MP2P_TRY(mp2p_b200_match_pt2pl(p->ctx, map, lx, ly, lz, n_local, local_on_device, pose, mprm, nullptr, pairs_device,
                                       capacity, 1, nullptr, nullptr));
        return peer_gn_loop(p, nullptr, 0, pairs_device, MP2P_B200_COUNT_ON_DEVICE, sprm, pose, pose_out, solved, iterations_done);
    }
}
