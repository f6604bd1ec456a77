// Matcher kernels (product code): the per-ICP-iteration correspondence search.
//
//   k_match_pt2pt  = transform_local_to_global (Matcher_Points_Base.cpp:183-220) +
//                    nn_single_search / nn_multiple_search (MRPT/nanoflann, external) +
//                    distance-threshold test (Matcher_Points_DistanceThreshold.cpp:214-259) +
//                    first-claim proposal on the global point.
//   k_compact_pt2pt = lambdaAddPair (…DistanceThreshold.cpp:94-121): bounding-box gate (:73-75),
//                    first-claim acceptance, stream compaction into 36-byte TMatchingPair records
//                    in ascending (localIdx, rank) order.
//   k_match_pt2pl / k_compact_pt2pl = Matcher_Point2Plane.cpp:41-114 with the k-NN + PCA plane fit;
//                    in line mode the same pipeline is Matcher_Point2Line.cpp:46-163.
//   k_match_pt2pt_nn1 = the k = 1 search (one lane per query, warp work redistribution);
//   k_iterate_nn1_horn<SHARDED> = a whole pt2pt + Horn iteration in one cooperative launch, optionally
//                    with the exchanges of a query-sharded iteration through the peers' mailboxes.
//   k_ir_*          = Matcher_Points_InlierRatio.cpp:41-143 (sort by distance, keep a ratio).
//   k_ad_*          = Matcher_Adaptive.cpp:59-314 (histogram of the 1st / 2nd errors out, adaptive
//                    threshold in, per-point plane / pt2pt decision).
//
// Query tiles (256 points x 3 axes) are staged into shared memory with TMA bulk copies
// (cp.async.bulk … mbarrier::complete_tx) so the per-thread search starts from on-chip data.
//
// First-claim semantics on a parallel machine: in the serial reference a proposal (i,k) -> g is
// rejected iff g was taken by a lexicographically smaller (i',k'). Every proposer does
// atomicMin(claim[g], tag | (i*K+k)); the proposal whose word survives is the accepted one.
// `tag` = (0xFFFFFFFF - epoch) << 32 decreases with every call, so words of earlier calls always
// lose and the claim array never needs clearing.
#include <cuda/std/limits>

#include <cstdlib>

#include "grid_search.cuh"
#include "plane_fit.cuh"
#include "peer.cuh"
#include "radix_sort.cuh"
#include "reduce.cuh"

namespace mp2p
{
namespace
{
// ------------------------------------------------------------------------------------------
// TMA bulk-copy helpers (sm_90+/sm_100a): 1-D global -> shared copies completing on an mbarrier.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                            uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// A CTA of kQueryTile threads serves kQueryTile / G queries in the group-cooperative kernels
// (G = 8, 16 or 32 lanes per query) and kNN1Threads queries in the lane-per-query K = 1 kernel.
constexpr uint32_t kNN1Threads = 128;

template <uint32_t NQ>
struct QueryTile
{
    alignas(128) float x[NQ];
    alignas(128) float y[NQ];
    alignas(128) float z[NQ];
    alignas(8) uint64_t bar;
};

// All threads of the CTA call this; afterwards tile.x/y/z hold queries [base, base+NQ).
// The staging arrays are padded to a multiple of kQueryTile, so the copy size is constant.
// Full tiles of 16-byte aligned arrays come in by TMA; the ragged last tile (or a caller's
// unaligned device arrays, used in place without a staging copy) by plain predicated loads.
template <uint32_t NQ>
__device__ __forceinline__ void load_query_tile(QueryTile<NQ>& tile, const float* lx, const float* ly,
                                                const float* lz, size_t base, size_t n_total, bool tma_ok)
{
    const bool use_tma = (NQ % 4 == 0) && tma_ok && base + NQ <= n_total;  // CTA-uniform; bulk copies move multiples of 16 bytes
    if (use_tma && threadIdx.x == 0)
    {
        mbar_init(&tile.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (!use_tma)
        for (uint32_t t = threadIdx.x; t < NQ; t += blockDim.x)
        {
            const size_t idx = base + t;
            const bool   in  = idx < n_total;
            tile.x[t] = in ? lx[idx] : 0.f, tile.y[t] = in ? ly[idx] : 0.f, tile.z[t] = in ? lz[idx] : 0.f;
        }
    __syncthreads();
    if (!use_tma) return;
    if (threadIdx.x == 0)
    {
        constexpr uint32_t bytes = NQ * sizeof(float);
        mbar_expect_tx(&tile.bar, 3 * bytes);
        tma_load_1d(tile.x, lx + base, bytes, &tile.bar);
        tma_load_1d(tile.y, ly + base, bytes, &tile.bar);
        tma_load_1d(tile.z, lz + base, bytes, &tile.bar);
    }
    mbar_wait(&tile.bar, 0);
}

struct PoseArg
{
    double m[12];
};

// CPose3D::composePoint(float…): double arithmetic left to right, one rounding to float.
__device__ __forceinline__ void compose_point_f(const PoseArg& T, float lx, float ly, float lz,
                                                float& gx, float& gy, float& gz)
{
    const double x = lx, y = ly, z = lz;
    gx = __double2float_rn(__dadd_rn(
        __dadd_rn(__dadd_rn(__dmul_rn(T.m[0], x), __dmul_rn(T.m[1], y)), __dmul_rn(T.m[2], z)), T.m[3]));
    gy = __double2float_rn(__dadd_rn(
        __dadd_rn(__dadd_rn(__dmul_rn(T.m[4], x), __dmul_rn(T.m[5], y)), __dmul_rn(T.m[6], z)), T.m[7]));
    gz = __double2float_rn(__dadd_rn(
        __dadd_rn(__dadd_rn(__dmul_rn(T.m[8], x), __dmul_rn(T.m[9], y)), __dmul_rn(T.m[10], z)), T.m[11]));
}

// Bounding box of the transformed local cloud (TransformedLocalPointCloud::localMin/localMax,
// Matcher_Points_Base.h:98-112) as 6 order-preserving 32-bit words (min xyz, max xyz), without
// any CTA barrier: warp redux -> shared atomics -> the last warp of the CTA to arrive pushes the
// CTA's extrema to the global words with (fire-and-forget) atomics. Warps never wait for each
// other, so a warp whose queries are done retires immediately.
__device__ __forceinline__ uint32_t f2ord(float f)
{
    const uint32_t u = __float_as_uint(f);
    return u ^ ((uint32_t)((int32_t)u >> 31) | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o)
{
    return __uint_as_float(o ^ (((o >> 31) - 1u) | 0x80000000u));
}
struct BBoxAcc
{
    uint32_t v[6];
    uint32_t warps_done;
};
// call by every thread BEFORE the CTA's first __syncthreads
__device__ __forceinline__ void bbox_init(BBoxAcc& b)
{
    if (threadIdx.x < 3) b.v[threadIdx.x] = 0xFFFFFFFFu;
    if (threadIdx.x >= 3 && threadIdx.x < 6) b.v[threadIdx.x] = 0u;
    if (threadIdx.x == 6) b.warps_done = 0u;
}
// call by all 32 lanes of every warp of the CTA, exactly once
__device__ __forceinline__ void bbox_accumulate(BBoxAcc& b, float gx, float gy, float gz, bool contributes,
                                                uint32_t* __restrict__ g_words)
{
    const uint32_t lo[3] = {contributes ? f2ord(gx) : 0xFFFFFFFFu, contributes ? f2ord(gy) : 0xFFFFFFFFu,
                            contributes ? f2ord(gz) : 0xFFFFFFFFu};
    const uint32_t hi[3] = {contributes ? f2ord(gx) : 0u, contributes ? f2ord(gy) : 0u,
                            contributes ? f2ord(gz) : 0u};
    uint32_t r[6];
#pragma unroll
    for (int d = 0; d < 3; d++)
    {
        r[d]     = __reduce_min_sync(0xffffffffu, lo[d]);
        r[3 + d] = __reduce_max_sync(0xffffffffu, hi[d]);
    }
    if ((threadIdx.x & 31) == 0)
    {
#pragma unroll
        for (int d = 0; d < 3; d++) atomicMin(&b.v[d], r[d]), atomicMax(&b.v[3 + d], r[3 + d]);
        __threadfence_block();
        if (atomicAdd(&b.warps_done, 1u) == (blockDim.x >> 5) - 1)
        {
            __threadfence_block();
#pragma unroll
            for (int d = 0; d < 3; d++)
            {
                atomicMin(g_words + d, *(volatile uint32_t*)&b.v[d]);
                atomicMax(g_words + 3 + d, *(volatile uint32_t*)&b.v[3 + d]);
            }
        }
    }
}

__device__ __forceinline__ bool bit_set(const uint32_t* bits, uint32_t i)
{
    return bits && ((bits[i >> 5] >> (i & 31)) & 1u);
}

// first-claim word of global point gi: the map's own array, or the owner rank's part over NVLink
__device__ __forceinline__ void claim_propose(unsigned long long* claim, unsigned long long* const* parts, uint32_t world,
                                              uint32_t gi, unsigned long long word)
{
    if (parts)
        atomicMin_system(parts[gi % world] + gi / world, word);
    else
        atomicMin(claim + gi, word);
}
__device__ __forceinline__ unsigned long long claim_read(const unsigned long long* claim, unsigned long long* const* parts,
                                                         uint32_t world, uint32_t gi)
{
    if (parts) return *reinterpret_cast<const volatile unsigned long long*>(parts[gi % world] + gi / world);
    return __ldcg(claim + gi);
}

struct Pt2PtArgs
{
    PoseArg  pose;
    float    maxDistSq, angSq;
    uint32_t n_local, K;
    int      allowLocal, allowGlobal;
    unsigned long long tag;  // (0xFFFFFFFF - epoch) << 32
    int      tma_ok;         // local arrays are 16-byte aligned
    int      rl_start;       // relative level the search starts from (start_level())
    int      n_phases;       // k > 1 search: 2 = centre, then all neighbours; 3 = centre, faces, edges + corners
    int      cand_sorted;    // candidate words go to the query's SORTED position (pt2pl path)
    uint32_t tile_stride;    // CTA b serves query tile (b * tile_stride) % n_tiles: spreads expensive
                             // neighbourhoods (sparse map regions cluster in any spatial order) over the grid
    unsigned long long slot_offset;  // K = 1 sharded single-launch iteration: global proposal slot of local point 0
    // K = 1, local cloud read straight from the caller's pinned host arrays (zero copy): the search
    // kernel leaves a device copy here for the compaction (NULL = the arrays are device memory)
    float *stage_x, *stage_y, *stage_z;
    // k > 1 search over a resident cloud: tile served by CTA b (longest tiles of the previous call first;
    // NULL = the strided order above) and where every tile reports how long it took (NULL = nowhere)
    // query-sharded run with owner-partitioned claims (OwnerClaims, common.cuh): NULL = the map's own claim array
    unsigned long long* const* claim_parts;
    uint32_t                   claim_world;
    // thread-per-query search: where over-budget queries hand themselves over (DeferList, grid_search.cuh), and the
    // group kernel serving that list: query positions come from qlist[0 .. min(*qlist_count, qlist_cap)) instead of
    // the tile; it also clears *defer_reset (the counter the NEXT call's thread pass will use)
    DeferList       defer;
    const uint32_t* qlist;
    const uint32_t* qlist_count;
    uint32_t        qlist_cap;
    uint32_t*       defer_reset;
    const uint32_t* tile_order;
    uint32_t*       tile_cost;
    uint32_t*       tile_trace;  // measurement hook: 8 words per CTA {SM, start ns, end ns, tile, rounds, steps, inserts, probes | levels << 16 of warp 0}
};

// ------------------------------------------------------------------------------------------
// pt2pl: the search kernel lists the queries whose k-NN holds at least `need` points (list order =
// arrival order: results are written under the query's index, so any order is fine) and clears the
// accepted-flag of the others; the plane fit then runs one thread per LISTED query on full warps.
struct FitList
{
    uint32_t* list;   // NULL = not requested
    uint32_t* count;  // zeroed by the host before the launch
    uint8_t*  ok_flags;
    int       need;
};

#ifndef MP2P_MATCH_MIN_BLOCKS
#define MP2P_MATCH_MIN_BLOCKS 4  // CTAs of 256 threads per SM the register allocation must allow
#endif
template <int G, bool V1 = false, int NT = (int)kQueryTile>
__global__ void __launch_bounds__(NT, MP2P_MATCH_MIN_BLOCKS * (int)kQueryTile / NT)
    k_match_pt2pt(GridView g, Pt2PtArgs a, const float* __restrict__ lx, const float* __restrict__ ly,
                  const float* __restrict__ lz, const uint32_t* __restrict__ perm,
                  const uint32_t* __restrict__ lbits,
                  const uint32_t* __restrict__ gbits, unsigned long long* __restrict__ claim,
                  unsigned long long* __restrict__ cand, uint32_t* __restrict__ bbox_words,
                  unsigned long long* __restrict__ stats, FitList fit)
{
    constexpr uint32_t NQ = NT / G;  // queries per CTA
    __shared__ QueryTile<NQ> tile;
    __shared__ BBoxAcc       bacc;
    __shared__ KnnShared<V1 ? 32 : G, NT> ks;  // (V1: the old search keeps no tables; smallest instantiation)
    const long long    t_start = clock64();
    unsigned long long t_trace0 = 0;
    if (a.tile_trace) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_trace0));
    const bool      listed  = a.qlist != nullptr;  // serving the deferred list of a thread-per-query pass
    const uint32_t  tile_id = listed ? blockIdx.x : (a.tile_order ? __ldg(a.tile_order + blockIdx.x) : (uint32_t)(((unsigned long long)blockIdx.x * a.tile_stride) % gridDim.x));
    const size_t    base    = (size_t)tile_id * NQ;
    uint32_t        n_listed = 0;
    if (listed)
    {
        n_listed = min(__ldg(a.qlist_count), a.qlist_cap);
        if (a.defer_reset && blockIdx.x == 0 && threadIdx.x == 0) *a.defer_reset = 0u;
        if (base >= n_listed) return;  // (CTA-uniform)
    }
    bbox_init(bacc);
    if (!V1) knn_shared_init(ks);
    const int sub = threadIdx.x % G, ql = threadIdx.x / G;
    if (!listed)
        load_query_tile(tile, lx, ly, lz, base, a.n_local, a.tma_ok != 0);  // (holds the __syncthreads)
    else
    {
        const bool     in = base + ql < n_listed;
        const uint32_t q  = in ? __ldg(a.qlist + base + ql) : 0u;
        if (sub == 0) tile.x[ql] = in ? lx[q] : 0.f, tile.y[ql] = in ? ly[q] : 0.f, tile.z[ql] = in ? lz[q] : 0.f;
        __syncthreads();
    }
    const uint32_t qpos  = listed ? (base + ql < n_listed ? __ldg(a.qlist + base + ql) : 0xFFFFFFFFu) : (uint32_t)base + ql;  // position in the array walked (sorted if perm)
    const bool     valid = qpos < a.n_local;
    const uint32_t i     = (perm && valid) ? __ldg(perm + qpos) : qpos;  // the caller's index of this query
    const uint32_t co    = a.cand_sorted ? qpos : i;                     // where its candidate words go

    float gx = 0, gy = 0, gz = 0;
    if (valid) compose_point_f(a.pose, tile.x[ql], tile.y[ql], tile.z[ql], gx, gy, gz);
    bbox_accumulate(bacc, gx, gy, gz, valid && sub == 0 && !listed, bbox_words);  // (the thread pass counted the listed queries)
    const int K = (int)a.K;
    // …DistanceThreshold.cpp:230,256-257 (float, unfused)
    const float normSq = __fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), __fmul_rn(gz, gz));
    const float thr2   = __fadd_rn(a.maxDistSq, __fmul_rn(a.angSq, normSq));
    const unsigned long long sentinel = (unsigned long long)__float_as_uint(thr2) << 32;

    unsigned long long mine;  // lane `sub` ends up with the sub-th best key
    SearchCounters     sc;
    // all lanes take part (warp-uniform search); lanes past the end and already paired locals
    // (:218-220) are disabled
    if constexpr (V1)
        knn_search_v1<G>(g, valid && (a.allowLocal || !bit_set(lbits, i)), gx, gy, gz, thr2, K, a.rl_start, mine, sub, sc);
    else
        knn_search<G>(g, valid && (a.allowLocal || !bit_set(lbits, i)), gx, gy, gz, thr2, K, a.rl_start, mine, sub, sc,
                      ks.tb, ks.nb, a.n_phases);
    if (valid)
    {
        // lane r < K writes rank r; unused ranks are marked with an impossible map index (all ones)
        const unsigned long long c = (sub < K && mine < sentinel) ? mine : ~0ull;
        if (sub < K)
        {
            cand[(size_t)co * K + sub] = c;
            if (c != ~0ull && !a.allowGlobal)
            {
                const uint32_t gi = (uint32_t)c;
                if (!bit_set(gbits, gi)) claim_propose(claim, a.claim_parts, a.claim_world, gi, a.tag | ((unsigned long long)(i * (uint32_t)K + sub) + a.slot_offset));
            }
        }
        flush_search_stats(sc, c != ~0ull ? 1u : 0u, stats);
    }
    if (fit.list)  // all lanes: warp-aggregated append of the queries that qualify for a plane fit
    {
        const bool     q_lane = valid && sub == fit.need - 1;  // the lane holding rank need-1
        const bool     has    = q_lane && fit.need <= K && mine < sentinel;
        const unsigned m      = __ballot_sync(0xffffffffu, has);
        const int      lane   = threadIdx.x & 31;
        uint32_t       at     = 0;
        if (m && lane == __ffs(m) - 1) at = atomicAdd(fit.count, (uint32_t)__popc(m));
        at = __shfl_sync(0xffffffffu, at, m ? __ffs(m) - 1 : 0);
        if (has) fit.list[at + __popc(m & ((1u << lane) - 1u))] = qpos;
        if (q_lane && !has) fit.ok_flags[i] = 0;
        if (valid && fit.need > G && sub == 0) fit.ok_flags[i] = 0;  // cannot hold `need` neighbours at all
    }
    // how long this tile took (its slowest warp, in units of 64 clocks): the next call over the same cloud
    // starts the long tiles first (k_tile_rank)
    if (a.tile_cost && (threadIdx.x & 31) == 0) atomicMax(a.tile_cost + tile_id, (uint32_t)min((clock64() - t_start) >> 6, 0xffffffffll));
    if (a.tile_trace)  // measurement hook ($MP2P_KNN_TRACE=1): where and when every CTA ran
    {
        __syncthreads();
        if (threadIdx.x == 0)
        {
            unsigned long long t1;
            uint32_t           smid;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            a.tile_trace[8 * blockIdx.x + 0] = smid, a.tile_trace[8 * blockIdx.x + 1] = (uint32_t)t_trace0;
            a.tile_trace[8 * blockIdx.x + 2] = (uint32_t)t1, a.tile_trace[8 * blockIdx.x + 3] = tile_id;
            a.tile_trace[8 * blockIdx.x + 4] = sc.rounds, a.tile_trace[8 * blockIdx.x + 5] = sc.steps;
            a.tile_trace[8 * blockIdx.x + 6] = sc.inserts, a.tile_trace[8 * blockIdx.x + 7] = sc.probes | (sc.levels << 16);
        }
    }
}

// The same contract served by ONE THREAD per query (knn_search_thread, grid_search.cuh): a CTA of NT threads
// takes a tile of NT queries; thread t leaves rank r of its query at cand[.. * K + r] like lane r of a group.
template <int KM, int NT>
__global__ void __launch_bounds__(NT)
    k_match_knn_thread(GridView g, Pt2PtArgs a, const float* __restrict__ lx, const float* __restrict__ ly,
                       const float* __restrict__ lz, const uint32_t* __restrict__ perm, const uint32_t* __restrict__ lbits,
                       const uint32_t* __restrict__ gbits, unsigned long long* __restrict__ claim,
                       unsigned long long* __restrict__ cand, uint32_t* __restrict__ bbox_words,
                       unsigned long long* __restrict__ stats, FitList fit)
{
    __shared__ QueryTile<NT>       tile;
    __shared__ BBoxAcc             bacc;
    __shared__ KnnThreadShared<NT> ks;
    const long long    t_start  = clock64();
    unsigned long long t_trace0 = 0;
    if (a.tile_trace) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_trace0));
    const uint32_t tile_id = a.tile_order ? __ldg(a.tile_order + blockIdx.x) : (uint32_t)(((unsigned long long)blockIdx.x * a.tile_stride) % gridDim.x);
    const size_t   base    = (size_t)tile_id * NT;
    bbox_init(bacc);
    load_query_tile(tile, lx, ly, lz, base, a.n_local, a.tma_ok != 0);  // (holds the __syncthreads)
    const uint32_t qpos  = (uint32_t)base + threadIdx.x;  // position in the array walked (sorted if perm)
    const bool     valid = qpos < a.n_local;
    const uint32_t i     = (perm && valid) ? __ldg(perm + qpos) : qpos;  // the caller's index of this query
    const uint32_t co    = a.cand_sorted ? qpos : i;                     // where its candidate words go

    float gx = 0, gy = 0, gz = 0;
    if (valid) compose_point_f(a.pose, tile.x[threadIdx.x], tile.y[threadIdx.x], tile.z[threadIdx.x], gx, gy, gz);
    bbox_accumulate(bacc, gx, gy, gz, valid, bbox_words);
    const int K = (int)a.K;
    // …DistanceThreshold.cpp:230,256-257 (float, unfused)
    const float normSq = __fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), __fmul_rn(gz, gz));
    const float thr2   = __fadd_rn(a.maxDistSq, __fmul_rn(a.angSq, normSq));
    const unsigned long long sentinel = (unsigned long long)__float_as_uint(thr2) << 32;

    unsigned long long best[KM];  // descending: best[K - 1 - r] = the r-th best key
    SearchCounters     sc;
    // lanes past the end and already paired locals (:218-220) do not search
    const bool deferred = knn_search_thread<KM, NT>(g, valid && (a.allowLocal || !bit_set(lbits, i)), gx, gy, gz, thr2, K, a.rl_start, best, sc, ks, a.defer, qpos);
    __syncwarp();
    unsigned long long need_key = ~0ull;  // the key of rank fit.need - 1
    if (deferred) flush_search_stats(sc, 0u, stats);  // (the work done before handing over is work done)
    if (valid && !deferred)
    {
        uint32_t n_found = 0;
#pragma unroll
        for (int j = 0; j < KM; j++)
        {
            if (j >= K) break;
            const int                r = K - 1 - j;
            const unsigned long long c = best[j] < sentinel ? best[j] : ~0ull;  // unused ranks: an impossible map index (all ones)
            cand[(size_t)co * K + r]   = c;
            if (r == fit.need - 1) need_key = c;
            n_found += c != ~0ull;
        }
        if (!a.allowGlobal)  // first claims (one rolled loop over the words just written: the unrolled one was half of the kernel's code)
#pragma unroll 1
            for (int r = 0; r < K; r++)
            {
                const unsigned long long c = cand[(size_t)co * K + r];
                if (c == ~0ull) continue;
                const uint32_t gi = (uint32_t)c;
                if (!bit_set(gbits, gi)) claim_propose(claim, a.claim_parts, a.claim_world, gi, a.tag | ((unsigned long long)(i * (uint32_t)K + r) + a.slot_offset));
            }
        flush_search_stats(sc, n_found, stats);
    }
    __syncwarp();
    if (fit.list)  // all lanes: warp-aggregated append of the queries that qualify for a plane fit
    {
        const bool     has  = valid && !deferred && fit.need <= K && need_key != ~0ull;
        const unsigned m    = __ballot_sync(0xffffffffu, has);
        const int      lane = threadIdx.x & 31;
        uint32_t       at   = 0;
        if (m && lane == __ffs(m) - 1) at = atomicAdd(fit.count, (uint32_t)__popc(m));
        at = __shfl_sync(0xffffffffu, at, m ? __ffs(m) - 1 : 0);
        if (has) fit.list[at + __popc(m & ((1u << lane) - 1u))] = qpos;
        if (valid && !deferred && !has) fit.ok_flags[i] = 0;
    }
    // how long this tile took (its slowest warp, in units of 64 clocks): see k_match_pt2pt
    if (a.tile_cost && (threadIdx.x & 31) == 0) atomicMax(a.tile_cost + tile_id, (uint32_t)min((clock64() - t_start) >> 6, 0xffffffffll));
    if (a.tile_trace)  // measurement hook ($MP2P_KNN_TRACE=1): where and when every CTA ran
    {
        __syncthreads();
        if (threadIdx.x == 0)
        {
            unsigned long long t1;
            uint32_t           smid;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            a.tile_trace[8 * blockIdx.x + 0] = smid, a.tile_trace[8 * blockIdx.x + 1] = (uint32_t)t_trace0;
            a.tile_trace[8 * blockIdx.x + 2] = (uint32_t)t1, a.tile_trace[8 * blockIdx.x + 3] = tile_id;
            a.tile_trace[8 * blockIdx.x + 4] = sc.levels, a.tile_trace[8 * blockIdx.x + 5] = sc.cands;
            a.tile_trace[8 * blockIdx.x + 6] = 0u, a.tile_trace[8 * blockIdx.x + 7] = sc.probes | (sc.levels << 16);
        }
    }
}

// Tile order for the next call: tiles by DESCENDING cost, 256 cost classes (counting sort by one CTA; the order
// inside a class is arbitrary — it only decides which CTA serves which tile, never a result).
__global__ void __launch_bounds__(1024) k_tile_rank(uint32_t* __restrict__ cost, uint32_t n_tiles, uint32_t* __restrict__ order)
{
    __shared__ uint32_t hist[256], cursor[256], mx;
    if (threadIdx.x < 256) hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) mx = 1;
    __syncthreads();
    uint32_t m = 0;
    for (uint32_t t = threadIdx.x; t < n_tiles; t += blockDim.x) m = max(m, cost[t]);
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0) atomicMax(&mx, m);
    __syncthreads();
    const float scale = 255.f / (float)mx;
    for (uint32_t t = threadIdx.x; t < n_tiles; t += blockDim.x) atomicAdd(&hist[255 - min(255u, (uint32_t)((float)cost[t] * scale))], 1u);
    __syncthreads();
    if (threadIdx.x == 0)
    {
        uint32_t acc = 0;
        for (int b = 0; b < 256; b++) cursor[b] = acc, acc += hist[b];
    }
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < n_tiles; t += blockDim.x)
    {
        order[atomicAdd(&cursor[255 - min(255u, (uint32_t)((float)cost[t] * scale))], 1u)] = t;
        cost[t] = 0u;  // re-armed for the next search (atomicMax accumulates into it)
    }
}

// ------------------------------------------------------------------------------------------
// K = 1 matcher: one LANE per query for the scalar part (transform, threshold, centre voxel), then
// the neighbour voxels that survive the box bound — a few per query, 0 for most — are pooled over
// the warp and dealt out again one (query, voxel) item per lane ("warp work redistribution"), so
// the hash probes and point reads of 32 different items are in flight together instead of one
// thread walking 27 voxels. Item results flow back to the owner through a shared-memory
// atomicMin on the 64-bit (d2, index) key. Exactness argument unchanged: every voxel with box
// bound <= current best distance is visited, candidates compare on (d2, index).
// ------------------------------------------------------------------------------------------
// scan `count` consecutive points for the smallest (d2, index) key: loads go out four at a time
// (clamped to the last point of the run — a duplicate never wins the strict compare), so a voxel
// of <= 4 points costs ONE memory round trip; returns the key, *best_j = position inside the run
template <int W = 4>
__device__ __forceinline__ unsigned long long scan_run_min(const float4* __restrict__ run, uint32_t count, float qx,
                                                           float qy, float qz, unsigned long long m, uint32_t& best_j,
                                                           uint32_t first = 0)
{
    const uint32_t last = count - 1;
    for (uint32_t j0 = first; j0 < count; j0 += W)
    {
        float4 q[W];
#pragma unroll
        for (int k = 0; k < W; k++) q[k] = __ldg(run + min(j0 + k, last));
#pragma unroll
        for (int k = 0; k < W; k++)
        {
            const unsigned long long c = point_key(qx, qy, qz, q[k]);
            if (c < m) m = c, best_j = min(j0 + k, last);
        }
    }
    return m;
}

// Body of the K = 1 matcher for a CTA of NT threads (NT queries); shared by the stand-alone kernel
// and by the fused single-launch iteration (k_iterate_nn1_horn).
// R = neighbour items a lane keeps in flight per pass (their hash probes go out together, then the
// first W points of each run): the dependent memory round trips of a warp's item list shrink from
// 2 x ceil(items / 32) to 2 x ceil(items / (32 R)). CW = points per step of the centre-voxel scan.
template <int NT, int R, int W, int CW>
__device__ __forceinline__ void nn1_body(const GridView& g, const Pt2PtArgs& a, const float* __restrict__ lx,
                                         const float* __restrict__ ly, const float* __restrict__ lz,
                                         const uint32_t* __restrict__ perm, const uint32_t* __restrict__ lbits,
                                         const uint32_t* __restrict__ gbits, unsigned long long* __restrict__ claim,
                                         unsigned long long* __restrict__ cand, float4* __restrict__ cand_xyz,
                                         uint32_t* __restrict__ bbox_words, unsigned long long* __restrict__ stats)
{
    constexpr int kWarpsNN1 = NT / 32;
    __shared__ QueryTile<NT> tile;
    __shared__ BBoxAcc                bacc;
    __shared__ unsigned long long     s_best[kWarpsNN1][32];
    __shared__ uint32_t               s_run[kWarpsNN1][32];         // map position of s_best's point
    __shared__ uint16_t               s_item[kWarpsNN1][32 * 26];   // (owner lane << 5) | neighbour bit
    const size_t                      base = (size_t)blockIdx.x * NT;
    bbox_init(bacc);
    load_query_tile(tile, lx, ly, lz, base, a.n_local, a.tma_ok != 0);
    const int      lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t qpos  = (uint32_t)base + threadIdx.x;  // position in the array walked (sorted if perm)
    const bool     valid = qpos < a.n_local;
    const uint32_t i     = (perm && valid) ? __ldg(perm + qpos) : qpos;  // the caller's index of this query
    const unsigned FULL  = 0xffffffffu;

    float gx = 0, gy = 0, gz = 0;
    if (valid) compose_point_f(a.pose, tile.x[threadIdx.x], tile.y[threadIdx.x], tile.z[threadIdx.x], gx, gy, gz);
    if (a.stage_x && valid)
        a.stage_x[qpos] = tile.x[threadIdx.x], a.stage_y[qpos] = tile.y[threadIdx.x], a.stage_z[qpos] = tile.z[threadIdx.x];
    bbox_accumulate(bacc, gx, gy, gz, valid, bbox_words);

    // …DistanceThreshold.cpp:230,256-257 (float, unfused); skipped locals (:218-220) get radius 0
    float thr2 = 0.f;
    if (valid && (a.allowLocal || !bit_set(lbits, i)))
    {
        const float normSq = __fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), __fmul_rn(gz, gz));
        thr2               = __fadd_rn(a.maxDistSq, __fmul_rn(a.angSq, normSq));
    }
    const unsigned long long sentinel = (unsigned long long)__float_as_uint(thr2) << 32;
    unsigned long long       best     = sentinel;
    uint32_t                 best_pos = 0;  // position in g.pts of the best candidate
    float                    kth      = thr2;
    bool                     active   = thr2 > 0.f;
    if (active)
    {
        const float ex = fmaxf(fmaxf(g.bbmin[0] - gx, gx - g.bbmax[0]), 0.f);
        const float ey = fmaxf(fmaxf(g.bbmin[1] - gy, gy - g.bbmax[1]), 0.f);
        const float ez = fmaxf(fmaxf(g.bbmin[2] - gz, gz - g.bbmax[2]), 0.f);
        if ((ex * ex + ey * ey + ez * ez) * 0.999999f > thr2) active = false;
    }
    const float lim = 4194304.f;  // 2^22
    const float ux  = fminf(fmaxf(grid_u(gx, g.ox, g.inv_s0), -lim), lim);
    const float uy  = fminf(fmaxf(grid_u(gy, g.oy, g.inv_s0), -lim), lim);
    const float uz  = fminf(fmaxf(grid_u(gz, g.oz, g.inv_s0), -lim), lim);
    const int   Ix = (int)floorf(ux), Iy = (int)floorf(uy), Iz = (int)floorf(uz);
    const float q2 = g.s0_lo * g.s0_lo * 0.999999f;
    SearchCounters sc;

    for (int rl = 0; rl < g.n_levels; rl++)
    {
        if (!__any_sync(FULL, active)) break;
        const int L = g.level_first + rl;
        if (L == kGridBits)
        {
            // top level = every point (see grid_search.cuh); reached only by queries whose radius
            // exceeds half the map extent
            if (active)
            {
                sc.probes++, sc.cands += g.n_points, sc.levels++;
                uint32_t j = 0;
                best       = scan_run_min(g.pts, g.n_points, gx, gy, gz, best, j);
                if (best < sentinel) best_pos = j;  // j is only meaningful if something won
            }
            break;
        }
        const int   cmax = ((1 << kGridBits) - 1) >> L;
        const float s    = (float)(1 << L);
        const int   cx = Ix >> L, cy = Iy >> L, cz = Iz >> L;
        const float fx = ux - (float)cx * s, fy = uy - (float)cy * s, fz = uz - (float)cz * s;

        // ---- centre voxel, own query
        if (active && (unsigned)cx <= (unsigned)cmax && (unsigned)cy <= (unsigned)cmax && (unsigned)cz <= (unsigned)cmax)
        {
            uint32_t start, count;
            sc.probes++;
            if (grid_lookup(g, rl, (uint32_t)cx, (uint32_t)cy, (uint32_t)cz, start, count))
            {
                sc.cands += count;
                uint32_t                 j = 0;
                const unsigned long long m = scan_run_min<CW>(g.pts + start, count, gx, gy, gz, best, j);
                if (m < best) best = m, best_pos = start + j;
            }
        }
        kth = fminf(kth, __uint_as_float((uint32_t)(best >> 32)));

        // ---- which of the 26 neighbours can hold something better: bit b = dz*9 + dy*3 + dx (each 0..2)
        uint32_t mask = 0;
        if (active)
        {
            // squared conservative gaps (metres^2) to the -1 / 0 / +1 slabs per axis
            const float glx = fmaxf(fx - 4.f, 0.f), ghx = fmaxf(s - fx - 4.f, 0.f);
            const float gly = fmaxf(fy - 4.f, 0.f), ghy = fmaxf(s - fy - 4.f, 0.f);
            const float glz = fmaxf(fz - 4.f, 0.f), ghz = fmaxf(s - fz - 4.f, 0.f);
            const float ax[3] = {glx * glx * q2, 0.f, ghx * ghx * q2};
            const float ay[3] = {gly * gly * q2, 0.f, ghy * ghy * q2};
            const float az[3] = {glz * glz * q2, 0.f, ghz * ghz * q2};
#pragma unroll
            for (int dz = 0; dz < 3; dz++)
            {
                if (az[dz] > kth || (unsigned)(cz + dz - 1) > (unsigned)cmax) continue;
#pragma unroll
                for (int dy = 0; dy < 3; dy++)
                {
                    const float t = az[dz] + ay[dy];
                    if (t > kth || (unsigned)(cy + dy - 1) > (unsigned)cmax) continue;
#pragma unroll
                    for (int dx = 0; dx < 3; dx++)
                    {
                        if (dx == 1 && dy == 1 && dz == 1) continue;
                        if (t + ax[dx] > kth || (unsigned)(cx + dx - 1) > (unsigned)cmax) continue;  // strict >
                        mask |= 1u << (dz * 9 + dy * 3 + dx);
                    }
                }
            }
        }
        // ---- pool the (query, voxel) items of the warp: every owner lists its items in shared
        // memory at its exclusive offset, then item `id` goes to lane id % 32
        const uint32_t cnt  = __popc(mask);
        uint32_t       incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t t = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += t;
        }
        const uint32_t total = __shfl_sync(FULL, incl, 31);
        {
            uint32_t pos = incl - cnt, m = mask;
            while (m)
            {
                const uint32_t bit = __ffs(m) - 1;
                m &= m - 1;
                s_item[warp][pos++] = (uint16_t)((lane << 5) | bit);
            }
        }
        s_best[warp][lane] = best;
        s_run[warp][lane]  = best_pos;
        __syncwarp();
        const uint32_t   shift = g.level_shift[rl], hmask = (1u << (64 - shift)) - 1u;
        const CellEntry* tab   = g.table + g.level_off[rl];
        for (uint32_t b0 = 0; b0 < total; b0 += 32 * R)
        {
            // ---- A: decode up to R items of this lane, first hash probe of each goes out
            float              oq[R][3];
            int                owner[R];
            bool               have[R];
            uint4              raw[R];
            unsigned long long ckey[R];
            uint32_t           h[R];
#pragma unroll
            for (int r = 0; r < R; r++)
            {
                have[r] = false, owner[r] = 0, ckey[r] = 0ull, h[r] = 0u, raw[r] = make_uint4(0u, 0u, 0u, 0u);
                oq[r][0] = oq[r][1] = oq[r][2] = 0.f;
                if (b0 + r * 32 >= total) continue;  // warp-uniform
                const uint32_t id = b0 + r * 32 + lane;
                // the shuffles below are executed by all lanes; lanes past the end mirror item 0's owner
                have[r]            = id < total;
                const uint32_t it  = have[r] ? s_item[warp][id] : 0u;
                owner[r]           = (int)(it >> 5);
                const uint32_t bit = it & 31u;
                oq[r][0] = __shfl_sync(FULL, gx, owner[r]), oq[r][1] = __shfl_sync(FULL, gy, owner[r]), oq[r][2] = __shfl_sync(FULL, gz, owner[r]);
                const int ocx = __shfl_sync(FULL, cx, owner[r]), ocy = __shfl_sync(FULL, cy, owner[r]), ocz = __shfl_sync(FULL, cz, owner[r]);
                const int dz = (int)(bit / 9u), dy = (int)((bit % 9u) / 3u), dx = (int)(bit % 3u);
                ckey[r] = cell_key((uint32_t)(ocx + dx - 1), (uint32_t)(ocy + dy - 1), (uint32_t)(ocz + dz - 1));
                h[r]    = cell_hash(ckey[r], shift);
                if (have[r]) raw[r] = __ldg(reinterpret_cast<const uint4*>(tab + h[r]));
            }
            // ---- B: resolve the probes (linear probing continues on a collision)
            uint32_t start[R], count[R];
#pragma unroll
            for (int r = 0; r < R; r++)
            {
                start[r] = 0u, count[r] = 0u;
                if (!have[r]) continue;
                sc.probes++;
                while (true)
                {
                    const unsigned long long k = (unsigned long long)raw[r].x | ((unsigned long long)raw[r].y << 32);
                    if (k == ckey[r])
                    {
                        start[r] = raw[r].z, count[r] = raw[r].w;
                        break;
                    }
                    if (k == kEmptyKey) break;
                    h[r]   = (h[r] + 1) & hmask;
                    raw[r] = __ldg(reinterpret_cast<const uint4*>(tab + h[r]));
                }
            }
            // ---- C: the first W points of every run found (index clamped to the run)
            float4 pw[R][W];
#pragma unroll
            for (int r = 0; r < R; r++)
#pragma unroll
                for (int k = 0; k < W; k++)
                    pw[r][k] = count[r] ? __ldg(g.pts + start[r] + min((uint32_t)k, count[r] - 1u)) : make_float4(0.f, 0.f, 0.f, 0.f);
            // ---- D: keys; longer runs continue four points at a time; results to the owners
            unsigned long long m[R];
            uint32_t           pos[R];
#pragma unroll
            for (int r = 0; r < R; r++)
            {
                m[r] = ~0ull, pos[r] = 0u;
                if (!count[r]) continue;
                sc.cands += count[r];
                uint32_t j = 0;
#pragma unroll
                for (int k = 0; k < W; k++)
                {
                    const unsigned long long c = point_key(oq[r][0], oq[r][1], oq[r][2], pw[r][k]);
                    if (c < m[r]) m[r] = c, j = min((uint32_t)k, count[r] - 1u);
                }
                if (count[r] > (uint32_t)W) m[r] = scan_run_min<4>(g.pts + start[r], count[r], oq[r][0], oq[r][1], oq[r][2], m[r], j, W);
                pos[r] = start[r] + j;
                atomicMin(&s_best[warp][owner[r]], m[r]);
            }
            // keys are unique (the map index is part of the key): once all items of the pass have
            // offered theirs, at most one of them equals the owner's slot — that one records where
            // its point sits (the owner's own centre candidate keeps its position otherwise)
            __syncwarp();
#pragma unroll
            for (int r = 0; r < R; r++)
                if (m[r] != ~0ull && s_best[warp][owner[r]] == m[r]) s_run[warp][owner[r]] = pos[r];
            __syncwarp();
        }
        best     = s_best[warp][lane];
        best_pos = s_run[warp][lane];
        __syncwarp();
        kth = fminf(kth, __uint_as_float((uint32_t)(best >> 32)));
        // everything outside the 3x3x3 block is at least m quanta away
        const float mx = s + fminf(fx, s - fx), my = s + fminf(fy, s - fy), mz = s + fminf(fz, s - fz);
        const float m  = fmaxf(fminf(mx, fminf(my, mz)) - 4.f, 0.f);
        if (active)
        {
            sc.levels++;
            if (kth <= m * m * q2) active = false;
        }
    }

    uint32_t n_valid = 0;
    if (valid)
    {
        // unused ranks are marked with an impossible map index (all ones)
        const unsigned long long c = best < sentinel ? best : ~0ull;
        n_valid                    = (c != ~0ull);
        cand[i]                    = c;
        float4 bp                  = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n_valid) bp = __ldg(g.pts + best_pos);  // just read by some lane of this warp: L1/L2 hit
        cand_xyz[i] = make_float4(bp.x, bp.y, bp.z, 0.f);
        if (c != ~0ull && !a.allowGlobal)
        {
            const uint32_t gi = (uint32_t)c;
            if (!bit_set(gbits, gi)) claim_propose(claim, a.claim_parts, a.claim_world, gi, a.tag | ((unsigned long long)i + a.slot_offset));
        }
    }
    flush_search_stats(sc, n_valid, stats);
}

#ifndef MP2P_NN1_FUSED_R
#define MP2P_NN1_FUSED_R 1  // variant built into the single-launch iteration (register budget: 85)
#define MP2P_NN1_FUSED_W 4
#define MP2P_NN1_FUSED_CW 4
#endif
template <int R, int W, int CW>
__global__ void __launch_bounds__(kNN1Threads)
    k_match_pt2pt_nn1(GridView g, Pt2PtArgs a, const float* __restrict__ lx, const float* __restrict__ ly,
                      const float* __restrict__ lz, const uint32_t* __restrict__ perm,
                      const uint32_t* __restrict__ lbits,
                      const uint32_t* __restrict__ gbits, unsigned long long* __restrict__ claim,
                      unsigned long long* __restrict__ cand, float4* __restrict__ cand_xyz,
                      uint32_t* __restrict__ bbox_words, unsigned long long* __restrict__ stats)
{
    nn1_body<kNN1Threads, R, W, CW>(g, a, lx, ly, lz, perm, lbits, gbits, claim, cand, cand_xyz, bbox_words, stats);
}
// variant of the stand-alone K = 1 kernel: $MP2P_NN1_VARIANT (A/B on the device). Measured on C2
// (gpurun visit 19, profiles/r01_nn1_variants_ab.txt): 0 -> 33.0 us, 1 -> 33.4, 2 -> 34.0, 3 -> 43.3,
// 4 -> 46.8 (register pressure costs the single wave): more items in flight per lane do not pay,
// the kernel is bound by the L1 tag rate of its divergent 16-byte loads, not by the chain length.
inline int nn1_variant()
{
    static int v = -1;
    if (v < 0)
    {
        const char* e = getenv("MP2P_NN1_VARIANT");
        v             = e ? atoi(e) : 0;
        if (v < 0 || v > 4) v = 0;
    }
    return v;
}
#define MP2P_LAUNCH_NN1(grid, st, ...)                                                       \
    switch (nn1_variant())                                                                   \
    {                                                                                        \
    case 0: k_match_pt2pt_nn1<1, 4, 4><<<grid, kNN1Threads, 0, st>>>(__VA_ARGS__); break;    \
    case 1: k_match_pt2pt_nn1<2, 4, 4><<<grid, kNN1Threads, 0, st>>>(__VA_ARGS__); break;    \
    case 2: k_match_pt2pt_nn1<2, 4, 8><<<grid, kNN1Threads, 0, st>>>(__VA_ARGS__); break;    \
    case 3: k_match_pt2pt_nn1<4, 2, 8><<<grid, kNN1Threads, 0, st>>>(__VA_ARGS__); break;    \
    default: k_match_pt2pt_nn1<4, 4, 8><<<grid, kNN1Threads, 0, st>>>(__VA_ARGS__); break;   \
    }

// ------------------------------------------------------------------------------------------
// Single-pass stream compaction (decoupled look-back over 1024-slot tiles).
// status word: [63:62] 1 = tile aggregate, 2 = inclusive prefix; [61:40] call epoch (22 bits);
// [39:0] value. A word whose epoch is not the current call's reads as "not ready", so the status
// array is never cleared between calls.
// ------------------------------------------------------------------------------------------
constexpr int      kScanThreads = 256;
constexpr int      kScanItems   = 1;  // the fused iteration, inlier-ratio and adaptive compactions: one slot per thread
constexpr uint32_t kScanTile    = kScanThreads * kScanItems;
// The stand-alone compactions take several CONSECUTIVE slots per thread. With one slot per thread the 1M slots
// of C5 are 3,907 tiles whose look-back walks (32 tiles per step, one L2 round trip each) run ~120 steps deep
// when every tile starts at once: 119 us for 44 MB. Four slots per thread quarter the chain; the loads of the
// four slots are in flight together.
constexpr int      kCompactItems2p = 4;  // 36-byte records: 36 KB of staging per CTA
constexpr int      kCompactItems2l = 2;  // 72-byte records: 36 KB
inline uint32_t compact_tiles_2p(uint64_t n_slots) { return (uint32_t)((n_slots + kScanThreads * kCompactItems2p - 1) / (kScanThreads * kCompactItems2p)); }
inline uint32_t compact_tiles_2l(uint64_t n_local) { return (uint32_t)((n_local + kScanThreads * kCompactItems2l - 1) / (kScanThreads * kCompactItems2l)); }

struct ScanSmem
{
    uint32_t warp_sums[kScanThreads / 32];
    uint32_t tile_id;
    uint32_t tile_total;           // outputs of this tile
    unsigned long long tile_base;  // outputs of all earlier tiles
};

// returns this thread's exclusive offset inside the grid-wide output; *total_out written by the
// last tile. `local` = number of outputs of this thread (its kScanItems consecutive slots).
__device__ __forceinline__ unsigned long long grid_exclusive_scan(
    ScanSmem& sm, uint32_t tile, uint32_t n_tiles, uint32_t local, unsigned long long* status,
    unsigned long long* total_out, uint32_t epoch22)
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t       incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += t;
    }
    if (lane == 31) sm.warp_sums[warp] = incl;
    __syncthreads();
    uint32_t warp_off = 0, block_total = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; w++)
    {
        const uint32_t s = sm.warp_sums[w];
        if (w < (int)warp) warp_off += s;
        block_total += s;
    }
    if (warp == 0)
    {
        // decoupled look-back, one WARP wide: lane l inspects tile (t - l); the nearest tile that
        // already published an inclusive prefix ends the walk.
        constexpr unsigned long long kVal = (1ull << 40) - 1;
        const unsigned long long     ep   = (unsigned long long)(epoch22 & 0x3FFFFFu) << 40;
        volatile unsigned long long* vstatus = status;
        if (tile > 0 && lane == 0)
        {
            __threadfence();
            vstatus[tile] = (1ull << 62) | ep | block_total;  // aggregate available
        }
        // (a window of 128 tiles per step — four statuses per lane — was measured and is slower: 68.9 us against
        // 62.9 us on the 1M slots of C5; once the tiles hold four slots per thread the walk is not the bound)
        unsigned long long base = 0;
        int                t    = (int)tile - 1;
        while (t >= 0)
        {
            const int          idx = t - (int)lane;
            unsigned long long sw  = (2ull << 62) | ep;  // tiles before the first: prefix 0
            if (idx >= 0)
            {
                do
                {
                    sw = vstatus[idx];
                } while ((sw >> 62) == 0 || (sw & (0x3FFFFFull << 40)) != ep);
            }
            const unsigned pm   = __ballot_sync(0xffffffffu, (sw >> 62) == 2);
            const int      stop = pm ? __ffs(pm) - 1 : 31;  // nearest published prefix, if any
            unsigned long long v = ((int)lane <= stop) ? (sw & kVal) : 0ull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            base += v;
            if (pm) break;  // (tiles before the first count as prefix 0, so this always triggers)
            t -= 32;
        }
        if (lane == 0)
        {
            __threadfence();
            vstatus[tile] = (2ull << 62) | ep | (base + block_total);
            sm.tile_base  = base;
            sm.tile_total = block_total;
            if (tile == n_tiles - 1) *total_out = base + block_total;
        }
    }
    __syncthreads();
    return sm.tile_base + warp_off + (incl - local);
}

struct CoopSync
{
    unsigned long long* arrivals;  // monotonically increasing across launches (never reset)
    unsigned long long  target;    // arrivals value that completes this launch's FIRST barrier (+ grid per further one)
    unsigned int*       sums_flag; // epoch of the launch whose HORN1 packet is complete
    unsigned int        epoch;
    double*             host_out;  // mapped pinned host memory: 64 packet doubles, then the flag word
    unsigned int        scan_barrier;  // number of the barrier the resident scan uses (1; 3 in the sharded launch)
};

// barrier number `which` (0, 1, ...) of a cooperative launch: all CTAs are resident, so spinning is safe
__device__ __forceinline__ void grid_barrier(const CoopSync& cs, unsigned which)
{
    __syncthreads();
    if (threadIdx.x == 0)
    {
        const unsigned long long target = cs.target + (unsigned long long)which * gridDim.x;
        __threadfence();
        atomicAdd(cs.arrivals, 1ull);
        while (*reinterpret_cast<volatile unsigned long long*>(cs.arrivals) < target) __nanosleep(20);
        __threadfence();
    }
    __syncthreads();
}

// Exclusive scan for CO-RESIDENT tiles (cooperative launch): every tile publishes its count, one grid
// barrier, every tile adds up the counts before it — two memory round trips instead of a look-back
// chain that is fully serialised when all tiles start at the same moment.
__device__ __forceinline__ unsigned long long grid_exclusive_scan_resident(
    ScanSmem& sm, uint32_t tile, uint32_t n_tiles, uint32_t local, unsigned long long* counts,
    unsigned long long* total_out, const CoopSync& cs, unsigned barrier_no)
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t       incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += t;
    }
    if (lane == 31) sm.warp_sums[warp] = incl;
    __syncthreads();
    uint32_t warp_off = 0, block_total = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; w++)
    {
        const uint32_t s = sm.warp_sums[w];
        if (w < (int)warp) warp_off += s;
        block_total += s;
    }
    if (threadIdx.x == 0) counts[tile] = block_total;
    grid_barrier(cs, barrier_no);
    unsigned long long part = 0;
    for (uint32_t t = threadIdx.x; t < tile; t += kScanThreads) part += __ldcg(counts + t);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    __shared__ unsigned long long s_part[kScanThreads / 32];
    if (lane == 0) s_part[warp] = part;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        unsigned long long base = 0;
#pragma unroll
        for (int w = 0; w < kScanThreads / 32; w++) base += s_part[w];
        sm.tile_base  = base;
        sm.tile_total = block_total;
        if (tile == n_tiles - 1) *total_out = base + block_total;
    }
    __syncthreads();
    return sm.tile_base + warp_off + (incl - local);
}


struct CompactArgs
{
    uint32_t n_local, K;
    int      allowGlobal;
    unsigned long long tag;
    float    gate_eps;  // threshold + bounding_box_intersection_check_epsilon (float)
    uint64_t capacity;
    uint32_t scan_epoch;    // stamps the look-back status words of this call
    uint64_t slot_offset;   // sharded runs: first global proposal slot of this shard (index_offset*K)
    uint32_t index_offset;  // sharded runs: first global local-point index of this shard
    // zero-copy output: the caller's pinned host buffer (device alias) also receives the records, and
    // the pairing count goes to pinned memory — no D2H copy behind the kernel
    uint32_t*           out_host;
    unsigned long long* count_host;
    // owner-partitioned claims of a query-sharded run (NULL = the map's own claim array)
    unsigned long long* const* claim_parts;
    uint32_t                   claim_world;
};

__device__ __forceinline__ bool bbox_gate(const GridView& g, const uint32_t* __restrict__ words, float eps)
{
    // mrpt TBoundingBoxf::intersection(other, epsilon) has a value (…DistanceThreshold.cpp:73-75)
#pragma unroll
    for (int d = 0; d < 3; d++)
    {
        if (__fsub_rn(ord2f(__ldcg(words + d)), eps) > g.bbmax[d]) return false;
        if (__fadd_rn(ord2f(__ldcg(words + 3 + d)), eps) < g.bbmin[d]) return false;
    }
    return true;
}
// the two bbox slots alternate between calls: the consumer of slot e re-arms slot e^1
__device__ __forceinline__ void bbox_rearm(uint32_t* __restrict__ next_words)
{
    if (blockIdx.x == 0 && threadIdx.x < 6) next_words[threadIdx.x] = threadIdx.x < 3 ? 0xFFFFFFFFu : 0u;
}

// sums (optional): HORN1 packet (sum local xyz, sum global xyz, count) of the accepted pairings,
// i.e. eval_centroids_robust (Pairings.cpp:68-110) folded into the compaction when the solver is
// known to follow (fused iteration) — saves one pass over the pairings.
struct FusedSums
{
    double*       partials;  // one row of 8 per tile
    unsigned int* ticket;
    double*       packet;    // NULL = not requested
};

// Body of the pt2pt compaction for tile `tile` of a CTA of kScanThreads threads. On return the
// tile's accepted records sit in s_rec[0 .. n_rec*9) (n_rec returned); shared by the stand-alone
// kernel and by the fused single-launch iteration.
template <int ITEMS = 1>
__device__ __forceinline__ uint32_t compact_pt2pt_body(
    const GridView& g, const CompactArgs& a, const float* __restrict__ lx, const float* __restrict__ ly,
    const float* __restrict__ lz, const uint32_t* __restrict__ gbits, const unsigned long long* __restrict__ claim,
    const unsigned long long* __restrict__ cand, const float4* __restrict__ cand_xyz, const uint32_t* __restrict__ bbox,
    unsigned long long* __restrict__ status, mp2p_b200_pair_pt2pt* __restrict__ out,
    unsigned long long* __restrict__ out_count, const FusedSums& fs, uint32_t tile, ScanSmem& sm, uint32_t* s_rec,
    bool* folded_sums = nullptr, const CoopSync* resident = nullptr)
{
    constexpr uint32_t TILE = kScanThreads * ITEMS;
    const uint64_t n_slots = (uint64_t)a.n_local * a.K;
    const uint32_t n_tiles = (uint32_t)((n_slots + TILE - 1) / TILE);
    const bool     gate    = bbox_gate(g, bbox, a.gate_eps);

    // Everything a record needs is requested up front (cand -> {claim word, global point, local
    // point} in parallel, for all of the thread's slots) so that only ONE dependent memory round trip
    // precedes the scan.
    const uint64_t     slot0 = (uint64_t)tile * TILE + (uint64_t)threadIdx.x * ITEMS;
    bool               ok[ITEMS];
    unsigned long long c[ITEMS];
    uint32_t           i[ITEMS], gi[ITEMS];
    float4             gp[ITEMS];
    float              px[ITEMS], py[ITEMS], pz[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; j++)
    {
        const uint64_t slot = slot0 + j;
        ok[j] = false, c[j] = ~0ull, i[j] = 0, gi[j] = 0, gp[j] = make_float4(0.f, 0.f, 0.f, 0.f), px[j] = py[j] = pz[j] = 0.f;
        if (gate && slot < n_slots)
        {
            c[j]  = cand[slot];
            ok[j] = ((uint32_t)c[j] != 0xFFFFFFFFu);  // unused ranks carry an all-ones map index
        }
    }
#pragma unroll
    for (int j = 0; j < ITEMS; j++)
        if (ok[j])
        {
            const uint64_t slot = slot0 + j;
            gi[j] = (uint32_t)c[j];
            i[j]  = (uint32_t)(a.K == 1 ? slot : slot / a.K);
            unsigned long long cw = a.tag | (unsigned long long)(slot + a.slot_offset);
            if (!a.allowGlobal) cw = claim_read(claim, a.claim_parts, a.claim_world, gi[j]);
            // the K = 1 matcher hands the matched point's coordinates over with the candidate
            // (sequential read); the K > 1 matcher does not (random gather from the map)
            gp[j] = cand_xyz ? __ldcs(cand_xyz + slot) : __ldg(g.pts_orig + gi[j]);
            px[j] = lx[i[j]], py[j] = ly[i[j]], pz[j] = lz[i[j]];
            if (!a.allowGlobal)
                ok[j] = !bit_set(gbits, gi[j]) && (cw == (a.tag | (unsigned long long)(slot + a.slot_offset)));
        }
    uint32_t local = 0;
#pragma unroll
    for (int j = 0; j < ITEMS; j++) local += ok[j] ? 1u : 0u;
    const unsigned long long w =
        resident ? grid_exclusive_scan_resident(sm, tile, n_tiles, local, status, out_count, *resident, resident->scan_barrier)
                 : grid_exclusive_scan(sm, tile, n_tiles, local, status, out_count, a.scan_epoch);
    // The tile's records are consecutive in the output: stage them in shared memory and store the
    // byte range with fully coalesced 4-byte words (full sectors: no read-for-ownership fills),
    // instead of nine strided stores per thread.
    const unsigned long long tile_base = sm.tile_base;
    {
        uint32_t at = (uint32_t)(w - tile_base);
#pragma unroll
        for (int j = 0; j < ITEMS; j++)
            if (ok[j])
            {
                uint32_t* o = s_rec + at * 9;
                at++;
                o[0] = gi[j], o[1] = i[j] + a.index_offset;
                o[2] = __float_as_uint(gp[j].x), o[3] = __float_as_uint(gp[j].y), o[4] = __float_as_uint(gp[j].z);
                o[5] = __float_as_uint(px[j]), o[6] = __float_as_uint(py[j]), o[7] = __float_as_uint(pz[j]);
                o[8] = (uint32_t)(c[j] >> 32);
            }
    }
    __syncthreads();
    const unsigned long long room  = a.capacity > tile_base ? a.capacity - tile_base : 0ull;
    const uint32_t           n_rec = (uint32_t)min((unsigned long long)sm.tile_total, room);
    {
        uint32_t* dst = reinterpret_cast<uint32_t*>(out) + tile_base * 9;
        for (uint32_t k = threadIdx.x; k < n_rec * 9; k += kScanThreads) dst[k] = s_rec[k];
    }
    if (a.out_host)
    {
        uint32_t* dst = a.out_host + tile_base * 9;
        for (uint32_t k = threadIdx.x; k < n_rec * 9; k += kScanThreads) dst[k] = s_rec[k];
        if (tile == n_tiles - 1 && threadIdx.x == 0) *a.count_host = tile_base + sm.tile_total;
    }
    if (fs.packet)
    {
        double             acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        unsigned long long at     = w;
#pragma unroll
        for (int j = 0; j < ITEMS; j++)  // (sequential per thread, slot order)
            if (ok[j])
            {
                if (at < a.capacity)
                {
                    acc[0] += (double)px[j], acc[1] += (double)py[j], acc[2] += (double)pz[j];
                    acc[3] += (double)gp[j].x, acc[4] += (double)gp[j].y, acc[5] += (double)gp[j].z;
                    acc[6] += 1.0, acc[7] += 1.0;
                }
                at++;
            }
        const bool folded = block_reduce_to_packet<8>(acc, fs.partials, fs.ticket, fs.packet, tile, n_tiles);
        if (folded_sums) *folded_sums = folded;
    }
    return n_rec;
}

__global__ void __launch_bounds__(kScanThreads)
    k_compact_pt2pt(GridView g, CompactArgs a, const float* __restrict__ lx,
                    const float* __restrict__ ly, const float* __restrict__ lz,
                    const uint32_t* __restrict__ gbits, const unsigned long long* __restrict__ claim,
                    const unsigned long long* __restrict__ cand, const float4* __restrict__ cand_xyz,
                    const uint32_t* __restrict__ bbox, uint32_t* __restrict__ bbox_next,
                    unsigned long long* __restrict__ status, uint32_t* __restrict__ tile_counter,
                    mp2p_b200_pair_pt2pt* __restrict__ out, unsigned long long* __restrict__ out_count,
                    FusedSums fs)
{
    static_assert(kScanThreads == kReduceThreads, "the sums use the solvers' block reduction");
    constexpr uint32_t TILE = kScanThreads * kCompactItems2p;
    __shared__ ScanSmem sm;
    __shared__ uint32_t s_rec[TILE * 9];  // this tile's records, staged for coalesced stores
    bbox_rearm(bbox_next);
    const uint64_t n_slots = (uint64_t)a.n_local * a.K;
    const uint32_t n_tiles = (uint32_t)((n_slots + TILE - 1) / TILE);
    if (threadIdx.x == 0)
    {
        sm.tile_id = atomicAdd(tile_counter, 1u);
        if (sm.tile_id == n_tiles - 1) *tile_counter = 0u;  // last ticket handed out: re-arm
    }
    __syncthreads();
    compact_pt2pt_body<kCompactItems2p>(g, a, lx, ly, lz, gbits, claim, cand, cand_xyz, bbox, status, out, out_count, fs, sm.tile_id,
                                        sm, s_rec);
}

// ------------------------------------------------------------------------------------------
// One ICP iteration (Matcher_Points_DistanceThreshold, k = 1, + Solver_Horn without robust kernel /
// weights / scale-outlier pass) in ONE cooperative launch of co-resident CTAs:
//   phase 1  nn1_body: search, candidate words, first-claim proposals, bounding box
//   -------- grid barrier (claims and the box are complete)
//   phase 2  compact_pt2pt_body: acceptance, look-back scan, records out, HORN1 sums -> packet[0..32)
//   -------- every CTA waits for the folded sums (flag written by the folding CTA)
//   phase 3  S moments of the tile's records straight from shared memory -> packet[32..64)
// Saves two launches, the solver's re-read of the pairings and the launch gaps; the numbers are the
// same sums in a different grouping (poses agree to ~1e-15 with the three-kernel path).
// ------------------------------------------------------------------------------------------
// SHARDED (one process per GPU, csrc/peer.cuh): the same launch also carries the exchanges of a
// query-sharded iteration through the peers' mailboxes over NVLink —
//   phase 1   the shard's proposals claim directly under their GLOBAL slot numbers; candidate words
//             go into this rank's record slot of its own mailbox
//   -------- grid barrier 0
//   phase 1b  every CTA stores a slice of the record (words, padding, bounding box) into EVERY peer's
//             mailbox; grid barrier 1; CTA 0 releases the record flags; every CTA acquires the peers'
//   phase 2a  replay of the OTHER shards' proposals into the claim array, fold of the bounding boxes
//   -------- grid barrier 2
//   phase 2   compaction (scan barrier 3) + HORN1 sums; the folding CTA all-reduces the packet with the
//             peers before it publishes it
//   phase 3   moments; the folding CTA all-reduces them and hands both packets to the host
// so a sharded pt2pt + Horn iteration is ONE launch and no collective call.
template <bool SHARDED>
__global__ void __launch_bounds__(kScanThreads, 3)
    k_iterate_nn1_horn(GridView g, Pt2PtArgs a, CompactArgs ca, const float* __restrict__ qx, const float* __restrict__ qy,
                       const float* __restrict__ qz, const float* __restrict__ lx, const float* __restrict__ ly,
                       const float* __restrict__ lz, const uint32_t* __restrict__ perm, unsigned long long* __restrict__ claim,
                       unsigned long long* __restrict__ cand, float4* __restrict__ cand_xyz, uint32_t* __restrict__ bbox,
                       uint32_t* __restrict__ bbox_next, unsigned long long* __restrict__ status,
                       mp2p_b200_pair_pt2pt* __restrict__ out, unsigned long long* __restrict__ out_count, FusedSums fs,
                       double* __restrict__ mom_partials, unsigned int* __restrict__ mom_ticket, double w_pt2pt, CoopSync cs,
                       PeerLaunch pl, unsigned long long per_k, uint32_t* __restrict__ cloud_bbox)
{
    __shared__ ScanSmem sm;
    __shared__ uint32_t s_rec[kScanThreads * 9];
    const PeerView& pv = pl.view;
    if (SHARDED)  // the candidate words of this shard live in its record slot of the own mailbox
        cand = reinterpret_cast<unsigned long long*>(pv.box[pv.rank] + rec_offset(pv.rec_words, pv.world, pl.rec_epoch & 1u, pv.rank));
    // ---- phase 1
    nn1_body<kScanThreads, MP2P_NN1_FUSED_R, MP2P_NN1_FUSED_W, MP2P_NN1_FUSED_CW>(g, a, qx, qy, qz, perm, nullptr, nullptr, claim, cand, cand_xyz, bbox, nullptr);
    grid_barrier(cs, 0);
    const uint32_t* gate_box = bbox;
    if (SHARDED)
    {
        const uint32_t parity = pl.rec_epoch & 1u;
        // ---- phase 1b: record = [per_k words | 6 bbox words | pad] into every mailbox
        for (unsigned long long j = (unsigned long long)blockIdx.x * kScanThreads + threadIdx.x; j < pv.rec_words;
             j += (unsigned long long)gridDim.x * kScanThreads)
        {
            unsigned long long w = 0ull;
            if (j < a.n_local)
                w = __ldcg(cand + j);
            else if (j < per_k)
                w = ~0ull;  // slots of a short shard: "no candidate"
            else if (j < per_k + 3)
                w = (unsigned long long)__ldcg(bbox + 2 * (j - per_k)) | ((unsigned long long)__ldcg(bbox + 2 * (j - per_k) + 1) << 32);
            if (j >= a.n_local) cand[j] = w;
            for (uint32_t p = 1; p < pv.world; p++)
            {
                const uint32_t dst = (pv.rank + p) % pv.world;
                reinterpret_cast<unsigned long long*>(pv.box[dst] + rec_offset(pv.rec_words, pv.world, parity, pv.rank))[j] = w;
            }
        }
        __threadfence_system();
        grid_barrier(cs, 1);
        if (blockIdx.x == 0 && threadIdx.x < pv.world) st_release_sys(rec_flag(pv.box[threadIdx.x], parity, pv.rank), pl.rec_epoch);
        if (threadIdx.x < pv.world) wait_flag(rec_flag(pv.box[pv.rank], parity, threadIdx.x), pl.rec_epoch);
        __syncthreads();
        // ---- phase 2a: the whole cloud's bounding box (every CTA writes the same six words), the
        // other shards' proposals
        const unsigned long long* recs =
            reinterpret_cast<const unsigned long long*>(pv.box[pv.rank] + rec_offset(pv.rec_words, pv.world, parity, 0));
        if (threadIdx.x < 6)
        {
            const int d = threadIdx.x;
            uint32_t  r = d < 3 ? 0xFFFFFFFFu : 0u;
            for (uint32_t p = 0; p < pv.world; p++)
            {
                const uint32_t v = __ldcg(reinterpret_cast<const uint32_t*>(recs + (size_t)p * pv.rec_words + per_k) + d);
                r                = d < 3 ? min(r, v) : max(r, v);
            }
            cloud_bbox[d] = r;
        }
        if (!ca.allowGlobal)
        {
            const unsigned long long n_all = per_k * pv.world;
            for (unsigned long long s = (unsigned long long)blockIdx.x * kScanThreads + threadIdx.x; s < n_all;
                 s += (unsigned long long)gridDim.x * kScanThreads)
            {
                const unsigned long long r = s / per_k, j = s - r * per_k;
                if (r == pv.rank) continue;  // claimed in phase 1
                const unsigned long long c = __ldcg(recs + r * pv.rec_words + j);
                if ((uint32_t)c == 0xFFFFFFFFu) continue;
                atomicMin(claim + (uint32_t)c, ca.tag | s);
            }
        }
        __threadfence();
        grid_barrier(cs, 2);
        gate_box = cloud_bbox;
    }
    // ---- phase 2 (tile = CTA index: all CTAs are resident, the look-back cannot starve)
    bbox_rearm(bbox_next);
    bool           folded = false;  // CTA-uniform
    const uint32_t n_rec  = compact_pt2pt_body(g, ca, lx, ly, lz, nullptr, claim, cand, cand_xyz, gate_box, status, out, out_count, fs,
                                               blockIdx.x, sm, s_rec, &folded, &cs);
    // ---- the CTA that folded the HORN1 packet publishes it; everybody waits for it
    __syncthreads();
    if (SHARDED && folded && threadIdx.x < 32) peer_allreduce_warp(pv, pl.pkt_epoch, fs.packet, threadIdx.x);
    __syncthreads();
    if (threadIdx.x == 0)
    {
        if (folded)
        {
            __threadfence();
            atomicExch(cs.sums_flag, cs.epoch);
        }
        while (*reinterpret_cast<volatile unsigned int*>(cs.sums_flag) != cs.epoch) __nanosleep(32);
        __threadfence();
    }
    __syncthreads();
    // ---- phase 3: visit_correspondences (visit_correspondences.h:100-212) over the tile's records
    const double* sums = fs.packet;
    const double  wc   = 1.0 / __ldcg(sums + 6);
    const double  cl[3] = {__ldcg(sums + 0) * wc, __ldcg(sums + 1) * wc, __ldcg(sums + 2) * wc};
    const double  cg[3] = {__ldcg(sums + 3) * wc, __ldcg(sums + 4) * wc, __ldcg(sums + 5) * wc};
    const double  waPoints = w_pt2pt / (w_pt2pt * __ldcg(sums + 7));  // :85-87
    double        acc[12];
#pragma unroll
    for (int v = 0; v < 12; v++) acc[v] = 0;
    if (threadIdx.x < n_rec)
    {
        const uint32_t* rec = s_rec + threadIdx.x * 9;
        const double    bi[3] = {(double)__uint_as_float(rec[2]) - cg[0], (double)__uint_as_float(rec[3]) - cg[1],
                                 (double)__uint_as_float(rec[4]) - cg[2]};
        const double    ri[3] = {(double)__uint_as_float(rec[5]) - cl[0], (double)__uint_as_float(rec[6]) - cl[1],
                                 (double)__uint_as_float(rec[7]) - cl[2]};
        const double bn = sqrt(bi[0] * bi[0] + bi[1] * bi[1] + bi[2] * bi[2]);
        const double rn = sqrt(ri[0] * ri[0] + ri[1] * ri[1] + ri[2] * ri[2]);
        if (!(bn < 1e-4 || rn < 1e-4))  // :141-146
        {
            acc[9] += waPoints;
            acc[11] += 1.0;
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int c = 0; c < 3; c++) acc[3 * r + c] += waPoints * ri[r] * bi[c];  // S += w r b^T
        }
    }
    const bool last = block_reduce_to_packet<12>(acc, mom_partials, mom_ticket, fs.packet + MP2P_B200_PACKET_DOUBLES,
                                                 blockIdx.x, gridDim.x);
    if (SHARDED && last)
    {
        __syncthreads();
        if (threadIdx.x < 32) peer_allreduce_warp(pv, pl.pkt_epoch + 1u, fs.packet + MP2P_B200_PACKET_DOUBLES, threadIdx.x);
    }
    if (last && cs.host_out)
    {
        // the CTA that folded the moments hands both packets to the host through mapped pinned memory
        // and raises the flag: the caller polls it instead of paying a DMA copy + stream synchronise
        __syncthreads();
        if (threadIdx.x < 2 * MP2P_B200_PACKET_DOUBLES)
            reinterpret_cast<volatile double*>(cs.host_out)[threadIdx.x] = __ldcg(fs.packet + threadIdx.x);
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0)
        {
            *reinterpret_cast<volatile unsigned int*>(cs.host_out + 2 * MP2P_B200_PACKET_DOUBLES) = cs.epoch;
            __threadfence_system();
        }
    }
}

// ------------------------------------------------------------------------------------------
// pt2pl
// ------------------------------------------------------------------------------------------
struct Pt2PlArgs
{
    PoseArg  pose;
    float    radiusSq, distThr;
    double   planeEigenThreshold;
    uint32_t n_local, K, minPts;
    int      allowLocal;
    float    gate_eps;
    uint64_t capacity;
    int      tma_ok;
    // Matcher_Point2Line: line fit instead of plane fit; planeEigenThreshold = lineEigenThreshold,
    // minPts = minimumLinePoints counted over the neighbours with d2 <= lineMaxSqr
    int      line_mode;
    float    lineMaxSqr;
};

// Plane fit of the pt2pl matcher, one THREAD per LISTED query (FitList: the k-NN search listed the
// queries whose neighbourhood holds at least max(3, minimumPlanePoints) points; the grid is sized
// for the worst case and CTAs past the list's end leave at once). Full warps in the fp64 Jacobi
// instead of the ~half of the lanes whose query qualifies. Per query: gather the K neighbour points
// (ascending (d2, index)), estimate_points_eigen + planarity + distance tests (plane_fit.cuh).
constexpr int kFitThreads = 128;

template <int KT>
__global__ void __launch_bounds__(kFitThreads)
    k_plane_fit(GridView g, Pt2PlArgs a, const float* __restrict__ qx, const float* __restrict__ qy,
                const float* __restrict__ qz, const uint32_t* __restrict__ perm,
                const unsigned long long* __restrict__ cand, const uint32_t* __restrict__ fit_list,
                const uint32_t* __restrict__ fit_count, PlaneCandidate* __restrict__ plc,
                uint8_t* __restrict__ ok_flags)
{
    const uint32_t t = blockIdx.x * kFitThreads + threadIdx.x;
    if (t >= *fit_count) return;
    const int      K    = (int)a.K;
    const uint32_t qpos = fit_list[t];  // position in the (possibly Morton-sorted) array the search walked
    const uint32_t i    = perm ? __ldg(perm + qpos) : qpos;
    int            cnt  = 0, within = 0;
    uint32_t       idx[KT];
#pragma unroll
    for (int k = 0; k < KT; k++)
        if (k < K)
        {
            const unsigned long long c = cand[(size_t)qpos * K + k];
            idx[k]                     = (uint32_t)c;
            cnt += ((uint32_t)c != 0xFFFFFFFFu);  // valid ranks come first
            // Matcher_Point2Line.cpp:110-127: the list is cut at the first distance > threshold^2
            if (a.line_mode && (uint32_t)c != 0xFFFFFFFFu && within == k && !(__uint_as_float((uint32_t)(c >> 32)) > a.lineMaxSqr)) within++;
        }
    float px[KT], py[KT], pz[KT];
#pragma unroll
    for (int k = 0; k < KT; k++)
        if (k < cnt)
        {
            const float4 p = __ldg(g.pts_orig + idx[k]);
            px[k] = p.x, py[k] = p.y, pz[k] = p.z;
        }
    float gx, gy, gz;
    compose_point_f(a.pose, qx[qpos], qy[qpos], qz[qpos], gx, gy, gz);
    PlaneCandidate pc;
    uint8_t        ok = 0;
    // line mode: >= minimumLinePoints neighbours within the threshold (:130), PCA over ALL cnt (:132-135)
    if (a.line_mode ? (within >= (int)a.minPts && fit_line<KT>(px, py, pz, cnt, a.planeEigenThreshold, pc))
                    : fit_plane<KT>(px, py, pz, cnt, gx, gy, gz, a.planeEigenThreshold, a.distThr, pc))
    {
        plc[i] = pc;
        ok     = 1;
    }
    ok_flags[i] = ok;
}

__global__ void __launch_bounds__(kScanThreads)
    k_compact_pt2pl(GridView g, uint32_t n_local, float gate_eps, uint64_t capacity,
                    const float* __restrict__ lx, const float* __restrict__ ly,
                    const float* __restrict__ lz, const PlaneCandidate* __restrict__ plc,
                    const uint8_t* __restrict__ ok_flags, const uint32_t* __restrict__ bbox,
                    uint32_t* __restrict__ bbox_next, unsigned long long* __restrict__ status,
                    uint32_t* __restrict__ tile_counter,
                    mp2p_b200_pair_pt2pl* __restrict__ out, unsigned long long* __restrict__ out_count,
                    uint32_t scan_epoch, int line_mode, uint32_t* __restrict__ out_host,
                    unsigned long long* __restrict__ count_host)
{
    constexpr int      ITEMS = kCompactItems2l;
    constexpr uint32_t TILE  = kScanThreads * ITEMS;
    __shared__ ScanSmem sm;
    __shared__ uint32_t s_rec[TILE * 18];  // this tile's 72-byte records, staged for coalesced stores
    bbox_rearm(bbox_next);
    const uint32_t n_tiles = (n_local + TILE - 1) / TILE;
    if (threadIdx.x == 0)
    {
        sm.tile_id = atomicAdd(tile_counter, 1u);
        if (sm.tile_id == n_tiles - 1) *tile_counter = 0u;
    }
    __syncthreads();
    const uint32_t tile = sm.tile_id;
    const bool     gate = bbox_gate(g, bbox, gate_eps);
    const uint32_t i0   = tile * TILE + threadIdx.x * ITEMS;  // the thread's consecutive local points
    // everything the records need is requested before the scan (one dependent round trip)
    bool           ok[ITEMS];
    PlaneCandidate pc[ITEMS];
    float          px[ITEMS], py[ITEMS], pz[ITEMS];
    uint32_t       local = 0;
#pragma unroll
    for (int j = 0; j < ITEMS; j++)
    {
        const uint32_t i = i0 + j;
        ok[j] = gate && i < n_local && ok_flags[i];
        pc[j] = PlaneCandidate{}, px[j] = py[j] = pz[j] = 0.f;
        if (ok[j]) pc[j] = plc[i], px[j] = lx[i], py[j] = ly[i], pz[j] = lz[i];  // ORIGINAL local point (Matcher_Point2Plane.cpp:105)
        local += ok[j] ? 1u : 0u;
    }
    const unsigned long long w = grid_exclusive_scan(sm, tile, n_tiles, local, status, out_count, scan_epoch);
    const unsigned long long tile_base = sm.tile_base;
    uint32_t                 at        = (uint32_t)(w - tile_base);
#pragma unroll
    for (int j = 0; j < ITEMS; j++)
        if (ok[j])
        {
            double* o = reinterpret_cast<double*>(s_rec + at * 18);
            at++;
            if (line_mode)  // point_line_pair_t: TLine3D {pBase, director}, TPoint3D pt_local (Pairings.h:61-73; Matcher_Point2Line.cpp:152)
            {
                o[0] = pc[j].centroid[0], o[1] = pc[j].centroid[1], o[2] = pc[j].centroid[2];
                o[3] = pc[j].coefs[0], o[4] = pc[j].coefs[1], o[5] = pc[j].coefs[2];
                o[6] = px[j], o[7] = py[j], o[8] = pz[j];
            }
            else  // point_plane_pair_t: plane_patch_t {TPlane coefs[4], TPoint3D centroid}, TPoint3Df pt_local, pad
            {
                o[0] = pc[j].coefs[0], o[1] = pc[j].coefs[1], o[2] = pc[j].coefs[2], o[3] = pc[j].coefs[3];
                o[4] = pc[j].centroid[0], o[5] = pc[j].centroid[1], o[6] = pc[j].centroid[2];
                uint32_t* t = reinterpret_cast<uint32_t*>(o + 7);
                t[0] = __float_as_uint(px[j]), t[1] = __float_as_uint(py[j]);
                t[2] = __float_as_uint(pz[j]), t[3] = 0u;
            }
        }
    __syncthreads();
    const unsigned long long room  = capacity > tile_base ? capacity - tile_base : 0ull;
    const uint32_t           n_rec = (uint32_t)min((unsigned long long)sm.tile_total, room);
    {
        uint32_t* dst = reinterpret_cast<uint32_t*>(out) + tile_base * 18;
        for (uint32_t k = threadIdx.x; k < n_rec * 18; k += kScanThreads) dst[k] = s_rec[k];
    }
    if (out_host)  // zero-copy output: the caller's pinned buffer receives the records, pinned memory the count
    {
        uint32_t* dst = out_host + tile_base * 18;
        for (uint32_t k = threadIdx.x; k < n_rec * 18; k += kScanThreads) dst[k] = s_rec[k];
        if (tile == n_tiles - 1 && threadIdx.x == 0) *count_host = tile_base + sm.tile_total;
    }
}

// ------------------------------------------------------------------------------------------
// raw k-NN (already transformed queries) — nn_* parity tests
// ------------------------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(256)
    k_knn(GridView g, const float* __restrict__ qx, const float* __restrict__ qy,
          const float* __restrict__ qz, uint32_t nq, uint32_t K, float radius2, int rl_start, int n_phases,
          uint32_t* __restrict__ out_idx, float* __restrict__ out_d2, int32_t* __restrict__ out_found)
{
    __shared__ KnnShared<G> ks;
    knn_shared_init(ks);
    __syncthreads();
    const int      sub   = threadIdx.x % G;
    const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << (G & 31)) - 1u)) << ((threadIdx.x & 31) / G * G);
    const uint32_t i     = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const bool     have  = i < nq;
    unsigned long long mine;
    SearchCounters     sc;
    knn_search<G>(g, have, have ? qx[i] : 0.f, have ? qy[i] : 0.f, have ? qz[i] : 0.f, radius2, (int)K, rl_start, mine, sub, sc, ks.tb, ks.nb, n_phases);
    if (!have) return;  // whole groups leave together
    const unsigned long long sentinel = (unsigned long long)__float_as_uint(radius2) << 32;
    const bool               f        = sub < (int)K && mine < sentinel;
    if (sub < (int)K)
    {
        out_idx[(size_t)i * K + sub] = f ? (uint32_t)mine : 0u;
        out_d2[(size_t)i * K + sub]  = f ? __uint_as_float((uint32_t)(mine >> 32)) : cuda::std::numeric_limits<float>::infinity();
    }
    const unsigned found = __ballot_sync(gmask, f) & gmask;
    if (sub == 0) out_found[i] = __popc(found);
}

// lanes per query of the group-cooperative kernels: the smallest supported group that holds k keys
int group_size(uint32_t K) { return K <= 8 ? 8 : (K <= 16 ? 16 : 32); }

int pick_kt(uint32_t K)
{
    if (K <= 1) return 1;
    if (K <= 4) return 4;
    if (K <= 8) return 8;
    if (K <= 16) return 16;
    return 32;
}

// group-cooperative kernels are instantiated for G = 8, 16, 32 lanes per query (group_size(k))
#define MP2P_DISPATCH_G(K, CALL)  \
    switch (group_size(K))        \
    {                             \
        case 8: CALL(8) break;   \
        case 16: CALL(16) break; \
        default: CALL(32) break; \
    }

// finest table whose voxels hold at least ~0.75 k points on average (k = 1: the finest table)
// odd stride near 0.618 n, coprime to n: b -> (b * stride) % n is a permutation of the tiles
uint32_t tile_stride_for(uint64_t n_tiles)
{
    static const bool off = [] {
        const char* e = getenv("MP2P_TILE_STRIDE");
        return e && atoi(e) == 0;
    }();
    if (off || n_tiles < 8) return 1;
    uint64_t s = (uint64_t)((double)n_tiles * 0.6180339887) | 1ull;
    auto     gcd = [](uint64_t a, uint64_t b) {
        while (b)
        {
            const uint64_t t = a % b;
            a = b, b = t;
        }
        return a;
    };
    while (gcd(s, n_tiles) != 1) s += 2;
    return (uint32_t)(s % n_tiles);
}
// measurement knobs of the k > 1 search (A/B on the device): MP2P_KNN_PHASES = 2 | 3, MP2P_KNN_V1 = 1
int knn_phases()
{
    static const int v = [] {
        const char* e = getenv("MP2P_KNN_PHASES");
        const int   x = e ? atoi(e) : 3;
        return x == 2 ? 2 : 3;
    }();
    return v;
}
bool knn_v1()
{
    static const bool v = [] {
        const char* e = getenv("MP2P_KNN_V1");
        return e && atoi(e) == 1;
    }();
    return v;
}
// CTA size of the k > 1 search (threads; G lanes per query): $MP2P_KNN_NT = 64 | 128 | 256
int knn_nt()
{
    static const int v = [] {
        const char* e = getenv("MP2P_KNN_NT");
        const int   x = e ? atoi(e) : 256;
        return (x == 64 || x == 128) ? x : 256;
    }();
    return v;
}
// Scheduling hint of a search over a RESIDENT cloud (knn_tile_hint): the kernel reports per-tile durations
// into the cloud handle, k_tile_rank turns them into the next call's tile order (longest first). The pose
// moves little between ICP iterations, so the expensive tiles of one call — queries that climb through empty
// space — are the expensive tiles of the next; started first they overlap with everything else instead of
// forming the tail of the launch. $MP2P_KNN_LPT=0 switches the hint off.
bool knn_lpt()
{
    static const bool v = [] {
        const char* e = getenv("MP2P_KNN_LPT");
        return !(e && atoi(e) == 0);
    }();
    return v;
}
int knn_tile_hint(mp2p_b200_ctx* ctx, uint32_t n_tiles, int G, int NT, Pt2PtArgs& a)
{
    a.tile_order = nullptr, a.tile_cost = nullptr, a.tile_trace = nullptr;
    mp2p_b200_cloud* c = ctx->cur_cloud;
    static const char* trace_path = getenv("MP2P_KNN_TRACE");
    if (trace_path && c)
    {
        MP2P_TRY(ctx->d_trace.ensure((size_t)n_tiles * 32));
        a.tile_trace = ctx->d_trace.as<uint32_t>(), ctx->trace_tiles = n_tiles;
    }
    if (!c || !knn_lpt() || !ctx->aux_stream || n_tiles < 2 * 148) return 0;  // (a launch of less than two CTAs per SM has no tail to hide)
    const uint64_t key = ((uint64_t)n_tiles << 20) | ((uint64_t)G << 12) | (uint64_t)NT;
    if (c->hint_key == key)
    {
        // the order was made on the side stream behind the previous search: this search waits for it
        a.tile_order = c->d_tile_order.as<uint32_t>();
        MP2P_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->ev_rank, 0));
    }
    else
    {
        if (c->hint_key) MP2P_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->ev_rank, 0));  // a rank kernel may still read / clear the arrays
        MP2P_TRY(c->d_tile_cost.ensure((size_t)n_tiles * 4));
        MP2P_TRY(c->d_tile_order.ensure((size_t)n_tiles * 4));
        MP2P_CUDA_TRY(cudaMemsetAsync(c->d_tile_cost.p, 0, (size_t)n_tiles * 4, ctx->stream));  // later calls: k_tile_rank re-arms it
    }
    c->hint_key = key;
    a.tile_cost = c->d_tile_cost.as<uint32_t>();
    return 0;
}
// behind the search, on the side stream: the next call's tile order (nothing of this call waits for it)
int knn_tile_rank(mp2p_b200_ctx* ctx, uint32_t n_tiles, const Pt2PtArgs& a)
{
    if (!a.tile_cost) return 0;
    MP2P_CUDA_TRY(cudaEventRecord(ctx->ev_rank_fork, ctx->stream));
    MP2P_CUDA_TRY(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_rank_fork, 0));
    k_tile_rank<<<1, 1024, 0, ctx->aux_stream>>>(a.tile_cost, n_tiles, ctx->cur_cloud->d_tile_order.as<uint32_t>());
    MP2P_CUDA_TRY(cudaEventRecord(ctx->ev_rank, ctx->aux_stream));
    count_launch(ctx);
    return 0;
}
// k <= kThreadKMax: one thread per query, A/B variant ($MP2P_KNN_THREAD=1; $MP2P_KNN_TNT = 64 | 128 | 256 queries per
// CTA). NOT the default: measured on C3 (profiles/r02_knn_ab.txt) it executes 2.3x fewer instructions than the group
// search and is still slower — a 119k-query scan is 3.7k warps, 25 per SM, one wave, and every SM waits for its
// slowest warp; with G = 8 lanes per query the same scan is 30k warps and the tail hides behind the bulk.
bool knn_thread_mode()
{
    static const bool v = [] {
        const char* e = getenv("MP2P_KNN_THREAD");
        return e && atoi(e) == 1;
    }();
    return v && !knn_v1();
}
int knn_thread_nt()
{
    static const int v = [] {
        const char* e = getenv("MP2P_KNN_TNT");
        const int   x = e ? atoi(e) : 128;
        return (x == 64 || x == 256) ? x : 128;
    }();
    return v;
}
template <int KM, int NT>
int launch_knn_thread_nt(mp2p_b200_ctx* ctx, size_t nq, cudaStream_t st, const GridView& g, Pt2PtArgs& a, const float* lx, const float* ly,
                         const float* lz, const uint32_t* perm, const uint32_t* lbits, const uint32_t* gbits, unsigned long long* claim,
                         unsigned long long* cand, uint32_t* bbox_words, unsigned long long* stats, const FitList& fit)
{
    const uint32_t nb = (uint32_t)((nq + NT - 1) / NT);
    a.tile_stride     = tile_stride_for(nb);
    MP2P_TRY(knn_tile_hint(ctx, nb, 1, NT, a));
    // over-budget queries: two counters used in turn (this call's, cleared by the previous call's second kernel)
    static const uint32_t max_probes = [] { const char* e = getenv("MP2P_KNN_DEFER_PROBES"); return e ? (uint32_t)atoi(e) : 24u; }();
    static const uint32_t max_cands  = [] { const char* e = getenv("MP2P_KNN_DEFER_CANDS"); return e ? (uint32_t)atoi(e) : 160u; }();
    const bool     defer = max_probes && max_cands && nq >= 4096;
    const uint32_t cap   = (uint32_t)(nq / 8);
    a.defer = DeferList{nullptr, nullptr, 0, 0, 0}, a.qlist = nullptr, a.qlist_count = nullptr, a.qlist_cap = 0, a.defer_reset = nullptr;
    uint32_t* cnt = nullptr;
    if (defer)
    {
        const bool fresh = ctx->d_defer.bytes < (size_t)(cap + 2) * 4;
        MP2P_TRY(ctx->d_defer.ensure((size_t)(cap + 2) * 4));
        cnt = ctx->d_defer.as<uint32_t>();
        if (fresh) MP2P_CUDA_TRY(cudaMemsetAsync(cnt, 0, 8, st));
        ctx->defer_turn ^= 1u;
        a.defer = DeferList{cnt + 2, cnt + ctx->defer_turn, cap, max_probes, max_cands};
    }
    k_match_knn_thread<KM, NT><<<nb, NT, 0, st>>>(g, a, lx, ly, lz, perm, lbits, gbits, claim, cand, bbox_words, stats, fit);
    if (defer)
    {
        // the deferred queries, one WARP per query (the group search with G = 32: probes, candidates and list are
        // spread over the lanes), eight queries per CTA
        Pt2PtArgs b = a;
        b.defer = DeferList{nullptr, nullptr, 0, 0, 0}, b.tile_order = nullptr, b.tile_cost = nullptr, b.tile_trace = nullptr;
        b.qlist = cnt + 2, b.qlist_count = cnt + ctx->defer_turn, b.qlist_cap = cap, b.defer_reset = cnt + (ctx->defer_turn ^ 1u);
        k_match_pt2pt<32, false, 256><<<(cap + 7) / 8, 256, 0, st>>>(g, b, lx, ly, lz, perm, lbits, gbits, claim, cand, bbox_words, stats, fit);
        count_launch(ctx);
    }
    MP2P_TRY(knn_tile_rank(ctx, nb, a));
    return 0;
}
template <int KM, class... Args>
int launch_knn_thread_km(mp2p_b200_ctx* ctx, size_t nq, cudaStream_t st, Args&&... args)
{
    switch (knn_thread_nt())
    {
        case 64: return launch_knn_thread_nt<KM, 64>(ctx, nq, st, args...);
        case 256: return launch_knn_thread_nt<KM, 256>(ctx, nq, st, args...);
        default: return launch_knn_thread_nt<KM, 128>(ctx, nq, st, args...);
    }
}
template <class... Args>
int launch_knn_thread(mp2p_b200_ctx* ctx, uint32_t K, size_t nq, cudaStream_t st, Args&&... args)
{
    if (K <= 4) return launch_knn_thread_km<4>(ctx, nq, st, args...);
    if (K <= 8) return launch_knn_thread_km<8>(ctx, nq, st, args...);
    if (K <= 12) return launch_knn_thread_km<12>(ctx, nq, st, args...);
    if (K <= 16) return launch_knn_thread_km<16>(ctx, nq, st, args...);
    return launch_knn_thread_km<20>(ctx, nq, st, args...);
}
#define MP2P_LAUNCH_KMATCH_NT(G, V1, NT, nq, st, ...)                                              \
    {                                                                                              \
        const uint32_t nb_ = (uint32_t)(((uint64_t)(nq) * G + NT - 1) / NT);                       \
        a_.tile_stride     = tile_stride_for(nb_);                                                 \
        MP2P_TRY(knn_tile_hint(ctx, nb_, G, NT, a_));                                              \
        k_match_pt2pt<G, V1, NT><<<nb_, NT, 0, st>>>(__VA_ARGS__);                                 \
        MP2P_TRY(knn_tile_rank(ctx, nb_, a_));                                                     \
    }
// `args` = the Pt2PtArgs lvalue passed in __VA_ARGS__ (its tile_stride is set here, per grid size)
#define MP2P_LAUNCH_KMATCH(G, args, nq, st, ...)                                                   \
    {                                                                                              \
        auto& a_ = args;                                                                           \
        if (knn_thread_mode() && a_.K <= (uint32_t)kThreadKMax)                                    \
        {                                                                                          \
            MP2P_TRY(launch_knn_thread(ctx, a_.K, nq, st, __VA_ARGS__));                           \
        }                                                                                          \
        else if (knn_v1())                                                                         \
        {                                                                                          \
            if (knn_nt() == 64) MP2P_LAUNCH_KMATCH_NT(G, true, 64, nq, st, __VA_ARGS__)            \
            else if (knn_nt() == 128) MP2P_LAUNCH_KMATCH_NT(G, true, 128, nq, st, __VA_ARGS__)     \
            else MP2P_LAUNCH_KMATCH_NT(G, true, 256, nq, st, __VA_ARGS__)                          \
        }                                                                                          \
        else                                                                                       \
        {                                                                                          \
            if (knn_nt() == 64) MP2P_LAUNCH_KMATCH_NT(G, false, 64, nq, st, __VA_ARGS__)           \
            else if (knn_nt() == 128) MP2P_LAUNCH_KMATCH_NT(G, false, 128, nq, st, __VA_ARGS__)    \
            else MP2P_LAUNCH_KMATCH_NT(G, false, 256, nq, st, __VA_ARGS__)                         \
        }                                                                                          \
    }
int start_level(const GridView& v, uint32_t K)
{
    if (K <= 1) return 0;
    static const float factor = [] {  // tuning knob (measurement only): MP2P_START_OCC, default 0.75
        const char* e = getenv("MP2P_START_OCC");
        const float f = e ? (float)atof(e) : 0.f;
        return f > 0.01f ? f : 0.75f;
    }();
    for (int rl = 0; rl < v.n_levels; rl++)
        if (v.level_occupancy[rl] >= factor * (float)K) return rl;
    return v.n_levels - 1;
}

// Host clouds are copied into (aligned) staging arrays; device-resident clouds are used in place.
// Sets ctx->cur_l{x,y,z} (what the kernels read) and ctx->cur_tma_ok.
// Device alias of a pinned (page-locked, mapped) host pointer, NULL for anything else. Zero copy can
// be switched off with MP2P_ZERO_COPY=0.
template <class T>
T* mapped_alias(T* host)
{
    static int enabled = -1;
    if (enabled < 0)
    {
        const char* e = getenv("MP2P_ZERO_COPY");
        enabled       = (e && atoi(e) == 0) ? 0 : 1;
    }
    if (!enabled || !host) return nullptr;
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess)
    {
        cudaGetLastError();
        return nullptr;
    }
    // (with unified addressing, page-locked host memory is portable and mapped for every device of the
    // process, whichever device was current when it was allocated)
    if (at.type != cudaMemoryTypeHost || !at.devicePointer) return nullptr;
    return static_cast<T*>(at.devicePointer);
}

// mapped_ok (optional, out): the caller can let its search kernel read pinned host arrays in place
// (then cur_q* = the host arrays' device aliases, cur_l* = staging arrays the kernel must fill)
int stage_local(mp2p_b200_ctx* ctx, const float* lx, const float* ly, const float* lz, uint64_t n,
                int on_device, bool* mapped_ok = nullptr)
{
    ctx->cur_perm = nullptr, ctx->cur_cloud = nullptr;
    if (mapped_ok && on_device != 0) mapped_ok = nullptr;
    if (mapped_ok) *mapped_ok = false;
    if (on_device == 2)  // resident cloud: search kernels walk the Morton-sorted copy
    {
        const auto* c = reinterpret_cast<const mp2p_b200_cloud*>(lx);
        if (!ctx->owns_cloud(c))
        {
            set_error("local cloud handle is not a live cloud of this context (destroyed, or created on another context)");
            return MP2P_B200_ERR_ARG;
        }
        if (c->ctx != ctx || c->n != n)
        {
            set_error("local cloud handle belongs to another context, or n_local differs from its size");
            return MP2P_B200_ERR_ARG;
        }
        ctx->cur_lx = c->d_x.as<float>(), ctx->cur_ly = c->d_y.as<float>(), ctx->cur_lz = c->d_z.as<float>();
        ctx->cur_qx = c->d_sx.as<float>(), ctx->cur_qy = c->d_sy.as<float>(), ctx->cur_qz = c->d_sz.as<float>();
        ctx->cur_perm   = c->d_perm.as<uint32_t>();
        ctx->cur_cloud  = const_cast<mp2p_b200_cloud*>(c);
        ctx->cur_tma_ok = true;
        return 0;
    }
    if (on_device)
    {
        ctx->cur_lx = ctx->cur_qx = lx, ctx->cur_ly = ctx->cur_qy = ly, ctx->cur_lz = ctx->cur_qz = lz;
        ctx->cur_tma_ok = ((reinterpret_cast<uintptr_t>(lx) | reinterpret_cast<uintptr_t>(ly) |
                            reinterpret_cast<uintptr_t>(lz)) & 15u) == 0;
        return 0;
    }
    const size_t bytes = n * sizeof(float);
    MP2P_TRY(ctx->d_lx.ensure(bytes));
    MP2P_TRY(ctx->d_ly.ensure(bytes));
    MP2P_TRY(ctx->d_lz.ensure(bytes));
    if (mapped_ok)
    {
        // pinned host arrays are read by the search kernel itself over PCIe (the transfer overlaps
        // the search); it leaves a device copy in the staging arrays for the compaction
        const float *mx = mapped_alias(lx), *my = mapped_alias(ly), *mz = mapped_alias(lz);
        if (mx && my && mz)
        {
            ctx->cur_qx = mx, ctx->cur_qy = my, ctx->cur_qz = mz;
            ctx->cur_lx = ctx->d_lx.as<float>(), ctx->cur_ly = ctx->d_ly.as<float>(), ctx->cur_lz = ctx->d_lz.as<float>();
            ctx->cur_tma_ok = false;
            *mapped_ok      = true;
            return 0;
        }
    }
    MP2P_TRY(copy_to_device(ctx, ctx->d_lx.p, lx, bytes, ctx->stream));
    MP2P_TRY(copy_to_device(ctx, ctx->d_ly.p, ly, bytes, ctx->stream));
    MP2P_TRY(copy_to_device(ctx, ctx->d_lz.p, lz, bytes, ctx->stream));
    ctx->cur_lx = ctx->cur_qx = ctx->d_lx.as<float>(), ctx->cur_ly = ctx->cur_qy = ctx->d_ly.as<float>();
    ctx->cur_lz = ctx->cur_qz = ctx->d_lz.as<float>();
    ctx->cur_tma_ok = true;
    return 0;
}

int upload_bits(mp2p_b200_ctx* ctx, DevBuf& buf, const uint32_t* bits, uint64_t n_bits,
                const uint32_t** d_out)
{
    *d_out = nullptr;
    if (!bits) return 0;
    const size_t bytes = ((n_bits + 31) / 32) * 4;
    MP2P_TRY(buf.ensure(bytes));
    MP2P_TRY(copy_to_device(ctx, buf.p, bits, bytes, ctx->stream));
    *d_out = buf.as<uint32_t>();
    return 0;
}

struct SmallView
{
    uint32_t*           bbox;       // this call's 6 ordered words (min xyz, max xyz)
    uint32_t*           bbox_next;  // the other slot, re-armed by this call's compaction kernel
    unsigned long long* count;
    uint32_t*           tile_counter;
};
int prepare_small(mp2p_b200_ctx* ctx, uint64_t n_tiles, SmallView& sv, unsigned long long** status)
{
    // counters re-arm themselves inside the kernels and the scan status words carry a call epoch,
    // so nothing is cleared per call: buffers are initialised once when they are (re)allocated.
    if (ctx->d_small.bytes < 128)
    {
        MP2P_TRY(ctx->d_small.ensure(128));
        MP2P_CUDA_TRY(cudaMemsetAsync(ctx->d_small.p, 0, ctx->d_small.bytes, ctx->stream));
        const uint32_t init[12] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0, 0, 0,
                                   0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0, 0, 0};
        MP2P_CUDA_TRY(cudaMemcpyAsync(ctx->d_small.p, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    }
    const size_t need_scan = (n_tiles + 1) * 8;
    if (need_scan > ctx->d_scan.bytes)
    {
        MP2P_TRY(ctx->d_scan.ensure(need_scan));
        MP2P_CUDA_TRY(cudaMemsetAsync(ctx->d_scan.p, 0, ctx->d_scan.bytes, ctx->stream));
    }
    ctx->scan_epoch = (ctx->scan_epoch % 0x3FFFFEu) + 1;  // 1 .. 2^22-2, never 0 (= cleared memory)
    char* base      = ctx->d_small.as<char>();
    sv.bbox         = reinterpret_cast<uint32_t*>(base) + 6 * (ctx->scan_epoch & 1u);
    sv.bbox_next    = reinterpret_cast<uint32_t*>(base) + 6 * ((ctx->scan_epoch & 1u) ^ 1u);
    sv.count        = reinterpret_cast<unsigned long long*>(base + 64);
    sv.tile_counter = reinterpret_cast<uint32_t*>(base + 72);
    *status         = ctx->d_scan.as<unsigned long long>();
    return 0;
}

int prepare_stats(mp2p_b200_ctx* ctx, unsigned long long** stats)
{
    *stats = nullptr;
    if (!ctx->prof_stats) return 0;
    MP2P_TRY(ctx->d_stats.ensure(64));
    MP2P_CUDA_TRY(cudaMemsetAsync(ctx->d_stats.p, 0, 64, ctx->stream));
    *stats = ctx->d_stats.as<unsigned long long>();
    return 0;
}

// Pairings back to the host with ONE synchronisation in the steady state: the record copy is
// issued speculatively right behind the compaction, sized from the previous call's count (+12.5 %),
// together with the count; only if the guess was too small a second copy fetches the remainder.
// pinned host scratch of the speculative solve (h_pinned + 1024): [0..64) Horn packets | GN: pose 12, state
inline double* spec_host(mp2p_b200_ctx* ctx) { return reinterpret_cast<double*>(static_cast<char*>(ctx->h_pinned) + 1024); }

// Enqueues, on the compute stream, the solver the caller is expected to ask for next over the
// pairings this matcher call leaves on the device (see mp2p_b200_ctx::SpecWant). `sums` = HORN1
// packet produced by the compaction (pt2pt only).
template <class Rec>
int enqueue_speculation(mp2p_b200_ctx* ctx, const Rec* d_pairs, const unsigned long long* d_count, uint64_t capacity,
                        const double pose[12], const double* sums)
{
    constexpr bool is2p = sizeof(Rec) == sizeof(mp2p_b200_pair_pt2pt);
    auto&          w    = ctx->spec_want;
    ctx->spec_res.valid = false;
    if (w.kind && ctx->spec_unused >= 2) w.kind = 0;
    double* h = spec_host(ctx);
    if (w.kind == 1 && is2p && sums)
    {
        double* mom = ctx->d_packet.as<double>() + 6 * MP2P_B200_PACKET_DOUBLES;
        MP2P_TRY(run_horn_moments(ctx, reinterpret_cast<const mp2p_b200_pair_pt2pt*>(d_pairs), capacity, &w.horn, sums, capacity,
                                  nullptr, nullptr, 0, nullptr, mom, d_count, 1));
        MP2P_CUDA_TRY(cudaMemcpyAsync(h, sums, MP2P_B200_PACKET_DOUBLES * 8, cudaMemcpyDeviceToHost, ctx->stream));
        MP2P_CUDA_TRY(cudaMemcpyAsync(h + MP2P_B200_PACKET_DOUBLES, mom, MP2P_B200_PACKET_DOUBLES * 8, cudaMemcpyDeviceToHost, ctx->stream));
        ctx->spec_res.kind = 1, ctx->spec_res.list = 1;
    }
    else if (w.kind == 2 && w.list == (is2p ? 1 : 2))
    {
        double*   d_pose  = ctx->d_spec.as<double>();
        uint32_t* d_state = reinterpret_cast<uint32_t*>(ctx->d_spec.as<char>() + 128);
        double*   hp      = h + 64;  // start pose staged in pinned memory, result comes back over it
        std::memcpy(hp, pose, 96);
        MP2P_CUDA_TRY(cudaMemcpyAsync(d_pose, hp, 96, cudaMemcpyHostToDevice, ctx->stream));
        MP2P_TRY(run_gn_device_loop(ctx, is2p ? reinterpret_cast<const mp2p_b200_pair_pt2pt*>(d_pairs) : nullptr, is2p ? capacity : 0,
                                    is2p ? nullptr : reinterpret_cast<const mp2p_b200_pair_pt2pl*>(d_pairs), is2p ? 0 : capacity,
                                    &w.gn, d_pose, d_state, ctx->d_packet.as<double>(), is2p ? d_count : nullptr,
                                    is2p ? nullptr : d_count));
        MP2P_CUDA_TRY(cudaMemcpyAsync(hp, d_pose, 96, cudaMemcpyDeviceToHost, ctx->stream));
        MP2P_CUDA_TRY(cudaMemcpyAsync(hp + 12, d_state, 8, cudaMemcpyDeviceToHost, ctx->stream));
        ctx->spec_res.kind = 2, ctx->spec_res.list = is2p ? 1 : 2;
    }
    else
        return 0;
    std::memcpy(ctx->spec_res.pose_in, pose, 96);
    ctx->spec_res.valid = true;  // n is filled in by fetch_results once the count is known
    ctx->spec_unused++;
    return 0;
}

template <class Rec>
int fetch_results(mp2p_b200_ctx* ctx, const unsigned long long* d_count, const Rec* d_pairs,
                  Rec* out, uint64_t capacity, int out_on_device, uint64_t* out_count, uint64_t* hint,
                  const double* pose = nullptr, const double* sums = nullptr, bool zero_copy = false)
{
    unsigned long long* h_count = static_cast<unsigned long long*>(ctx->h_pinned);
    uint64_t spec = 0;
    ctx->spec_res.valid = false, ctx->spec_res.pending = false;
    if (zero_copy)
    {
        // the compaction kernel stored records and count straight into the caller's pinned memory;
        // the solver the caller is expected to ask for next is enqueued behind it and runs while the
        // caller looks at the pairings
        MP2P_CUDA_TRY(cudaEventRecord(ctx->ev_fork, ctx->stream));
        MP2P_CUDA_TRY(cudaEventSynchronize(ctx->ev_fork));
        MP2P_CUDA_TRY(cudaGetLastError());
        // (enqueued only now: an earlier speculation nobody collected has drained, so its pinned
        // scratch can be reused)
        if (pose && ctx->spec_want.kind)
        {
            MP2P_TRY(enqueue_speculation(ctx, d_pairs, d_count, capacity, pose, sums));
            ctx->spec_res.pending = ctx->spec_res.valid;
        }
        spec = ~0ull;  // nothing left to copy
    }
    else if (!out_on_device && capacity && ctx->copy_stream && pose && ctx->spec_want.kind)
    {
        // records to the host on the copy stream, the expected solver on the compute stream
        spec = *hint == ~0ull ? capacity : std::min<uint64_t>(capacity, *hint + *hint / 8 + 1024);
        MP2P_CUDA_TRY(cudaEventRecord(ctx->ev_fork, ctx->stream));
        MP2P_CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_fork, 0));
        MP2P_CUDA_TRY(cudaMemcpyAsync(h_count, d_count, 8, cudaMemcpyDeviceToHost, ctx->copy_stream));
        if (copy_wants_helpers(ctx, out, spec * sizeof(Rec)))
        {
            // pageable output: the solver is enqueued first, then this thread and the helpers drain the bounce buffer
            MP2P_TRY(enqueue_speculation(ctx, d_pairs, d_count, capacity, pose, sums));
            MP2P_TRY(copy_to_host_sync(ctx, out, d_pairs, spec * sizeof(Rec), ctx->copy_stream));
        }
        else
        {
            MP2P_CUDA_TRY(cudaMemcpyAsync(out, d_pairs, spec * sizeof(Rec), cudaMemcpyDeviceToHost, ctx->copy_stream));
            MP2P_TRY(enqueue_speculation(ctx, d_pairs, d_count, capacity, pose, sums));
            MP2P_CUDA_TRY(cudaStreamSynchronize(ctx->copy_stream));
        }
    }
    else
    {
        MP2P_CUDA_TRY(cudaMemcpyAsync(h_count, d_count, 8, cudaMemcpyDeviceToHost, ctx->stream));
        if (!out_on_device && capacity)
        {
            spec = *hint == ~0ull ? capacity : std::min<uint64_t>(capacity, *hint + *hint / 8 + 1024);
            if (copy_wants_helpers(ctx, out, spec * sizeof(Rec)))
                MP2P_TRY(copy_to_host_sync(ctx, out, d_pairs, spec * sizeof(Rec), ctx->stream));
            else
                MP2P_CUDA_TRY(cudaMemcpyAsync(out, d_pairs, spec * sizeof(Rec), cudaMemcpyDeviceToHost, ctx->stream));
        }
    }
    if (!zero_copy)
    {
        MP2P_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        MP2P_CUDA_TRY(cudaGetLastError());
    }
    ctx->spec_res.n = *h_count;
    const uint64_t cnt = *h_count;
    *out_count         = cnt;
    *hint              = cnt;
    if (!out_on_device)  // the device copy outlives the call: solver calls may name it (PAIRS_LAST_MATCH)
    {
        mp2p_b200_ctx::LastMatch& lm = sizeof(Rec) == sizeof(mp2p_b200_pair_pt2pt) ? ctx->last2p : ctx->last2l;
        lm.dev = d_pairs, lm.n = cnt, lm.valid = cnt <= capacity;
    }
    if (cnt > capacity)
    {
        set_error("output capacity %llu too small for %llu pairings", (unsigned long long)capacity,
                  (unsigned long long)cnt);
        return MP2P_B200_ERR_CAPACITY;
    }
    if (!out_on_device && cnt > spec)
    {
        MP2P_CUDA_TRY(cudaMemcpyAsync(out + spec, d_pairs + spec, (cnt - spec) * sizeof(Rec), cudaMemcpyDeviceToHost, ctx->stream));
        MP2P_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    return 0;
}
}  // namespace

// Launches k_iterate_nn1_horn if the whole grid can be co-resident; returns 1 if it cannot (the
// caller then takes the three-kernel path), 0 after a successful launch. `peer` != NULL: the sharded
// instantiation (exchanges through the peers' mailboxes inside the launch), per_k = per_shard * K.
static int launch_iterate_nn1_horn(mp2p_b200_ctx* ctx, mp2p_b200_map* map, const Pt2PtArgs& a, const CompactArgs& c,
                                   const SmallView& sv, unsigned long long* status, unsigned long long* cand,
                                   float4* cand_xyz, mp2p_b200_pair_pt2pt* d_out, double* d_packets, double w_pt2pt,
                                   uint32_t n_tiles, mp2p_b200_peer* peer = nullptr, unsigned long long per_k = 0)
{
    // co-residency bound, cached per DEVICE and per instantiation (contexts on different GPUs of one process,
    // e.g. the plugin's `device:` parameter, must not share it)
    static int per_dev[64][2], sm_dev[64];
    static bool init_dev[64][2] = {};
    const int  which = peer ? 1 : 0;
    const int  devi  = ctx->device;
    if (devi < 0 || devi >= 64) return 1;
    int* blocks_per_sm = per_dev[devi];
    int& n_sm          = sm_dev[devi];
    if (!init_dev[devi][which])
    {
        init_dev[devi][which] = true;
        int dev = devi, coop = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        blocks_per_sm[which] = 0;
        if (coop)
        {
            if (peer)
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm[which], k_iterate_nn1_horn<true>, kScanThreads, 0);
            else
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm[which], k_iterate_nn1_horn<false>, kScanThreads, 0);
        }
        const char* e = getenv(peer ? "MP2P_FUSED_SHARDED_ITERATION" : "MP2P_FUSED_ITERATION");
        if (e && atoi(e) == 0) blocks_per_sm[which] = 0;
    }
    if ((uint64_t)n_tiles > (uint64_t)blocks_per_sm[which] * (uint64_t)n_sm) return 1;
    if (!ctx->d_coop.p)
    {
        MP2P_TRY(ctx->d_coop.ensure(64));
        MP2P_CUDA_TRY(cudaMemsetAsync(ctx->d_coop.p, 0, 64, ctx->stream));
        ctx->coop_arrivals = 0, ctx->coop_epoch = 0;
    }
    FusedSums fs{};
    MP2P_TRY(solve_scratch(ctx, 2 * (size_t)n_tiles, &fs.ticket, &fs.partials));
    fs.packet                 = d_packets;
    double*       mom_partials = fs.partials + (size_t)n_tiles * 32;
    unsigned int* mom_ticket   = fs.ticket + 1;
    // grid barriers of one launch: search | scan (2), sharded: search | records out | replay | scan (4)
    const unsigned n_barriers = peer ? 4u : 2u;
    CoopSync      cs{};
    cs.arrivals  = ctx->d_coop.as<unsigned long long>();
    cs.target    = ctx->coop_arrivals + n_tiles;  // barrier 0; barrier b completes at + (b + 1) n_tiles
    cs.sums_flag = reinterpret_cast<unsigned int*>(ctx->d_coop.as<char>() + 16);
    cs.epoch     = ctx->coop_epoch + 1;
    cs.host_out  = ctx->h_mapped_dev;  // NULL if the mapping is not available: the caller then copies
    cs.scan_barrier = n_barriers - 1;
    PeerLaunch pl{};
    if (peer)
    {
        pl.view      = peer->view;
        pl.rec_epoch = peer->rec_epoch + 1;
        pl.pkt_epoch = peer->pkt_epoch + 1;
    }
    uint32_t*      cloud_bbox = reinterpret_cast<uint32_t*>(ctx->d_coop.as<char>() + 32);  // 6 words
    GridView       gv = map->view;
    Pt2PtArgs      aa = a;
    CompactArgs    cc = c;
    const float *  qx = ctx->cur_qx, *qy = ctx->cur_qy, *qz = ctx->cur_qz, *lx = ctx->cur_lx, *ly = ctx->cur_ly, *lz = ctx->cur_lz;
    const uint32_t*     perm  = ctx->cur_perm;
    unsigned long long* claim = map->d_claim.as<unsigned long long>();
    uint32_t *          bbox = sv.bbox, *bbox_next = sv.bbox_next;
    unsigned long long* count = sv.count;
    void* args[] = {&gv, &aa, &cc, &qx, &qy, &qz, &lx, &ly, &lz, &perm, &claim, &cand, &cand_xyz, &bbox, &bbox_next, &status,
                    &d_out, &count, &fs, &mom_partials, &mom_ticket, &w_pt2pt, &cs, &pl, &per_k, &cloud_bbox};
    void* fn = peer ? reinterpret_cast<void*>(k_iterate_nn1_horn<true>) : reinterpret_cast<void*>(k_iterate_nn1_horn<false>);
    const cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(n_tiles), dim3(kScanThreads), args, 0, ctx->stream);
    if (e == cudaErrorCooperativeLaunchTooLarge)
    {
        cudaGetLastError();  // the grid does not fit after all (another context holds SM resources): three-kernel path
        return 1;
    }
    if (e != cudaSuccess)
    {
        cudaGetLastError();
        set_error("cooperative launch failed: %s", cudaGetErrorString(e));
        return MP2P_B200_ERR_CUDA;
    }
    ctx->coop_arrivals += (unsigned long long)n_barriers * n_tiles, ctx->coop_epoch += 1;
    if (peer) peer->rec_epoch += 1, peer->pkt_epoch += 2;
    count_launch(ctx);
    return 0;
}

// ==========================================================================================
int run_match_pt2pt(mp2p_b200_ctx* ctx, mp2p_b200_map* map, const float* lx, const float* ly,
                    const float* lz, uint64_t n_local, int local_on_device, const double pose[12],
                    const mp2p_b200_pt2pt_params* prm, const uint32_t* lbits, const uint32_t* gbits,
                    mp2p_b200_pair_pt2pt* out, uint64_t capacity, int out_on_device,
                    uint64_t* out_count, DeviceMatch* keep_on_device)
{
    *out_count          = 0;
    if (!ctx->owns_map(map))
    {
        set_error("map handle is not a live map of this context (destroyed, or created on another context)");
        return MP2P_B200_ERR_ARG;
    }
    if (keep_on_device) keep_on_device->d_count = nullptr, keep_on_device->d_pairs = nullptr, keep_on_device->capacity = 0;
    ctx->last2p.valid = false, ctx->last2p.sums = nullptr, ctx->spec_res.valid = false;
    const uint32_t K    = prm->pairingsPerPoint;
    const uint64_t nmap = map->view.n_points;
    if (nmap == 0 || n_local == 0) return 0;  // …DistanceThreshold.cpp:67
    if (n_local * (uint64_t)K >= 0xFFFFFFFFull)
    {
        set_error("n_local*pairingsPerPoint must be < 2^32-1");
        return MP2P_B200_ERR_ARG;
    }
    cudaStream_t st = ctx->stream;
    bool local_mapped = false;
    MP2P_TRY(stage_local(ctx, lx, ly, lz, n_local, local_on_device, K == 1 ? &local_mapped : nullptr));
    const uint32_t *d_lbits, *d_gbits;
    MP2P_TRY(upload_bits(ctx, ctx->d_lbits, lbits, n_local, &d_lbits));
    MP2P_TRY(upload_bits(ctx, ctx->d_gbits, gbits, nmap, &d_gbits));
    const uint64_t n_slots = n_local * K;
    const uint64_t n_tiles = (n_slots + kScanTile - 1) / kScanTile;
    SmallView           sv;
    unsigned long long* status;
    MP2P_TRY(prepare_small(ctx, n_tiles, sv, &status));
    MP2P_TRY(ctx->d_cand.ensure(n_slots * 8));

    if (++map->epoch == 0xFFFFFFFFu)  // tags exhausted: restart the claim words
    {
        MP2P_CUDA_TRY(cudaMemsetAsync(map->d_claim.p, 0xff, nmap * 8, st));
        map->epoch = 1;
    }

    Pt2PtArgs a{};
    for (int k = 0; k < 12; k++) a.pose.m[k] = pose[k];
    a.maxDistSq = (float)(prm->threshold * prm->threshold);  // …DistanceThreshold.cpp:82
    const double ang = prm->thresholdAngularDeg * 3.14159265358979323846 / 180.0;
    a.angSq     = (float)(ang * ang);  // :83
    a.n_local = (uint32_t)n_local, a.K = K;
    a.allowLocal = prm->allowMatchAlreadyMatchedPoints, a.allowGlobal = prm->allowMatchAlreadyMatchedGlobalPoints;
    a.tag = (unsigned long long)(0xFFFFFFFFu - map->epoch) << 32;
    a.tma_ok = ctx->cur_tma_ok;
    a.rl_start = start_level(map->view, K);
    a.n_phases = knn_phases();
    if (local_mapped) a.stage_x = ctx->d_lx.as<float>(), a.stage_y = ctx->d_ly.as<float>(), a.stage_z = ctx->d_lz.as<float>();

    const float *  dlx = ctx->cur_lx, *dly = ctx->cur_ly, *dlz = ctx->cur_lz;  // caller order (records)
    const float *  dqx = ctx->cur_qx, *dqy = ctx->cur_qy, *dqz = ctx->cur_qz;  // what the search walks
    auto*          claim = map->d_claim.as<unsigned long long>();
    auto*          cand  = ctx->d_cand.as<unsigned long long>();
    unsigned long long* stats = nullptr;
    MP2P_TRY(prepare_stats(ctx, &stats));
    float4* cand_xyz = nullptr;

    mp2p_b200_pair_pt2pt* d_out = out;
    if (!out_on_device)
    {
        MP2P_TRY(ctx->d_out2p.ensure(std::min<uint64_t>(capacity, n_slots) * sizeof(mp2p_b200_pair_pt2pt)));
        d_out = ctx->d_out2p.as<mp2p_b200_pair_pt2pt>();
    }
    CompactArgs c{};
    c.n_local = (uint32_t)n_local, c.K = K, c.allowGlobal = a.allowGlobal, c.tag = a.tag;
    c.gate_eps = (float)(prm->threshold + prm->bounding_box_intersection_check_epsilon);
    c.capacity = std::min<uint64_t>(capacity, n_slots);
    c.scan_epoch = ctx->scan_epoch;
    if (!out_on_device && !keep_on_device && c.capacity)
    {
        // pinned host output: the compaction stores the records there itself (no copy behind it)
        c.out_host = reinterpret_cast<uint32_t*>(mapped_alias(out));
        if (c.out_host)
        {
            if (!ctx->h_pinned_dev) ctx->h_pinned_dev = mapped_alias(ctx->h_pinned);
            c.count_host = static_cast<unsigned long long*>(ctx->h_pinned_dev);
            if (!c.count_host) c.out_host = nullptr;
        }
    }

    // whole iteration in one cooperative launch (k = 1, Horn sums AND moments wanted, no MatchState
    // bits, no search statistics), when the grid fits on the device
    mp2p_b200_peer* peer = keep_on_device ? keep_on_device->peer : nullptr;
    if (peer)  // shard of a bigger cloud: proposals and records carry whole-cloud numbers
    {
        const uint64_t per_k = keep_on_device->per_shard * K;
        a.slot_offset = (unsigned long long)peer->view.rank * per_k;
        c.slot_offset = a.slot_offset, c.index_offset = (uint32_t)(peer->view.rank * keep_on_device->per_shard);
    }
    if (K == 1 && keep_on_device && keep_on_device->want_horn_sums && keep_on_device->fuse_moments_w > 0.0 && !lbits &&
        !gbits && !stats)
    {
        MP2P_TRY(ctx->d_candxyz.ensure(n_slots * sizeof(float4)));
        cand_xyz = ctx->d_candxyz.as<float4>();
        prof_begin(ctx, 0);
        const int rc = launch_iterate_nn1_horn(ctx, map, a, c, sv, status, cand, cand_xyz, d_out, keep_on_device->want_horn_sums,
                                               keep_on_device->fuse_moments_w, (uint32_t)n_tiles, peer,
                                               keep_on_device->per_shard * K);
        if (rc < 0) return rc;
        if (rc == 0)
        {
            prof_end(ctx, 0);
            keep_on_device->d_count = sv.count, keep_on_device->d_pairs = d_out, keep_on_device->capacity = c.capacity;
            keep_on_device->moments_done = true;
            return 0;
        }
    }
    if (peer) return 1;  // nothing enqueued that the multi-kernel path does not redo
    // Host output into pinned memory + a plain Solver_Horn expected next (the caller's previous solver
    // call named the last matcher output): the single launch does search, compaction straight into
    // the caller's buffer AND both Horn passes; the solver call that follows costs no GPU work.
    if (K == 1 && !keep_on_device && c.out_host && !lbits && !gbits && !stats && ctx->spec_want.kind == 1 &&
        ctx->spec_unused < 2 && ctx->spec_want.horn.robust_kernel == 0 && ctx->spec_want.horn.w_pt2pt > 0.0 &&
        !ctx->spec_want.horn.use_scale_outlier_detector && ctx->h_mapped)
    {
        MP2P_TRY(ctx->d_candxyz.ensure(n_slots * sizeof(float4)));
        cand_xyz        = ctx->d_candxyz.as<float4>();
        double* packets = ctx->d_packet.as<double>() + 6 * MP2P_B200_PACKET_DOUBLES;  // [6],[7]: sums, moments
        prof_begin(ctx, 0);
        const int rc = launch_iterate_nn1_horn(ctx, map, a, c, sv, status, cand, cand_xyz, d_out, packets,
                                               ctx->spec_want.horn.w_pt2pt, (uint32_t)n_tiles);
        if (rc < 0) return rc;
        if (rc == 0)
        {
            prof_end(ctx, 0);
            MP2P_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            MP2P_CUDA_TRY(cudaGetLastError());
            const uint64_t cnt = *static_cast<unsigned long long*>(ctx->h_pinned);
            *out_count = cnt, ctx->hint_pt2pt = cnt;
            ctx->last2p.dev = d_out, ctx->last2p.n = cnt, ctx->last2p.valid = cnt <= c.capacity;
            ctx->last2p.sums = packets;
            if (cnt > c.capacity)
            {
                set_error("output capacity %llu too small for %llu pairings", (unsigned long long)c.capacity, (unsigned long long)cnt);
                return MP2P_B200_ERR_CAPACITY;
            }
            double* hp = nullptr;  // packets -> the speculation's pinned scratch
            MP2P_TRY(read_iteration_packets(ctx, true, packets, &hp));
            std::memcpy(spec_host(ctx), hp, 2 * MP2P_B200_PACKET_DOUBLES * 8);
            ctx->spec_res.valid = true, ctx->spec_res.pending = false, ctx->spec_res.kind = 1, ctx->spec_res.list = 1;
            ctx->spec_res.n = cnt;
            std::memcpy(ctx->spec_res.pose_in, pose, 96);
            ctx->spec_unused++;
            return 0;
        }
    }
    prof_begin(ctx, 0);
#define LAUNCH_MATCH(G)                                                                                    \
    {                                                                                                      \
        MP2P_LAUNCH_KMATCH(G, a, n_local, st, map->view, a, dqx, dqy, dqz, ctx->cur_perm, d_lbits, d_gbits, claim, cand, sv.bbox, stats, FitList{nullptr, nullptr, nullptr, 0}) \
    }
    if (K == 1)
    {
        MP2P_TRY(ctx->d_candxyz.ensure(n_slots * sizeof(float4)));
        cand_xyz = ctx->d_candxyz.as<float4>();
        MP2P_LAUNCH_NN1((uint32_t)((n_local + kNN1Threads - 1) / kNN1Threads), st, map->view, a, dqx, dqy, dqz, ctx->cur_perm,
                        d_lbits, d_gbits, claim, cand, cand_xyz, sv.bbox, stats);
    }
    else
    {
        MP2P_DISPATCH_G(K, LAUNCH_MATCH)
    }
#undef LAUNCH_MATCH
    prof_end(ctx, 0);
    count_launch(ctx);

    prof_begin(ctx, 1);
    FusedSums fs{nullptr, nullptr, nullptr};
    if (keep_on_device && keep_on_device->want_horn_sums)
    {
        MP2P_TRY(solve_scratch(ctx, n_tiles, &fs.ticket, &fs.partials));
        fs.packet = keep_on_device->want_horn_sums;
    }
    else if (!keep_on_device && !out_on_device)
    {
        // host output: the records also stay in d_out2p for a solver call that names them
        // (PAIRS_LAST_MATCH); the centroid sums of Solver_Horn's first pass cost nothing here
        MP2P_TRY(solve_scratch(ctx, n_tiles, &fs.ticket, &fs.partials));
        fs.packet        = ctx->d_packet.as<double>() + 4 * MP2P_B200_PACKET_DOUBLES;
        ctx->last2p.sums = fs.packet;
    }
    k_compact_pt2pt<<<compact_tiles_2p(n_slots), kScanThreads, 0, st>>>(map->view, c, dlx, dly, dlz, d_gbits,
                                                                       claim, cand, cand_xyz, sv.bbox, sv.bbox_next, status,
                                                                       sv.tile_counter, d_out, sv.count, fs);
    prof_end(ctx, 1);
    count_launch(ctx);
    if (keep_on_device)  // fused iteration: the caller enqueues the solver and synchronises once
    {
        keep_on_device->d_count = sv.count, keep_on_device->d_pairs = d_out, keep_on_device->capacity = c.capacity;
        return 0;
    }
    return fetch_results(ctx, sv.count, d_out, out, c.capacity, out_on_device, out_count, &ctx->hint_pt2pt, pose, fs.packet,
                         c.out_host != nullptr);
}

// ------------------------------------------------------------------------------------------
// Query-sharded matching (one process per GPU): phase A searches the shard, the caller all-gathers
// the candidate words of all shards, phase B replays EVERY shard's proposals into this GPU's claim
// array with one global slot numbering (slot = globalLocalIdx*K + rank) and compacts its own shard.
// The result is bit-identical to a single-GPU run over the whole cloud.
// ------------------------------------------------------------------------------------------
namespace
{
// records: n_shards exchange records of `rec_words` 64-bit words each, laid out
// [per_k candidate words | 3 words = 6 ordered bbox words | 1 pad]; global slot = shard*per_k + j
__global__ void __launch_bounds__(256)
    k_claim_all(const unsigned long long* __restrict__ records, uint32_t n_shards, uint64_t per_k, uint64_t rec_words,
                unsigned long long tag, const uint32_t* __restrict__ gbits, unsigned long long* __restrict__ claim)
{
    const uint64_t n_slots = per_k * n_shards;
    for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < n_slots; s += (uint64_t)gridDim.x * blockDim.x)
    {
        const uint64_t           r = s / per_k, j = s - r * per_k;
        const unsigned long long c = records[r * rec_words + j];
        if ((uint32_t)c == 0xFFFFFFFFu) continue;
        const uint32_t gi = (uint32_t)c;
        if (!bit_set(gbits, gi)) atomicMin(claim + gi, tag | s);
    }
}
// fold the per-shard bounding boxes (6 ordered words each: min xyz / max xyz) into one
__global__ void k_fold_bbox(const unsigned long long* __restrict__ records, uint32_t n_shards, uint64_t per_k,
                            uint64_t rec_words, uint32_t* __restrict__ out)
{
    const int d = threadIdx.x;
    if (d >= 6) return;
    uint32_t r = d < 3 ? 0xFFFFFFFFu : 0u;
    for (uint32_t p = 0; p < n_shards; p++)
    {
        const uint32_t v = reinterpret_cast<const uint32_t*>(records + p * rec_words + per_k)[d];
        r                = d < 3 ? min(r, v) : max(r, v);
    }
    out[d] = r;
}
}  // namespace

uint64_t shard_record_words(uint64_t per_shard, uint32_t K) { return per_shard * K + 4; }

int run_shard_search_pt2pt(mp2p_b200_ctx* ctx, mp2p_b200_map* map, const float* lx, const float* ly,
                           const float* lz, uint64_t n_local, int local_on_device, const double pose[12],
                           const mp2p_b200_pt2pt_params* prm, const uint32_t* lbits, uint64_t per_shard,
                           unsigned long long* d_record, const OwnerClaims* oc, uint32_t shard_rank)
{
    const uint32_t K  = prm->pairingsPerPoint;
    cudaStream_t   st = ctx->stream;
    if (!ctx->owns_map(map))
    {
        set_error("map handle is not a live map of this context (destroyed, or created on another context)");
        return MP2P_B200_ERR_ARG;
    }
    if (n_local > per_shard)
    {
        set_error("shard_search: n_local exceeds per_shard");
        return MP2P_B200_ERR_ARG;
    }
    uint32_t* d_bbox6 = reinterpret_cast<uint32_t*>(d_record + per_shard * K);
    // bbox = inverted (nothing seen yet); slots of a short or empty shard = "no candidate"
    const uint32_t init[8] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0, 0, 0, 0, 0};
    MP2P_CUDA_TRY(cudaMemcpyAsync(d_bbox6, init, sizeof(init), cudaMemcpyHostToDevice, st));
    const bool nothing = n_local == 0 || map->view.n_points == 0;
    const uint64_t first_pad = nothing ? 0 : n_local;
    if (first_pad < per_shard)
        MP2P_CUDA_TRY(cudaMemsetAsync(d_record + first_pad * K, 0xff, (per_shard - first_pad) * K * 8, st));
    if (nothing) return 0;
    MP2P_TRY(stage_local(ctx, lx, ly, lz, n_local, local_on_device));
    const uint32_t* d_lbits;
    MP2P_TRY(upload_bits(ctx, ctx->d_lbits, lbits, n_local, &d_lbits));
    Pt2PtArgs a{};
    for (int k = 0; k < 12; k++) a.pose.m[k] = pose[k];
    a.maxDistSq = (float)(prm->threshold * prm->threshold);
    const double ang = prm->thresholdAngularDeg * 3.14159265358979323846 / 180.0;
    a.angSq     = (float)(ang * ang);
    a.n_local = (uint32_t)n_local, a.K = K;
    a.allowLocal = prm->allowMatchAlreadyMatchedPoints, a.allowGlobal = 1;  // claims happen in phase B
    a.tag = 0;
    if (oc && !prm->allowMatchAlreadyMatchedGlobalPoints)
    {
        // owner-partitioned claims: the search proposes straight into the owners' memory, under the GLOBAL slot numbers
        a.allowGlobal = 0, a.tag = oc->tag, a.claim_parts = oc->parts, a.claim_world = oc->world;
        a.slot_offset = (unsigned long long)shard_rank * per_shard * K;
    }
    a.tma_ok = ctx->cur_tma_ok;
    a.rl_start = start_level(map->view, K);
    a.n_phases = knn_phases();
    const float *dqx = ctx->cur_qx, *dqy = ctx->cur_qy, *dqz = ctx->cur_qz;
    unsigned long long* stats = nullptr;
    MP2P_TRY(prepare_stats(ctx, &stats));
    prof_begin(ctx, 0);
#define LAUNCH_MATCH(G)                                                                                    \
    {                                                                                                      \
        MP2P_LAUNCH_KMATCH(G, a, n_local, st, map->view, a, dqx, dqy, dqz, ctx->cur_perm, d_lbits, nullptr, nullptr, d_record, d_bbox6, stats, FitList{nullptr, nullptr, nullptr, 0}) \
    }
    if (K == 1)
    {
        MP2P_TRY(ctx->d_candxyz.ensure(n_local * sizeof(float4)));
        const uint32_t nb = (uint32_t)((n_local + kNN1Threads - 1) / kNN1Threads);
        MP2P_LAUNCH_NN1(nb, st, map->view, a, dqx, dqy, dqz, ctx->cur_perm, d_lbits, nullptr, nullptr, d_record,
                        ctx->d_candxyz.as<float4>(), d_bbox6, stats);
    }
    else
    {
        MP2P_DISPATCH_G(K, LAUNCH_MATCH)
    }
#undef LAUNCH_MATCH
    prof_end(ctx, 0);
    count_launch(ctx);
    return 0;
}

// out_count == NULL: nothing is read back and the stream is not synchronised — the count stays on
// the device (ctx->last_count) for solver calls that pass MP2P_B200_COUNT_ON_DEVICE.
int run_shard_resolve_pt2pt(mp2p_b200_ctx* ctx, mp2p_b200_map* map, uint64_t n_local, uint32_t shard_rank,
                            uint32_t n_shards, uint64_t per_shard, const unsigned long long* d_records,
                            const mp2p_b200_pt2pt_params* prm, const uint32_t* gbits,
                            mp2p_b200_pair_pt2pt* out, uint64_t capacity, int out_on_device,
                            uint64_t* out_count, double* d_horn_sums, const OwnerClaims* oc)
{
    if (out_count) *out_count = 0;
    if (!ctx->owns_map(map))
    {
        set_error("map handle is not a live map of this context (destroyed, or created on another context)");
        return MP2P_B200_ERR_ARG;
    }
    ctx->last_count = nullptr, ctx->last_capacity = 0;
    ctx->last2p.valid = false, ctx->last2p.sums = nullptr, ctx->spec_res.valid = false;
    const uint32_t K    = prm->pairingsPerPoint;
    const uint64_t nmap = map->view.n_points;
    const uint64_t per_k = per_shard * K, rec_words = shard_record_words(per_shard, K);
    if (per_k * n_shards >= 0xFFFFFFFFull || shard_rank >= n_shards || n_local > per_shard)
    {
        set_error("shard_resolve: need n_shards*per_shard*pairingsPerPoint < 2^32-1, shard_rank < n_shards, n_local <= per_shard");
        return MP2P_B200_ERR_ARG;
    }
    cudaStream_t st = ctx->stream;
    if (nmap == 0 || n_local == 0)
    {
        if (d_horn_sums) MP2P_CUDA_TRY(cudaMemsetAsync(d_horn_sums, 0, MP2P_B200_PACKET_DOUBLES * 8, st));
        return 0;  // last_count stays NULL: the solver calls that follow see zero pairings
    }
    const uint32_t* d_gbits;
    MP2P_TRY(upload_bits(ctx, ctx->d_gbits, gbits, nmap, &d_gbits));
    const uint64_t      n_slots = n_local * K;
    const uint64_t      n_tiles = (n_slots + kScanTile - 1) / kScanTile;
    SmallView           sv;
    unsigned long long* status;
    MP2P_TRY(prepare_small(ctx, n_tiles, sv, &status));
    k_fold_bbox<<<1, 32, 0, st>>>(d_records, n_shards, per_k, rec_words, sv.bbox);
    count_launch(ctx);
    if (++map->epoch == 0xFFFFFFFFu)
    {
        MP2P_CUDA_TRY(cudaMemsetAsync(map->d_claim.p, 0xff, nmap * 8, st));
        map->epoch = 1;
    }
    const unsigned long long tag   = oc ? oc->tag : (unsigned long long)(0xFFFFFFFFu - map->epoch) << 32;
    auto*                    claim = map->d_claim.as<unsigned long long>();
    if (!prm->allowMatchAlreadyMatchedGlobalPoints && !oc)  // (owner-partitioned claims were proposed by the searches)
    {
        const uint64_t all_slots = per_k * n_shards;
        const uint32_t blocks    = (uint32_t)std::min<uint64_t>((all_slots + 255) / 256, 148 * 16);
        k_claim_all<<<blocks, 256, 0, st>>>(d_records, n_shards, per_k, rec_words, tag, d_gbits, claim);
        count_launch(ctx);
    }
    mp2p_b200_pair_pt2pt* d_out = out;
    if (!out_on_device)
    {
        MP2P_TRY(ctx->d_out2p.ensure(std::min<uint64_t>(capacity, n_slots) * sizeof(mp2p_b200_pair_pt2pt)));
        d_out = ctx->d_out2p.as<mp2p_b200_pair_pt2pt>();
    }
    CompactArgs c{};
    c.n_local = (uint32_t)n_local, c.K = K, c.allowGlobal = prm->allowMatchAlreadyMatchedGlobalPoints, c.tag = tag;
    c.gate_eps = (float)(prm->threshold + prm->bounding_box_intersection_check_epsilon);
    c.capacity = std::min<uint64_t>(capacity, n_slots);
    c.slot_offset = (uint64_t)shard_rank * per_k, c.index_offset = (uint32_t)(shard_rank * per_shard);
    c.scan_epoch = ctx->scan_epoch;
    if (oc) c.claim_parts = oc->parts, c.claim_world = oc->world;
    FusedSums fs{nullptr, nullptr, nullptr};
    if (d_horn_sums)
    {
        MP2P_TRY(solve_scratch(ctx, n_tiles, &fs.ticket, &fs.partials));
        fs.packet = d_horn_sums;
    }
    prof_begin(ctx, 1);
    k_compact_pt2pt<<<compact_tiles_2p(n_slots), kScanThreads, 0, st>>>(
        map->view, c, ctx->cur_lx, ctx->cur_ly, ctx->cur_lz, d_gbits, claim,
        d_records + (uint64_t)shard_rank * rec_words, K == 1 ? ctx->d_candxyz.as<float4>() : nullptr, sv.bbox,
        sv.bbox_next, status, sv.tile_counter, d_out, sv.count, fs);
    prof_end(ctx, 1);
    count_launch(ctx);
    ctx->last_count = sv.count, ctx->last_capacity = c.capacity;
    if (!out_count) return 0;
    return fetch_results(ctx, sv.count, d_out, out, c.capacity, out_on_device, out_count, &ctx->hint_pt2pt);
}

int run_match_pt2pl(mp2p_b200_ctx* ctx, mp2p_b200_map* map, const float* lx, const float* ly,
                    const float* lz, uint64_t n_local, int local_on_device, const double pose[12],
                    const mp2p_b200_pt2pl_params* prm, const uint32_t* lbits,
                    mp2p_b200_pair_pt2pl* out, uint64_t capacity, int out_on_device,
                    uint64_t* out_count, DeviceMatch* keep_on_device, const LineMode* line)
{
    *out_count          = 0;
    if (!ctx->owns_map(map))
    {
        set_error("map handle is not a live map of this context (destroyed, or created on another context)");
        return MP2P_B200_ERR_ARG;
    }
    if (keep_on_device) *keep_on_device = DeviceMatch{};
    ctx->last2l.valid = false, ctx->spec_res.valid = false;
    const uint64_t nmap = map->view.n_points;
    if (nmap == 0 || n_local == 0) return 0;
    if (n_local >= 0xFFFFFFFFull || prm->knn < 1 || prm->knn > MP2P_B200_MAX_KNN)
    {
        set_error("pt2pl: knn must be in [1,%d] and n_local < 2^32-1", MP2P_B200_MAX_KNN);
        return MP2P_B200_ERR_ARG;
    }
    cudaStream_t st = ctx->stream;
    MP2P_TRY(stage_local(ctx, lx, ly, lz, n_local, local_on_device));
    const uint32_t* d_lbits;
    MP2P_TRY(upload_bits(ctx, ctx->d_lbits, lbits, n_local, &d_lbits));
    const uint64_t      n_tiles = (n_local + kScanTile - 1) / kScanTile;
    SmallView           sv;
    unsigned long long* status;
    MP2P_TRY(prepare_small(ctx, n_tiles, sv, &status));
    MP2P_TRY(ctx->d_plcand.ensure(n_local * sizeof(PlaneCandidate)));
    MP2P_TRY(ctx->d_cand.ensure(n_local * prm->knn * 8));  // k-NN keys
    MP2P_TRY(ctx->d_okflags.ensure(n_local));

    Pt2PlArgs a{};
    for (int k = 0; k < 12; k++) a.pose.m[k] = pose[k];
    a.radiusSq = (float)(prm->searchRadius * prm->searchRadius);
    a.distThr  = (float)prm->distanceThreshold;
    a.planeEigenThreshold = prm->planeEigenThreshold;
    a.n_local = (uint32_t)n_local, a.K = prm->knn, a.minPts = prm->minimumPlanePoints;
    a.allowLocal = prm->allowMatchAlreadyMatchedPoints;
    a.tma_ok     = ctx->cur_tma_ok;
    if (line)  // Matcher_Point2Line: unbounded nn_multiple_search (:103-105), then the cut at threshold^2
    {
        a.line_mode = 1, a.radiusSq = __builtin_inff();
        a.lineMaxSqr = (float)(prm->distanceThreshold * prm->distanceThreshold);  // :77
        a.planeEigenThreshold = line->lineEigenThreshold, a.minPts = line->minimumLinePoints;
    }
    const float gate_eps = (float)(prm->distanceThreshold + prm->bounding_box_intersection_check_epsilon);

    const float *  dlx = ctx->cur_lx, *dly = ctx->cur_ly, *dlz = ctx->cur_lz;  // caller order (records)
    const float *  dqx = ctx->cur_qx, *dqy = ctx->cur_qy, *dqz = ctx->cur_qz;  // what the search walks
    auto*          plc = ctx->d_plcand.as<PlaneCandidate>();
    auto*          okf = ctx->d_okflags.as<uint8_t>();
    auto*          cand = ctx->d_cand.as<unsigned long long>();
    unsigned long long* stats = nullptr;
    MP2P_TRY(prepare_stats(ctx, &stats));
    // k-NN within searchRadius: the pt2pt search kernel with threshold = searchRadius, no claims
    Pt2PtArgs sa{};
    sa.pose = a.pose, sa.maxDistSq = a.radiusSq, sa.angSq = 0.f, sa.n_local = a.n_local, sa.K = a.K;
    sa.allowLocal = a.allowLocal, sa.allowGlobal = 1, sa.tag = 0, sa.tma_ok = a.tma_ok;
    sa.cand_sorted = 1;  // the plane fit walks the same order
    MP2P_TRY(ctx->d_fitlist.ensure((n_local + 1) * 4));
    MP2P_CUDA_TRY(cudaMemsetAsync(ctx->d_fitlist.p, 0, 4, st));
    const FitList fit{ctx->d_fitlist.as<uint32_t>() + 1, ctx->d_fitlist.as<uint32_t>(), okf,
                      line ? (int)std::max<uint32_t>(1u, line->minimumLinePoints) : (int)std::max<uint32_t>(3u, prm->minimumPlanePoints)};
    prof_begin(ctx, 0);
#define LAUNCH_SEARCH(G)                                                                                   \
    {                                                                                                      \
        MP2P_LAUNCH_KMATCH(G, sa, n_local, st, map->view, sa, dqx, dqy, dqz, ctx->cur_perm, d_lbits, nullptr, nullptr, cand, sv.bbox, stats, fit) \
    }
#define LAUNCH_FIT(KT) \
    k_plane_fit<KT><<<(uint32_t)((n_local + kFitThreads - 1) / kFitThreads), kFitThreads, 0, st>>>(map->view, a, dqx, dqy, dqz, ctx->cur_perm, cand, fit.list, fit.count, plc, okf)
    sa.rl_start = start_level(map->view, prm->knn);
    sa.n_phases = knn_phases();
    MP2P_DISPATCH_G(prm->knn, LAUNCH_SEARCH)
    prof_end(ctx, 0);
    prof_begin(ctx, 6);
    switch (pick_kt(prm->knn))
    {
        case 1:
        case 4: LAUNCH_FIT(4); break;
        case 8: LAUNCH_FIT(8); break;
        case 16: LAUNCH_FIT(16); break;
        default: LAUNCH_FIT(32); break;
    }
#undef LAUNCH_SEARCH
#undef LAUNCH_FIT
    prof_end(ctx, 6);
    count_launch(ctx, 2);

    mp2p_b200_pair_pt2pl* d_out = out;
    const uint64_t        cap   = std::min<uint64_t>(capacity, n_local);
    if (!out_on_device)
    {
        MP2P_TRY(ctx->d_out2l.ensure(cap * sizeof(mp2p_b200_pair_pt2pl)));
        d_out = ctx->d_out2l.as<mp2p_b200_pair_pt2pl>();
    }
    // pinned host output: the compaction CAN store records and count straight into the caller's buffer like the
    // pt2pt matcher does — measured on C3 (4.5 MB of 72-byte records) the stores over PCIe take 55 us longer than
    // the DMA copy behind the kernel (round 2, visit 13: e2e 0.534 vs 0.478 ms), so it is off unless
    // $MP2P_PT2PL_ZERO_COPY=1
    uint32_t*           out_host   = nullptr;
    unsigned long long* count_host = nullptr;
    static const bool   zc_pt2pl   = [] {
        const char* e = getenv("MP2P_PT2PL_ZERO_COPY");
        return e && atoi(e) == 1;
    }();
    if (zc_pt2pl && !keep_on_device && !out_on_device)
    {
        out_host = reinterpret_cast<uint32_t*>(mapped_alias(out));
        if (out_host)
        {
            if (!ctx->h_pinned_dev) ctx->h_pinned_dev = mapped_alias(ctx->h_pinned);
            count_host = static_cast<unsigned long long*>(ctx->h_pinned_dev);
            if (!count_host) out_host = nullptr;
        }
    }
    prof_begin(ctx, 1);
    k_compact_pt2pl<<<compact_tiles_2l(n_local), kScanThreads, 0, st>>>(map->view, (uint32_t)n_local, gate_eps,
                                                               cap, dlx, dly, dlz, plc, okf, sv.bbox, sv.bbox_next,
                                                               status, sv.tile_counter, d_out, sv.count,
                                                               ctx->scan_epoch, line ? 1 : 0, out_host, count_host);
    prof_end(ctx, 1);
    count_launch(ctx);
    if (keep_on_device)
    {
        keep_on_device->d_count = sv.count, keep_on_device->d_pairs = d_out, keep_on_device->capacity = cap;
        return 0;
    }
    if (line)
    {
        const int rc = fetch_results(ctx, sv.count, d_out, out, cap, out_on_device, out_count, &ctx->hint_pt2ln, nullptr, nullptr,
                                     out_host != nullptr);
        ctx->last2l.valid = false;  // the device copy holds LINE records: not a pt2pl list a solver may name
        return rc;
    }
    return fetch_results(ctx, sv.count, d_out, out, cap, out_on_device, out_count, &ctx->hint_pt2pl, pose, nullptr,
                         out_host != nullptr);
}

int run_knn(mp2p_b200_ctx* ctx, const mp2p_b200_map* map, const float* qx, const float* qy,
            const float* qz, uint64_t nq, uint32_t K, float radius2, uint32_t* out_idx,
            float* out_d2, int32_t* out_found)
{
    if (nq == 0) return 0;
    if (!ctx->owns_map(map))
    {
        set_error("map handle is not a live map of this context (destroyed, or created on another context)");
        return MP2P_B200_ERR_ARG;
    }
    if (K < 1 || K > MP2P_B200_MAX_KNN || nq >= 0xFFFFFFFFull)
    {
        set_error("knn: k must be in [1,%d]", MP2P_B200_MAX_KNN);
        return MP2P_B200_ERR_ARG;
    }
    cudaStream_t st = ctx->stream;
    MP2P_TRY(stage_local(ctx, qx, qy, qz, nq, 0));
    MP2P_TRY(ctx->d_knn_idx.ensure(nq * K * 4));
    MP2P_TRY(ctx->d_knn_d2.ensure(nq * K * 4));
    MP2P_TRY(ctx->d_knn_found.ensure(nq * 4));
    if (map->view.n_points == 0)
    {
        MP2P_CUDA_TRY(cudaMemsetAsync(ctx->d_knn_found.p, 0, nq * 4, st));
    }
    else
    {
        const float *  dqx = ctx->cur_lx, *dqy = ctx->cur_ly, *dqz = ctx->cur_lz;
        auto *oi = ctx->d_knn_idx.as<uint32_t>();
        auto *od = ctx->d_knn_d2.as<float>();
        auto *of = ctx->d_knn_found.as<int32_t>();
        const int rl0 = start_level(map->view, K);
#define LAUNCH_KNN(G) k_knn<G><<<(uint32_t)((nq * G + 255) / 256), 256, 0, st>>>(map->view, dqx, dqy, dqz, (uint32_t)nq, K, radius2, rl0, knn_phases(), oi, od, of);
        MP2P_DISPATCH_G(K, LAUNCH_KNN)
#undef LAUNCH_KNN
        count_launch(ctx);
    }
    MP2P_CUDA_TRY(cudaMemcpyAsync(out_idx, ctx->d_knn_idx.p, nq * K * 4, cudaMemcpyDeviceToHost, st));
    MP2P_CUDA_TRY(cudaMemcpyAsync(out_d2, ctx->d_knn_d2.p, nq * K * 4, cudaMemcpyDeviceToHost, st));
    MP2P_CUDA_TRY(cudaMemcpyAsync(out_found, ctx->d_knn_found.p, nq * 4, cudaMemcpyDeviceToHost, st));
    MP2P_CUDA_TRY(cudaStreamSynchronize(st));
    MP2P_CUDA_TRY(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// Matcher_Points_InlierRatio (mp2p_icp/src/Matcher_Points_InlierRatio.cpp:41-143):
// unbounded 1-NN of every local point (the K = 1 search kernel with an infinite radius), all
// tentative pairings ordered by errorSquareAfterTransformation, the first
// mrpt::round(nTotal * inliersRatio) kept and emitted IN THAT ORDER, a global point going to the
// first pairing (in sorted order) that names it.
//   k_ir_keys     32-bit sort keys (the distance bits; "no candidate" sorts last) laid out by
//                 DESCENDING local index + count of the tentative pairings
//   rs::sort_pairs  stable LSD radix sort, 4 passes: equal distances stay in descending local index
//                 = reverse insertion order = what multimap::emplace_hint(begin()) leaves (:104)
//   k_ir_claim    positions p < nKeep propose  claim[g] = min(tag | p)
//   k_ir_compact  acceptance + scan + records staged in shared memory, coalesced stores
// ------------------------------------------------------------------------------------------
namespace
{
__device__ __forceinline__ unsigned long long ir_keep(const unsigned long long* __restrict__ n_total, double ratio)
{
    // mrpt::round(double(nTotal) * inliersRatio), :119 — lrint: nearest, ties to even
    return (unsigned long long)__double2ll_rn(__dmul_rn((double)__ldg(n_total), ratio));
}

__global__ void __launch_bounds__(256)
    k_ir_keys(const unsigned long long* __restrict__ cand, uint32_t n, unsigned long long* __restrict__ keys,
              uint32_t* __restrict__ vals, unsigned long long* __restrict__ n_total)
{
    const uint32_t j     = blockIdx.x * 256u + threadIdx.x;
    bool           valid = false;
    if (j < n)
    {
        const uint32_t           i = n - 1u - j;
        const unsigned long long c = cand[i];
        valid                      = (uint32_t)c != 0xFFFFFFFFu;
        keys[j]                    = valid ? (c >> 32) : 0xFFFFFFFFull;
        vals[j]                    = i;
    }
    const unsigned m = __ballot_sync(0xffffffffu, valid);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_total, (unsigned long long)__popc(m));
}

__global__ void __launch_bounds__(256)
    k_ir_claim(const unsigned long long* __restrict__ cand, const uint32_t* __restrict__ vals,
               const unsigned long long* __restrict__ n_total, double ratio, unsigned long long tag,
               const uint32_t* __restrict__ gbits, unsigned long long* __restrict__ claim)
{
    const unsigned long long p = (unsigned long long)blockIdx.x * 256u + threadIdx.x;
    if (p >= ir_keep(n_total, ratio)) return;
    const uint32_t g = (uint32_t)cand[vals[p]];
    if (!bit_set(gbits, g)) atomicMin(claim + g, tag | p);
}

__global__ void __launch_bounds__(kScanThreads)
    k_ir_compact(GridView g, uint32_t n_local, int allowGlobal, unsigned long long tag, float gate_eps, uint64_t capacity,
                 uint32_t scan_epoch, double ratio, const float* __restrict__ lx, const float* __restrict__ ly,
                 const float* __restrict__ lz, const uint32_t* __restrict__ gbits,
                 const unsigned long long* __restrict__ claim, const unsigned long long* __restrict__ cand,
                 const float4* __restrict__ cand_xyz, const uint32_t* __restrict__ vals,
                 unsigned long long* __restrict__ n_total /* [0] tentative pairings, [1] out: bbox gate */,
                 const uint32_t* __restrict__ bbox, uint32_t* __restrict__ bbox_next, unsigned long long* __restrict__ status,
                 uint32_t* __restrict__ tile_counter, mp2p_b200_pair_pt2pt* __restrict__ out,
                 unsigned long long* __restrict__ out_count)
{
    __shared__ ScanSmem sm;
    __shared__ uint32_t s_rec[kScanThreads * 9];
    bbox_rearm(bbox_next);
    const uint32_t n_tiles = (n_local + kScanTile - 1) / kScanTile;
    if (threadIdx.x == 0)
    {
        sm.tile_id = atomicAdd(tile_counter, 1u);
        if (sm.tile_id == n_tiles - 1) *tile_counter = 0u;
    }
    __syncthreads();
    const uint32_t tile = sm.tile_id;
    const bool     gate = bbox_gate(g, bbox, gate_eps);
    if (tile == 0 && threadIdx.x == 0) n_total[1] = gate ? 1ull : 0ull;
    const unsigned long long p  = (unsigned long long)tile * kScanTile + threadIdx.x;
    bool                     ok = gate && p < ir_keep(n_total, ratio);
    uint32_t                 i = 0, gi = 0;
    unsigned long long       c  = 0;
    float4                   gp = make_float4(0.f, 0.f, 0.f, 0.f);
    float                    px = 0.f, py = 0.f, pz = 0.f;
    if (ok)
    {
        i  = vals[p];
        c  = cand[i];
        gi = (uint32_t)c;
        gp = __ldg(cand_xyz + i);
        px = lx[i], py = ly[i], pz = lz[i];
        if (!allowGlobal) ok = !bit_set(gbits, gi) && __ldcg(claim + gi) == (tag | p);  // :126-128
    }
    const unsigned long long w         = grid_exclusive_scan(sm, tile, n_tiles, ok ? 1u : 0u, status, out_count, scan_epoch);
    const unsigned long long tile_base = sm.tile_base;
    if (ok)
    {
        uint32_t* o = s_rec + (uint32_t)(w - tile_base) * 9;
        o[0] = gi, o[1] = i;
        o[2] = __float_as_uint(gp.x), o[3] = __float_as_uint(gp.y), o[4] = __float_as_uint(gp.z);
        o[5] = __float_as_uint(px), o[6] = __float_as_uint(py), o[7] = __float_as_uint(pz);
        o[8] = (uint32_t)(c >> 32);
    }
    __syncthreads();
    const unsigned long long room  = capacity > tile_base ? capacity - tile_base : 0ull;
    const uint32_t           n_rec = (uint32_t)min((unsigned long long)sm.tile_total, room);
    uint32_t*                dst   = reinterpret_cast<uint32_t*>(out) + tile_base * 9;
    for (uint32_t k = threadIdx.x; k < n_rec * 9; k += kScanThreads) dst[k] = s_rec[k];
}
}  // namespace

// *status_out: 0 ok, 1 = the reference would have thrown ASSERT_(nTotal > 0) (:117)
int run_match_inlier_ratio(mp2p_b200_ctx* ctx, mp2p_b200_map* map, const float* lx, const float* ly, const float* lz,
                           uint64_t n_local, int local_on_device, const double pose[12], double ratio, int allowLocal,
                           int allowGlobal, double bbox_eps, const uint32_t* lbits, const uint32_t* gbits,
                           mp2p_b200_pair_pt2pt* out, uint64_t capacity, int out_on_device, uint64_t* out_count)
{
    *out_count = 0;
    if (!ctx->owns_map(map))
    {
        set_error("map handle is not a live map of this context (destroyed, or created on another context)");
        return MP2P_B200_ERR_ARG;
    }
    ctx->last2p.valid = false, ctx->last2p.sums = nullptr, ctx->spec_res.valid = false;
    const uint64_t nmap = map->view.n_points;
    if (nmap == 0 || n_local == 0) return 0;  // :58
    if (n_local >= 0x7FFFFFFFull)
    {
        set_error("match_inlier_ratio: n_local must be < 2^31");
        return MP2P_B200_ERR_ARG;
    }
    cudaStream_t st = ctx->stream;
    MP2P_TRY(stage_local(ctx, lx, ly, lz, n_local, local_on_device));
    const uint32_t *d_lbits, *d_gbits;
    MP2P_TRY(upload_bits(ctx, ctx->d_lbits, lbits, n_local, &d_lbits));
    MP2P_TRY(upload_bits(ctx, ctx->d_gbits, gbits, nmap, &d_gbits));
    const uint64_t      n_tiles = (n_local + kScanTile - 1) / kScanTile;
    SmallView           sv;
    unsigned long long* status;
    MP2P_TRY(prepare_small(ctx, n_tiles, sv, &status));
    MP2P_TRY(ctx->d_cand.ensure(n_local * 8));
    MP2P_TRY(ctx->d_candxyz.ensure(n_local * sizeof(float4)));
    const uint32_t sort_tiles = (uint32_t)((n_local + rs::kTile - 1) / rs::kTile);
    MP2P_TRY(ctx->d_irk0.ensure(n_local * 8));
    MP2P_TRY(ctx->d_irk1.ensure(n_local * 8));
    MP2P_TRY(ctx->d_irv0.ensure(n_local * 4));
    MP2P_TRY(ctx->d_irv1.ensure(n_local * 4));
    MP2P_TRY(ctx->d_irtmp.ensure(((size_t)256 * sort_tiles + 256) * 4 + 64));
    if (++map->epoch == 0xFFFFFFFFu)
    {
        MP2P_CUDA_TRY(cudaMemsetAsync(map->d_claim.p, 0xff, nmap * 8, st));
        map->epoch = 1;
    }
    const unsigned long long tag = (unsigned long long)(0xFFFFFFFFu - map->epoch) << 32;

    Pt2PtArgs a{};
    for (int k = 0; k < 12; k++) a.pose.m[k] = pose[k];
    a.maxDistSq = __builtin_inff(), a.angSq = 0.f;  // nn_single_search is unbounded (:89-91)
    a.n_local = (uint32_t)n_local, a.K = 1;
    a.allowLocal = allowLocal, a.allowGlobal = 1;  // no proposals from the search: claims follow the sort
    a.tma_ok   = ctx->cur_tma_ok;
    a.rl_start = start_level(map->view, 1);
    auto*               cand     = ctx->d_cand.as<unsigned long long>();
    auto*               cand_xyz = ctx->d_candxyz.as<float4>();
    unsigned long long* stats    = nullptr;
    MP2P_TRY(prepare_stats(ctx, &stats));
    prof_begin(ctx, 0);
    MP2P_LAUNCH_NN1((uint32_t)((n_local + kNN1Threads - 1) / kNN1Threads), st, map->view, a, ctx->cur_qx, ctx->cur_qy,
                    ctx->cur_qz, ctx->cur_perm, d_lbits, nullptr, nullptr, cand, cand_xyz, sv.bbox, stats);
    prof_end(ctx, 0);
    count_launch(ctx);

    prof_begin(ctx, 1);
    // counters live behind the sort scratch: [0] tentative pairings, [1] bbox gate
    auto* n_total = reinterpret_cast<unsigned long long*>(ctx->d_irtmp.as<char>() + ((size_t)256 * sort_tiles + 256) * 4);
    n_total       = reinterpret_cast<unsigned long long*>((reinterpret_cast<uintptr_t>(n_total) + 15) & ~uintptr_t(15));
    MP2P_CUDA_TRY(cudaMemsetAsync(n_total, 0, 16, st));
    const uint32_t nb = (uint32_t)((n_local + 255) / 256);
    k_ir_keys<<<nb, 256, 0, st>>>(cand, (uint32_t)n_local, ctx->d_irk0.as<unsigned long long>(), ctx->d_irv0.as<uint32_t>(), n_total);
    count_launch(ctx);
    MP2P_TRY(rs::sort_pairs(ctx, ctx->d_irk0.as<unsigned long long>(), ctx->d_irv0.as<uint32_t>(),
                            ctx->d_irk1.as<unsigned long long>(), ctx->d_irv1.as<uint32_t>(), (uint32_t)n_local, 32,
                            ctx->d_irtmp.as<uint32_t>()));
    auto* claim = map->d_claim.as<unsigned long long>();
    if (!allowGlobal)
    {
        k_ir_claim<<<nb, 256, 0, st>>>(cand, ctx->d_irv0.as<uint32_t>(), n_total, ratio, tag, d_gbits, claim);
        count_launch(ctx);
    }
    mp2p_b200_pair_pt2pt* d_out = out;
    const uint64_t        cap   = std::min<uint64_t>(capacity, n_local);
    if (!out_on_device)
    {
        MP2P_TRY(ctx->d_out2p.ensure(cap * sizeof(mp2p_b200_pair_pt2pt)));
        d_out = ctx->d_out2p.as<mp2p_b200_pair_pt2pt>();
    }
    k_ir_compact<<<(uint32_t)n_tiles, kScanThreads, 0, st>>>(map->view, (uint32_t)n_local, allowGlobal, tag, (float)bbox_eps, cap,
                                                            ctx->scan_epoch, ratio, ctx->cur_lx, ctx->cur_ly, ctx->cur_lz, d_gbits,
                                                            claim, cand, cand_xyz, ctx->d_irv0.as<uint32_t>(), n_total, sv.bbox,
                                                            sv.bbox_next, status, sv.tile_counter, d_out, sv.count);
    prof_end(ctx, 1);
    count_launch(ctx);
    unsigned long long* h_tot = reinterpret_cast<unsigned long long*>(static_cast<char*>(ctx->h_pinned) + 64);
    MP2P_CUDA_TRY(cudaMemcpyAsync(h_tot, n_total, 16, cudaMemcpyDeviceToHost, st));
    MP2P_TRY(fetch_results(ctx, sv.count, d_out, out, cap, out_on_device, out_count, &ctx->hint_ir));
    if (h_tot[1] && h_tot[0] == 0)
    {
        set_error("match_inlier_ratio: no tentative pairing at all (the reference asserts nTotal > 0, Matcher_Points_InlierRatio.cpp:117)");
        return MP2P_B200_ERR_ARG;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// Matcher_Adaptive (mp2p_icp/src/Matcher_Adaptive.cpp:59-314), in two device phases
// with the 50-bin histogram crossing to the host in between — the confidence-interval step is
// mrpt::math code (CHistogram / confidenceIntervalsFromHistogram) that a plugin built against MRPT
// calls itself; the library carries a restatement for everybody else (api.cu).
//   phase 1  k-NN (k = 1: k_match_pt2pt_nn1, radius <= absMax; k > 1: k_match_pt2pt<G>, radius <
//            absMax, at most MAX_CORRS_PER_LOCAL = 10 kept), k_ad_minmax (min / max / count of the
//            1st and 2nd neighbour errors, :168-181), k_ad_hist (CHistogram::add, bin =
//            size_t(binSizeInv * (x - min)) in fp64, + the bounding-box gate)
//   phase 2  k_ad_decide (per local point: plane through its neighbours -> pt2pl, else the pt2pt
//            candidates that pass the adaptive threshold and the 1st-to-2nd ratio, :219-298), then
//            the two ordinary compactions
// ------------------------------------------------------------------------------------------
namespace
{
constexpr int kAdBins = MP2P_B200_ADAPTIVE_BINS;

struct AdResult  // device block read back by the host after phase 1
{
    unsigned int       emin_bits, emax_bits;  // float bits (errors are >= 0: the bit patterns order like the values)
    unsigned long long n_samples;
    unsigned long long bins[kAdBins];
    unsigned int       gate;
};

__global__ void __launch_bounds__(256)
    k_ad_minmax(const unsigned long long* __restrict__ cand, uint32_t n_local, uint32_t K, AdResult* __restrict__ res)
{
    unsigned int lo = 0xFFFFFFFFu, hi = 0u, cnt = 0u;
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < n_local; i += gridDim.x * 256u)
        for (uint32_t r = 0; r < min(K, 2u); r++)
        {
            const unsigned long long c = cand[(size_t)i * K + r];
            if ((uint32_t)c == 0xFFFFFFFFu) break;  // valid ranks come first
            const unsigned int e = (unsigned int)(c >> 32);
            lo = min(lo, e), hi = max(hi, e), cnt++;
        }
    lo  = __reduce_min_sync(0xffffffffu, lo);
    hi  = __reduce_max_sync(0xffffffffu, hi);
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt)
    {
        atomicMin(&res->emin_bits, lo), atomicMax(&res->emax_bits, hi);
        atomicAdd(&res->n_samples, (unsigned long long)cnt);
    }
}

__global__ void __launch_bounds__(256)
    k_ad_hist(GridView g, const unsigned long long* __restrict__ cand, uint32_t n_local, uint32_t K, const uint32_t* __restrict__ bbox,
              uint32_t* __restrict__ bbox_next, float gate_eps, AdResult* __restrict__ res)
{
    __shared__ unsigned int sh[kAdBins];
    bbox_rearm(bbox_next);  // phase 2 may never run (gate closed, reference assertion): the next call's slot is re-armed here
    if (threadIdx.x < kAdBins) sh[threadIdx.x] = 0u;
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) res->gate = bbox_gate(g, bbox, gate_eps) ? 1u : 0u;
    const double hmin = (double)__uint_as_float(res->emin_bits), hmax = (double)__uint_as_float(res->emax_bits);
    if (res->n_samples && hmax > hmin)
    {
        const double inv = __ddiv_rn((double)kAdBins - 1.0, __dsub_rn(hmax, hmin));  // CHistogram: (nBins - 1) / (max - min)
        for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < n_local; i += gridDim.x * 256u)
            for (uint32_t r = 0; r < min(K, 2u); r++)
            {
                const unsigned long long c = cand[(size_t)i * K + r];
                if ((uint32_t)c == 0xFFFFFFFFu) break;
                const double x = (double)__uint_as_float((uint32_t)(c >> 32));
                if (x < hmin || x > hmax) continue;
                const unsigned long long b = (unsigned long long)__double2ull_rz(__dmul_rn(inv, __dsub_rn(x, hmin)));
                atomicAdd(&sh[min(b, (unsigned long long)(kAdBins - 1))], 1u);
            }
    }
    __syncthreads();
    if (threadIdx.x < kAdBins && sh[threadIdx.x]) atomicAdd(&res->bins[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

struct AdDecideArgs
{
    uint32_t n_local, K, Kp;  // neighbours kept per local, pt2pt slots per local (Kp <= K)
    int      planes;
    uint32_t planeMinFound;
    double   planeEigenThreshold, planeMinimumDistance;
    double   maxCorrDistSqr;
    float    maxSqr1to2;
    int      allowGlobal;
};

template <int KT>
__global__ void __launch_bounds__(128)
    k_ad_decide(GridView g, AdDecideArgs a, const float* __restrict__ lx, const float* __restrict__ ly, const float* __restrict__ lz,
                const unsigned long long* __restrict__ cand, const uint32_t* __restrict__ gbits, unsigned long long* __restrict__ sel,
                PlaneCandidate* __restrict__ plc, uint8_t* __restrict__ ok_flags)
{
    const uint32_t i = blockIdx.x * 128u + threadIdx.x;
    if (i >= a.n_local) return;
    unsigned long long c[KT];
    int                cnt = 0;
#pragma unroll
    for (int k = 0; k < KT; k++)
    {
        c[k] = (k < (int)a.K) ? cand[(size_t)i * a.K + k] : ~0ull;
        cnt += ((uint32_t)c[k] != 0xFFFFFFFFu);
    }
    uint8_t plane_ok = 0;
    if (a.planes && cnt >= (int)a.planeMinFound)  // :222-268
    {
        float px[KT], py[KT], pz[KT];
#pragma unroll
        for (int k = 0; k < KT; k++)
            if (k < cnt)
            {
                const float4 p = __ldg(g.pts_orig + (uint32_t)c[k]);
                px[k] = p.x, py[k] = p.y, pz[k] = p.z;
            }
        PlaneCandidate pc;
        if (fit_plane_adaptive<KT>(px, py, pz, cnt, lx[i], ly[i], lz[i], a.planeEigenThreshold, a.planeMinimumDistance, pc))
        {
            plc[i]   = pc;
            plane_ok = 1;
        }
    }
    if (ok_flags) ok_flags[i] = plane_ok;
    // :270-297
    const float e0     = __uint_as_float((uint32_t)(c[0] >> 32));
    bool        broken = plane_ok != 0;
#pragma unroll
    for (int k = 0; k < KT; k++)
        if (k < (int)a.Kp)
        {
            unsigned long long w = ~0ull;
            if (!broken && k < cnt)
            {
                const float e    = __uint_as_float((uint32_t)(c[k] >> 32));
                bool        emit = true;
                if (!a.allowGlobal && bit_set(gbits, (uint32_t)c[k])) emit = false;  // :276-278
                if (emit && (double)e >= a.maxCorrDistSqr) emit = false;               // :281
                if (emit && k != 0 && e > __fmul_rn(e0, a.maxSqr1to2)) emit = false, broken = true;  // :283-287
                if (emit) w = c[k];
            }
            sel[(size_t)i * a.Kp + k] = w;
        }
}
}  // namespace

int run_adaptive_search(mp2p_b200_ctx* ctx, mp2p_b200_map* map, const float* lx, const float* ly, const float* lz, uint64_t n_local,
                        int local_on_device, const double pose[12], const mp2p_b200_adaptive_params* prm, const uint32_t* lbits,
                        uint64_t hist_out[MP2P_B200_ADAPTIVE_BINS], double* err_min, double* err_max, uint64_t* n_samples,
                        int* gate_out)
{
    auto& S = ctx->adaptive;
    S       = mp2p_b200_ctx::AdaptiveState{};
    for (int b = 0; b < kAdBins; b++) hist_out[b] = 0;
    *err_min = *err_max = 0.0, *n_samples = 0, *gate_out = 0;
    if (!ctx->owns_map(map))
    {
        set_error("map handle is not a live map of this context (destroyed, or created on another context)");
        return MP2P_B200_ERR_ARG;
    }
    const uint64_t nmap = map->view.n_points;
    if (nmap == 0 || n_local == 0) return 0;  // :71
    const uint32_t nnMax = prm->enableDetectPlanes ? prm->planeSearchPoints : prm->maxPt2PtCorrespondences;  // :119-120
    const uint32_t K     = std::min<uint32_t>(nnMax, 10u);                                                    // MAX_CORRS_PER_LOCAL
    if (n_local * (uint64_t)std::max(K, 1u) >= 0xFFFFFFFFull || K < 1)
    {
        set_error("match_adaptive: need 1 <= neighbours per point and n_local * neighbours < 2^32-1");
        return MP2P_B200_ERR_ARG;
    }
    cudaStream_t st = ctx->stream;
    MP2P_TRY(stage_local(ctx, lx, ly, lz, n_local, local_on_device));
    const uint32_t* d_lbits;
    MP2P_TRY(upload_bits(ctx, ctx->d_lbits, lbits, n_local, &d_lbits));
    const uint64_t      n_tiles = (n_local + kScanTile - 1) / kScanTile;
    SmallView           sv;
    unsigned long long* status;
    MP2P_TRY(prepare_small(ctx, n_tiles, sv, &status));
    MP2P_TRY(ctx->d_cand.ensure(n_local * K * 8));
    MP2P_TRY(ctx->d_adres.ensure(sizeof(AdResult)));
    const float absMaxSqr = (float)(prm->absoluteMaxSearchDistance * prm->absoluteMaxSearchDistance);  // :87
    Pt2PtArgs   a{};
    for (int k = 0; k < 12; k++) a.pose.m[k] = pose[k];
    // nn_single_search keeps d2 <= absMax^2 (:166), nn_radius_search d2 < absMax^2; the kernels test d2 < maxDistSq
    a.maxDistSq = nnMax == 1 ? nextafterf(absMaxSqr, __builtin_inff()) : absMaxSqr;
    a.angSq     = 0.f;
    a.n_local = (uint32_t)n_local, a.K = K;
    a.allowLocal = prm->allowMatchAlreadyMatchedPoints, a.allowGlobal = 1;  // no first-claim dedup in this matcher
    a.tma_ok   = ctx->cur_tma_ok;
    a.rl_start = start_level(map->view, K);
    a.n_phases = knn_phases();
    auto*               cand  = ctx->d_cand.as<unsigned long long>();
    unsigned long long* stats = nullptr;
    MP2P_TRY(prepare_stats(ctx, &stats));
    prof_begin(ctx, 0);
    if (K == 1)
    {
        MP2P_TRY(ctx->d_candxyz.ensure(n_local * sizeof(float4)));
        MP2P_LAUNCH_NN1((uint32_t)((n_local + kNN1Threads - 1) / kNN1Threads), st, map->view, a, ctx->cur_qx, ctx->cur_qy, ctx->cur_qz,
                        ctx->cur_perm, d_lbits, nullptr, nullptr, cand, ctx->d_candxyz.as<float4>(), sv.bbox, stats);
    }
    else
    {
#define LAUNCH_MATCH(G)                                                                                    \
    {                                                                                                      \
        MP2P_LAUNCH_KMATCH(G, a, n_local, st, map->view, a, ctx->cur_qx, ctx->cur_qy, ctx->cur_qz, ctx->cur_perm, d_lbits, nullptr, nullptr, cand, sv.bbox, stats, FitList{nullptr, nullptr, nullptr, 0}) \
    }
        MP2P_DISPATCH_G(K, LAUNCH_MATCH)
#undef LAUNCH_MATCH
    }
    prof_end(ctx, 0);
    count_launch(ctx);
    auto* res = ctx->d_adres.as<AdResult>();
    {
        AdResult init{};
        init.emin_bits = 0xFFFFFFFFu;
        AdResult* h = reinterpret_cast<AdResult*>(static_cast<char*>(ctx->h_pinned) + 3072);
        *h          = init;
        MP2P_CUDA_TRY(cudaMemcpyAsync(res, h, sizeof(AdResult), cudaMemcpyHostToDevice, st));
    }
    const uint32_t nb       = (uint32_t)std::min<uint64_t>((n_local + 255) / 256, 148 * 8);
    const float    gate_eps = (float)prm->bounding_box_intersection_check_epsilon;  // :77-80
    k_ad_minmax<<<nb, 256, 0, st>>>(cand, (uint32_t)n_local, K, res);
    k_ad_hist<<<nb, 256, 0, st>>>(map->view, cand, (uint32_t)n_local, K, sv.bbox, sv.bbox_next, gate_eps, res);
    count_launch(ctx, 2);
    AdResult* h = reinterpret_cast<AdResult*>(static_cast<char*>(ctx->h_pinned) + 3072);
    MP2P_CUDA_TRY(cudaMemcpyAsync(h, res, sizeof(AdResult), cudaMemcpyDeviceToHost, st));
    MP2P_CUDA_TRY(cudaStreamSynchronize(st));
    MP2P_CUDA_TRY(cudaGetLastError());
    *gate_out = (int)h->gate;
    *n_samples = h->n_samples;
    if (h->n_samples)
    {
        float lo, hi;
        std::memcpy(&lo, &h->emin_bits, 4), std::memcpy(&hi, &h->emax_bits, 4);
        *err_min = lo, *err_max = hi;
        for (int b = 0; b < kAdBins; b++) hist_out[b] = h->bins[b];
    }
    S.valid = true, S.map = map, S.n_local = n_local, S.K = K, S.sv_bbox = sv.bbox, S.sv_bbox_next = sv.bbox_next;
    S.sv_count = sv.count, S.sv_tile_counter = sv.tile_counter, S.status = status, S.scan_epoch = ctx->scan_epoch;
    return 0;
}

int run_adaptive_emit(mp2p_b200_ctx* ctx, mp2p_b200_map* map, const mp2p_b200_adaptive_params* prm, double maxCorrDistSqr,
                      const uint32_t* gbits, mp2p_b200_pair_pt2pt* out2p, uint64_t cap2p, mp2p_b200_pair_pt2pl* out2l,
                      uint64_t cap2l, int out_on_device, uint64_t* n2p, uint64_t* n2l)
{
    *n2p = *n2l = 0;
    ctx->last2p.valid = false, ctx->last2p.sums = nullptr, ctx->last2l.valid = false, ctx->spec_res.valid = false;
    auto& S = ctx->adaptive;
    if (!S.valid || S.map != map)
    {
        set_error("adaptive_emit: no matching adaptive_search on this context / map (or it reported gate = 0)");
        return MP2P_B200_ERR_ARG;
    }
    S.valid                 = false;
    cudaStream_t   st       = ctx->stream;
    const uint64_t n_local  = S.n_local;
    const uint32_t K        = S.K;
    const uint32_t Kp       = std::min<uint32_t>(std::max<uint32_t>(prm->maxPt2PtCorrespondences, 1u), K);
    const uint64_t nmap     = map->view.n_points;
    const uint32_t* d_gbits;
    MP2P_TRY(upload_bits(ctx, ctx->d_gbits, gbits, nmap, &d_gbits));
    MP2P_TRY(ctx->d_adsel.ensure(n_local * Kp * 8));
    MP2P_TRY(ctx->d_plcand.ensure(n_local * sizeof(PlaneCandidate)));
    MP2P_TRY(ctx->d_okflags.ensure(n_local));
    AdDecideArgs d{};
    d.n_local = (uint32_t)n_local, d.K = K, d.Kp = Kp;
    d.planes = prm->enableDetectPlanes, d.planeMinFound = prm->planeMinimumFoundPoints;
    d.planeEigenThreshold = prm->planeEigenThreshold, d.planeMinimumDistance = prm->planeMinimumDistance;
    d.maxCorrDistSqr = maxCorrDistSqr;
    d.maxSqr1to2     = (float)(prm->firstToSecondDistanceMax * prm->firstToSecondDistanceMax);  // :216
    d.allowGlobal    = prm->allowMatchAlreadyMatchedGlobalPoints;
    auto*          cand = ctx->d_cand.as<unsigned long long>();
    auto*          sel  = ctx->d_adsel.as<unsigned long long>();
    auto*          plc  = ctx->d_plcand.as<PlaneCandidate>();
    auto*          okf  = ctx->d_okflags.as<uint8_t>();
    const uint32_t nbd  = (uint32_t)((n_local + 127) / 128);
    prof_begin(ctx, 1);
#define LAUNCH_DECIDE(KT) k_ad_decide<KT><<<nbd, 128, 0, st>>>(map->view, d, ctx->cur_lx, ctx->cur_ly, ctx->cur_lz, cand, d_gbits, sel, plc, okf)
    if (K <= 1)
        LAUNCH_DECIDE(1);
    else if (K <= 4)
        LAUNCH_DECIDE(4);
    else if (K <= 8)
        LAUNCH_DECIDE(8);
    else
        LAUNCH_DECIDE(10);
#undef LAUNCH_DECIDE
    count_launch(ctx);
    const float gate_eps = (float)prm->bounding_box_intersection_check_epsilon;
    // ---- point-to-plane pairings (ascending local index)
    mp2p_b200_pair_pt2pl* d_out2l = out2l;
    const uint64_t        capl    = std::min<uint64_t>(cap2l, n_local);
    if (prm->enableDetectPlanes)
    {
        if (!out_on_device)
        {
            MP2P_TRY(ctx->d_out2l.ensure(std::max<uint64_t>(capl, 1) * sizeof(mp2p_b200_pair_pt2pl)));
            d_out2l = ctx->d_out2l.as<mp2p_b200_pair_pt2pl>();
        }
        k_compact_pt2pl<<<compact_tiles_2l(n_local), kScanThreads, 0, st>>>(map->view, (uint32_t)n_local, gate_eps, capl, ctx->cur_lx, ctx->cur_ly,
                                                                   ctx->cur_lz, plc, okf, S.sv_bbox, S.sv_bbox_next, S.status,
                                                                   S.sv_tile_counter, d_out2l, S.sv_count, S.scan_epoch, 0, nullptr, nullptr);
        count_launch(ctx);
    }
    // ---- point-to-point pairings: ordinary compaction over the selected candidate words, its own
    //      status array and count word (two scans in one call)
    const uint64_t n_slots  = n_local * Kp;
    const uint64_t n_tiles2 = (n_slots + kScanTile - 1) / kScanTile;
    if ((n_tiles2 + 1) * 8 > ctx->d_scan2.bytes)
    {
        MP2P_TRY(ctx->d_scan2.ensure((n_tiles2 + 1) * 8));
        MP2P_CUDA_TRY(cudaMemsetAsync(ctx->d_scan2.p, 0, ctx->d_scan2.bytes, st));
    }
    auto*                 count2  = reinterpret_cast<unsigned long long*>(ctx->d_small.as<char>() + 80);
    mp2p_b200_pair_pt2pt* d_out2p = out2p;
    const uint64_t        capp    = std::min<uint64_t>(cap2p, n_slots);
    if (!out_on_device)
    {
        MP2P_TRY(ctx->d_out2p.ensure(std::max<uint64_t>(capp, 1) * sizeof(mp2p_b200_pair_pt2pt)));
        d_out2p = ctx->d_out2p.as<mp2p_b200_pair_pt2pt>();
    }
    CompactArgs c{};
    c.n_local = (uint32_t)n_local, c.K = Kp, c.allowGlobal = 1, c.tag = 0;
    c.gate_eps = gate_eps, c.capacity = capp, c.scan_epoch = S.scan_epoch;
    k_compact_pt2pt<<<compact_tiles_2p(n_slots), kScanThreads, 0, st>>>(map->view, c, ctx->cur_lx, ctx->cur_ly, ctx->cur_lz, nullptr,
                                                                map->d_claim.as<unsigned long long>(), sel,
                                                                K == 1 ? ctx->d_candxyz.as<float4>() : nullptr, S.sv_bbox, S.sv_bbox_next,
                                                                ctx->d_scan2.as<unsigned long long>(), S.sv_tile_counter, d_out2p, count2,
                                                                FusedSums{nullptr, nullptr, nullptr});
    prof_end(ctx, 1);
    count_launch(ctx);
    // ---- counts and records back
    unsigned long long* hc = reinterpret_cast<unsigned long long*>(static_cast<char*>(ctx->h_pinned) + 3072 + 512);
    hc[0] = hc[1] = 0;
    MP2P_CUDA_TRY(cudaMemcpyAsync(hc, count2, 8, cudaMemcpyDeviceToHost, st));
    if (prm->enableDetectPlanes) MP2P_CUDA_TRY(cudaMemcpyAsync(hc + 1, S.sv_count, 8, cudaMemcpyDeviceToHost, st));
    MP2P_CUDA_TRY(cudaStreamSynchronize(st));
    MP2P_CUDA_TRY(cudaGetLastError());
    *n2p = hc[0], *n2l = hc[1];
    if (*n2p > capp || *n2l > capl)
    {
        set_error("match_adaptive: output capacity too small (%llu pt2pt, %llu pt2pl pairings)", (unsigned long long)*n2p,
                  (unsigned long long)*n2l);
        return MP2P_B200_ERR_CAPACITY;
    }
    if (!out_on_device)
    {
        if (*n2p) MP2P_CUDA_TRY(cudaMemcpyAsync(out2p, d_out2p, *n2p * sizeof(mp2p_b200_pair_pt2pt), cudaMemcpyDeviceToHost, st));
        if (*n2l) MP2P_CUDA_TRY(cudaMemcpyAsync(out2l, d_out2l, *n2l * sizeof(mp2p_b200_pair_pt2pl), cudaMemcpyDeviceToHost, st));
        MP2P_CUDA_TRY(cudaStreamSynchronize(st));
        ctx->last2p.dev = d_out2p, ctx->last2p.n = *n2p, ctx->last2p.valid = true;
        ctx->last2l.dev = d_out2l, ctx->last2l.n = *n2l, ctx->last2l.valid = prm->enableDetectPlanes != 0;
    }
    return 0;
}
}  // namespace mp2p
