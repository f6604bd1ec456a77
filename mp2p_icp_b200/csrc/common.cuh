// Shared declarations of the B200 hot-path library (product code; never includes oracle/).
#pragma once
#include <unordered_set>
#include <vector>
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "mp2p_b200.h"

namespace mp2p
{
void set_error(const char* fmt, ...);

#define MP2P_CUDA_TRY(expr)                                                                   \
    do                                                                                        \
    {                                                                                         \
        cudaError_t e_ = (expr);                                                              \
        if (e_ != cudaSuccess)                                                                \
        {                                                                                     \
            ::mp2p::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                   \
                              cudaGetErrorString(e_));                                        \
            return MP2P_B200_ERR_CUDA;                                                        \
        }                                                                                     \
    } while (0)

#define MP2P_TRY(expr)            \
    do                            \
    {                             \
        int rc_ = (expr);         \
        if (rc_ != 0) return rc_; \
    } while (0)

constexpr int      kGridBits  = 21;  // finest quantisation: 2^21 cells along the longest axis
constexpr int      kMaxLevels = 22;  // levels 0..21 (level L: cell = s0 * 2^L)
constexpr uint32_t kQueryTile = 256; // queries per CTA tile (TMA bulk copies of 1 KiB per axis)

// One occupied voxel of one level: 16 bytes, fetched with a single 128-bit load.
struct __align__(16) CellEntry
{
    unsigned long long key;  // cx | cy<<21 | cz<<42 ; ~0 = empty slot
    uint32_t           start;  // first point (Morton-sorted order)
    uint32_t           count;
};
constexpr unsigned long long kEmptyKey = ~0ull;

// Read-only view of a map index, passed BY VALUE to kernels (lives in constant/param space).
struct GridView
{
    const float4*    pts;       // Morton-sorted points: x,y,z, original index (int bits)
    const float4*    pts_orig;  // original order: x,y,z,0  (for emitting TMatchingPair::global)
    const CellEntry* table;     // all level tables back to back
    // per table slot: the TIGHT box of the voxel's points, 8 bits per bound relative to the voxel (box_lo / box_hi
    // in grid_search.cuh). A voxel is a cube, the points in it usually a patch of surface: its cube says a far
    // query must look at it, its box says it need not. NULL = not built (the default; $MP2P_INDEX_BOX=1 builds them)
    const uint2*     box;
    float            ox, oy, oz;  // grid origin = map bbox min
    float            inv_s0;      // 1 / finest quantum
    float            s0_lo;       // finest quantum, rounded DOWN (conservative bounds)
    int              level_first; // absolute level of table 0
    int              n_levels;    // tables for levels level_first .. level_first+n_levels-1
    uint32_t         level_off[kMaxLevels];    // entry offset of each table
    uint32_t         level_shift[kMaxLevels];  // 64 - log2(capacity)
    float            bbmin[3], bbmax[3];
    uint32_t         n_points;
    float            level_occupancy[kMaxLevels];  // mean points per occupied voxel, per table
};

class PageableCopier;

struct DevBuf
{
    void*  p     = nullptr;
    size_t bytes = 0;
    int    ensure(size_t need)
    {
        if (need <= bytes) return 0;
        if (p) cudaFree(p);
        p     = nullptr;
        bytes = 0;
        // grow geometrically to avoid re-allocations when cloud sizes jitter between calls
        size_t want = need + need / 4 + 256;
        if (cudaMalloc(&p, want) != cudaSuccess)
        {
            set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(cudaGetLastError()));
            return MP2P_B200_ERR_NOMEM;
        }
        bytes = want;
        return 0;
    }
    void release()
    {
        if (p) cudaFree(p);
        p     = nullptr;
        bytes = 0;
    }
    template <class T>
    T* as() const
    {
        return static_cast<T*>(p);
    }
};

}  // namespace mp2p

// ---- opaque handles -------------------------------------------------------------------------
struct mp2p_b200_ctx
{
    int          device     = 0;
    cudaStream_t stream     = nullptr;
    bool         own_stream = false;
    uint64_t     launches   = 0;
    uint32_t     scan_epoch = 0;  // stamps look-back status words (match.cu)
    uint64_t     hint_pt2pt = ~0ull, hint_pt2pl = ~0ull, hint_ir = ~0ull, hint_pt2ln = ~0ull;  // previous pairing counts (speculative D2H size)
    // pairing count a matcher call left on the device without reading it back (shard_resolve with
    // out_count == NULL), consumed by solver calls given n = MP2P_B200_COUNT_ON_DEVICE
    const unsigned long long* last_count    = nullptr;
    uint64_t                  last_capacity = 0;
    // device-resident copy of the pairings the last matcher call returned to the HOST (d_out2p /
    // d_out2l): a solver call given pairs_on_device = MP2P_B200_PAIRS_LAST_MATCH reads it instead of
    // uploading the same records again
    struct LastMatch
    {
        const void*   dev   = nullptr;
        uint64_t      n     = 0;
        bool          valid = false;
        const double* sums  = nullptr;  // pt2pt: HORN1 packet the compaction produced on the way (device)
    } last2p, last2l;
    // Speculative solve: when the previous solver call named the last matcher output
    // (PAIRS_LAST_MATCH), the NEXT matcher call that returns pairings to the host also enqueues that
    // solver over the device copy, on the compute stream, while the copy stream carries the records
    // to the host; the solver call that follows — same parameters, same start pose — finds the
    // result ready. A guess that does not come true costs a few idle-GPU microseconds; after two
    // unused guesses in a row the library stops guessing until the pattern shows again.
    struct SpecWant
    {
        int                   kind = 0;  // 0 none, 1 Solver_Horn (pt2pt), 2 Solver_GaussNewton
        int                   list = 0;  // GN: 1 = pt2pt list, 2 = pt2pl list
        mp2p_b200_horn_params horn{};
        mp2p_b200_gn_params   gn{};
    } spec_want;
    int spec_unused = 0;  // speculations in a row nobody asked for: two and the guessing stops
    struct SpecResult
    {
        bool     valid = false;
        bool     pending = false;  // enqueued on the compute stream, not yet synchronised (zero-copy matcher output)
        int      kind = 0, list = 0;
        uint64_t n = 0;
        double   pose_in[12] = {};
    } spec_res;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t  ev_fork = nullptr;
    // side stream of the k > 1 search's scheduling hint (k_tile_rank runs behind the search, match.cu)
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t  ev_rank_fork = nullptr, ev_rank = nullptr;
    mp2p::DevBuf d_spec;  // GN speculation: 12 doubles pose + state words
    // live handles created on this context: a stale or foreign map / cloud handle is refused instead of
    // being dereferenced (the objects are not thread-safe, like the reference's: no locking here)
    std::unordered_set<const void*> live_maps, live_clouds;
    // host layers the library keeps a device copy of (mp2p_b200_map_cached / mp2p_b200_cloud_cached)
    struct CachedLayer
    {
        const float*     x = nullptr;
        uint64_t         n = 0, fingerprint = 0, last_use = 0;
        mp2p_b200_map*   map   = nullptr;
        mp2p_b200_cloud* cloud = nullptr;
    };
    std::vector<CachedLayer> layer_cache;
    uint64_t                 layer_clock = 0;
    bool owns_map(const void* m) const { return m && live_maps.count(m) != 0; }
    bool owns_cloud(const void* c) const { return c && live_clouds.count(c) != 0; }
    // Matcher_Adaptive: what phase 2 (adaptive_emit) needs from phase 1 (adaptive_search)
    struct AdaptiveState
    {
        bool                valid   = false;
        mp2p_b200_map*      map     = nullptr;
        uint64_t            n_local = 0;
        uint32_t            K       = 0;
        uint32_t *          sv_bbox = nullptr, *sv_bbox_next = nullptr, *sv_tile_counter = nullptr;
        unsigned long long *sv_count = nullptr, *status = nullptr;
        uint32_t            scan_epoch = 0;
    } adaptive;
    mp2p::DevBuf d_adres, d_adsel, d_scan2;  // phase-1 result block, selected pt2pt candidate words, 2nd scan status array
    // grid barrier state of the single-launch iteration (match.cu): counters only ever grow
    mp2p::DevBuf       d_coop;
    unsigned long long coop_arrivals = 0;
    unsigned int       coop_epoch    = 0;
    cudaEvent_t  ev0 = nullptr, ev1 = nullptr;
    // measurement hooks
    bool         prof_timings = false, prof_stats = false;
    cudaEvent_t  pev[16]      = {};     // pairs (2k, 2k+1) bracket timing slot k
    bool         pev_used[8]  = {};
    float        timings[MP2P_B200_N_TIMINGS] = {};
    mp2p::DevBuf d_stats;               // 8 x u64 search counters
    mp2p::DevBuf d_trace;               // measurement hook ($MP2P_KNN_TRACE): per-CTA {SM, start, end, tile} of the last k > 1 search
    uint32_t     trace_tiles = 0;

    // matcher scratch
    const float *cur_lx = nullptr, *cur_ly = nullptr, *cur_lz = nullptr;  // local cloud, caller's order
    // what the SEARCH kernels read: the same arrays, or a resident cloud's Morton-sorted copy
    // (then cur_perm[j] = caller's index of sorted position j)
    const float *   cur_qx = nullptr, *cur_qy = nullptr, *cur_qz = nullptr;
    const uint32_t* cur_perm   = nullptr;
    mp2p_b200_cloud* cur_cloud = nullptr;  // the resident cloud being searched (scheduling hint), or NULL
    bool            cur_tma_ok = false;
    mp2p::DevBuf d_lx, d_ly, d_lz;     // local cloud staging (padded to kQueryTile)
    mp2p::DevBuf d_cand;               // u64 [n_local*K]  (d2 bits << 32 | map index)
    mp2p::DevBuf d_candxyz;            // float4 [n_local] coordinates of the K = 1 candidate
    mp2p::DevBuf d_lbits, d_gbits;     // MatchState bitfields
    mp2p::DevBuf d_scan;               // tile status words + counters
    mp2p::DevBuf d_small;              // bbox (6 u32) + count (u64) + misc
    mp2p::DevBuf d_out2p, d_out2l;     // compacted pairs when the caller wants them on the host
    mp2p::DevBuf d_plcand, d_okflags;  // per-query plane candidates + accepted flags (pt2pl)
    mp2p::DevBuf d_fitlist;            // [0] count, [1..] queries that qualify for a plane fit
    mp2p::DevBuf d_defer;              // [0], [1] counters used in turn, [2..] queries the thread-per-query search handed over
    uint32_t     defer_turn = 0;
    mp2p::DevBuf d_knn_idx, d_knn_d2, d_knn_found;
    mp2p::DevBuf d_irk0, d_irk1, d_irv0, d_irv1, d_irtmp;  // Matcher_Points_InlierRatio: sort keys / values / scratch
    // solver scratch
    mp2p::DevBuf d_pairs2p, d_pairs2l, d_pairs2ln; // H2D staging of host pairings
    mp2p::DevBuf d_partials;           // per-block partial sums
    mp2p::DevBuf d_packet;             // 4 packets of 32 doubles
    mp2p::DevBuf d_pose;               // 12 doubles + flags
    mp2p::DevBuf d_weights;            // run-length point weights
    mp2p::DevBuf d_outlier;            // Horn scale-outlier flags
    mp2p::DevBuf d_conv;               // pt2pl -> pt2pt conversion scratch and output
    // FilterDecimateVoxels scratch (filter.cu): extrema + count | sort keys a/b | values a/b | flags + tile sums | radix scratch | staged input / output
    mp2p::DevBuf d_fd_small, d_fd_keys, d_fd_vals, d_fd_flags, d_fd_rs, d_fd_in, d_fd_out;
    // pinned host scratch
    void* h_pinned = nullptr;  // 4 KiB: counts, packets, poses
    mp2p::PageableCopier* copier = nullptr;  // pageable host buffers: bounce buffer + helper threads (hostcopy.hpp), made on first use
    void* h_pinned_dev = nullptr;  // its device alias (kernels that hand a count to the host themselves)
    // 1 KiB of MAPPED pinned memory a kernel writes results into directly (host view / device view)
    double* h_mapped = nullptr;
    double* h_mapped_dev = nullptr;
};

struct mp2p_b200_map
{
    mp2p_b200_ctx*     ctx = nullptr;
    mp2p::GridView     view{};
    mp2p_b200_map_info info{};
    mp2p::DevBuf       d_pts, d_pts_orig, d_table, d_box, d_claim;
    uint32_t           epoch = 0;  // first-claim epoch (see match.cu)
};

// A local cloud kept resident on the device for a whole align(): caller-order SoA plus a
// Morton-sorted copy (spatially coherent warps in the search kernels) and the permutation.
struct mp2p_b200_cloud
{
    mp2p_b200_ctx* ctx = nullptr;
    uint64_t       n   = 0;
    mp2p::DevBuf   d_x, d_y, d_z;     // caller order
    mp2p::DevBuf   d_sx, d_sy, d_sz;  // sorted by Morton code of the cloud's own bounding box
    mp2p::DevBuf   d_perm;            // u32 [n]: sorted position -> caller index
    float          build_ms = 0.f;
    // scheduling hint of the k > 1 search (match.cu): how long every query tile took in the previous call
    // over this cloud, and the tile order (longest first) derived from it for the next one
    mp2p::DevBuf   d_tile_cost, d_tile_order;
    uint64_t       hint_key = 0;  // (tiles, lanes per query, CTA size) the order was made for; 0 = none
};

namespace mp2p
{
inline void count_launch(mp2p_b200_ctx* c, uint64_t n = 1) { c->launches += n; }
// timing slot brackets (no-ops unless profiling is on)
inline void prof_begin(mp2p_b200_ctx* c, int slot)
{
    if (c->prof_timings) cudaEventRecord(c->pev[2 * slot], c->stream);
}
inline void prof_end(mp2p_b200_ctx* c, int slot)
{
    if (c->prof_timings) cudaEventRecord(c->pev[2 * slot + 1], c->stream), c->pev_used[slot] = true;
}
void prof_reset(mp2p_b200_ctx* c);    // api.cu: start of a public call
void prof_collect(mp2p_b200_ctx* c);  // api.cu: after the call's final synchronize

// result of a matcher call whose pairings stay on the device (fused iteration path)
struct DeviceMatch
{
    const unsigned long long* d_count  = nullptr;  // number of pairings, device memory
    const void*               d_pairs  = nullptr;  // compacted records, device memory
    uint64_t                  capacity = 0;        // upper bound of *d_count
    double*                   want_horn_sums = nullptr;  // in: device packet to receive the HORN1 sums
    // in: > 0 = pair weight pt2pt: the caller also wants the HORN2 moments (plain Horn: no robust
    // kernel, weights or scale-outlier pass) in want_horn_sums[32..64) if the matcher can fuse them
    double                    fuse_moments_w = 0.0;
    bool                      moments_done   = false;    // out: the matcher produced the moments too
    // in: query-sharded single-launch iteration (peer.cu): this call is the shard `peer->view.rank` of
    // a cloud cut into blocks of per_shard points; ONLY the single launch is acceptable — if it cannot
    // be used run_match_pt2pt returns 1 without having enqueued anything and the caller takes the
    // multi-kernel path (same mailbox protocol, so ranks may differ in their choice)
    struct mp2p_b200_peer*    peer      = nullptr;
    uint64_t                  per_shard = 0;
};

// api.cu
int read_iteration_packets(mp2p_b200_ctx* ctx, bool polled, const double* d_packets, double** hp_out);
// index.cu
int build_index(mp2p_b200_ctx* ctx, mp2p_b200_map* map, const float* x, const float* y,
                const float* z, uint64_t n, int on_device);
int build_cloud(mp2p_b200_ctx* ctx, mp2p_b200_cloud* cloud, const float* x, const float* y, const float* z,
                uint64_t n, int on_device);
// match.cu — `local_on_device`: 0 host arrays, 1 device arrays, 2 = lx is a mp2p_b200_cloud*
int run_match_pt2pt(mp2p_b200_ctx* ctx, mp2p_b200_map* map, const float* lx, const float* ly,
                    const float* lz, uint64_t n_local, int local_on_device, const double pose[12],
                    const mp2p_b200_pt2pt_params* prm, const uint32_t* lbits, const uint32_t* gbits,
                    mp2p_b200_pair_pt2pt* out, uint64_t capacity, int out_on_device,
                    uint64_t* out_count, DeviceMatch* keep_on_device = nullptr);
// Matcher_Point2Line rides on the pt2pl pipeline (k-NN search -> per-query fit -> compaction)
struct LineMode
{
    uint32_t minimumLinePoints;
    double   lineEigenThreshold;
};
int run_match_pt2pl(mp2p_b200_ctx* ctx, mp2p_b200_map* map, const float* lx, const float* ly,
                    const float* lz, uint64_t n_local, int local_on_device, const double pose[12],
                    const mp2p_b200_pt2pl_params* prm, const uint32_t* lbits,
                    mp2p_b200_pair_pt2pl* out, uint64_t capacity, int out_on_device,
                    uint64_t* out_count, DeviceMatch* keep_on_device = nullptr, const LineMode* line = nullptr);
uint64_t shard_record_words(uint64_t per_shard, uint32_t K);
// OWNER-PARTITIONED first claims of a query-sharded run (peer.cu): global point g belongs to rank g % world,
// which keeps its claim word at parts[g % world][g / world] in memory every rank has mapped (CUDA IPC over
// NVLink). Proposals are system-scope atomicMin's straight into the owner's HBM, acceptance reads the word back
// from there: nothing is gathered and nothing replayed, per-GPU work is the shard's own proposals.
struct OwnerClaims
{
    unsigned long long* const* parts = nullptr;  // device array of `world` pointers
    uint32_t                   world = 0;
    unsigned long long         tag   = 0;        // (0xFFFFFFFF - epoch) << 32, the same on every rank
};
int run_shard_search_pt2pt(mp2p_b200_ctx* ctx, mp2p_b200_map* map, const float* lx, const float* ly,
                           const float* lz, uint64_t n_local, int local_on_device, const double pose[12],
                           const mp2p_b200_pt2pt_params* prm, const uint32_t* lbits, uint64_t per_shard,
                           unsigned long long* d_record, const OwnerClaims* oc = nullptr, uint32_t shard_rank = 0);
int run_shard_resolve_pt2pt(mp2p_b200_ctx* ctx, mp2p_b200_map* map, uint64_t n_local, uint32_t shard_rank,
                            uint32_t n_shards, uint64_t per_shard, const unsigned long long* d_records,
                            const mp2p_b200_pt2pt_params* prm, const uint32_t* gbits,
                            mp2p_b200_pair_pt2pt* out, uint64_t capacity, int out_on_device,
                            uint64_t* out_count, double* d_horn_sums, const OwnerClaims* oc = nullptr);
int run_match_inlier_ratio(mp2p_b200_ctx* ctx, mp2p_b200_map* map, const float* lx, const float* ly, const float* lz,
                           uint64_t n_local, int local_on_device, const double pose[12], double ratio, int allowLocal,
                           int allowGlobal, double bbox_eps, const uint32_t* lbits, const uint32_t* gbits,
                           mp2p_b200_pair_pt2pt* out, uint64_t capacity, int out_on_device, uint64_t* out_count);
int run_adaptive_search(mp2p_b200_ctx* ctx, mp2p_b200_map* map, const float* lx, const float* ly, const float* lz, uint64_t n_local,
                        int local_on_device, const double pose[12], const mp2p_b200_adaptive_params* prm, const uint32_t* lbits,
                        uint64_t hist_out[MP2P_B200_ADAPTIVE_BINS], double* err_min, double* err_max, uint64_t* n_samples,
                        int* gate_out);
int run_adaptive_emit(mp2p_b200_ctx* ctx, mp2p_b200_map* map, const mp2p_b200_adaptive_params* prm, double maxCorrDistSqr,
                      const uint32_t* gbits, mp2p_b200_pair_pt2pt* out2p, uint64_t cap2p, mp2p_b200_pair_pt2pl* out2l,
                      uint64_t cap2l, int out_on_device, uint64_t* n2p, uint64_t* n2l);
int run_knn(mp2p_b200_ctx* ctx, const mp2p_b200_map* map, const float* qx, const float* qy,
            const float* qz, uint64_t nq, uint32_t k, float radius2, uint32_t* out_idx,
            float* out_d2, int32_t* out_found);
// filter.cu — FilterDecimateVoxels over device arrays; synchronises, *h_count = points produced
// Host <-> device copies of caller buffers: pinned memory and small transfers go straight to cudaMemcpyAsync,
// large PAGEABLE buffers through the context's bounce buffer and helper threads (hostcopy.hpp).
// copy_to_host_sync = cudaMemcpyAsync(DeviceToHost, st) + cudaStreamSynchronize(st);
// copy_to_device    = cudaMemcpyAsync(HostToDevice, st): the source may be reused when it returns.
int copy_to_host_sync(mp2p_b200_ctx* ctx, void* dst, const void* src_dev, size_t bytes, cudaStream_t st);
int copy_to_device(mp2p_b200_ctx* ctx, void* dst_dev, const void* src, size_t bytes, cudaStream_t st);
bool copy_wants_helpers(mp2p_b200_ctx* ctx, const void* host, size_t bytes);

int run_decimate_voxels(mp2p_b200_ctx* ctx, const float* dx, const float* dy, const float* dz, uint64_t n,
                        const mp2p_b200_decimate_params* prm, float* d_ox, float* d_oy, float* d_oz, long long* d_osrc,
                        uint64_t capacity, uint64_t* h_count);
// solve.cu
// `d_n*` (optional): pair counts read from DEVICE memory at kernel time (n* then are upper bounds
// used for the grid size) — lets a solver be enqueued behind a matcher without a host round trip.
int run_gn_accumulate(mp2p_b200_ctx* ctx, const mp2p_b200_pair_pt2pt* d2p, uint64_t n2p,
                      const mp2p_b200_pair_pt2pl* d2l, uint64_t n2l, const mp2p_b200_gn_params* prm,
                      const double* d_pose, double* d_packet, const unsigned long long* d_n2p = nullptr,
                      const unsigned long long* d_n2l = nullptr, const uint32_t* d_done = nullptr,
                      uint32_t* d_step_state = nullptr /* != NULL: the launch also applies the GN update */,
                      const mp2p_b200_pair_pt2ln* d2ln = nullptr, uint64_t n2ln = 0, double w_pt2ln = 1.0);
// whole inner loop of optimal_tf_gauss_newton on the device: (accumulate, solve+update) x maxIter,
// no host synchronisation; d_pose in/out, d_state = {done flag, iterations done}
int run_gn_device_loop(mp2p_b200_ctx* ctx, const mp2p_b200_pair_pt2pt* d2p, uint64_t n2p,
                       const mp2p_b200_pair_pt2pl* d2l, uint64_t n2l, const mp2p_b200_gn_params* prm,
                       double* d_pose, uint32_t* d_state, double* d_packet,
                       const unsigned long long* d_n2p = nullptr, const unsigned long long* d_n2l = nullptr,
                       const mp2p_b200_pair_pt2ln* d2ln = nullptr, uint64_t n2ln = 0, double w_pt2ln = 1.0);
int run_gn_coop_loop(mp2p_b200_ctx* ctx, const mp2p_b200_pair_pt2pt* d2p, uint64_t n2p, const mp2p_b200_pair_pt2pl* d2l, uint64_t n2l,
                     const mp2p_b200_gn_params* prm, double* d_pose, uint32_t* d_state, double* d_packet,
                     const unsigned long long* d_n2p, const unsigned long long* d_n2l, struct mp2p_b200_peer* peer);
int run_gn_step(mp2p_b200_ctx* ctx, const double* d_packet, const mp2p_b200_gn_params* prm, double* d_pose,
                uint32_t* d_state);
// pt2ln_pl_to_pt2pt (plane part) on the device; synchronises, *h_total = records kept
int run_pt2pl_to_pt2pt(mp2p_b200_ctx* ctx, const mp2p_b200_pair_pt2pl* d_in, uint64_t n, const double pose[12],
                       mp2p_b200_pair_pt2pt* d_out, uint64_t capacity, uint64_t* h_total);
int run_horn_sums(mp2p_b200_ctx* ctx, const mp2p_b200_pair_pt2pt* d2p, uint64_t n,
                  const uint8_t* d_outlier, double* d_packet, const unsigned long long* d_n = nullptr);
int run_horn_moments(mp2p_b200_ctx* ctx, const mp2p_b200_pair_pt2pt* d2p, uint64_t n,
                     const mp2p_b200_horn_params* prm, const double* d_sums_packet,
                     uint64_t n_total_pairs, const uint64_t* d_wcount_prefix, const double* d_wvalue,
                     uint32_t n_wblocks, uint8_t* d_outlier, double* d_packet,
                     const unsigned long long* d_n = nullptr, int n_total_mode = 0);
// covariance(): J^T J of the numerically differentiated error vector (solve.cu)
int run_cov_accumulate(mp2p_b200_ctx* ctx, const mp2p_b200_pair_pt2pt* d2p, uint64_t n2p, const mp2p_b200_pair_pt2pl* d2l,
                       uint64_t n2l, const mp2p_b200_pair_pt2ln* d2ln, uint64_t n2ln, const double poses[12][12],
                       const double inv2h[6], double* d_packet);
}  // namespace mp2p
