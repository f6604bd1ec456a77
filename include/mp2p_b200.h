/* mp2p_b200.h — C ABI of the B200-native mp2p_icp Matcher+Solver hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, `int` status codes, no
 * exceptions, no torch / MRPT types. The C++ plugin classes that derive from the reference's
 * `mp2p_icp::Matcher_Points_Base` / `mp2p_icp::Solver` (mp2p_icp_b200/host/mrpt_plugin.cpp, built
 * only where MRPT exists) and the MRPT-free host mirror (mp2p_icp_b200/host/) call exactly these
 * entry points. Every entry point names the reference interface it replaces (paths relative to the
 * reference checkout).
 *
 * Conventions
 *  - poses: `double[12]`, row-major 3x4 [R | t]  (mrpt::poses::CPose3D rotation + translation).
 *  - `*_on_device` flags: 0 = the pointer is HOST memory (pinned memory is DMA'd directly, pageable
 *    memory goes through the CUDA staging path), 1 = the pointer is DEVICE memory on the context's
 *    GPU (lets a caller keep pairings resident between matcher and solver).
 *  - `pairs_on_device` additionally accepts MP2P_B200_PAIRS_LAST_MATCH (2): the HOST records handed
 *    in are, unmodified, what the last matcher call on this context returned to the host (true
 *    between run_matchers and run_solvers, mp2p_icp/src/ICP.cpp:143-170: nothing touches the
 *    Pairings in between). The solver then reads the copy that matcher call left in device memory
 *    instead of uploading the same bytes again; the count must match or the call fails with
 *    MP2P_B200_ERR_ARG. The caller vouches for the identity — the library only checks the count.
 *  - return value: 0 = MP2P_B200_OK, negative = error; text via mp2p_b200_last_error().
 *  - There is NO CPU fallback: without a CUDA device every compute entry point fails with
 *    MP2P_B200_ERR_CUDA.
 */
#ifndef MP2P_B200_H
#define MP2P_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MP2P_B200_OK 0
#define MP2P_B200_ERR_ARG (-1)
#define MP2P_B200_ERR_CUDA (-2)
#define MP2P_B200_ERR_CAPACITY (-3)
#define MP2P_B200_ERR_NOMEM (-4)

#define MP2P_B200_MAX_KNN 32
#define MP2P_B200_PAIRS_LAST_MATCH 2

typedef struct mp2p_b200_ctx mp2p_b200_ctx; /* one per (process, GPU): stream + scratch */
typedef struct mp2p_b200_map mp2p_b200_map; /* device-resident global layer + NN index   */

/* mrpt::tfest::TMatchingPair — 36 bytes, the element of Pairings::paired_pt2pt
 * (fields written at mp2p_icp/src/Matcher_Points_DistanceThreshold.cpp:106-113). */
#pragma pack(push, 1)
typedef struct
{
    uint32_t globalIdx, localIdx;
    float    global_x, global_y, global_z;
    float    local_x, local_y, local_z; /* ORIGINAL (untransformed) local point */
    float    errorSquareAfterTransformation;
} mp2p_b200_pair_pt2pt;
#pragma pack(pop)

/* mp2p_icp::point_plane_pair_t — 72 bytes, the element of Pairings::paired_pt2pl
 * (mp2p_icp_map/include/mp2p_icp/point_plane_pair_t.h:34-38, plane_patch.h:30-34). */
typedef struct
{
    double plane_coefs[4]; /* TPlane: A x + B y + C z + D = 0, unit normal */
    double centroid[3];
    float  local_x, local_y, local_z; /* ORIGINAL local point */
    float  _pad;
} mp2p_b200_pair_pt2pl;

/* mp2p_icp::point_line_pair_t (mp2p_icp/include/mp2p_icp/Pairings.h:61-73): mrpt::math::TLine3D
 * {TPoint3D pBase, director[3]} + TPoint3D pt_local — 72 bytes, all doubles. */
typedef struct
{
    double pBase[3];
    double director[3];
    double local[3];
} mp2p_b200_pair_pt2ln;

/* Parameters of Matcher_Points_DistanceThreshold (+ Matcher_Points_Base):
 * mp2p_icp/src/Matcher_Points_DistanceThreshold.cpp:39-46, Matcher_Points_Base.cpp:132-181. */
typedef struct
{
    double   threshold;
    double   thresholdAngularDeg;
    uint32_t pairingsPerPoint;
    int32_t  allowMatchAlreadyMatchedPoints;
    int32_t  allowMatchAlreadyMatchedGlobalPoints;
    double   bounding_box_intersection_check_epsilon; /* default 0.20 */
} mp2p_b200_pt2pt_params;

/* Parameters of Matcher_Point2Plane (mp2p_icp/src/Matcher_Point2Plane.cpp:35-39) plus the
 * plane-fit parameters of the NearestPlaneCapable implementation over a point layer
 * (names from tests/test-mp2p_matcher_pt2pl.cpp:74-81; definition SURVEY.md §8a-7'). */
typedef struct
{
    double   distanceThreshold;
    double   searchRadius;
    uint32_t knn;
    uint32_t minimumPlanePoints;
    double   planeEigenThreshold;
    int32_t  allowMatchAlreadyMatchedPoints;
    double   bounding_box_intersection_check_epsilon;
} mp2p_b200_pt2pl_params;

/* mp2p_icp::WeightParameters as used by Solver_Horn
 * (mp2p_icp/include/mp2p_icp/WeightParameters.h:36-63). robust_kernel: 0 None, 1 GemanMcClure,
 * 2 Cauchy (mp2p_icp/include/mp2p_icp/robust_kernels.h:33-44). */
typedef struct
{
    int32_t use_scale_outlier_detector;
    double  scale_outlier_threshold;
    double  w_pt2pt;
    int32_t robust_kernel;
    double  robust_kernel_param;
    double  currentEstimateForRobust[12];
} mp2p_b200_horn_params;

/* mp2p_icp::OptimalTF_GN_Parameters (mp2p_icp/include/mp2p_icp/optimal_tf_gauss_newton.h:32-61),
 * restricted to the pt2pt and pt2pl terms; the prior term stays host-side. */
typedef struct
{
    uint32_t maxInnerLoopIterations;
    double   minDelta;
    double   maxCost;
    double   w_pt2pt, w_pt2pl;
    int32_t  kernel;
    double   kernelParam;
} mp2p_b200_gn_params;

typedef struct
{
    uint64_t n_points;
    float    bbox_min[3], bbox_max[3];
    float    finest_cell_size; /* metres */
    uint32_t n_levels;
    uint64_t n_finest_cells;
    uint64_t index_bytes;
    float    build_ms; /* device time of the index build (CUDA events) */
} mp2p_b200_map_info;

typedef struct
{
    uint64_t     n_points;
    float        build_ms;                       /* device time of upload + sort (CUDA events) */
    const float *x_device, *y_device, *z_device; /* the cloud in the caller's order, device memory */
} mp2p_b200_cloud_info;

/* Accumulator packets (32 doubles each; what a multi-GPU caller all-reduces with SUM):
 *  GN   : [0..20] upper triangle of H row-major, [21..26] g, [27] sum w|e|^2, [28] pair count
 *  HORN1: [0..2] sum local, [3..5] sum global, [6] count (non-outlier pairs), [7] pairs (all)
 *  HORN2: [0..8] S row-major (sum w r b^T), [9] w_sum, [10] new outliers, [11] pairs used   */
#define MP2P_B200_PACKET_DOUBLES 32

/* ------------------------------------------------------------------------------------------ */
const char* mp2p_b200_last_error(void);
int         mp2p_b200_device_count(void);

/* `cuda_stream` = a cudaStream_t the caller owns (e.g. torch's current stream) or NULL to let the
 * context create its own non-blocking stream. */
int  mp2p_b200_ctx_create(int device, void* cuda_stream, mp2p_b200_ctx** out);
void mp2p_b200_ctx_destroy(mp2p_b200_ctx* ctx);
int  mp2p_b200_ctx_synchronize(mp2p_b200_ctx* ctx);
/* number of kernels this context launched so far (bench.py's `gpu_launches`). */
uint64_t mp2p_b200_ctx_launch_count(const mp2p_b200_ctx* ctx);

/* Number of pairings the last matcher call of this context left on the DEVICE without reading it back
 * (out_count == NULL, the fused and the query-sharded iterations); synchronises. */
int mp2p_b200_ctx_last_count(mp2p_b200_ctx* ctx, uint64_t* n_pairs);

/* Upload a global map layer and build its NN index. Replaces
 * mrpt::maps::NearestNeighborsCapable::nn_prepare_for_3d_queries() on a CPointsMap
 * (call site mp2p_icp/src/Matcher_Points_DistanceThreshold.cpp:92); rebuilt whenever the layer is
 * modified (Matcher_Points_Base.cpp:105-114). x/y/z = CPointsMap::getPointsBufferRef_x/y/z(). */
int  mp2p_b200_map_create(mp2p_b200_ctx* ctx, const float* x, const float* y, const float* z,
                          uint64_t n, int on_device, mp2p_b200_map** out);
void mp2p_b200_map_destroy(mp2p_b200_map* map);
int  mp2p_b200_map_get_info(const mp2p_b200_map* map, mp2p_b200_map_info* out);

/* A LOCAL cloud kept on the device for a whole ICP::align(): the local layer does not change
 * between iterations (ICP.cpp:123-308 only moves the pose), so it is uploaded once and a second
 * copy is sorted along a Morton curve — warps of the search kernels then work on spatially
 * neighbouring queries and share hash cells / map points in L1/L2. Results are always reported
 * under the caller's original indices and in the reference's order; the sort is invisible.
 * Every entry point below that takes `lx, ly, lz, n_local, local_on_device` accepts
 *   local_on_device = 0  host arrays (copied on every call),
 *                     1  device arrays (used in place, caller's order),
 *                     2  `lx` is a mp2p_b200_cloud* (cast), `ly`/`lz` ignored, n_local = its size. */
typedef struct mp2p_b200_cloud mp2p_b200_cloud;
#define MP2P_B200_LOCAL_HOST 0
#define MP2P_B200_LOCAL_DEVICE 1
#define MP2P_B200_LOCAL_CLOUD 2
int  mp2p_b200_cloud_create(mp2p_b200_ctx* ctx, const float* x, const float* y, const float* z,
                            uint64_t n, int on_device, mp2p_b200_cloud** out);
void mp2p_b200_cloud_destroy(mp2p_b200_cloud* cloud);
int  mp2p_b200_cloud_get_info(const mp2p_b200_cloud* cloud, mp2p_b200_cloud_info* out);

/* Device copies of HOST layers, cached by the library (what a plugin whose methods are `const` and whose
 * callers hand it mrpt point maps needs; SURVEY.md §8b "Ownership"). The reference relies on MRPT's private
 * "kd-tree up to date" flag; from outside only the buffers are visible, so a layer is taken for unchanged while
 * its x-buffer address, its size and a FINGERPRINT (FNV-1a over the bit patterns of ~4096 evenly spaced points
 * plus the first and the last one, ~3 us) are the same; anything else rebuilds the copy (*rebuilt = 1).
 * An in-place edit that touches none of the sampled points is NOT seen: call mp2p_b200_layer_invalidate (or
 * use the non-cached entry points, which never assume anything). The handles stay owned by the context (at
 * most MP2P_B200_LAYER_CACHE_SLOTS layers per kind, least recently used first out) — do not destroy them. */
#define MP2P_B200_LAYER_CACHE_SLOTS 8
int  mp2p_b200_map_cached(mp2p_b200_ctx* ctx, const float* x, const float* y, const float* z, uint64_t n,
                          mp2p_b200_map** out, int32_t* rebuilt);
int  mp2p_b200_cloud_cached(mp2p_b200_ctx* ctx, const float* x, const float* y, const float* z, uint64_t n,
                            mp2p_b200_cloud** out, int32_t* rebuilt);
void mp2p_b200_layer_invalidate(mp2p_b200_ctx* ctx, const float* x);
uint64_t mp2p_b200_layer_fingerprint(const float* x, const float* y, const float* z, uint64_t n);

/* Raw k-NN of already-transformed query points (ascending (d2, index), d2 < radius2 strictly;
 * out_idx/out_d2 are [nq*k], out_found [nq]); replaces nn_single_search / nn_multiple_search /
 * nn_radius_search (Matcher_Points_DistanceThreshold.cpp:161-163,174-177,246-248). Host pointers. */
int mp2p_b200_knn(mp2p_b200_ctx* ctx, const mp2p_b200_map* map, const float* qx, const float* qy,
                  const float* qz, uint64_t nq, uint32_t k, float radius2, uint32_t* out_idx,
                  float* out_d2, int32_t* out_found);

/* Matcher_Points_DistanceThreshold::implMatchOneLayer
 * (mp2p_icp/src/Matcher_Points_DistanceThreshold.cpp:48-269, serial-branch semantics) including
 * transform_local_to_global (Matcher_Points_Base.cpp:183-249) and the bounding-box gate.
 *  local_paired_bits / global_paired_bits: MatchState bitfields on entry (bit i of word i/32;
 *  NULL = none set), HOST memory (Matcher.cpp:46-88, pointcloud_bitfield.h:46-133). The caller
 *  marks the bits of the returned pairs (lambdaAddPair, :116-120).
 *  out_pairs: capacity records; *out_count = pairs produced (ascending localIdx, then rank);
 *  *potential_pairings is incremented by n_local*pairingsPerPoint (:64). */
int mp2p_b200_match_pt2pt(mp2p_b200_ctx* ctx, mp2p_b200_map* map, const float* lx, const float* ly,
                          const float* lz, uint64_t n_local, int local_on_device,
                          const double pose[12], const mp2p_b200_pt2pt_params* params,
                          const uint32_t* local_paired_bits, const uint32_t* global_paired_bits,
                          mp2p_b200_pair_pt2pt* out_pairs, uint64_t capacity, int out_on_device,
                          uint64_t* out_count, uint64_t* potential_pairings);

/* Matcher_Point2Plane::implMatchOneLayer (mp2p_icp/src/Matcher_Point2Plane.cpp:41-114) with
 * NearestPlaneCapable::nn_search_pt2pl (mp2p_icp_map/include/mp2p_icp/NearestPlaneCapable.h:39-51)
 * realised over the point layer as k-NN + estimate_points_eigen
 * (mp2p_icp_map/src/estimate_points_eigen.cpp:27-123) + planarity test
 * (mp2p_icp/src/Matcher_Adaptive.cpp:229-253). */
int mp2p_b200_match_pt2pl(mp2p_b200_ctx* ctx, mp2p_b200_map* map, const float* lx, const float* ly,
                          const float* lz, uint64_t n_local, int local_on_device,
                          const double pose[12], const mp2p_b200_pt2pl_params* params,
                          const uint32_t* local_paired_bits, mp2p_b200_pair_pt2pl* out_pairs,
                          uint64_t capacity, int out_on_device, uint64_t* out_count,
                          uint64_t* potential_pairings);

/* Parameters of Matcher_Point2Line (mp2p_icp/src/Matcher_Point2Line.cpp:35-45,
 * mp2p_icp/include/mp2p_icp/Matcher_Point2Line.h:61-64) + Matcher_Points_Base. */
typedef struct
{
    double   distanceThreshold;
    uint32_t knn;
    uint32_t minimumLinePoints; /* >= 2 */
    double   lineEigenThreshold;
    int32_t  allowMatchAlreadyMatchedPoints;
    double   bounding_box_intersection_check_epsilon;
} mp2p_b200_pt2ln_params;

/* Matcher_Point2Line::implMatchOneLayer (mp2p_icp/src/Matcher_Point2Line.cpp:46-163):
 * the `knn` nearest global points of every local point not yet paired (unbounded search, :103-105); at
 * least minimumLinePoints of them within distanceThreshold (:110-130); estimate_points_eigen over all
 * knn points (:132-135, as written upstream); line test e0, e1 <= lineEigenThreshold * e2 (:148-149);
 * record = {mean, unit eigenvector of the largest eigenvalue, ORIGINAL local point}. Global points are
 * never marked (:92-95); the caller marks the local bits of the returned pairs (:159). Output in
 * ascending local index; *potential_pairings += n_local (:58). */
int mp2p_b200_match_pt2ln(mp2p_b200_ctx* ctx, mp2p_b200_map* map, const float* lx, const float* ly,
                          const float* lz, uint64_t n_local, int local_on_device, const double pose[12],
                          const mp2p_b200_pt2ln_params* params, const uint32_t* local_paired_bits,
                          mp2p_b200_pair_pt2ln* out_pairs, uint64_t capacity, int out_on_device,
                          uint64_t* out_count, uint64_t* potential_pairings);

/* Parameters of Matcher_Adaptive (mp2p_icp/include/mp2p_icp/Matcher_Adaptive.h:66-75,
 * mp2p_icp/src/Matcher_Adaptive.cpp:32-57) + Matcher_Points_Base. */
typedef struct
{
    double   confidenceInterval;        /* (0,1) */
    double   firstToSecondDistanceMax;
    double   absoluteMaxSearchDistance; /* m */
    double   minimumCorrDist;           /* m */
    int32_t  enableDetectPlanes;
    uint32_t planeSearchPoints;
    uint32_t planeMinimumFoundPoints; /* >= 3 */
    uint32_t maxPt2PtCorrespondences; /* >= 1 */
    double   planeEigenThreshold;
    double   planeMinimumDistance;
    int32_t  allowMatchAlreadyMatchedPoints;
    int32_t  allowMatchAlreadyMatchedGlobalPoints;
    double   bounding_box_intersection_check_epsilon;
} mp2p_b200_adaptive_params;
#define MP2P_B200_ADAPTIVE_BINS 50 /* mrpt::math::CHistogram hist(min, max, 50), Matcher_Adaptive.cpp:188 */

/* Matcher_Adaptive::implMatchOneLayer (mp2p_icp/src/Matcher_Adaptive.cpp:59-314) in two
 * device phases around the one step that is mrpt::math code:
 *   adaptive_search    per local point the nearest neighbour(s) within absoluteMaxSearchDistance (at most
 *                      MAX_CORRS_PER_LOCAL = 10 kept, Matcher_Adaptive.h:84); returns the 50-bin histogram of
 *                      the 1st / 2nd neighbour squared errors exactly as CHistogram::add bins them
 *                      (:168-194), its limits and sample count; *gate = 0 if the bounding boxes do not
 *                      overlap (:77-80: the reference returns before anything else — ignore the histogram)
 *   (host)             ci_high = upper confidence bound of that histogram (:196-199,
 *                      mrpt::math::confidenceIntervalsFromHistogram); maxCorrDistSqr =
 *                      max(minimumCorrDist^2, ci_high) (:214). A plugin built against MRPT calls MRPT here;
 *                      mp2p_b200_adaptive_threshold is the library's restatement of the two MRPT helpers
 *                      (recalled from MRPT 2.x, NOT in the reference tree: parity unpinned)
 *   adaptive_emit      per local point a plane through its neighbours -> point-to-plane pairing (:222-268),
 *                      else the point-to-point pairings below the threshold (:270-297); both lists in
 *                      ascending local index. Must directly follow adaptive_search on the same context.
 * mp2p_b200_match_adaptive = the three steps in one call; where the reference throws (no neighbour at all,
 * all errors equal: CHistogram asserts max > min) it returns MP2P_B200_ERR_ARG. Global points are never
 * marked by this matcher (:303-311); the caller marks the local bits of the returned pairings.
 * *potential_pairings += n_local * maxPt2PtCorrespondences (:68). */
int mp2p_b200_adaptive_search(mp2p_b200_ctx* ctx, mp2p_b200_map* map, const float* lx, const float* ly,
                              const float* lz, uint64_t n_local, int local_on_device, const double pose[12],
                              const mp2p_b200_adaptive_params* params, const uint32_t* local_paired_bits,
                              uint64_t histogram_out[MP2P_B200_ADAPTIVE_BINS], double* err_min, double* err_max,
                              uint64_t* n_samples, int32_t* gate, uint64_t* potential_pairings);
int mp2p_b200_adaptive_threshold(const uint64_t histogram[MP2P_B200_ADAPTIVE_BINS], double err_min, double err_max,
                                 uint64_t n_samples, double confidenceInterval, double minimumCorrDist,
                                 double* ci_high, double* maxCorrDistSqr);
int mp2p_b200_adaptive_emit(mp2p_b200_ctx* ctx, mp2p_b200_map* map, const mp2p_b200_adaptive_params* params,
                            double maxCorrDistSqr, const uint32_t* global_paired_bits,
                            mp2p_b200_pair_pt2pt* out_pt2pt, uint64_t capacity_pt2pt,
                            mp2p_b200_pair_pt2pl* out_pt2pl, uint64_t capacity_pt2pl, int out_on_device,
                            uint64_t* n_pt2pt, uint64_t* n_pt2pl);
int mp2p_b200_match_adaptive(mp2p_b200_ctx* ctx, mp2p_b200_map* map, const float* lx, const float* ly,
                             const float* lz, uint64_t n_local, int local_on_device, const double pose[12],
                             const mp2p_b200_adaptive_params* params, const uint32_t* local_paired_bits,
                             const uint32_t* global_paired_bits, mp2p_b200_pair_pt2pt* out_pt2pt,
                             uint64_t capacity_pt2pt, mp2p_b200_pair_pt2pl* out_pt2pl, uint64_t capacity_pt2pl,
                             int out_on_device, uint64_t* n_pt2pt, uint64_t* n_pt2pl, double* ci_high,
                             uint64_t* potential_pairings);

/* Parameters of Matcher_Points_InlierRatio (mp2p_icp/include/mp2p_icp/Matcher_Points_InlierRatio.h:50-56,
 * mp2p_icp/src/Matcher_Points_InlierRatio.cpp:35-39) + Matcher_Points_Base. */
typedef struct
{
    double  inliersRatio; /* (0,1); class default 0.80 */
    int32_t allowMatchAlreadyMatchedPoints;
    int32_t allowMatchAlreadyMatchedGlobalPoints;
    double  bounding_box_intersection_check_epsilon; /* default 0.20 */
} mp2p_b200_inlier_ratio_params;

/* Matcher_Points_InlierRatio::implMatchOneLayer (mp2p_icp/src/Matcher_Points_InlierRatio.cpp:41-143;
 * SURVEY.md §8f N1): the unbounded nearest neighbour of every local point not yet paired, the tentative
 * pairings ordered by errorSquareAfterTransformation (equal distances in reverse local order, as the
 * reference's multimap::emplace_hint(begin()) leaves them, :104), the first
 * mrpt::round(nTotal * inliersRatio) of them emitted IN THAT ORDER, skipping global points already paired
 * on entry or named by an earlier pairing of this call (:121-137). Arguments as mp2p_b200_match_pt2pt;
 * *potential_pairings is incremented by n_local (:55). Where the reference throws — inliersRatio outside
 * (0,1) (:49-50), no tentative pairing at all (:117) — the call returns MP2P_B200_ERR_ARG. */
int mp2p_b200_match_inlier_ratio(mp2p_b200_ctx* ctx, mp2p_b200_map* map, const float* lx, const float* ly,
                                 const float* lz, uint64_t n_local, int local_on_device, const double pose[12],
                                 const mp2p_b200_inlier_ratio_params* params, const uint32_t* local_paired_bits,
                                 const uint32_t* global_paired_bits, mp2p_b200_pair_pt2pt* out_pairs,
                                 uint64_t capacity, int out_on_device, uint64_t* out_count,
                                 uint64_t* potential_pairings);

/* ---- query-sharded pt2pt matching, one process per GPU --------------------------
 * The local cloud is split in contiguous shards of `per_shard` points (the last may be shorter, or
 * empty); shard r owns the local indices [r*per_shard, r*per_shard + n_local_r). The map (and its
 * index) is replicated on every GPU. The shards talk through fixed-size EXCHANGE RECORDS of
 * mp2p_b200_shard_record_words(per_shard, pairingsPerPoint) 64-bit words:
 *     [per_shard*pairingsPerPoint candidate words | 24 opaque bytes: the shard's bounding box | pad]
 *  phase A `..._shard_search`: transform + NN search of the shard; writes the shard's record to
 *     DEVICE memory of the caller (asynchronous: no host synchronisation);
 *  (caller) all-gather the records of all shards, rank order, contiguous — ONE collective
 *     (NCCL all_gather over NVLink), nothing to repack on either side;
 *  phase B `..._shard_resolve`: replays every shard's proposals on this GPU's first-claim array
 *     with the global proposal numbering, applies the bounding-box gate of the WHOLE cloud and
 *     compacts this shard's accepted pairs (localIdx = index in the whole cloud). Optionally the
 *     HORN1 sums of the shard's pairs are produced in the same pass (horn_sums_packet_device).
 *     With out_count == NULL (device output only; also accepted by mp2p_b200_match_pt2pt and
 *     mp2p_b200_match_pt2pl) the call is asynchronous as well: the pairing
 *     count stays on the device and the solver building blocks below take it from there when
 *     given n = MP2P_B200_COUNT_ON_DEVICE — a whole sharded iteration then needs ONE host
 *     synchronisation (reading the final packets).
 * Concatenating the shards' outputs in rank order gives exactly the single-GPU result.
 * Phase B must follow phase A on the same context (the staged shard is reused). */
#define MP2P_B200_COUNT_ON_DEVICE UINT64_MAX
uint64_t mp2p_b200_shard_record_words(uint64_t per_shard, uint32_t pairingsPerPoint);
int mp2p_b200_match_pt2pt_shard_search(mp2p_b200_ctx* ctx, mp2p_b200_map* map, const float* lx,
                                       const float* ly, const float* lz, uint64_t n_local,
                                       int local_on_device, const double pose[12],
                                       const mp2p_b200_pt2pt_params* params,
                                       const uint32_t* local_paired_bits, uint64_t per_shard,
                                       uint64_t* record_out_device);
int mp2p_b200_match_pt2pt_shard_resolve(mp2p_b200_ctx* ctx, mp2p_b200_map* map, uint64_t n_local,
                                        uint32_t shard_rank, uint32_t n_shards, uint64_t per_shard,
                                        const uint64_t* records_device,
                                        const mp2p_b200_pt2pt_params* params,
                                        const uint32_t* global_paired_bits,
                                        mp2p_b200_pair_pt2pt* out_pairs, uint64_t capacity,
                                        int out_on_device, uint64_t* out_count /* may be NULL */,
                                        double* horn_sums_packet_device /* may be NULL */);

/* optimal_tf_horn (mp2p_icp/src/optimal_tf_horn.cpp:201-252) over pt2pt pairings:
 * eval_centroids_robust (Pairings.cpp:68-110) + visit_correspondences S accumulation
 * (visit_correspondences.h:39-221) on the GPU; 4x4 eigen-solve on the host.
 * weight_counts/values: Pairings::point_weights run-length blocks (NULL/0 = all 1.0).
 * *solved = 0 mirrors `return false` (fewer than 3 pairings, optimal_tf_horn.cpp:96). */
int mp2p_b200_solve_horn(mp2p_b200_ctx* ctx, const mp2p_b200_pair_pt2pt* pairs, uint64_t n,
                         int pairs_on_device, const mp2p_b200_horn_params* params,
                         const uint64_t* weight_counts, const double* weight_values,
                         uint64_t n_weight_blocks, double pose_out[12], int32_t* solved);

/* pt2ln_pl_to_pt2pt, plane part (mp2p_icp/src/pt2ln_pl_to_pt2pt.cpp:25-83): what Solver_Horn does to
 * pt2pl pairings before optimal_tf_horn (Solver_Horn.cpp:51-55) — each pairing becomes the pt2pt
 * pairing {local point -> its projection on the plane, computed at guess_pose}; those whose
 * |distance| reaches 25 % of the largest one are kept, or the 3 largest. Indices are 0 (dummies) and
 * errorSquareAfterTransformation is 0, as in the reference. ORDER: the reference emits the records by
 * descending |distance| (multimap walk); this call keeps the input order — Horn's sums do not depend
 * on it. mp2p_b200_solve_horn_pt2pl = conversion + optimal_tf_horn without leaving the device. */
int mp2p_b200_pt2pl_to_pt2pt(mp2p_b200_ctx* ctx, const mp2p_b200_pair_pt2pl* pairs, uint64_t n,
                             int pairs_on_device, const double guess_pose[12],
                             mp2p_b200_pair_pt2pt* out_pairs, uint64_t capacity, int out_on_device,
                             uint64_t* out_count);
int mp2p_b200_solve_horn_pt2pl(mp2p_b200_ctx* ctx, const mp2p_b200_pair_pt2pl* pairs, uint64_t n,
                               int pairs_on_device, const double guess_pose[12],
                               const mp2p_b200_horn_params* params, double pose_out[12],
                               int32_t* solved);

/* optimal_tf_gauss_newton (mp2p_icp/src/optimal_tf_gauss_newton.cpp:36-372), pt2pt + pt2pl terms
 * (error_point2point / error_point2plane, errorTerms.cpp:36-66,115-161; robust_kernels.h:57-94).
 * H and g are zeroed every inner iteration. */
int mp2p_b200_solve_gauss_newton(mp2p_b200_ctx* ctx, const mp2p_b200_pair_pt2pt* pairs_pt2pt,
                                 uint64_t n_pt2pt, const mp2p_b200_pair_pt2pl* pairs_pt2pl,
                                 uint64_t n_pt2pl, int pairs_on_device,
                                 const mp2p_b200_gn_params* params, const double pose_init[12],
                                 double pose_out[12], uint32_t* iterations_done, int32_t* solved);

/* The same solver with the point-to-line term added (error_point2line, errorTerms.cpp:67-113;
 * optimal_tf_gauss_newton.cpp:182-203: weight = w_pt2ln * robust(|e|^2), cost term weight^2 |e|^2 as
 * written there). pairs_on_device: 0 = host lists, 1 = device lists (MP2P_B200_PAIRS_LAST_MATCH is not
 * offered for this form). Reference tests: tests/test-mp2p_optimize_pt2ln.cpp. */
int mp2p_b200_solve_gauss_newton_ex(mp2p_b200_ctx* ctx, const mp2p_b200_pair_pt2pt* pairs_pt2pt, uint64_t n_pt2pt,
                                    const mp2p_b200_pair_pt2pl* pairs_pt2pl, uint64_t n_pt2pl,
                                    const mp2p_b200_pair_pt2ln* pairs_pt2ln, uint64_t n_pt2ln, int pairs_on_device,
                                    const mp2p_b200_gn_params* params, double w_pt2ln, const double pose_init[12],
                                    double pose_out[12], uint32_t* iterations_done, int32_t* solved);

/* ---- fused iterations: when BOTH plugins of an ICP iteration are ours, run_matchers + run_solvers
 * (mp2p_icp/src/ICP.cpp:143,170) are enqueued back to back on the device — the pairings never
 * leave HBM and the call synchronises ONCE. `pairs_device` (optional, DEVICE memory, `capacity`
 * records) receives the pairings so that the caller can still fetch them (Results::finalPairings,
 * hooks, logging: ICP.cpp:232-241,286-303,331); NULL = library scratch. `*solved` = 0 mirrors a
 * solver returning false (no / too few pairings). MatchState bitfields are not taken: one matcher
 * per iteration starts from an empty MatchState (Matcher.cpp:58-66). ---- */
int mp2p_b200_iterate_pt2pt_horn(mp2p_b200_ctx* ctx, mp2p_b200_map* map, const float* lx, const float* ly,
                                 const float* lz, uint64_t n_local, int local_on_device,
                                 const double pose[12], const mp2p_b200_pt2pt_params* matcher_params,
                                 const mp2p_b200_horn_params* solver_params,
                                 mp2p_b200_pair_pt2pt* pairs_device, uint64_t capacity, double pose_out[12],
                                 int32_t* solved, uint64_t* n_pairs, uint64_t* potential_pairings);
int mp2p_b200_iterate_pt2pl_gn(mp2p_b200_ctx* ctx, mp2p_b200_map* map, const float* lx, const float* ly,
                               const float* lz, uint64_t n_local, int local_on_device, const double pose[12],
                               const mp2p_b200_pt2pl_params* matcher_params,
                               const mp2p_b200_gn_params* solver_params, mp2p_b200_pair_pt2pl* pairs_device,
                               uint64_t capacity, double pose_out[12], int32_t* solved, uint64_t* n_pairs,
                               uint32_t* iterations_done, uint64_t* potential_pairings);

/* ---- building blocks for query-sharded multi-GPU runs: each rank accumulates
 * over its shard, the caller all-reduces the 32-double packet (SUM), every rank finishes the
 * solve redundantly. `packet` may be host or device memory (packet_on_device); with device pairs
 * AND a device packet the calls only enqueue work. `n` / `n_pt2pt` = MP2P_B200_COUNT_ON_DEVICE: the
 * pairs are the device output of the preceding shard_resolve, whose count is read on the device.
 * horn_moments: n_total_pairs = 0 takes the pair count of the WHOLE cloud from the reduced HORN1
 * packet ([7]) on the device instead of from the host. ---- */
int mp2p_b200_gn_accumulate(mp2p_b200_ctx* ctx, const mp2p_b200_pair_pt2pt* pairs_pt2pt,
                            uint64_t n_pt2pt, const mp2p_b200_pair_pt2pl* pairs_pt2pl,
                            uint64_t n_pt2pl, int pairs_on_device, const mp2p_b200_gn_params* params,
                            const double pose[12], double* packet, int packet_on_device);
/* Gauss-Newton inner loop with the pose kept on the DEVICE (no host round trip per inner iteration;
 * a multi-GPU caller enqueues  begin, then maxInnerLoopIterations x { accumulate, all_reduce(packet),
 * step }  and reads the state back once). state_device: MP2P_B200_GN_STATE_DOUBLES doubles =
 * [0..11] pose, then two uint32 {done flag, pose updates applied}. Once `done` is set (converged,
 * optimal_tf_gauss_newton.cpp:344-365) accumulate and step become no-ops, so every rank can run the
 * same fixed number of rounds. Pairings must be device memory; n = MP2P_B200_COUNT_ON_DEVICE allowed
 * for the list the last matcher call produced. All three calls only enqueue work. */
#define MP2P_B200_GN_STATE_DOUBLES 16
int mp2p_b200_gn_device_begin(mp2p_b200_ctx* ctx, const double pose[12], double* state_device);
int mp2p_b200_gn_device_accumulate(mp2p_b200_ctx* ctx, const mp2p_b200_pair_pt2pt* pairs_pt2pt,
                                   uint64_t n_pt2pt, const mp2p_b200_pair_pt2pl* pairs_pt2pl,
                                   uint64_t n_pt2pl, const mp2p_b200_gn_params* params,
                                   const double* state_device, double* packet_device);
int mp2p_b200_gn_device_step(mp2p_b200_ctx* ctx, const double* packet_device,
                             const mp2p_b200_gn_params* params, double* state_device);
/* host-side step from a reduced GN packet: delta = -H^{-1} g (LDLT), pose_out = pose (+) exp(delta);
 * *converged = 1 if |delta| < minDelta or sqrt(err) <= maxCost (optimal_tf_gauss_newton.cpp:344-365). */
int mp2p_b200_gn_step_from_packet(const double packet[MP2P_B200_PACKET_DOUBLES],
                                  const mp2p_b200_gn_params* params, const double pose[12],
                                  double pose_out[12], int32_t* converged);
int mp2p_b200_horn_sums(mp2p_b200_ctx* ctx, const mp2p_b200_pair_pt2pt* pairs, uint64_t n,
                        int pairs_on_device, double* packet, int packet_on_device);
int mp2p_b200_horn_moments(mp2p_b200_ctx* ctx, const mp2p_b200_pair_pt2pt* pairs, uint64_t n,
                           int pairs_on_device, const mp2p_b200_horn_params* params,
                           const double* sums_packet /* reduced HORN1 */, int sums_on_device,
                           uint64_t n_total_pairs, double* packet, int packet_on_device);
/* host-side finish from reduced HORN1 + HORN2 packets (optimal_tf_horn.cpp:132-174,238-247). */
int mp2p_b200_horn_finish(const double sums_packet[MP2P_B200_PACKET_DOUBLES],
                          const double moments_packet[MP2P_B200_PACKET_DOUBLES], double pose_out[12],
                          int32_t* solved);

/* ---- peer exchange over NVLink for the query-sharded path (one process per GPU) ----------------
 * The two collectives of a sharded iteration — all-gather of the exchange records, all-reduce (SUM)
 * of a 32-double packet — done by small kernels that write straight into the peers' HBM (CUDA IPC
 * mappings) instead of by a collective library: no host call per collective, a few microseconds on
 * the device (csrc/peer.cu). Setup: every rank calls peer_create (allocates its mailbox, returns an
 * opaque handle of MP2P_B200_PEER_HANDLE_BYTES bytes), the caller all-gathers the handles by any
 * means (torch.distributed, MPI, files) and passes all of them, rank order, to peer_connect.
 * Every rank must issue the same sequence of exchanges. All calls only enqueue work on the context
 * stream. record_words = mp2p_b200_shard_record_words(per_shard, pairingsPerPoint).
 *   peer_record_slot        where THIS rank's next exchange record has to be written (pass it as
 *                           record_out_device of ..._shard_search)
 *   peer_allgather_records  pushes that record to every peer and waits for theirs; *records_device =
 *                           the gathered records, rank order, contiguous (pass to ..._shard_resolve)
 *   peer_allreduce_packet   packet_device[0..32) <- sum over ranks, in rank order: bit-identical on
 *                           every rank. A peer that does not show up within 10 s traps the kernel. */
#define MP2P_B200_PEER_HANDLE_BYTES 64
typedef struct mp2p_b200_peer mp2p_b200_peer;
int  mp2p_b200_peer_create(mp2p_b200_ctx* ctx, uint32_t rank, uint32_t world, uint64_t record_words,
                           uint8_t handle_out[MP2P_B200_PEER_HANDLE_BYTES], mp2p_b200_peer** out);
int  mp2p_b200_peer_connect(mp2p_b200_peer* peer, const uint8_t* all_handles /* world x 64 bytes */);
void mp2p_b200_peer_destroy(mp2p_b200_peer* peer);
/* OWNER-PARTITIONED first claims (optional, after peer_connect): global point g belongs to rank g % world, which
 * keeps its claim word in memory every rank maps over NVLink. peer_iterate_pt2pt then proposes straight into the
 * owners' HBM (system-scope atomicMin), gathers only the shards' 32-byte bounding boxes (the flags of that exchange
 * are also the barrier "everybody has proposed") and reads each claim word back from its owner: no record
 * all-gather, no replay of the other shards' proposals — NVLink bytes and atomics per GPU are the shard's own,
 * whatever the world size. Same results, bit for bit. Setup like the mailboxes: every rank calls claims_create
 * (n_map_points = size of the replicated map; allocates 8 * ceil(n / world) bytes), the caller all-gathers the
 * 64-byte handles, every rank calls claims_connect with all of them in rank order. */
int  mp2p_b200_peer_claims_create(mp2p_b200_peer* peer, uint64_t n_map_points, uint8_t handle_out[MP2P_B200_PEER_HANDLE_BYTES]);
int  mp2p_b200_peer_claims_connect(mp2p_b200_peer* peer, const uint8_t* all_handles /* world x 64 bytes */);
int  mp2p_b200_peer_record_slot(mp2p_b200_peer* peer, uint64_t** slot_device);
int  mp2p_b200_peer_allgather_records(mp2p_b200_peer* peer, const uint64_t** records_device);
int  mp2p_b200_peer_allreduce_packet(mp2p_b200_peer* peer, double* packet_device);

/* Whole query-sharded iterations over the peer exchange, enqueued natively with ONE host
 * synchronisation (run_matchers + run_solvers, mp2p_icp/src/ICP.cpp:143,170, for the shard of this
 * rank; every rank gets the same pose). peer_iterate_pt2pt: Matcher_Points_DistanceThreshold with
 * exact cross-shard first-claim dedup, then Solver_Horn (`horn` != NULL; *n_pairs_total = pairings
 * of the whole cloud) or Solver_GaussNewton (`gn` != NULL) — exactly one of the two.
 * peer_iterate_pt2pl_gn: Matcher_Point2Plane + Solver_GaussNewton. `pairs_device` (device memory,
 * `capacity` >= n_local * pairingsPerPoint records) receives this shard's pairings. */
int mp2p_b200_peer_iterate_pt2pt(mp2p_b200_peer* peer, mp2p_b200_map* map, const float* lx, const float* ly,
                                 const float* lz, uint64_t n_local, int local_on_device,
                                 const double pose[12], const mp2p_b200_pt2pt_params* matcher_params,
                                 const mp2p_b200_horn_params* horn, const mp2p_b200_gn_params* gn,
                                 uint64_t per_shard, mp2p_b200_pair_pt2pt* pairs_device, uint64_t capacity,
                                 double pose_out[12], int32_t* solved, uint64_t* n_pairs_total,
                                 uint32_t* iterations_done);
int mp2p_b200_peer_iterate_pt2pl_gn(mp2p_b200_peer* peer, mp2p_b200_map* map, const float* lx, const float* ly,
                                    const float* lz, uint64_t n_local, int local_on_device,
                                    const double pose[12], const mp2p_b200_pt2pl_params* matcher_params,
                                    const mp2p_b200_gn_params* solver_params,
                                    mp2p_b200_pair_pt2pl* pairs_device, uint64_t capacity,
                                    double pose_out[12], int32_t* solved, uint32_t* iterations_done);

/* ---- measurement hooks (bench.py): per-kernel CUDA-event timing and search statistics ----
 * With profiling on, every public call records CUDA events around its kernels on the context
 * stream; mp2p_b200_ctx_get_timings returns the durations (ms) of the LAST call:
 *   [0] NN search kernel (k_match_pt2pt_nn1 / k_match_pt2pt<G>)   [1] compaction kernel
 *   [2] Horn sums  [3] Horn moments  [4] GN accumulate (the LAST inner iteration's launch)
 *   [5] whole call, device time   [6] plane / line fit kernel (pt2pl, pt2ln)   (unused slots = 0)
 * mp2p_b200_ctx_get_search_stats returns counters of the LAST match call when stats are on:
 *   [0] hash-table probes (16 B each)  [1] candidate points read (16 B each)
 *   [2] valid candidates written       [3] queries that climbed above the finest level
 *   [4] largest candidate count of one query  [5] largest probe count  [6] most levels visited
 *   [7] warps holding a query with more than 2000 candidates (the stragglers of a launch)  */
#define MP2P_B200_N_TIMINGS 8
int mp2p_b200_ctx_set_profiling(mp2p_b200_ctx* ctx, int timings_on, int search_stats_on);
int mp2p_b200_ctx_get_timings(mp2p_b200_ctx* ctx, float ms[MP2P_B200_N_TIMINGS]);
int mp2p_b200_ctx_get_search_stats(mp2p_b200_ctx* ctx, uint64_t stats[8]);

/* Per-CTA trace of the last k > 1 search over a resident cloud, recorded only when the environment variable
 * MP2P_KNN_TRACE is set: 8 words per CTA {SM id, start ns (low 32 bits of %globaltimer), end ns, query tile,
 * rounds, scan steps, list insertions, probes | levels << 16 of its first warp}. */
int mp2p_b200_ctx_get_tile_trace(mp2p_b200_ctx* ctx, uint32_t* out, uint64_t capacity_tiles, uint64_t* n_tiles);

/* ---- KITTI .bin straight to the device ------------------------------------------
 * A KITTI velodyne scan is a flat file of float32 (x, y, z, intensity) records — what the reference's
 * kitti2mm reads through mrpt::obs::CObservationPointCloud / CPointsMapXYZI
 * (apps/kitti2mm/main.cpp:55-69). read_kitti_bin loads such a file into PINNED host memory
 * (release with mp2p_b200_host_free); map_create_xyzi / cloud_create_xyzi take the interleaved records
 * (host or device memory) and split them into the library's SoA layout on the device — no host-side
 * de-interleaving pass, and from pinned memory the upload is one DMA transfer. The intensity channel is
 * not used by the matchers and is dropped. Results are those of the SoA entry points. */
int mp2p_b200_read_kitti_bin(const char* path, float** xyzi_pinned_out, uint64_t* n_points_out);
int mp2p_b200_map_create_xyzi(mp2p_b200_ctx* ctx, const float* xyzi, uint64_t n, int on_device, mp2p_b200_map** out);
int mp2p_b200_cloud_create_xyzi(mp2p_b200_ctx* ctx, const float* xyzi, uint64_t n, int on_device,
                                mp2p_b200_cloud** out);

/* ---- covariance() on the device -------------------------------------------------
 * mp2p_icp::covariance (mp2p_icp/src/covariance.cpp:28-141; call site ICP.cpp:337, once per align()): the
 * stacked error vector of the final pairings (pt2pt :75-82, pt2ln :85-93, pt2pl :107-115; three rows per
 * pairing) differentiated numerically w.r.t. (x, y, z, yaw, pitch, roll) by central differences with the
 * steps of CovarianceParameters (covariance.h:27-31: finDif_xyz, finDif_angles, default 1e-7; the
 * mrpt::math::estimateJacobian scheme, recalled from MRPT 2.x: parity unpinned for that helper),
 * hessian = J^T J, cov = hessian^-1 (inverse_LLt). The 12 sweeps over the pairings run in ONE launch.
 *   x6 = the vector the Jacobian is taken at. AS WRITTEN UPSTREAM the z slot is never assigned
 *        (covariance.cpp:41-47 sets [0], [1], [0] again, [3], [4], [5]) and a default-constructed
 *        CMatrixDouble61 is zero-filled, so the reference evaluates at (x, y, 0, yaw, pitch, roll): a caller
 *        that wants the reference's number passes x6[2] = 0 (the plugin and the host mirror do), one that wants
 *        the covariance at the solution passes its z.
 *   No pairings at all: cov = diag(1e6) (:33-38). ln2ln / pl2pl pairings are not taken (host-side upstream).
 *   *positive_definite = 0 if the Cholesky factorisation of the hessian fails (upstream: undefined result);
 *   cov_out then holds the hessian's pseudo-diagonal inverse and the call still returns MP2P_B200_OK.
 * pairs_on_device: 0 host lists, 1 device lists. cov_out / hessian_out: 6x6 row-major (hessian_out may be NULL). */
int mp2p_b200_covariance(mp2p_b200_ctx* ctx, const mp2p_b200_pair_pt2pt* pairs_pt2pt, uint64_t n_pt2pt,
                         const mp2p_b200_pair_pt2pl* pairs_pt2pl, uint64_t n_pt2pl, const mp2p_b200_pair_pt2ln* pairs_pt2ln,
                         uint64_t n_pt2ln, int pairs_on_device, const double x6[6], double finDif_xyz, double finDif_angles,
                         double cov_out[36], double hessian_out[36], int32_t* positive_definite);

/* ---- FilterDecimateVoxels on the device -----------------------------------------
 * mp2p_icp_filters::FilterDecimateVoxels::filter over ONE input layer
 * (mp2p_icp_filters/src/FilterDecimateVoxels.cpp:109-378; parameters FilterDecimateVoxels.h:84-116) — the
 * step right before ICP::align() in the reference's pipelines (demos/icp-settings-kitti.yaml:76-82).
 * Voxel index per axis = int32(coordinate / resolution): float division, truncation toward zero
 * (PointCloudToVoxelGridSingle.h:105). decimate_method: 0 FirstPoint (the first point inserted into the
 * voxel), 1 ClosestToAverage (the member closest to the voxel mean, first on ties), 2 VoxelAverage (float
 * sums in insertion order times float(1/n)); RandomPoint (an unseeded generator upstream) is refused.
 * has_flatten_to: z is replaced by flatten_to and only the first voxel of every (cx, cy) column emits
 * (:210-224, :335-349). OUTPUT ORDER: ascending (cx, cy, cz) — the order of the reference's std::map walk
 * (use_tsl_robin_map = false); with the default tsl::robin_map the reference's own order is
 * implementation-defined and only the set of points is comparable. out_src_index[i] = index of the input
 * point output i is a copy of, -1 for an averaged point (may be NULL). Returns MP2P_B200_ERR_CAPACITY
 * (with *out_count = the number needed) if capacity is too small; capacity = n always suffices.
 * mp2p_b200_cloud_create_decimated = the filter followed by mp2p_b200_cloud_create without leaving the
 * device: the decimated local layer of an align() goes from the filter to the matchers in HBM. */
typedef struct
{
    float   voxel_filter_resolution; /* metres, > 0 */
    int32_t decimate_method;
    int32_t has_flatten_to;
    float   flatten_to;
} mp2p_b200_decimate_params;
int mp2p_b200_filter_decimate_voxels(mp2p_b200_ctx* ctx, const float* x, const float* y, const float* z, uint64_t n,
                                     int on_device, const mp2p_b200_decimate_params* params, float* out_x, float* out_y,
                                     float* out_z, int64_t* out_src_index, uint64_t capacity, int out_on_device,
                                     uint64_t* out_count);
int mp2p_b200_cloud_create_decimated(mp2p_b200_ctx* ctx, const float* x, const float* y, const float* z, uint64_t n,
                                     int on_device, const mp2p_b200_decimate_params* params, mp2p_b200_cloud** out,
                                     uint64_t* out_count);

/* Pinned host memory helpers (so callers in any language can give the library DMA-able buffers). */
int  mp2p_b200_host_alloc(size_t bytes, void** out);
void mp2p_b200_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* MP2P_B200_H */
