// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the mp2p_icp Matcher+Solver hot path, written from the
// reference's documented behaviour; every function cites the reference file:line it follows
// (paths relative to /root/reference). It exists to CHECK the CUDA product path and to serve as
// the timed CPU baseline; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs may load it. The product library never links or calls it.
//
// Parity pinning: MRPT (the reference's hard dependency) is not installable here, so the real
// reference cannot be executed; this oracle is pinned against every known-answer fixture the
// reference's own tests hold for this path (tests/test_oracle_golden.py lists them one by one).
// pt2pl matching over a plain point layer is NOT pinned upstream (no in-tree NearestPlaneCapable
// implementer; its unit test is disabled, tests/CMakeLists.txt:37): for that function alone the
// header says "parity unpinned beyond the disabled test's known answers".
//
// Build: see oracle/Makefile (g++ -O3 -ffp-contract=off -fopenmp; no -march=native, matching the
// reference's default flags, 3rdparty/mola_common/cmake/mola_cmake_functions.cmake:162-172).
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <vector>

#include "kdtree.hpp"
#include "se3.hpp"

using namespace orc;

extern "C"
{
// mrpt::tfest::TMatchingPair, 36 bytes (fields used at Matcher_Points_DistanceThreshold.cpp:106-113)
#pragma pack(push, 1)
    struct orc_pair_pt2pt
    {
        uint32_t globalIdx, localIdx;
        float    gx, gy, gz;
        float    lx, ly, lz;
        float    errSq;
    };
#pragma pack(pop)
    // mp2p_icp::point_plane_pair_t (point_plane_pair_t.h:34-38, plane_patch.h:30-34), 72 bytes
    struct orc_pair_pt2pl
    {
        double coefs[4];
        double centroid[3];
        float  lx, ly, lz;
        float  _pad;
    };

    // mp2p_icp::point_line_pair_t (Pairings.h:61-73): TLine3D {pBase, director} + TPoint3D pt_local, 72 bytes
    struct orc_pair_pt2ln
    {
        double pBase[3];
        double director[3];
        double lx, ly, lz;
    };

    // Matcher_Point2Line (Matcher_Point2Line.cpp:35-45) + Matcher_Points_Base
    struct orc_match_pt2ln_params
    {
        double   distanceThreshold;
        uint32_t knn;
        uint32_t minimumLinePoints;
        double   lineEigenThreshold;
        int32_t  allowMatchAlreadyMatchedPoints;
        double   bounding_box_intersection_check_epsilon;
    };

    struct orc_match_pt2pt_params
    {
        double   threshold;
        double   thresholdAngularDeg;
        uint32_t pairingsPerPoint;
        int32_t  allowMatchAlreadyMatchedPoints;
        int32_t  allowMatchAlreadyMatchedGlobalPoints;
        double   bounding_box_intersection_check_epsilon;
    };

    struct orc_match_pt2pl_params
    {
        double   distanceThreshold;
        double   searchRadius;
        uint32_t knn;
        uint32_t minimumPlanePoints;
        double   planeEigenThreshold;
        int32_t  allowMatchAlreadyMatchedPoints;
        double   bounding_box_intersection_check_epsilon;
    };

    // Matcher_Points_InlierRatio (mp2p_icp/include/mp2p_icp/Matcher_Points_InlierRatio.h) + Matcher_Points_Base
    struct orc_match_inlier_params
    {
        double  inliersRatio;
        int32_t allowMatchAlreadyMatchedPoints;
        int32_t allowMatchAlreadyMatchedGlobalPoints;
        double  bounding_box_intersection_check_epsilon;
    };

    // Matcher_Adaptive (mp2p_icp/include/mp2p_icp/Matcher_Adaptive.h:66-75, src/Matcher_Adaptive.cpp:32-57) + base
    struct orc_match_adaptive_params
    {
        double   confidenceInterval;
        double   firstToSecondDistanceMax;
        double   absoluteMaxSearchDistance;
        double   minimumCorrDist;
        int32_t  enableDetectPlanes;
        uint32_t planeSearchPoints;
        uint32_t planeMinimumFoundPoints;
        uint32_t maxPt2PtCorrespondences;
        double   planeEigenThreshold;
        double   planeMinimumDistance;
        int32_t  allowMatchAlreadyMatchedPoints;
        int32_t  allowMatchAlreadyMatchedGlobalPoints;
        double   bounding_box_intersection_check_epsilon;
    };

    struct orc_horn_params
    {
        int32_t use_scale_outlier_detector;
        double  scale_outlier_threshold;
        double  w_pt2pt;           // PairWeights::pt2pt
        int32_t robust_kernel;     // 0 None, 1 GemanMcClure, 2 Cauchy
        double  robust_kernel_param;
        double  currentEstimateForRobust[12];
    };

    struct orc_gn_params
    {
        uint32_t maxInnerLoopIterations;
        double   minDelta;
        double   maxCost;
        double   w_pt2pt, w_pt2pl;
        int32_t  kernel;
        double   kernelParam;
    };
}

static_assert(sizeof(orc_pair_pt2pt) == 36, "TMatchingPair layout");
static_assert(sizeof(orc_pair_pt2pl) == 72, "point_plane_pair_t layout");

namespace
{
inline Pose to_pose(const double* T)
{
    Pose p;
    std::memcpy(p.m, T, sizeof(p.m));
    return p;
}

// robust_kernels.h:57-94 — returns the weight factor given the SQUARED error.
inline double robust_weight(int kernel, double param, double errSqr)
{
    const double p2 = param * param;
    switch (kernel)
    {
        case 1:  // GemanMcClure: c^2/(e^2+c)^2  (note `+c`, robust_kernels.h:76-77)
            return p2 / ((errSqr + param) * (errSqr + param));
        case 2:  // Cauchy: c^2/(e^2+c^2)
            return p2 / (errSqr + p2);
        default:
            return 1.0;
    }
}

// mrpt TBoundingBoxf::intersection(other, epsilon) -> has value?
// (Matcher_Points_DistanceThreshold.cpp:73-75; Matcher_Point2Plane.cpp:63-66)
inline bool bbox_intersects(const float amin[3], const float amax[3], const float bmin[3],
                            const float bmax[3], float eps)
{
    if (bmin[0] - eps > amax[0] || bmin[1] - eps > amax[1] || bmin[2] - eps > amax[2] ||
        bmax[0] + eps < amin[0] || bmax[1] + eps < amin[1] || bmax[2] + eps < amin[2])
        return false;
    return true;
}

// Matcher_Points_Base.cpp:183-220 (all-points branch; sub-sampling is out of scope)
void transform_local_to_global(const float* lx, const float* ly, const float* lz, size_t n,
                               const Pose& T, float* gx, float* gy, float* gz, float bbmin[3],
                               float bbmax[3])
{
    const float fMax = std::numeric_limits<float>::max();
    for (int d = 0; d < 3; d++) bbmin[d] = fMax, bbmax[d] = -fMax;
    for (size_t i = 0; i < n; i++)
    {
        compose_point_f(T, lx[i], ly[i], lz[i], gx[i], gy[i], gz[i]);
        bbmax[0] = std::max(bbmax[0], gx[i]), bbmax[1] = std::max(bbmax[1], gy[i]);
        bbmax[2] = std::max(bbmax[2], gz[i]);
        bbmin[0] = std::min(bbmin[0], gx[i]), bbmin[1] = std::min(bbmin[1], gy[i]);
        bbmin[2] = std::min(bbmin[2], gz[i]);
    }
}

// estimate_points_eigen.cpp:27-123 (totalCount branch): mean in FLOAT, centred second moments
// accumulated into a DOUBLE lower-triangular 3x3, scaled by float inv_n, eig ascending.
struct PlaneFit
{
    float  mean[3];
    double eigVals[3];
    double eigVec0[3];  // eigenvector of the smallest eigenvalue
    double eigVec2[3];  // eigenvector of the largest eigenvalue
};
PlaneFit estimate_points_eigen(const float* xs, const float* ys, const float* zs, size_t count)
{
    PlaneFit    r;
    float       mx = 0, my = 0, mz = 0;
    const float inv_n = 1.0f / static_cast<float>(count);
    for (size_t i = 0; i < count; i++) mx += xs[i], my += ys[i], mz += zs[i];
    mx *= inv_n, my *= inv_n, mz *= inv_n;
    double a00 = 0, a10 = 0, a20 = 0, a11 = 0, a21 = 0, a22 = 0;
    for (size_t i = 0; i < count; i++)
    {
        // TPoint3Df - TPoint3Df => float; float*float products are formed in float, then widened
        const float ax = xs[i] - mx, ay = ys[i] - my, az = zs[i] - mz;
        a00 += ax * ax, a10 += ax * ay, a20 += ax * az;
        a11 += ay * ay, a21 += ay * az, a22 += az * az;
    }
    a00 *= inv_n, a10 *= inv_n, a20 *= inv_n, a11 *= inv_n, a21 *= inv_n, a22 *= inv_n;
    const double A[9] = {a00, a10, a20, a10, a11, a21, a20, a21, a22};
    double       V[9], vals[3];
    eig_symmetric<3>(A, V, vals);
    r.mean[0] = mx, r.mean[1] = my, r.mean[2] = mz;
    for (int i = 0; i < 3; i++) r.eigVals[i] = vals[i];
    r.eigVec0[0] = V[0], r.eigVec0[1] = V[3], r.eigVec0[2] = V[6];
    r.eigVec2[0] = V[2], r.eigVec2[1] = V[5], r.eigVec2[2] = V[8];
    return r;
}
}  // namespace

extern "C"
{
    // ---------------------------------------------------------------- SE(3) helpers
    void orc_pose_from_xyzypr(const double v[6], double T[12])
    {
        const Pose p = pose_from_xyzypr(v[0], v[1], v[2], v[3], v[4], v[5]);
        std::memcpy(T, p.m, sizeof(p.m));
    }
    void orc_pose_compose(const double A[12], const double B[12], double out[12])
    {
        const Pose p = compose(to_pose(A), to_pose(B));
        std::memcpy(out, p.m, sizeof(p.m));
    }
    void orc_pose_inverse(const double A[12], double out[12])
    {
        const Pose p = inverse(to_pose(A));
        std::memcpy(out, p.m, sizeof(p.m));
    }
    void orc_se3_exp(const double xi[6], double T[12])
    {
        const Pose p = se3_exp(xi);
        std::memcpy(T, p.m, sizeof(p.m));
    }
    void orc_se3_log(const double T[12], double xi[6]) { se3_log(to_pose(T), xi); }
    void orc_eig_sym3(const double A[9], double V[9], double vals[3]) { eig_symmetric<3>(A, V, vals); }
    void orc_eig_sym4(const double A[16], double V[16], double vals[4])
    {
        eig_symmetric<4>(A, V, vals);
    }
    void orc_ldlt_solve6(const double H[36], const double b[6], double x[6]) { ldlt_solve6(H, b, x); }

    // ---------------------------------------------------------------- transform (a3)
    void orc_transform_local_to_global(const float* lx, const float* ly, const float* lz, size_t n,
                                       const double T[12], float* gx, float* gy, float* gz,
                                       float bbmin[3], float bbmax[3])
    {
        transform_local_to_global(lx, ly, lz, n, to_pose(T), gx, gy, gz, bbmin, bbmax);
    }

    // ---------------------------------------------------------------- NN index (a5)
    void* orc_kdtree_build(const float* x, const float* y, const float* z, size_t n, int leaf_max)
    {
        auto* t = new KDTree();
        t->build(x, y, z, n, leaf_max);
        return t;
    }
    void orc_kdtree_free(void* t) { delete static_cast<KDTree*>(t); }
    int  orc_knn(void* tree, const float q[3], int K, float radius2, uint32_t* idx, float* d2,
                 int bruteforce)
    {
        const auto* t = static_cast<const KDTree*>(tree);
        return bruteforce ? t->knn_bruteforce(q, K, radius2, idx, d2) : t->knn(q, K, radius2, idx, d2);
    }
    // batched variant for tests: nq queries -> idx[nq*K], d2[nq*K], found[nq]
    void orc_knn_batch(void* tree, const float* qx, const float* qy, const float* qz, size_t nq,
                       int K, float radius2, uint32_t* idx, float* d2, int32_t* found,
                       int bruteforce, int nthreads)
    {
        const auto* t = static_cast<const KDTree*>(tree);
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads > 0 ? nthreads : 1)
        for (long i = 0; i < static_cast<long>(nq); i++)
        {
            const float q[3] = {qx[i], qy[i], qz[i]};
            found[i]         = bruteforce ? t->knn_bruteforce(q, K, radius2, idx + i * K, d2 + i * K)
                                          : t->knn(q, K, radius2, idx + i * K, d2 + i * K);
        }
    }

    // ---------------------------------------------------------------- pt2pt matcher (a4)
    // Matcher_Points_DistanceThreshold::implMatchOneLayer, SERIAL branch semantics
    // (Matcher_Points_DistanceThreshold.cpp:48-121,208-265). `nthreads>1` only parallelises the
    // read-only NN searches (like the TBB branch, :128-201) and then replays the claims serially
    // in ascending local index, which is exactly the serial result.
    // `local_paired`/`global_paired`: one byte per point, in/out (MatchState bitfields,
    // Matcher.cpp:46-88, pointcloud_bitfield.h:46-133). Returns the number of pairs appended.
    size_t orc_match_pt2pt(void* tree, const float* lx, const float* ly, const float* lz,
                           size_t nLocal, const double T[12], const orc_match_pt2pt_params* prm,
                           uint8_t* local_paired, uint8_t* global_paired, orc_pair_pt2pt* out,
                           size_t out_capacity, uint64_t* potential_pairings, int nthreads)
    {
        const auto* kd = static_cast<const KDTree*>(tree);
        const Pose  pose = to_pose(T);
        const uint32_t K = prm->pairingsPerPoint;

        if (potential_pairings) *potential_pairings += nLocal * K;  // :64
        if (kd->n == 0 || nLocal == 0) return 0;                     // :67

        std::vector<float> gx(nLocal), gy(nLocal), gz(nLocal);
        float              lmin[3], lmax[3];
        transform_local_to_global(lx, ly, lz, nLocal, pose, gx.data(), gy.data(), gz.data(), lmin,
                                  lmax);  // :69

        // :73-75 bounding boxes must overlap within threshold + epsilon
        const float eps =
            static_cast<float>(prm->threshold + prm->bounding_box_intersection_check_epsilon);
        if (!bbox_intersects(kd->bbmin, kd->bbmax, lmin, lmax, eps)) return 0;

        // :82-83 (double -> float)
        const float maxDistSq = static_cast<float>(prm->threshold * prm->threshold);
        const double angRad   = prm->thresholdAngularDeg * M_PI / 180.0;
        const float  angSq    = static_cast<float>(angRad * angRad);

        // Stage 1: NN search per local point (read-only). For K==1 nn_single_search is unbounded
        // (:235); for K>1 the serial branch uses unbounded nn_multiple_search (:246-248). The
        // acceptance test below rejects everything with d2 >= finalThresSqr, so bounding the search
        // by finalThresSqr is result-equivalent and keeps the oracle fast on outliers.
        std::vector<uint32_t> nnIdx(nLocal * K);
        std::vector<float>    nnD2(nLocal * K);
        std::vector<int32_t>  nnCount(nLocal);
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads > 0 ? nthreads : 1)
        for (long i = 0; i < static_cast<long>(nLocal); i++)
        {
            const float q[3]         = {gx[i], gy[i], gz[i]};
            const float localNormSqr = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];  // :230
            const float finalThresSqr = maxDistSq + angSq * localNormSqr;        // :256-257
            nnCount[i] = kd->knn(q, K, finalThresSqr, &nnIdx[i * K], &nnD2[i * K]);
        }

        // Stage 2: serial claim replay (:214-265 + lambdaAddPair :94-121)
        size_t nOut = 0;
        for (size_t i = 0; i < nLocal; i++)
        {
            if (!prm->allowMatchAlreadyMatchedPoints && local_paired && local_paired[i]) continue;
            for (int k = 0; k < nnCount[i]; k++)
            {
                const uint32_t g  = nnIdx[i * K + k];
                const float    d2 = nnD2[i * K + k];
                // (d2 < finalThresSqr already guaranteed by the bounded search)
                if (!prm->allowMatchAlreadyMatchedGlobalPoints && global_paired && global_paired[g])
                    continue;
                if (nOut < out_capacity)
                {
                    orc_pair_pt2pt& p = out[nOut];
                    p.globalIdx = g, p.localIdx = static_cast<uint32_t>(i);
                    p.gx = kd->x[g], p.gy = kd->y[g], p.gz = kd->z[g];
                    p.lx = lx[i], p.ly = ly[i], p.lz = lz[i];  // ORIGINAL local coords (:111)
                    p.errSq = d2;
                }
                nOut++;
                if (!prm->allowMatchAlreadyMatchedGlobalPoints)
                {
                    if (local_paired) local_paired[i] = 1;
                    if (global_paired) global_paired[g] = 1;
                }
            }
        }
        return nOut;
    }

    // ---------------------------------------------------------------- inlier-ratio matcher
    // Matcher_Points_InlierRatio::implMatchOneLayer (mp2p_icp/src/Matcher_Points_InlierRatio.cpp:41-143):
    // unbounded nn_single_search per local point (:89-91), all tentative pairings sorted by
    // errorSquareAfterTransformation in a std::multimap filled with emplace_hint(begin()) (:104 —
    // equal keys therefore end up in REVERSE insertion order: the hint puts a new element at the
    // lower bound of its key), the first nKeep = mrpt::round(nTotal * inliersRatio) kept (:113),
    // then emitted in sorted order skipping global points already paired, marking both bitfields as
    // it goes (:121-137). The same std::multimap call is used here, so tie order is the reference's
    // by construction. mrpt::round(double) is recalled from MRPT 2.x <mrpt/core/round.h> as
    // lrint / _mm_cvtsd_si32, i.e. round-half-to-even in the default rounding mode (MRPT is not in
    // this container: unverifiable here). Returns the number of pairings, or -1
    // where the reference throws (ASSERT_(nTotal > 0), :111; ratio outside (0,1), :49-50).
    long orc_match_inlier_ratio(void* tree, const float* lx, const float* ly, const float* lz, size_t nLocal,
                                const double T[12], const orc_match_inlier_params* prm, uint8_t* local_paired,
                                uint8_t* global_paired, orc_pair_pt2pt* out, size_t out_capacity,
                                uint64_t* potential_pairings, int nthreads)
    {
        const auto* kd   = static_cast<const KDTree*>(tree);
        const Pose  pose = to_pose(T);
        if (!(prm->inliersRatio > 0.0) || !(prm->inliersRatio < 1.0)) return -1;  // :49-50
        if (potential_pairings) *potential_pairings += nLocal;                    // :55
        if (kd->n == 0 || nLocal == 0) return 0;                                   // :58

        std::vector<float> gx(nLocal), gy(nLocal), gz(nLocal);
        float              lmin[3], lmax[3];
        transform_local_to_global(lx, ly, lz, nLocal, pose, gx.data(), gy.data(), gz.data(), lmin, lmax);  // :60
        if (!bbox_intersects(kd->bbmin, kd->bbmax, lmin, lmax,
                             static_cast<float>(prm->bounding_box_intersection_check_epsilon)))
            return 0;  // :64-67

        std::vector<uint32_t> nnIdx(nLocal);
        std::vector<float>    nnD2(nLocal);
        std::vector<int32_t>  nnCount(nLocal, 0);
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads > 0 ? nthreads : 1)
        for (long i = 0; i < static_cast<long>(nLocal); i++)
        {
            if (!prm->allowMatchAlreadyMatchedPoints && local_paired && local_paired[i]) continue;  // :81-83
            const float q[3] = {gx[i], gy[i], gz[i]};
            nnCount[i]       = kd->knn(q, 1, std::numeric_limits<float>::infinity(), &nnIdx[i], &nnD2[i]);  // :89-91
        }
        std::multimap<double, orc_pair_pt2pt> sortedPairings;  // :76
        for (size_t i = 0; i < nLocal; i++)
        {
            if (nnCount[i] < 1) continue;
            orc_pair_pt2pt p;
            const uint32_t g = nnIdx[i];
            p.globalIdx = g, p.localIdx = static_cast<uint32_t>(i);
            p.gx = kd->x[g], p.gy = kd->y[g], p.gz = kd->z[g];
            p.lx = lx[i], p.ly = ly[i], p.lz = lz[i];
            p.errSq = nnD2[i];
            sortedPairings.emplace_hint(sortedPairings.begin(), static_cast<double>(nnD2[i]), p);  // :104
        }
        const size_t nTotal = sortedPairings.size();
        if (nTotal == 0) return -1;  // :111
        const long nKeep = lrint(static_cast<double>(nTotal) * prm->inliersRatio);  // :113
        auto       itEnd = sortedPairings.begin();
        std::advance(itEnd, nKeep);
        size_t nOut = 0;
        for (auto it = sortedPairings.begin(); it != itEnd; ++it)
        {
            const uint32_t li = it->second.localIdx, gi = it->second.globalIdx;
            if (!prm->allowMatchAlreadyMatchedGlobalPoints && global_paired && global_paired[gi]) continue;  // :126-128
            if (nOut < out_capacity) out[nOut] = it->second;
            nOut++;
            if (local_paired) local_paired[li] = 1;  // :133-135
            if (global_paired) global_paired[gi] = 1;
        }
        return static_cast<long>(nOut);
    }

    // ---------------------------------------------------------------- adaptive matcher
    // Matcher_Adaptive::implMatchOneLayer (mp2p_icp/src/Matcher_Adaptive.cpp:59-314), as written:
    //  1. per local point not yet paired: nn_single_search (unbounded, then dropped if d2 > absMax^2,
    //     :142-155,166) when one neighbour is wanted, else nn_radius_search(absMax^2, maxPoints)
    //     (:156-162); at most MAX_CORRS_PER_LOCAL = 10 neighbours are kept (:108, Matcher_Adaptive.h:84);
    //  2. min / max of the 1st and 2nd neighbour errors (:168-181) -> 50-bin histogram of those errors
    //     (:188-194) -> upper confidence bound ci_high (:196-199) -> maxCorrDistSqr =
    //     max(minimumCorrDist^2, ci_high) (:214);
    //  3. per local point: optional plane through ALL its kept neighbours (:222-268; the point-plane
    //     distance is evaluated with the LOCAL-frame point, :247, as written) -> pt2pl pairing; else
    //     up to maxPt2PtCorrespondences pt2pt pairings with errSq < maxCorrDistSqr, stopping at the
    //     first whose error exceeds firstToSecond^2 times the nearest's (:270-297). Global points are
    //     never marked (:303-311).
    // The two MRPT helpers the threshold rests on are NOT in the reference tree (mrpt-math, version
    // ">= 2.11.5", unpinned) and are restated here from the MRPT 2.x sources as recalled — their
    // parity is UNPINNED (no upstream test exercises this matcher either):
    //   mrpt::math::CHistogram(min, max, nBins): binSizeInv = (nBins - 1) / (max - min);
    //     add(x): ignored outside [min, max], bin = size_t(binSizeInv * (x - min));
    //     getHistogramNormalized: x = linspace(min, max, nBins), hits[i] = bins[i] * binSizeInv / count;
    //   mrpt::math::confidenceIntervalsFromHistogram(x, hits, lo, hi, ci): Hc = cumsum(hits) / max(Hc);
    //     lo = x[lower_bound(Hc, ci)], hi = x[upper_bound(Hc, 1 - ci)].
    // Returns 0, or -1 where the reference throws / dereferences an empty optional (no neighbour at
    // all; all first/second errors equal: CHistogram asserts max > min).
    long orc_match_adaptive(void* tree, const float* lx, const float* ly, const float* lz, size_t nLocal, const double T[12],
                            const orc_match_adaptive_params* prm, uint8_t* local_paired, const uint8_t* global_paired,
                            orc_pair_pt2pt* out2p, size_t cap2p, size_t* n2p_out, orc_pair_pt2pl* out2l, size_t cap2l,
                            size_t* n2l_out, uint64_t* potential_pairings, double* ci_high_out, int nthreads)
    {
        constexpr int MAX_CORRS = 10;
        const auto*   kd        = static_cast<const KDTree*>(tree);
        const Pose    pose      = to_pose(T);
        *n2p_out = *n2l_out = 0;
        if (potential_pairings) *potential_pairings += nLocal * prm->maxPt2PtCorrespondences;  // :68
        if (kd->n == 0 || nLocal == 0) return 0;                                                 // :71
        std::vector<float> gx(nLocal), gy(nLocal), gz(nLocal);
        float              lmin[3], lmax[3];
        transform_local_to_global(lx, ly, lz, nLocal, pose, gx.data(), gy.data(), gz.data(), lmin, lmax);
        if (!bbox_intersects(kd->bbmin, kd->bbmax, lmin, lmax, static_cast<float>(prm->bounding_box_intersection_check_epsilon)))
            return 0;  // :77-80
        const float absMaxSqr = static_cast<float>(prm->absoluteMaxSearchDistance * prm->absoluteMaxSearchDistance);  // :87
        const uint32_t nnMax  = prm->enableDetectPlanes ? prm->planeSearchPoints : prm->maxPt2PtCorrespondences;     // :119-120
        const int      K      = static_cast<int>(std::min<uint32_t>(nnMax, MAX_CORRS));
        // nn_single_search keeps d2 <= absMax^2 (:166), nn_radius_search d2 < absMax^2 (nanoflann)
        const float radius = nnMax == 1 ? std::nextafterf(absMaxSqr, std::numeric_limits<float>::infinity()) : absMaxSqr;
        std::vector<uint32_t> idx(nLocal * K);
        std::vector<float>    d2(nLocal * K);
        std::vector<int32_t>  cnt(nLocal, 0);
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads > 0 ? nthreads : 1)
        for (long i = 0; i < static_cast<long>(nLocal); i++)
        {
            if (!prm->allowMatchAlreadyMatchedPoints && local_paired && local_paired[i]) continue;  // :126-132
            const float q[3] = {gx[i], gy[i], gz[i]};
            cnt[i]           = kd->knn(q, K, radius, &idx[i * K], &d2[i * K]);
        }
        bool  any = false;
        float emin = 0, emax = 0;
        for (size_t i = 0; i < nLocal; i++)
            for (int k = 0; k < std::min(cnt[i], 2); k++)  // :168-181
            {
                const float e = d2[i * K + k];
                if (!any) emin = emax = e, any = true;
                emin = std::min(emin, e), emax = std::max(emax, e);
            }
        if (!any) return -1;  // *minSqrErrorForHistogram on an empty optional
        // ---- mrpt::math::CHistogram hist(min, max, 50) (:188-194)
        constexpr int NB = 50;
        const double  hmin = emin, hmax = emax;
        if (!(hmax > hmin)) return -1;  // CHistogram: ASSERT_(max > min)
        const double binSizeInv = (static_cast<double>(NB) - 1) / (hmax - hmin);
        size_t       bins[NB]   = {0}, count = 0;
        for (size_t i = 0; i < nLocal; i++)
            for (int k = 0; k < std::min(cnt[i], 2); k++)
            {
                const double x = d2[i * K + k];
                if (x < hmin || x > hmax) continue;
                bins[static_cast<size_t>(binSizeInv * (x - hmin))]++;
                count++;
            }
        double xs[NB], hits[NB], Hc[NB];
        for (int b = 0; b < NB; b++) xs[b] = hmin + b * (hmax - hmin) / (NB - 1), hits[b] = (binSizeInv / count) * bins[b];
        // ---- confidenceIntervalsFromHistogram(xs, hits, lo, hi, 1 - confidenceInterval) (:196-199)
        double acc = 0;
        for (int b = 0; b < NB; b++) acc += hits[b], Hc[b] = acc;
        const double mx = *std::max_element(Hc, Hc + NB);
        for (int b = 0; b < NB; b++) Hc[b] *= 1.0 / mx;
        const double ci      = 1.0 - prm->confidenceInterval;
        const double ci_high = xs[std::min<size_t>(NB - 1, std::upper_bound(Hc, Hc + NB, 1.0 - ci) - Hc)];
        if (ci_high_out) *ci_high_out = ci_high;
        const double maxCorrDistSqr = std::max(prm->minimumCorrDist * prm->minimumCorrDist, ci_high);  // :214
        const float  maxSqr1to2     = static_cast<float>(prm->firstToSecondDistanceMax * prm->firstToSecondDistanceMax);  // :216

        size_t n2p = 0, n2l = 0;
        for (size_t i = 0; i < nLocal; i++)  // :219-298
        {
            const int c = cnt[i];
            if (prm->enableDetectPlanes && c >= static_cast<int>(prm->planeMinimumFoundPoints))
            {
                float xs_[MAX_CORRS], ys_[MAX_CORRS], zs_[MAX_CORRS];
                for (int k = 0; k < c; k++) xs_[k] = kd->x[idx[i * K + k]], ys_[k] = kd->y[idx[i * K + k]], zs_[k] = kd->z[idx[i * K + k]];
                const PlaneFit f = estimate_points_eigen(xs_, ys_, zs_, c);
                if (f.eigVals[0] < prm->planeEigenThreshold * f.eigVals[2] && f.eigVals[0] < prm->planeEigenThreshold * f.eigVals[1])
                {
                    const double n[3]  = {f.eigVec0[0], f.eigVec0[1], f.eigVec0[2]};
                    const double cxyz[3] = {f.mean[0], f.mean[1], f.mean[2]};
                    const double D     = -(n[0] * cxyz[0] + n[1] * cxyz[1] + n[2] * cxyz[2]);  // TPlane(point, normal)
                    // :247 thePlane.distance(mspl.at(0).local): the LOCAL-frame point, as written
                    const double ev   = n[0] * lx[i] + n[1] * ly[i] + n[2] * lz[i] + D;
                    const double dist = std::fabs(std::fabs(ev) / std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]));
                    if (dist < prm->planeMinimumDistance)
                    {
                        if (n2l < cap2l)
                        {
                            orc_pair_pt2pl& p = out2l[n2l];
                            p.coefs[0] = n[0], p.coefs[1] = n[1], p.coefs[2] = n[2], p.coefs[3] = D;
                            p.centroid[0] = cxyz[0], p.centroid[1] = cxyz[1], p.centroid[2] = cxyz[2];
                            p.lx = lx[i], p.ly = ly[i], p.lz = lz[i], p._pad = 0;
                        }
                        n2l++;
                        if (local_paired) local_paired[i] = 1;  // :262
                        continue;
                    }
                }
            }
            for (int k = 0; k < std::min<int>(c, prm->maxPt2PtCorrespondences); k++)  // :270-297
            {
                const uint32_t g = idx[i * K + k];
                const float    e = d2[i * K + k];
                if (!prm->allowMatchAlreadyMatchedGlobalPoints && global_paired && global_paired[g]) continue;
                if (e >= maxCorrDistSqr) continue;
                if (k != 0 && e > d2[i * K] * maxSqr1to2) break;
                if (n2p < cap2p)
                {
                    orc_pair_pt2pt& p = out2p[n2p];
                    p.globalIdx = g, p.localIdx = static_cast<uint32_t>(i);
                    p.gx = kd->x[g], p.gy = kd->y[g], p.gz = kd->z[g];
                    p.lx = lx[i], p.ly = ly[i], p.lz = lz[i];
                    p.errSq = e;
                }
                n2p++;
                if (!prm->allowMatchAlreadyMatchedGlobalPoints && local_paired) local_paired[i] = 1;  // :291-295
            }
        }
        *n2p_out = n2p, *n2l_out = n2l;
        return 0;
    }

    // ---------------------------------------------------------------- pt2ln matcher
    // Matcher_Point2Line::implMatchOneLayer (mp2p_icp/src/Matcher_Point2Line.cpp:46-163): UNBOUNDED
    // nn_multiple_search of `knn` points (:103-105); the neighbour list is cut at the first distance
    // > distanceThreshold^2 (:110-127) and must keep >= minimumLinePoints entries (:130) — but only
    // the index and distance vectors are cut, kddPts is not, so estimate_points_eigen runs over ALL
    // knn points (:132-135). Restated as written. Line test :148-149, director = eigVectors[2]
    // unitarized, pBase = the (float) mean (:151-156); the local point is marked, global points never
    // (:92-95). No upstream test exercises this matcher: parity of this function is pinned only by the
    // hand-made known answers in tests/test_oracle_golden.py.
    size_t orc_match_pt2ln(void* tree, const float* lx, const float* ly, const float* lz, size_t nLocal,
                           const double T[12], const orc_match_pt2ln_params* prm, uint8_t* local_paired,
                           orc_pair_pt2ln* out, size_t out_capacity, uint64_t* potential_pairings, int nthreads)
    {
        const auto* kd   = static_cast<const KDTree*>(tree);
        const Pose  pose = to_pose(T);
        if (potential_pairings) *potential_pairings += nLocal;  // :58
        if (kd->n == 0 || nLocal == 0) return 0;                 // :61
        std::vector<float> gx(nLocal), gy(nLocal), gz(nLocal);
        float              lmin[3], lmax[3];
        transform_local_to_global(lx, ly, lz, nLocal, pose, gx.data(), gy.data(), gz.data(), lmin, lmax);
        const float eps = static_cast<float>(prm->distanceThreshold + prm->bounding_box_intersection_check_epsilon);
        if (!bbox_intersects(kd->bbmin, kd->bbmax, lmin, lmax, eps)) return 0;  // :67-70
        const int   K      = static_cast<int>(prm->knn);
        const float maxSqr = static_cast<float>(prm->distanceThreshold * prm->distanceThreshold);  // :77
        std::vector<uint8_t>        ok(nLocal, 0);
        std::vector<orc_pair_pt2ln> cand(nLocal);
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
        {
            std::vector<uint32_t> idx(K);
            std::vector<float>    d2(K), xs(K), ys(K), zs(K);
#pragma omp for schedule(dynamic, 256)
            for (long i = 0; i < static_cast<long>(nLocal); i++)
            {
                if (!prm->allowMatchAlreadyMatchedPoints && local_paired && local_paired[i]) continue;  // :88-90
                const float q[3] = {gx[i], gy[i], gz[i]};
                const int   cnt  = kd->knn(q, K, std::numeric_limits<float>::infinity(), idx.data(), d2.data());
                int within = cnt;  // :110-127
                for (int j = 0; j < cnt; j++)
                    if (d2[j] > maxSqr)
                    {
                        within = j;
                        break;
                    }
                if (within < static_cast<int>(prm->minimumLinePoints)) continue;  // :130
                for (int k = 0; k < cnt; k++) xs[k] = kd->x[idx[k]], ys[k] = kd->y[idx[k]], zs[k] = kd->z[idx[k]];
                const PlaneFit f = estimate_points_eigen(xs.data(), ys.data(), zs.data(), cnt);  // :134-135 (all cnt points)
                if (f.eigVals[0] > prm->lineEigenThreshold * f.eigVals[2]) continue;  // :148
                if (f.eigVals[1] > prm->lineEigenThreshold * f.eigVals[2]) continue;  // :149
                orc_pair_pt2ln& p = cand[i];
                p.lx = lx[i], p.ly = ly[i], p.lz = lz[i];  // :152
                for (int d = 0; d < 3; d++) p.pBase[d] = f.mean[d];
                const double* v = f.eigVec2;
                const double  inv = 1.0 / std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);  // TPoint3D::unitarize
                for (int d = 0; d < 3; d++) p.director[d] = v[d] * inv;
                ok[i] = 1;
            }
        }
        size_t nOut = 0;
        for (size_t i = 0; i < nLocal; i++)
        {
            if (!ok[i]) continue;
            if (nOut < out_capacity) out[nOut] = cand[i];
            nOut++;
            if (local_paired) local_paired[i] = 1;  // :159
        }
        return nOut;
    }

    // ---------------------------------------------------------------- pt2pl matcher (a7, a7')
    // Matcher_Point2Plane::implMatchOneLayer (Matcher_Point2Plane.cpp:41-114) over a plain point
    // layer, with nn_search_pt2pl defined as in SURVEY §8a-7' (k-NN within searchRadius ->
    // estimate_points_eigen -> planarity test -> TPlane(centroid, normal) -> |distance|), the
    // in-tree statement of which is Matcher_Adaptive.cpp:229-253. PARITY UNPINNED upstream beyond
    // tests/test-mp2p_matcher_pt2pl.cpp (disabled).
    size_t orc_match_pt2pl(void* tree, const float* lx, const float* ly, const float* lz,
                           size_t nLocal, const double T[12], const orc_match_pt2pl_params* prm,
                           uint8_t* local_paired, orc_pair_pt2pl* out, size_t out_capacity,
                           uint64_t* potential_pairings, int nthreads)
    {
        const auto* kd   = static_cast<const KDTree*>(tree);
        const Pose  pose = to_pose(T);
        if (potential_pairings) *potential_pairings += nLocal;  // :54
        if (kd->n == 0 || nLocal == 0) return 0;

        std::vector<float> gx(nLocal), gy(nLocal), gz(nLocal);
        float              lmin[3], lmax[3];
        transform_local_to_global(lx, ly, lz, nLocal, pose, gx.data(), gy.data(), gz.data(), lmin,
                                  lmax);
        const float eps = static_cast<float>(prm->distanceThreshold +
                                             prm->bounding_box_intersection_check_epsilon);
        if (!bbox_intersects(kd->bbmin, kd->bbmax, lmin, lmax, eps)) return 0;  // :63-66

        const int   K        = static_cast<int>(prm->knn);
        const float radiusSq = static_cast<float>(prm->searchRadius * prm->searchRadius);
        const float distThr  = static_cast<float>(prm->distanceThreshold);

        std::vector<uint8_t>        ok(nLocal, 0);
        std::vector<orc_pair_pt2pl> cand(nLocal);
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
        {
            std::vector<uint32_t> idx(K);
            std::vector<float>    d2(K), xs(K), ys(K), zs(K);
#pragma omp for schedule(dynamic, 256)
            for (long i = 0; i < static_cast<long>(nLocal); i++)
            {
                if (!prm->allowMatchAlreadyMatchedPoints && local_paired && local_paired[i]) continue;
                const float q[3] = {gx[i], gy[i], gz[i]};
                const int   cnt  = kd->knn(q, K, radiusSq, idx.data(), d2.data());
                if (cnt < 3 || cnt < static_cast<int>(prm->minimumPlanePoints)) continue;
                for (int k = 0; k < cnt; k++) xs[k] = kd->x[idx[k]], ys[k] = kd->y[idx[k]], zs[k] = kd->z[idx[k]];
                const PlaneFit f = estimate_points_eigen(xs.data(), ys.data(), zs.data(), cnt);
                // Matcher_Adaptive.cpp:232-233
                if (!(f.eigVals[0] < prm->planeEigenThreshold * f.eigVals[2] &&
                      f.eigVals[0] < prm->planeEigenThreshold * f.eigVals[1]))
                    continue;
                // TPlane(point, normal): unit normal, D = -n.p   (Matcher_Adaptive.cpp:250)
                const double cx = f.mean[0], cy = f.mean[1], cz = f.mean[2];
                double       nx = f.eigVec0[0], ny = f.eigVec0[1], nz = f.eigVec0[2];
                const double inv_nn = 1.0 / std::sqrt(nx * nx + ny * ny + nz * nz);
                nx *= inv_nn, ny *= inv_nn, nz *= inv_nn;
                const double D = -nx * cx - ny * cy - nz * cz;
                // TPlane::distance(p) = |A x + B y + C z + D| / |(A,B,C)|
                const double ev   = nx * q[0] + ny * q[1] + nz * q[2] + D;
                const double dist = std::fabs(ev) / std::sqrt(nx * nx + ny * ny + nz * nz);
                // NearestPlaneResult::distance is a float (NearestPlaneCapable.h:46);
                // Matcher_Point2Plane.cpp:101 rejects `np.distance > distanceThreshold`
                if (static_cast<float>(dist) > distThr) continue;
                orc_pair_pt2pl& p = cand[i];
                p.coefs[0] = nx, p.coefs[1] = ny, p.coefs[2] = nz, p.coefs[3] = D;
                p.centroid[0] = cx, p.centroid[1] = cy, p.centroid[2] = cz;
                p.lx = lx[i], p.ly = ly[i], p.lz = lz[i];  // ORIGINAL local point (:105)
                p._pad = 0;
                ok[i]  = 1;
            }
        }
        size_t nOut = 0;
        for (size_t i = 0; i < nLocal; i++)
        {
            if (!ok[i]) continue;
            if (nOut < out_capacity) out[nOut] = cand[i];
            nOut++;
            if (local_paired) local_paired[i] = 1;  // :109 (global never deduped, :87-90)
        }
        return nOut;
    }

    // ---------------------------------------------------------------- Horn solver (a14)
    // optimal_tf_horn (optimal_tf_horn.cpp:201-252) = eval_centroids_robust (Pairings.cpp:68-110)
    // + se3_l2_internal (optimal_tf_horn.cpp:77-199) over visit_correspondences
    // (visit_correspondences.h:39-221); pt2pt pairings only (ln2ln / pl2pl stay host-side,
    // SURVEY §2). `weights_count/weights_value`: Pairings::point_weights run-length blocks.
    // Returns 1 on success, 0 if fewer than 3 pairings (optimal_tf_horn.cpp:96).
    int orc_optimal_tf_horn(const orc_pair_pt2pt* pairs, size_t n, const orc_horn_params* wp,
                            const uint64_t* weights_count, const double* weights_value,
                            size_t n_weight_blocks, double T_out[12], uint64_t* n_outliers_out)
    {
        if (n < 3) return 0;
        std::vector<size_t> outliers;  // OutlierIndices::point2point (sorted ascending)
        const Pose          robustRef = to_pose(wp->currentEstimateForRobust);

        auto centroids = [&](const std::vector<size_t>& outl, double cl[3], double cg[3])
        {
            // Pairings.cpp:80-107
            const double w = 1.0 / static_cast<double>(n - outl.size());
            for (int d = 0; d < 3; d++) cl[d] = cg[d] = 0;
            size_t io = 0;
            for (size_t i = 0; i < n; i++)
            {
                if (io < outl.size() && i == outl[io])
                {
                    io++;
                    continue;
                }
                cg[0] += pairs[i].gx, cg[1] += pairs[i].gy, cg[2] += pairs[i].gz;
                cl[0] += pairs[i].lx, cl[1] += pairs[i].ly, cl[2] += pairs[i].lz;
            }
            for (int d = 0; d < 3; d++) cl[d] *= w, cg[d] *= w;
        };

        double q[4];
        auto   se3_l2 = [&](const double cl[3], const double cg[3]) -> bool
        {
            double S[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            // visit_correspondences.h:66-78
            std::vector<std::pair<uint64_t, double>> blocks;
            for (size_t b = 0; b < n_weight_blocks; b++) blocks.emplace_back(weights_count[b], weights_value[b]);
            if (blocks.empty()) blocks.emplace_back(n, 1.0);
            size_t       blk = 0, blk_start = 0;
            const double waPoints = wp->w_pt2pt / (wp->w_pt2pt * static_cast<double>(n));  // :85-87
            double       w_sum    = 0;
            std::vector<size_t> new_outliers;
            size_t              io = 0;
            for (size_t i = 0; i < n; i++)
            {
                if (io < outliers.size() && i == outliers[io])  // :108-114
                {
                    io++;
                    new_outliers.push_back(i);
                    continue;
                }
                double wi = waPoints;
                if (i >= blk_start + blocks[blk].first)  // :127-132
                {
                    blk++;
                    blk_start = i;
                }
                wi *= blocks[blk].second;
                const double bi[3] = {pairs[i].gx - cg[0], pairs[i].gy - cg[1], pairs[i].gz - cg[2]};
                const double ri[3] = {pairs[i].lx - cl[0], pairs[i].ly - cl[1], pairs[i].lz - cl[2]};
                const double bn = std::sqrt(bi[0] * bi[0] + bi[1] * bi[1] + bi[2] * bi[2]);
                const double rn = std::sqrt(ri[0] * ri[0] + ri[1] * ri[1] + ri[2] * ri[2]);
                if (bn < 1e-4 || rn < 1e-4) continue;  // :141-146
                if (wp->use_scale_outlier_detector)    // :158-169
                {
                    const double mism = std::max(bn, rn) / std::min(bn, rn);
                    if (mism > wp->scale_outlier_threshold)
                    {
                        new_outliers.push_back(i);
                        continue;
                    }
                }
                if (wp->robust_kernel != 0)  // :200-210 (CPose3D::composePoint(TVector3D))
                {
                    double rx, ry, rz;
                    compose_point(robustRef, ri[0], ri[1], ri[2], rx, ry, rz);
                    const double e2 = (rx - bi[0]) * (rx - bi[0]) + (ry - bi[1]) * (ry - bi[1]) +
                                      (rz - bi[2]) * (rz - bi[2]);
                    wi *= robust_weight(wp->robust_kernel, wp->robust_kernel_param, e2);
                }
                w_sum += wi;
                // optimal_tf_horn.cpp:101-117   S += w * r * b^T
                S[0] += wi * ri[0] * bi[0], S[1] += wi * ri[0] * bi[1], S[2] += wi * ri[0] * bi[2];
                S[3] += wi * ri[1] * bi[0], S[4] += wi * ri[1] * bi[1], S[5] += wi * ri[1] * bi[2];
                S[6] += wi * ri[2] * bi[0], S[7] += wi * ri[2] * bi[1], S[8] += wi * ri[2] * bi[2];
            }
            outliers = std::move(new_outliers);
            if (w_sum > 0)
                for (double& s : S) s *= 1.0 / w_sum;  // :121-124
            // :132-152
            double N[16];
            N[0]  = S[0] + S[4] + S[8];
            N[1]  = S[5] - S[7];
            N[2]  = S[6] - S[2];
            N[3]  = S[1] - S[3];
            N[4]  = N[1];
            N[5]  = S[0] - S[4] - S[8];
            N[6]  = S[1] + S[3];
            N[7]  = S[6] + S[2];
            N[8]  = N[2];
            N[9]  = N[6];
            N[10] = -S[0] + S[4] - S[8];
            N[11] = S[5] + S[7];
            N[12] = N[3];
            N[13] = N[7];
            N[14] = N[11];
            N[15] = -S[0] - S[4] + S[8];
            double Z[16], vals[4];
            eig_symmetric<4>(N, Z, vals);  // :156-160 ascending -> last column
            for (int i = 0; i < 4; i++) q[i] = Z[4 * i + 3];
            if (q[0] < 0)
                for (double& v : q) v = -v;  // :165-171
            return true;
        };

        double cl[3], cg[3];
        centroids(outliers, cl, cg);  // optimal_tf_horn.cpp:216
        if (!se3_l2(cl, cg)) return 0;
        if (wp->use_scale_outlier_detector && !outliers.empty())  // :224-235
        {
            if (outliers.size() >= n) return 0;
            const std::vector<size_t> o = outliers;
            centroids(o, cl, cg);
            if (!se3_l2(cl, cg)) return 0;
        }
        Pose   R = pose_from_quat(q[0], q[1], q[2], q[3]);  // :238
        double px, py, pz;
        compose_point(R, cl[0], cl[1], cl[2], px, py, pz);  // :242
        R.m[3] = cg[0] - px, R.m[7] = cg[1] - py, R.m[11] = cg[2] - pz;  // :245-247
        std::memcpy(T_out, R.m, sizeof(R.m));
        if (n_outliers_out) *n_outliers_out = outliers.size();
        return 1;
    }

    // pt2ln_pl_to_pt2pt (pt2ln_pl_to_pt2pt.cpp:25-113), plane part only: project each local point
    // on its plane, keep pairings whose |d| >= 25 % of the largest (or at least 3), largest first.
    size_t orc_pt2pl_to_pt2pt(const orc_pair_pt2pl* in, size_t n, const double T_guess[12],
                              orc_pair_pt2pt* out, size_t cap)
    {
        const Pose                               rel = to_pose(T_guess);
        std::multimap<double, orc_pair_pt2pt>    sorted;
        for (size_t i = 0; i < n; i++)
        {
            double gx, gy, gz;
            compose_point(rel, in[i].lx, in[i].ly, in[i].lz, gx, gy, gz);
            const double* c = in[i].coefs;
            const double  d = c[0] * gx + c[1] * gy + c[2] * gz + c[3];
            orc_pair_pt2pt p;
            p.globalIdx = p.localIdx = 0;
            p.gx = static_cast<float>(gx - c[0] * d), p.gy = static_cast<float>(gy - c[1] * d);
            p.gz = static_cast<float>(gz - c[2] * d);
            p.lx = in[i].lx, p.ly = in[i].ly, p.lz = in[i].lz;
            p.errSq = 0;
            sorted.insert({std::fabs(d), p});
        }
        size_t nOut = 0;
        if (sorted.empty()) return 0;
        const double thr = sorted.rbegin()->first * 0.25;
        for (auto it = sorted.rbegin(); it != sorted.rend(); ++it)
        {
            if (it->first < thr && nOut >= 3) break;
            if (nOut < cap) out[nOut] = it->second;
            nOut++;
        }
        return nOut;
    }

    // ---------------------------------------------------------------- Gauss-Newton solver (a11-a13)
    // One accumulation of the normal equations at pose T over pt2pt and pt2pl pairings:
    //   H = sum w J^T J (6x6), g = sum w J^T e (6), with Ji = J1 * jacob_dDexpe_de(T)
    // (optimal_tf_gauss_newton.cpp:149-180 pt2pt, :267-286 pt2pl; errorTerms.cpp:36-66,115-161).
    // J1*dDexpe_de is evaluated in the closed form Ji = A [R | -R [l]x], A = I (pt2pt) or
    // -n n^T/|n|^2 (pt2pl), validated by the finite-difference test the reference
    // uses (tests/test-mp2p_error_terms_jacobians.cpp:63-101).
    // out_errNormSqr follows the TBB/pt2pt convention  sum weight*|e|^2.
    // pt2ln term (optimal_tf_gauss_newton.cpp:182-203, errorTerms.cpp:67-113): e = q - u (u.q),
    // q = T(+)l - pBase, J1 = (I - u u^T)-like matrix AS WRITTEN (:92-97, no assumption |u| = 1) times
    // [l_x I .. I], weight = w_pt2ln * robust(|e|^2); the cost term is weight^2 |e|^2 there (:197).
    void orc_gn_accumulate_ex(const orc_pair_pt2pt* p2p, size_t n2p, const orc_pair_pt2pl* p2l, size_t n2l,
                              const orc_pair_pt2ln* p2ln, size_t n2ln, double w_pt2ln, const double T[12],
                              const orc_gn_params* prm, double H[36], double g[6], double* out_errNormSqr, int nthreads);
    void orc_gn_accumulate(const orc_pair_pt2pt* p2p, size_t n2p, const orc_pair_pt2pl* p2l,
                           size_t n2l, const double T[12], const orc_gn_params* prm, double H[36],
                           double g[6], double* out_errNormSqr, int nthreads)
    {
        orc_gn_accumulate_ex(p2p, n2p, p2l, n2l, nullptr, 0, 1.0, T, prm, H, g, out_errNormSqr, nthreads);
    }
    void orc_gn_accumulate_ex(const orc_pair_pt2pt* p2p, size_t n2p, const orc_pair_pt2pl* p2l, size_t n2l,
                              const orc_pair_pt2ln* p2ln, size_t n2ln, double w_pt2ln, const double T[12],
                              const orc_gn_params* prm, double H[36], double g[6], double* out_errNormSqr, int nthreads)
    {
        const Pose pose = to_pose(T);
        const int  nt   = nthreads > 0 ? nthreads : 1;
        std::vector<double> partial(static_cast<size_t>(nt) * 43, 0.0);
#pragma omp parallel num_threads(nt)
        {
            double* acc = &partial[static_cast<size_t>(omp_get_thread_num()) * 43];
            auto    add = [&](const double J[3][6], const double e[3], double w)
            {
                for (int a = 0; a < 6; a++)
                {
                    acc[36 + a] += w * (J[0][a] * e[0] + J[1][a] * e[1] + J[2][a] * e[2]);
                    for (int b = 0; b < 6; b++)
                        acc[6 * a + b] += w * (J[0][a] * J[0][b] + J[1][a] * J[1][b] + J[2][a] * J[2][b]);
                }
            };
#pragma omp for schedule(static)
            for (long i = 0; i < static_cast<long>(n2p); i++)
            {
                const double l[3] = {p2p[i].lx, p2p[i].ly, p2p[i].lz};
                double       gx, gy, gz;
                compose_point(pose, l[0], l[1], l[2], gx, gy, gz);
                const double e[3] = {gx - p2p[i].gx, gy - p2p[i].gy, gz - p2p[i].gz};
                double       J[3][6];
                for (int r = 0; r < 3; r++)
                {
                    const double R0 = pose.R(r, 0), R1 = pose.R(r, 1), R2 = pose.R(r, 2);
                    J[r][0] = R0, J[r][1] = R1, J[r][2] = R2;
                    // -R [l]x : row r = (R2*ly - R1*lz, R0*lz - R2*lx, R1*lx - R0*ly)
                    J[r][3] = R2 * l[1] - R1 * l[2];
                    J[r][4] = R0 * l[2] - R2 * l[0];
                    J[r][5] = R1 * l[0] - R0 * l[1];
                }
                const double e2 = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
                double       w  = prm->w_pt2pt;
                if (prm->kernel != 0) w *= robust_weight(prm->kernel, prm->kernelParam, e2);
                acc[42] += w * e2;
                add(J, e, w);
            }
#pragma omp for schedule(static)
            for (long i = 0; i < static_cast<long>(n2l); i++)
            {
                const double  l[3] = {p2l[i].lx, p2l[i].ly, p2l[i].lz};
                const double* c    = p2l[i].coefs;
                double        gx, gy, gz;
                compose_point(pose, l[0], l[1], l[2], gx, gy, gz);
                const double mod_n = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
                const double ev    = c[0] * gx + c[1] * gy + c[2] * gz + c[3];
                const double e[3]  = {-(c[0] / mod_n) * ev, -(c[1] / mod_n) * ev, -(c[2] / mod_n) * ev};
                double       B[3][6];
                for (int r = 0; r < 3; r++)
                {
                    const double R0 = pose.R(r, 0), R1 = pose.R(r, 1), R2 = pose.R(r, 2);
                    B[r][0] = R0, B[r][1] = R1, B[r][2] = R2;
                    B[r][3] = R2 * l[1] - R1 * l[2];
                    B[r][4] = R0 * l[2] - R2 * l[0];
                    B[r][5] = R1 * l[0] - R0 * l[1];
                }
                double J[3][6];
                for (int r = 0; r < 3; r++)
                    for (int a = 0; a < 6; a++)
                        J[r][a] = -(c[r] * c[0] * B[0][a] + c[r] * c[1] * B[1][a] + c[r] * c[2] * B[2][a]) / mod_n;
                const double e2 = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
                double       w  = prm->w_pt2pl;
                if (prm->kernel != 0) w *= robust_weight(prm->kernel, prm->kernelParam, e2);
                acc[42] += w * e2;
                add(J, e, w);
            }
#pragma omp for schedule(static)
            for (long i = 0; i < static_cast<long>(n2ln); i++)
            {
                const double  l[3] = {p2ln[i].lx, p2ln[i].ly, p2ln[i].lz};
                const double* u    = p2ln[i].director;
                double        gx, gy, gz;
                compose_point(pose, l[0], l[1], l[2], gx, gy, gz);
                const double q[3] = {gx - p2ln[i].pBase[0], gy - p2ln[i].pBase[1], gz - p2ln[i].pBase[2]};
                const double uq   = u[0] * q[0] + u[1] * q[1] + u[2] * q[2];
                const double e[3] = {q[0] - u[0] * uq, q[1] - u[1] * uq, q[2] - u[2] * uq};
                double       B[3][6];
                for (int r = 0; r < 3; r++)
                {
                    const double R0 = pose.R(r, 0), R1 = pose.R(r, 1), R2 = pose.R(r, 2);
                    B[r][0] = R0, B[r][1] = R1, B[r][2] = R2;
                    B[r][3] = R2 * l[1] - R1 * l[2];
                    B[r][4] = R0 * l[2] - R2 * l[0];
                    B[r][5] = R1 * l[0] - R0 * l[1];
                }
                double J[3][6];
                for (int r = 0; r < 3; r++)
                    for (int a = 0; a < 6; a++)
                    {
                        double v = 0;
                        for (int c = 0; c < 3; c++) v += ((r == c ? 1.0 : 0.0) - u[r] * u[c]) * B[c][a];
                        J[r][a] = v;
                    }
                const double e2 = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
                double       w  = w_pt2ln;
                if (prm->kernel != 0) w *= robust_weight(prm->kernel, prm->kernelParam, e2);
                acc[42] += w * w * e2;  // :197 (weight squared, as written)
                add(J, e, w);
            }
        }
        for (int k = 0; k < 36; k++) H[k] = 0;
        for (int k = 0; k < 6; k++) g[k] = 0;
        double err = 0;
        for (int t = 0; t < nt; t++)  // fixed-order combine => run-to-run stable
        {
            for (int k = 0; k < 36; k++) H[k] += partial[static_cast<size_t>(t) * 43 + k];
            for (int k = 0; k < 6; k++) g[k] += partial[static_cast<size_t>(t) * 43 + 36 + k];
            err += partial[static_cast<size_t>(t) * 43 + 42];
        }
        if (out_errNormSqr) *out_errNormSqr = err;
    }

    // optimal_tf_gauss_newton (optimal_tf_gauss_newton.cpp:36-372) restricted to pt2pt + pt2pl
    // terms, no prior. H and g are ZEROED every inner iteration (the TBB / mathematically intended
    // behaviour; the non-TBB build never zeroes them — SURVEY Q4, documented deviation).
    int orc_optimal_tf_gauss_newton_ex(const orc_pair_pt2pt* p2p, size_t n2p, const orc_pair_pt2pl* p2l, size_t n2l,
                                       const orc_pair_pt2ln* p2ln, size_t n2ln, double w_pt2ln, const orc_gn_params* prm,
                                       const double T_init[12], double T_out[12], uint32_t* iters_done, int nthreads);
    int orc_optimal_tf_gauss_newton(const orc_pair_pt2pt* p2p, size_t n2p, const orc_pair_pt2pl* p2l,
                                    size_t n2l, const orc_gn_params* prm, const double T_init[12],
                                    double T_out[12], uint32_t* iters_done, int nthreads)
    {
        return orc_optimal_tf_gauss_newton_ex(p2p, n2p, p2l, n2l, nullptr, 0, 1.0, prm, T_init, T_out, iters_done, nthreads);
    }
    int orc_optimal_tf_gauss_newton_ex(const orc_pair_pt2pt* p2p, size_t n2p, const orc_pair_pt2pl* p2l, size_t n2l,
                                       const orc_pair_pt2ln* p2ln, size_t n2ln, double w_pt2ln, const orc_gn_params* prm,
                                       const double T_init[12], double T_out[12], uint32_t* iters_done, int nthreads)
    {
        Pose     pose = to_pose(T_init);  // :50
        uint32_t it = 0, updates = 0;
        for (; it < prm->maxInnerLoopIterations; it++)
        {
            double H[36], g[6], errSq = 0;
            orc_gn_accumulate_ex(p2p, n2p, p2l, n2l, p2ln, n2ln, w_pt2ln, pose.m, prm, H, g, &errSq, nthreads);
            if (std::sqrt(errSq) <= prm->maxCost) break;  // :344-346
            double delta[6], mg[6];
            for (int k = 0; k < 6; k++) mg[k] = -g[k];
            ldlt_solve6(H, mg, delta);  // :351
            pose = compose(pose, se3_exp(delta));  // :354-356
            updates++;
            double nrm = 0;
            for (double d : delta) nrm += d * d;
            if (std::sqrt(nrm) < prm->minDelta) break;  // :365
        }
        std::memcpy(T_out, pose.m, sizeof(pose.m));
        if (iters_done) *iters_done = updates;  // number of pose updates applied
        return 1;
    }

    // Raw residual/Jacobian of one pairing, for the finite-difference test
    // (tests/test-mp2p_error_terms_jacobians.cpp).  kind 0 = pt2pt, 1 = pt2pl.
    void orc_error_and_jacobian_pt2ln(const orc_pair_pt2ln* pn, const double T[12], double e[3], double J[18])
    {
        const Pose   pose = to_pose(T);
        const double l[3] = {pn->lx, pn->ly, pn->lz};
        const double* u   = pn->director;
        double       gx, gy, gz;
        compose_point(pose, l[0], l[1], l[2], gx, gy, gz);
        const double q[3] = {gx - pn->pBase[0], gy - pn->pBase[1], gz - pn->pBase[2]};
        const double uq   = u[0] * q[0] + u[1] * q[1] + u[2] * q[2];
        for (int r = 0; r < 3; r++) e[r] = q[r] - u[r] * uq;
        for (int r = 0; r < 3; r++)
        {
            double B[3][6];
            for (int c = 0; c < 3; c++)
            {
                const double R0 = pose.R(c, 0), R1 = pose.R(c, 1), R2 = pose.R(c, 2);
                B[c][0] = R0, B[c][1] = R1, B[c][2] = R2;
                B[c][3] = R2 * l[1] - R1 * l[2], B[c][4] = R0 * l[2] - R2 * l[0], B[c][5] = R1 * l[0] - R0 * l[1];
            }
            for (int a = 0; a < 6; a++)
            {
                double v = 0;
                for (int c = 0; c < 3; c++) v += ((r == c ? 1.0 : 0.0) - u[r] * u[c]) * B[c][a];
                J[6 * r + a] = v;
            }
        }
    }

    void orc_error_and_jacobian(int kind, const orc_pair_pt2pt* pp, const orc_pair_pt2pl* pl,
                                const double T[12], double e[3], double J[18])
    {
        const Pose   pose = to_pose(T);
        const double l[3] = {kind == 0 ? pp->lx : pl->lx, kind == 0 ? pp->ly : pl->ly,
                             kind == 0 ? pp->lz : pl->lz};
        double       gx, gy, gz;
        compose_point(pose, l[0], l[1], l[2], gx, gy, gz);
        double B[3][6];
        for (int r = 0; r < 3; r++)
        {
            const double R0 = pose.R(r, 0), R1 = pose.R(r, 1), R2 = pose.R(r, 2);
            B[r][0] = R0, B[r][1] = R1, B[r][2] = R2;
            B[r][3] = R2 * l[1] - R1 * l[2];
            B[r][4] = R0 * l[2] - R2 * l[0];
            B[r][5] = R1 * l[0] - R0 * l[1];
        }
        if (kind == 0)
        {
            e[0] = gx - pp->gx, e[1] = gy - pp->gy, e[2] = gz - pp->gz;
            for (int r = 0; r < 3; r++)
                for (int a = 0; a < 6; a++) J[6 * r + a] = B[r][a];
        }
        else
        {
            const double* c     = pl->coefs;
            const double  mod_n = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
            const double  ev    = c[0] * gx + c[1] * gy + c[2] * gz + c[3];
            for (int r = 0; r < 3; r++) e[r] = -(c[r] / mod_n) * ev;
            for (int r = 0; r < 3; r++)
                for (int a = 0; a < 6; a++)
                    J[6 * r + a] = -(c[r] * c[0] * B[0][a] + c[r] * c[1] * B[1][a] + c[r] * c[2] * B[2][a]) / mod_n;
        }
    }

    // ---------------------------------------------------------------- voxel decimation (oracle first)
    // FilterDecimateVoxels::filter (mp2p_icp_filters/src/FilterDecimateVoxels.cpp:109-378) over ONE input layer:
    // voxel index per axis = int32(coordinate / resolution) — TRUNCATION toward zero, float division
    // (PointCloudToVoxelGridSingle.h:105, PointCloudToVoxelGrid.h:118) — and per occupied voxel
    //   method 0 FirstPoint        the first point that fell into it (PointCloudToVoxelGridSingle.cpp:60-95)
    //   method 1 ClosestToAverage  the member closest to the voxel mean, first on ties (strict <, :279-298)
    //   method 2 VoxelAverage      the mean itself: float sums, times float(1/n) (:265-277, :300-304)
    // (RandomPoint draws from an unseeded mrpt::random generator, :247-248: not restatable.)
    // flatten_to (optional): z is replaced and only the FIRST voxel visited of every (cx, cy) column emits (:210-224,
    // :335-349). Visiting order: the reference walks a tsl::robin_map (implementation-defined order) unless
    // use_tsl_robin_map = false, where FirstPoint walks a std::map ordered by (cx, cy, cz) (PointCloudToVoxelGridSingle.h:87-99);
    // this restatement always emits in that ascending (cx, cy, cz) order — the SET of output points is the
    // reference's in every configuration, the ORDER only in that one. Returns the number of output points;
    // out_src[i] = index of the source point, or -1 for an averaged point.
    size_t orc_decimate_voxels(const float* x, const float* y, const float* z, size_t n, float resolution, int method,
                               int has_flatten, float flatten_to, float* ox, float* oy, float* oz, int64_t* out_src,
                               size_t cap)
    {
        struct Key
        {
            int32_t cx, cy, cz;
            bool    operator<(const Key& o) const { return cx != o.cx ? cx < o.cx : (cy != o.cy ? cy < o.cy : cz < o.cz); }
        };
        std::map<Key, std::vector<size_t>> vox;
        for (size_t i = 0; i < n; i++)
        {
            const Key k{static_cast<int32_t>(x[i] / resolution), static_cast<int32_t>(y[i] / resolution),
                        static_cast<int32_t>(z[i] / resolution)};
            vox[k].push_back(i);
        }
        std::map<std::pair<int32_t, int32_t>, bool> usedColumn;
        size_t                                       nOut = 0;
        for (const auto& [k, idx] : vox)
        {
            float   px = 0, py = 0, pz = 0;
            int64_t src = -1;
            if (method == 0)
                src = static_cast<int64_t>(idx[0]);
            else
            {
                float       mx = 0, my = 0, mz = 0;
                const float inv_n = 1.0f / idx.size();
                for (size_t i : idx) mx += x[i], my += y[i], mz += z[i];
                mx *= inv_n, my *= inv_n, mz *= inv_n;
                if (method == 1)
                {
                    bool  have = false;
                    float best = 0;
                    for (size_t i : idx)
                    {
                        const float e = (x[i] - mx) * (x[i] - mx) + (y[i] - my) * (y[i] - my) + (z[i] - mz) * (z[i] - mz);
                        if (!have || e < best) have = true, best = e, src = static_cast<int64_t>(i);
                    }
                }
                else
                    px = mx, py = my, pz = mz;
            }
            if (src >= 0) px = x[src], py = y[src], pz = z[src];
            if (has_flatten)
            {
                auto& used = usedColumn[{k.cx, k.cy}];
                if (used) continue;
                used = true;
                pz   = flatten_to;
            }
            if (nOut < cap) ox[nOut] = px, oy[nOut] = py, oz[nOut] = pz, out_src[nOut] = src;
            nOut++;
        }
        return nOut;
    }

    // ---------------------------------------------------------------- covariance() (oracle first)
    // mp2p_icp::covariance (mp2p_icp/src/covariance.cpp:28-141): the stacked error vector of the final pairings
    // (pt2pt :75-82, pt2ln :85-93, pt2pl :107-115; 3 rows per pairing, error_point2point / error_point2line /
    // error_point2plane of errorTerms.cpp) is differentiated numerically w.r.t. (x, y, z, yaw, pitch, roll) by
    // mrpt::math::estimateJacobian — central differences, column i = (f(x + h_i e_i) - f(x - h_i e_i)) * (0.5 / h_i),
    // recalled from MRPT 2.x num_jacobian.h (not in the reference tree: parity unpinned for that helper) —
    // hessian = J^T J (:134), cov = hessian.inverse_LLt() (:136). `x6` is the vector the Jacobian is taken at.
    // NOTE (as written upstream, :41-47): the reference fills xInitial[0], [1], [0] again, [3], [4], [5] — slot 2 (z)
    // is never assigned; a default-constructed CMatrixDouble61 is zero-filled, so the reference evaluates at z = 0.
    // Callers that want the reference's number pass x6[2] = 0; this function takes the vector as given.
    // Empty pairings: diag(1e6) (:33-38). ln2ln / pl2pl terms are not restated (they stay host-side upstream).
    void orc_covariance(const orc_pair_pt2pt* p2p, size_t n2p, const orc_pair_pt2pl* p2l, size_t n2l,
                        const orc_pair_pt2ln* p2ln, size_t n2ln, const double x6[6], double finDif_xyz,
                        double finDif_angles, double cov_out[36], double hessian_out[36])
    {
        for (int k = 0; k < 36; k++) cov_out[k] = 0, hessian_out[k] = 0;
        if (n2p + n2l + n2ln == 0)
        {
            for (int k = 0; k < 6; k++) cov_out[7 * k] = 1e6;
            return;
        }
        const size_t        rows = 3 * (n2p + n2ln + n2l);
        std::vector<double> fp(rows), fm(rows), J(rows * 6);
        auto                eval = [&](const double x[6], std::vector<double>& err)
        {
            const Pose pose = pose_from_xyzypr(x[0], x[1], x[2], x[3], x[4], x[5]);  // CPose3D::setFromValues
            size_t     r    = 0;
            for (size_t i = 0; i < n2p; i++, r += 3)  // error_point2point, errorTerms.cpp:36-66
            {
                double gx, gy, gz;
                compose_point(pose, p2p[i].lx, p2p[i].ly, p2p[i].lz, gx, gy, gz);
                err[r] = gx - p2p[i].gx, err[r + 1] = gy - p2p[i].gy, err[r + 2] = gz - p2p[i].gz;
            }
            for (size_t i = 0; i < n2ln; i++, r += 3)  // error_point2line, errorTerms.cpp:67-113
            {
                double gx, gy, gz;
                compose_point(pose, p2ln[i].lx, p2ln[i].ly, p2ln[i].lz, gx, gy, gz);
                const double* u    = p2ln[i].director;
                const double  q[3] = {gx - p2ln[i].pBase[0], gy - p2ln[i].pBase[1], gz - p2ln[i].pBase[2]};
                const double  uq   = u[0] * q[0] + u[1] * q[1] + u[2] * q[2];
                for (int c = 0; c < 3; c++) err[r + c] = q[c] - u[c] * uq;
            }
            for (size_t i = 0; i < n2l; i++, r += 3)  // error_point2plane, errorTerms.cpp:115-161
            {
                double gx, gy, gz;
                compose_point(pose, p2l[i].lx, p2l[i].ly, p2l[i].lz, gx, gy, gz);
                const double* c     = p2l[i].coefs;
                const double  mod_n = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
                const double  ev    = c[0] * gx + c[1] * gy + c[2] * gz + c[3];
                for (int k = 0; k < 3; k++) err[r + k] = -(c[k] / mod_n) * ev;
            }
        };
        for (int i = 0; i < 6; i++)
        {
            const double h = i < 3 ? finDif_xyz : finDif_angles;
            double       xp[6], xm[6];
            for (int k = 0; k < 6; k++) xp[k] = xm[k] = x6[k];
            xp[i] = x6[i] + h, xm[i] = x6[i] - h;
            eval(xp, fp), eval(xm, fm);
            const double inv2h = 0.5 / h;
            for (size_t r = 0; r < rows; r++) J[r * 6 + i] = inv2h * (fp[r] - fm[r]);
        }
        double H[36] = {0};
        for (size_t r = 0; r < rows; r++)
            for (int a = 0; a < 6; a++)
                for (int b = 0; b < 6; b++) H[6 * a + b] += J[r * 6 + a] * J[r * 6 + b];
        for (int k = 0; k < 36; k++) hessian_out[k] = H[k];
        // inverse_LLt: H = L L^T (Cholesky), cov = L^-T L^-1
        double Lm[36] = {0};
        for (int i = 0; i < 6; i++)
            for (int j = 0; j <= i; j++)
            {
                double s = H[6 * i + j];
                for (int k = 0; k < j; k++) s -= Lm[6 * i + k] * Lm[6 * j + k];
                Lm[6 * i + j] = i == j ? std::sqrt(s) : s / Lm[6 * j + j];
            }
        double Li[36] = {0};  // L^-1 (lower)
        for (int c = 0; c < 6; c++)
            for (int i = c; i < 6; i++)
            {
                double s = i == c ? 1.0 : 0.0;
                for (int k = c; k < i; k++) s -= Lm[6 * i + k] * Li[6 * k + c];
                Li[6 * i + c] = s / Lm[6 * i + i];
            }
        for (int a = 0; a < 6; a++)
            for (int b = 0; b < 6; b++)
            {
                double s = 0;
                for (int k = 0; k < 6; k++) s += Li[6 * k + a] * Li[6 * k + b];
                cov_out[6 * a + b] = s;
            }
    }

    int orc_max_threads() { return omp_get_max_threads(); }
}
