// Reference-side binding: the mp2p_icp plugin classes that put the B200 hot path under the
// reference's own Matcher / Solver interfaces. BUILT only where MRPT and mp2p_icp headers exist
// (`make -C mp2p_icp_b200/host plugin MRPT=1`); this container has neither. Here the
// file is TYPE-CHECKED against shape stubs of every MRPT / mp2p_icp declaration it touches
// (tests/stubs/, `make -C mp2p_icp_b200/host shape`, run by __graft_entry__.build() and
// tests/test_plugin_shape.py): the virtual signatures, member names and record layouts it relies on
// are the ones cited there from the reference headers.
//
// Safety defaults (VERDICT r1 / ADVICE r1):
//  * the solvers UPLOAD the pairings they are given. Reading the device copy the last matcher call
//    left (MP2P_B200_PAIRS_LAST_MATCH) is an explicit opt-in, YAML `assumeUnmodifiedPairings: true`
//    on the solver, for pipelines where nothing edits the Pairings between run_matchers and
//    run_solvers (ICP.cpp:143-170) — then the count and 8 sampled records are still compared.
//  * a global layer is re-indexed (and a local layer re-uploaded) when its buffer address, size OR a
//    fingerprint of ~4096 sampled points + the first / last point changes (mp2p_b200_map_cached /
//    mp2p_b200_cloud_cached). MRPT exposes no modification counter for the point buffers, so an
//    in-place edit that touches none of the samples needs mp2p_icp::B200InvalidateLayer(layer) —
//    documented in INTEGRATION.md. YAML `cacheLocalCloud: false` uploads the local layer every call.
//  * the solvers' uploaded pairings are compared on the device, byte for byte, with the copy the last
//    matcher call left there; only on equality is the result that call computed ahead of time used.
//  * YAML `device: <n>` on every class selects the GPU (default 0); one context per device.
//
// Usage from a pipeline YAML (the reference loads the .so through its `plugin:` key,
// mp2p_icp_map/src/load_plugin.cpp:70-134, then creates the class by name, ICP.cpp:507-516):
//
//   matchers:
//     - class: mp2p_icp::Matcher_Points_DistanceThreshold_B200   (or Matcher_Points_InlierRatio_B200)
//       plugin: libmp2p_icp_b200_plugin.so
//       params: { threshold: 1.0, thresholdAngularDeg: 0, pairingsPerPoint: 1 }
//   solvers:
//     - class: mp2p_icp::Solver_Horn_B200
//       plugin: libmp2p_icp_b200_plugin.so
#if defined(MP2P_B200_WITH_MRPT)

#include <mp2p_icp/Matcher_Points_Base.h>
#include <mp2p_icp/QualityEvaluator.h>
#include <mp2p_icp/Solver.h>
#include <mp2p_icp/Solver_GaussNewton.h>
#include <mp2p_icp/Solver_Horn.h>
#include <mp2p_icp/metricmap.h>
#include <mp2p_icp_filters/FilterDecimateVoxels.h>
#include <mp2p_icp_filters/GetOrCreatePointLayer.h>
#include <mrpt/core/initializer.h>
#include <mrpt/maps/CPointsMap.h>
#include <mrpt/math/distributions.h>
#include <mrpt/math/utils.h>
#include <mrpt/rtti/CObject.h>

#include <algorithm>
#include <cstring>
#include <map>
#include <mutex>
#include <set>

#include "mp2p_b200.h"

namespace mp2p_icp
{
namespace b200_detail
{
inline void check(int rc)
{
    if (rc != MP2P_B200_OK) THROW_EXCEPTION_FMT("mp2p_b200: %s", mp2p_b200_last_error());
}
// one context per device, created on first use (YAML `device:` of the classes below)
inline mp2p_b200_ctx* ctx(int device)
{
    static std::mutex                     mtx;
    static std::map<int, mp2p_b200_ctx*>  all;
    std::lock_guard<std::mutex>           lk(mtx);
    auto                                  it = all.find(device);
    if (it != all.end()) return it->second;
    mp2p_b200_ctx* p = nullptr;
    check(mp2p_b200_ctx_create(device, nullptr, &p));
    all[device] = p;
    return p;
}
inline void pose12(const mrpt::poses::CPose3D& p, double out[12])
{
    const auto& R = p.getRotationMatrix();
    for (int r = 0; r < 3; r++)
    {
        for (int c = 0; c < 3; c++) out[4 * r + c] = R(r, c);
        out[4 * r + 3] = p.m_coords[r];
    }
}
// Device copies of the layers: kept by the LIBRARY (mp2p_b200_map_cached / mp2p_b200_cloud_cached), keyed on
// the layer's x-buffer address, its size and a fingerprint of ~4096 sampled points — the reference relies on
// MRPT's own "kd-tree up to date" flag, which is not visible from outside, and the methods here are const
// (SURVEY.md §8b "Ownership"). A global layer is re-indexed, a local layer re-uploaded, whenever one of the
// three changes; B200InvalidateLayer() covers in-place edits that miss every sample.
inline mp2p_b200_map* device_map(const mrpt::maps::CMetricMap& layer, int device)
{
    const auto* pts = mp2p_icp::MapToPointsMap(layer);
    ASSERTMSG_(pts, "B200 matchers need a CPointsMap global layer");
    const auto&    xs = pts->getPointsBufferRef_x();
    mp2p_b200_map* m  = nullptr;
    check(mp2p_b200_map_cached(ctx(device), xs.data(), pts->getPointsBufferRef_y().data(), pts->getPointsBufferRef_z().data(),
                               xs.size(), &m, nullptr));
    return m;
}
// the local layer of an align() does not change between iterations (ICP.cpp:123-308 only moves the pose): with
// `cacheLocalCloud` (default) it crosses PCIe once and is searched in Morton order; NULL = pass the host buffers
inline const float* device_cloud(const mrpt::maps::CPointsMap& pts, int device, bool cache)
{
    if (!cache) return nullptr;
    const auto&      xs = pts.getPointsBufferRef_x();
    mp2p_b200_cloud* c  = nullptr;
    check(mp2p_b200_cloud_cached(ctx(device), xs.data(), pts.getPointsBufferRef_y().data(), pts.getPointsBufferRef_z().data(),
                                 xs.size(), &c, nullptr));
    return reinterpret_cast<const float*>(c);
}
// Witness of the pairings a matcher call just returned to the host (count + 8 sample records). A
// solver handed a list with the same count and samples takes it for that output — nothing modifies
// the Pairings between run_matchers and run_solvers (ICP.cpp:143-170) — and passes
// MP2P_B200_PAIRS_LAST_MATCH, so the library reads the copy still resident on the device instead of
// uploading the same records again.
struct Witness
{
    size_t        n = ~size_t(0);
    unsigned char sample[8][sizeof(mp2p_b200_pair_pt2pl)];
    template <class Rec>
    void note(const Rec* recs, size_t cnt)
    {
        n = cnt;
        for (size_t k = 0; k < 8 && cnt; k++) std::memcpy(sample[k], &recs[k * (cnt - 1) / 7], sizeof(Rec));
    }
    template <class Rec>
    bool same(const Rec* recs, size_t cnt) const
    {
        if (!cnt || cnt != n) return false;
        for (size_t k = 0; k < 8; k++)
            if (std::memcmp(sample[k], &recs[k * (cnt - 1) / 7], sizeof(Rec)) != 0) return false;
        return true;
    }
};
inline Witness& witness2p(int device)
{
    static std::map<int, Witness> w;
    return w[device];
}
inline Witness& witness2l(int device)
{
    static std::map<int, Witness> w;
    return w[device];
}
// MatchState bit field -> the bit words of the C ABI (bit i of word i / 32). An all-clear field — every
// first matcher of a run_matchers call sees one (Matcher.cpp:58-66 builds a fresh MatchState) — comes back
// EMPTY, and the callers then pass NULL ("none set": nothing to upload). Words are assembled in a
// register, one store per 32 points; with libstdc++ and the dense form the vector<bool> storage is
// read word-wise instead of bit by bit.
struct BitFieldMirror  // the two data members of DenseOrSparseBitField (pointcloud_bitfield.h:86-91)
{
    std::optional<std::vector<bool>> dense_;
    std::set<uint64_t>               sparse_;
};
inline std::vector<uint32_t> to_bits(const pointcloud_bitfield_t::DenseOrSparseBitField& bf, size_t n)
{
    std::vector<uint32_t> w((n + 31) / 32, 0u);
    uint32_t              any = 0;
#if defined(__GLIBCXX__) && defined(MP2P_B200_BITFIELD_FAST_PATH)
    // opt-in (-DMP2P_B200_BITFIELD_FAST_PATH): reads the private members through a layout mirror
    static_assert(sizeof(BitFieldMirror) == sizeof(pointcloud_bitfield_t::DenseOrSparseBitField));
    const auto& m = reinterpret_cast<const BitFieldMirror&>(bf);
    if (m.dense_.has_value() && m.dense_->size() >= n)
    {
        const unsigned long* src = m.dense_->begin()._M_p;  // libstdc++: _Bit_type words, bit i at word i / 64
        for (size_t k = 0; k < w.size(); k++)
        {
            const unsigned long word = src[k >> 1];
            w[k] = static_cast<uint32_t>((k & 1) ? (word >> 32) : word);
            any |= w[k];
        }
        if (n & 31) w.back() &= (1u << (n & 31)) - 1u;
        if (!any) w.clear();
        return w;
    }
#endif
    for (size_t k = 0; k < w.size(); k++)
    {
        const size_t i0 = k << 5, i1 = std::min(n, i0 + 32);
        uint32_t     word = 0;
        for (size_t i = i0; i < i1; i++) word |= static_cast<uint32_t>(bf[i]) << (i - i0);
        w[k] = word, any |= word;
    }
    if (!any) w.clear();
    return w;
}
inline const uint32_t* bits_or_null(const std::vector<uint32_t>& w) { return w.empty() ? nullptr : w.data(); }
}  // namespace b200_detail

/** Tells the plugin that a layer (global or local) was edited IN PLACE — same buffers, same size: its device copy
 *  is rebuilt on the next matcher call. Edits that change the buffer address, the size or any of the ~4096
 *  sampled points are detected without this call. */
inline void B200InvalidateLayer(const mrpt::maps::CPointsMap& layer, int device = 0)
{
    mp2p_b200_layer_invalidate(b200_detail::ctx(device), layer.getPointsBufferRef_x().data());
}

/** Drop-in for Matcher_Points_DistanceThreshold (same parameters, same results). */
class Matcher_Points_DistanceThreshold_B200 : public Matcher_Points_Base
{
    DEFINE_MRPT_OBJECT(Matcher_Points_DistanceThreshold_B200, mp2p_icp)
   public:
    int  device = 0;              //!< YAML `device`: the GPU this object works on
    bool cacheLocalCloud = true;  //!< YAML: keep the local layer on the device while its fingerprint is unchanged
    void initialize(const mrpt::containers::yaml& params) override
    {
        Matcher_Points_Base::initialize(params);
        MCP_LOAD_OPT(params, device);
        MCP_LOAD_OPT(params, cacheLocalCloud);
        DECLARE_PARAMETER_REQ(params, threshold);
        DECLARE_PARAMETER_REQ(params, thresholdAngularDeg);
        DECLARE_PARAMETER_OPT(params, pairingsPerPoint);
    }
    double   threshold = 0.5, thresholdAngularDeg = 0.0;
    uint32_t pairingsPerPoint = 1;

   private:
    void implMatchOneLayer(const mrpt::maps::CMetricMap& pcGlobal, const mrpt::maps::CPointsMap& pcLocal,
                           const mrpt::poses::CPose3D& localPose, MatchState& ms,
                           const layer_name_t& globalName, const layer_name_t& localName,
                           Pairings& out) const override
    {
        using namespace b200_detail;
        checkAllParametersAreRealized();
        ASSERT_(pairingsPerPoint >= 1);
        ASSERT_GT_(threshold, .0);
        ASSERT_GE_(thresholdAngularDeg, .0);
        mp2p_b200_map* gmap = device_map(pcGlobal, device);
        const auto& lx = pcLocal.getPointsBufferRef_x();
        double T[12];
        pose12(localPose, T);
        mp2p_b200_pt2pt_params p{threshold, thresholdAngularDeg, pairingsPerPoint,
                                 allowMatchAlreadyMatchedPoints_, allowMatchAlreadyMatchedGlobalPoints_,
                                 bounding_box_intersection_check_epsilon_};
        auto&      lbf   = ms.localPairedBitField.point_layers[localName];
        auto&      gbf   = ms.globalPairedBitField.point_layers[globalName];
        const auto lbits = to_bits(lbf, lx.size());
        const auto gbits = to_bits(gbf, mp2p_icp::MapToNN(pcGlobal, true)->nn_index_count());
        const size_t before = out.paired_pt2pt.size();
        out.paired_pt2pt.resize(before + lx.size() * pairingsPerPoint);
        static_assert(sizeof(mrpt::tfest::TMatchingPair) == sizeof(mp2p_b200_pair_pt2pt));
        uint64_t cnt = 0, pot = 0;
        const float* resident = device_cloud(pcLocal, device, cacheLocalCloud);
        check(mp2p_b200_match_pt2pt(ctx(device), gmap, resident ? resident : lx.data(),
                                    resident ? nullptr : pcLocal.getPointsBufferRef_y().data(),
                                    resident ? nullptr : pcLocal.getPointsBufferRef_z().data(), lx.size(),
                                    resident ? MP2P_B200_LOCAL_CLOUD : MP2P_B200_LOCAL_HOST, T, &p,
                                    bits_or_null(lbits), bits_or_null(gbits),
                                    reinterpret_cast<mp2p_b200_pair_pt2pt*>(out.paired_pt2pt.data() + before),
                                    lx.size() * pairingsPerPoint, 0, &cnt, &pot));
        out.paired_pt2pt.resize(before + cnt);
        out.potential_pairings += pot;
        witness2p(device).note(out.paired_pt2pt.data() + before, before == 0 ? cnt : 0);
        if (!allowMatchAlreadyMatchedGlobalPoints_)  // lambdaAddPair, …DistanceThreshold.cpp:116-120
            for (size_t i = before; i < out.paired_pt2pt.size(); i++)
            {
                lbf.mark_as_set(out.paired_pt2pt[i].localIdx);
                gbf.mark_as_set(out.paired_pt2pt[i].globalIdx);
            }
    }
};
IMPLEMENTS_MRPT_OBJECT(Matcher_Points_DistanceThreshold_B200, Matcher, mp2p_icp)

/** Drop-in for Matcher_Points_InlierRatio (mp2p_icp/src/Matcher_Points_InlierRatio.cpp:35-143). */
class Matcher_Points_InlierRatio_B200 : public Matcher_Points_Base
{
    DEFINE_MRPT_OBJECT(Matcher_Points_InlierRatio_B200, mp2p_icp)
   public:
    int  device = 0;              //!< YAML `device`: the GPU this object works on
    bool cacheLocalCloud = true;  //!< YAML: keep the local layer on the device while its fingerprint is unchanged
    void initialize(const mrpt::containers::yaml& params) override
    {
        Matcher_Points_Base::initialize(params);
        MCP_LOAD_OPT(params, device);
        MCP_LOAD_OPT(params, cacheLocalCloud);
        MCP_LOAD_REQ(params, inliersRatio);
    }
    double inliersRatio = 0.80;

   private:
    void implMatchOneLayer(const mrpt::maps::CMetricMap& pcGlobal, const mrpt::maps::CPointsMap& pcLocal,
                           const mrpt::poses::CPose3D& localPose, MatchState& ms,
                           const layer_name_t& globalName, const layer_name_t& localName,
                           Pairings& out) const override
    {
        using namespace b200_detail;
        ASSERT_GT_(inliersRatio, 0.0);
        ASSERT_LT_(inliersRatio, 1.0);
        mp2p_b200_map* gmap = device_map(pcGlobal, device);
        const auto&    lx   = pcLocal.getPointsBufferRef_x();
        double         T[12];
        pose12(localPose, T);
        mp2p_b200_inlier_ratio_params p{inliersRatio, allowMatchAlreadyMatchedPoints_,
                                        allowMatchAlreadyMatchedGlobalPoints_,
                                        bounding_box_intersection_check_epsilon_};
        auto&        lbf    = ms.localPairedBitField.point_layers[localName];
        auto&        gbf    = ms.globalPairedBitField.point_layers[globalName];
        const auto   lbits  = to_bits(lbf, lx.size());
        const auto   gbits  = to_bits(gbf, mp2p_icp::MapToNN(pcGlobal, true)->nn_index_count());
        const size_t before = out.paired_pt2pt.size();
        out.paired_pt2pt.resize(before + lx.size());
        uint64_t     cnt = 0, pot = 0;
        const float* resident = device_cloud(pcLocal, device, cacheLocalCloud);
        check(mp2p_b200_match_inlier_ratio(ctx(device), gmap, resident ? resident : lx.data(),
                                           resident ? nullptr : pcLocal.getPointsBufferRef_y().data(),
                                           resident ? nullptr : pcLocal.getPointsBufferRef_z().data(), lx.size(),
                                           resident ? MP2P_B200_LOCAL_CLOUD : MP2P_B200_LOCAL_HOST, T, &p,
                                           bits_or_null(lbits), bits_or_null(gbits),
                                           reinterpret_cast<mp2p_b200_pair_pt2pt*>(out.paired_pt2pt.data() + before),
                                           lx.size(), 0, &cnt, &pot));
        out.paired_pt2pt.resize(before + cnt);
        out.potential_pairings += pot;
        witness2p(device).note(out.paired_pt2pt.data() + before, before == 0 ? cnt : 0);
        for (size_t i = before; i < out.paired_pt2pt.size(); i++)  // :133-135
        {
            lbf.mark_as_set(out.paired_pt2pt[i].localIdx);
            gbf.mark_as_set(out.paired_pt2pt[i].globalIdx);
        }
    }
};
IMPLEMENTS_MRPT_OBJECT(Matcher_Points_InlierRatio_B200, Matcher, mp2p_icp)

/** Drop-in for Solver_Horn: pt2pt accumulation on the GPU. pt2ln / pt2pl pairings are first
 *  projected by the reference's own pt2ln_pl_to_pt2pt (host). */
class Solver_Horn_B200 : public Solver_Horn
{
    DEFINE_MRPT_OBJECT(Solver_Horn_B200, mp2p_icp)
   public:
    int  device = 0;                         //!< YAML `device`: the GPU this object works on
    bool assumeUnmodifiedPairings = false;   //!< YAML opt-in: read the device copy the last matcher call left
    void initialize(const mrpt::containers::yaml& params) override
    {
        Solver_Horn::initialize(params);
        MCP_LOAD_OPT(params, device);
        MCP_LOAD_OPT(params, assumeUnmodifiedPairings);
    }

   protected:
    bool impl_optimal_pose(const Pairings& pairings, OptimalTF_Result& out, const SolverContext& sc) const override
    {
        using namespace b200_detail;
        if (!pairings.paired_pt2ln.empty() || !pairings.paired_pt2pl.empty() ||
            !pairings.paired_ln2ln.empty() || !pairings.paired_pl2pl.empty())
            return Solver_Horn::impl_optimal_pose(pairings, out, sc);  // terms that stay host-side
        out = OptimalTF_Result();
        const auto&           wp = pairingsWeightParameters;
        mp2p_b200_horn_params p{};
        p.use_scale_outlier_detector = wp.use_scale_outlier_detector;
        p.scale_outlier_threshold    = wp.scale_outlier_threshold;
        p.w_pt2pt                    = wp.pair_weights.pt2pt;
        p.robust_kernel              = static_cast<int>(wp.robust_kernel);
        p.robust_kernel_param        = wp.robust_kernel_param;
        if (wp.currentEstimateForRobust) pose12(*wp.currentEstimateForRobust, p.currentEstimateForRobust);
        std::vector<uint64_t> wc;
        std::vector<double>   wv;
        for (const auto& [cnt, w] : pairings.point_weights) wc.push_back(cnt), wv.push_back(w);
        double  T[12];
        int32_t solved = 0;
        check(mp2p_b200_solve_horn(ctx(device), reinterpret_cast<const mp2p_b200_pair_pt2pt*>(pairings.paired_pt2pt.data()),
                                   pairings.paired_pt2pt.size(),
                                   (assumeUnmodifiedPairings && witness2p(device).same(pairings.paired_pt2pt.data(), pairings.paired_pt2pt.size()))
                                       ? MP2P_B200_PAIRS_LAST_MATCH
                                       : 0,
                                   &p, wc.data(), wv.data(), wc.size(), T, &solved));
        if (!solved) return false;
        mrpt::math::CMatrixDouble44 M = mrpt::math::CMatrixDouble44::Identity();
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 4; c++) M(r, c) = T[4 * r + c];
        out.optimalPose = mrpt::poses::CPose3D(M);
        return true;
    }
};
IMPLEMENTS_MRPT_OBJECT(Solver_Horn_B200, Solver, mp2p_icp)

/** Drop-in for Matcher_Point2Plane over a plain point layer: NearestPlaneCapable::nn_search_pt2pl is
 *  realised on the GPU as k-NN + PCA plane (SURVEY.md §8a-7'; the reference ships no implementer of
 *  that interface, tests/CMakeLists.txt:37). `distanceThreshold` as in the reference
 *  (Matcher_Point2Plane.cpp:36-39); the plane-detection parameters are those of Matcher_Adaptive
 *  (Matcher_Adaptive.cpp:229-253) with the same names and defaults. */
class Matcher_Point2Plane_B200 : public Matcher_Points_Base
{
    DEFINE_MRPT_OBJECT(Matcher_Point2Plane_B200, mp2p_icp)
   public:
    int  device = 0;              //!< YAML `device`: the GPU this object works on
    bool cacheLocalCloud = true;  //!< YAML: keep the local layer on the device while its fingerprint is unchanged
    void initialize(const mrpt::containers::yaml& params) override
    {
        Matcher_Points_Base::initialize(params);
        MCP_LOAD_OPT(params, device);
        MCP_LOAD_OPT(params, cacheLocalCloud);
        DECLARE_PARAMETER_REQ(params, distanceThreshold);
        DECLARE_PARAMETER_OPT(params, searchRadius);
        DECLARE_PARAMETER_OPT(params, knn);
        DECLARE_PARAMETER_OPT(params, minimumPlanePoints);
        DECLARE_PARAMETER_OPT(params, planeEigenThreshold);
    }
    double   distanceThreshold = 0.50, searchRadius = 1.0, planeEigenThreshold = 0.01;
    uint32_t knn = 8, minimumPlanePoints = 5;

   private:
    void implMatchOneLayer(const mrpt::maps::CMetricMap& pcGlobal, const mrpt::maps::CPointsMap& pcLocal,
                           const mrpt::poses::CPose3D& localPose, MatchState& ms,
                           const layer_name_t& globalName, const layer_name_t& localName,
                           Pairings& out) const override
    {
        using namespace b200_detail;
        (void)globalName;
        checkAllParametersAreRealized();
        mp2p_b200_map* gmap = device_map(pcGlobal, device);
        const auto&    lx   = pcLocal.getPointsBufferRef_x();
        double         T[12];
        pose12(localPose, T);
        mp2p_b200_pt2pl_params p{distanceThreshold, searchRadius, knn, minimumPlanePoints, planeEigenThreshold,
                                 allowMatchAlreadyMatchedPoints_, bounding_box_intersection_check_epsilon_};
        auto&                  lbf   = ms.localPairedBitField.point_layers.at(localName);
        const auto             lbits = to_bits(lbf, lx.size());
        const size_t           before = out.paired_pt2pl.size();
        out.paired_pt2pl.resize(before + lx.size());
        static_assert(sizeof(mp2p_icp::point_plane_pair_t) == sizeof(mp2p_b200_pair_pt2pl));
        uint64_t     cnt = 0, pot = 0;
        const float* resident = device_cloud(pcLocal, device, cacheLocalCloud);
        check(mp2p_b200_match_pt2pl(ctx(device), gmap, resident ? resident : lx.data(),
                                    resident ? nullptr : pcLocal.getPointsBufferRef_y().data(),
                                    resident ? nullptr : pcLocal.getPointsBufferRef_z().data(), lx.size(),
                                    resident ? MP2P_B200_LOCAL_CLOUD : MP2P_B200_LOCAL_HOST, T, &p, bits_or_null(lbits),
                                    reinterpret_cast<mp2p_b200_pair_pt2pl*>(out.paired_pt2pl.data() + before),
                                    lx.size(), 0, &cnt, &pot));
        out.paired_pt2pl.resize(before + cnt);
        out.potential_pairings += pot;
        witness2l(device).note(out.paired_pt2pl.data() + before, before == 0 ? cnt : 0);
        // Matcher_Point2Plane.cpp:105-109: only the local point is marked. point_plane_pair_t carries
        // the local COORDINATES, not the index: re-identify by a parallel walk (the output is in
        // ascending local index and each local point pairs at most once)
        const auto& ly = pcLocal.getPointsBufferRef_y();
        const auto& lz = pcLocal.getPointsBufferRef_z();
        size_t      i  = 0;
        for (size_t k = before; k < out.paired_pt2pl.size(); k++)
        {
            const auto& pl = out.paired_pt2pl[k].pt_local;
            while (i < lx.size() && !(lx[i] == pl.x && ly[i] == pl.y && lz[i] == pl.z && !lbf[i])) i++;
            if (i < lx.size()) lbf.mark_as_set(i++);
        }
    }
};
IMPLEMENTS_MRPT_OBJECT(Matcher_Point2Plane_B200, Matcher, mp2p_icp)

/** Drop-in for Solver_GaussNewton: pt2pt and pt2pl terms accumulated on the GPU (whole inner loop on
 *  the device); any other term or a prior sends the call to the reference implementation. */
class Solver_GaussNewton_B200 : public Solver_GaussNewton
{
    DEFINE_MRPT_OBJECT(Solver_GaussNewton_B200, mp2p_icp)
   public:
    int  device = 0;                         //!< YAML `device`: the GPU this object works on
    bool assumeUnmodifiedPairings = false;   //!< YAML opt-in: read the device copy the last matcher call left
    void initialize(const mrpt::containers::yaml& params) override
    {
        Solver_GaussNewton::initialize(params);
        MCP_LOAD_OPT(params, device);
        MCP_LOAD_OPT(params, assumeUnmodifiedPairings);
    }

   protected:
    bool impl_optimal_pose(const Pairings& pairings, OptimalTF_Result& out, const SolverContext& sc) const override
    {
        using namespace b200_detail;
        // Pairings::point_weights re-weight the pt2pt term block by block (optimal_tf_gauss_newton.cpp:118-128):
        // the device accumulation carries one uniform pt2pt weight, so weighted layers go to the reference
        if (sc.prior.has_value() || !pairings.paired_ln2ln.empty() || !pairings.paired_pl2pl.empty() ||
            !pairings.point_weights.empty())
            return Solver_GaussNewton::impl_optimal_pose(pairings, out, sc);  // terms that stay host-side
        checkAllParametersAreRealized();
        out = OptimalTF_Result();
        ASSERT_(sc.guessRelativePose.has_value());
        mp2p_b200_gn_params p{};
        p.maxInnerLoopIterations = maxIterations;
        p.minDelta               = 1e-7;  // OptimalTF_GN_Parameters defaults (optimal_tf_gauss_newton.h:46-50)
        p.maxCost                = 0;
        p.w_pt2pt                = pairWeights.pt2pt;
        p.w_pt2pl                = pairWeights.pt2pl;
        p.kernel                 = static_cast<int>(robustKernel);
        p.kernelParam            = robustKernelParam;
        double T0[12], T[12];
        pose12(mrpt::poses::CPose3D(sc.guessRelativePose.value()), T0);
        int32_t  solved = 0;
        uint32_t iters  = 0;
        // every non-empty list must be the witnessed output of the last matcher call of its kind
        const auto& l2p  = pairings.paired_pt2pt;
        const auto& l2l  = pairings.paired_pt2pl;
        const bool  last = assumeUnmodifiedPairings && (l2p.empty() || witness2p(device).same(l2p.data(), l2p.size())) &&
                          (l2l.empty() || witness2l(device).same(l2l.data(), l2l.size())) && !(l2p.empty() && l2l.empty());
        if (const auto& l2n = pairings.paired_pt2ln; !l2n.empty())
        {
            // point-to-line term on the device too (optimal_tf_gauss_newton.cpp:182-203); point_line_pair_t
            // = TLine3D {pBase, director} + TPoint3D pt_local = nine doubles, the layout of mp2p_b200_pair_pt2ln
            static_assert(sizeof(mp2p_icp::point_line_pair_t) == sizeof(mp2p_b200_pair_pt2ln));
            check(mp2p_b200_solve_gauss_newton_ex(ctx(device), reinterpret_cast<const mp2p_b200_pair_pt2pt*>(l2p.data()), l2p.size(),
                                                  reinterpret_cast<const mp2p_b200_pair_pt2pl*>(l2l.data()), l2l.size(),
                                                  reinterpret_cast<const mp2p_b200_pair_pt2ln*>(l2n.data()), l2n.size(), 0, &p,
                                                  pairWeights.pt2ln, T0, T, &iters, &solved));
        }
        else
            check(mp2p_b200_solve_gauss_newton(ctx(device), reinterpret_cast<const mp2p_b200_pair_pt2pt*>(l2p.data()), l2p.size(),
                                           reinterpret_cast<const mp2p_b200_pair_pt2pl*>(l2l.data()), l2l.size(),
                                           last ? MP2P_B200_PAIRS_LAST_MATCH : 0, &p, T0, T, &iters, &solved));
        if (!solved) return false;
        mrpt::math::CMatrixDouble44 M = mrpt::math::CMatrixDouble44::Identity();
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 4; c++) M(r, c) = T[4 * r + c];
        out.optimalPose = mrpt::poses::CPose3D(M);
        return true;
    }
};
IMPLEMENTS_MRPT_OBJECT(Solver_GaussNewton_B200, Solver, mp2p_icp)

/** Drop-in for Matcher_Adaptive (mp2p_icp/src/Matcher_Adaptive.cpp:32-314). The neighbour search, the
 *  histogram binning and the per-point decisions run on the device; the ONE step in between — histogram
 *  -> upper confidence bound — is done here by MRPT itself (mrpt::math::confidenceIntervalsFromHistogram,
 *  :196-199), on the 50 numbers the device hands over, so the threshold is the reference's by construction. */
class Matcher_Adaptive_B200 : public Matcher_Points_Base
{
    DEFINE_MRPT_OBJECT(Matcher_Adaptive_B200, mp2p_icp)
   public:
    int  device = 0;              //!< YAML `device`: the GPU this object works on
    bool cacheLocalCloud = true;  //!< YAML: keep the local layer on the device while its fingerprint is unchanged
    void initialize(const mrpt::containers::yaml& params) override
    {
        Matcher_Points_Base::initialize(params);
        MCP_LOAD_OPT(params, device);
        MCP_LOAD_OPT(params, cacheLocalCloud);
        MCP_LOAD_REQ(params, confidenceInterval);
        MCP_LOAD_REQ(params, firstToSecondDistanceMax);
        MCP_LOAD_REQ(params, absoluteMaxSearchDistance);
        MCP_LOAD_OPT(params, minimumCorrDist);
        MCP_LOAD_REQ(params, enableDetectPlanes);
        MCP_LOAD_OPT(params, planeSearchPoints);
        MCP_LOAD_OPT(params, planeMinimumFoundPoints);
        MCP_LOAD_OPT(params, planeEigenThreshold);
        MCP_LOAD_OPT(params, maxPt2PtCorrespondences);
        MCP_LOAD_OPT(params, planeMinimumDistance);
        ASSERT_LT_(confidenceInterval, 1.0);
        ASSERT_GT_(confidenceInterval, 0.0);
        ASSERT_GE_(planeSearchPoints, planeMinimumFoundPoints);
        ASSERT_GE_(planeMinimumFoundPoints, 3);
        ASSERT_GT_(planeEigenThreshold, 0.0);
    }
    double   confidenceInterval = 0.80, firstToSecondDistanceMax = 1.2, absoluteMaxSearchDistance = 5.0;
    bool     enableDetectPlanes      = false;
    uint32_t maxPt2PtCorrespondences = 1, planeSearchPoints = 8, planeMinimumFoundPoints = 4;
    double   planeMinimumDistance = 0.10, planeEigenThreshold = 0.01, minimumCorrDist = 0.1;

   private:
    void implMatchOneLayer(const mrpt::maps::CMetricMap& pcGlobal, const mrpt::maps::CPointsMap& pcLocal,
                           const mrpt::poses::CPose3D& localPose, MatchState& ms, const layer_name_t& globalName,
                           const layer_name_t& localName, Pairings& out) const override
    {
        using namespace b200_detail;
        mp2p_b200_map* gmap = device_map(pcGlobal, device);
        const auto&    lx   = pcLocal.getPointsBufferRef_x();
        const auto&    ly   = pcLocal.getPointsBufferRef_y();
        const auto&    lz   = pcLocal.getPointsBufferRef_z();
        double         T[12];
        pose12(localPose, T);
        mp2p_b200_adaptive_params p{confidenceInterval, firstToSecondDistanceMax, absoluteMaxSearchDistance, minimumCorrDist,
                                    enableDetectPlanes, planeSearchPoints, planeMinimumFoundPoints, maxPt2PtCorrespondences,
                                    planeEigenThreshold, planeMinimumDistance, allowMatchAlreadyMatchedPoints_,
                                    allowMatchAlreadyMatchedGlobalPoints_, bounding_box_intersection_check_epsilon_};
        auto&      lbf   = ms.localPairedBitField.point_layers[localName];
        auto&      gbf   = ms.globalPairedBitField.point_layers[globalName];
        const auto lbits = to_bits(lbf, lx.size());
        const auto gbits = to_bits(gbf, mp2p_icp::MapToNN(pcGlobal, true)->nn_index_count());
        uint64_t   hist[MP2P_B200_ADAPTIVE_BINS], ns = 0, pot = 0;
        double     emin = 0, emax = 0;
        int32_t    gate = 0;
        const float* resident = device_cloud(pcLocal, device, cacheLocalCloud);
        check(mp2p_b200_adaptive_search(ctx(device), gmap, resident ? resident : lx.data(), resident ? nullptr : ly.data(),
                                        resident ? nullptr : lz.data(), lx.size(),
                                        resident ? MP2P_B200_LOCAL_CLOUD : MP2P_B200_LOCAL_HOST, T, &p, bits_or_null(lbits), hist, &emin,
                                        &emax, &ns, &gate, &pot));
        out.potential_pairings += pot;
        if (!gate) return;  // :71, :77-80
        // :188-199 — the histogram of the 1st / 2nd neighbour errors in MRPT's own normalisation, then MRPT's
        // confidence interval (CHistogram(min, max, 50): binSizeInv = (nBins - 1) / (max - min))
        ASSERT_(ns > 0);
        ASSERT_(emax > emin);
        std::vector<double> histXs, histValues(MP2P_B200_ADAPTIVE_BINS);
        mrpt::math::linspace(emin, emax, size_t(MP2P_B200_ADAPTIVE_BINS), histXs);
        const double K = ((MP2P_B200_ADAPTIVE_BINS - 1) / (emax - emin)) / double(ns);
        for (int b = 0; b < MP2P_B200_ADAPTIVE_BINS; b++) histValues[b] = K * double(hist[b]);
        double ci_low = 0, ci_high = 0;
        mrpt::math::confidenceIntervalsFromHistogram(histXs, histValues, ci_low, ci_high, 1.0 - confidenceInterval);
        const double maxCorrDistSqr = std::max(mrpt::square(minimumCorrDist), ci_high);  // :214

        const size_t b2p = out.paired_pt2pt.size(), b2l = out.paired_pt2pl.size();
        const size_t cap2p = lx.size() * maxPt2PtCorrespondences, cap2l = lx.size();
        out.paired_pt2pt.resize(b2p + cap2p), out.paired_pt2pl.resize(b2l + cap2l);
        uint64_t n2p = 0, n2l = 0;
        check(mp2p_b200_adaptive_emit(ctx(device), gmap, &p, maxCorrDistSqr, bits_or_null(gbits),
                                      reinterpret_cast<mp2p_b200_pair_pt2pt*>(out.paired_pt2pt.data() + b2p), cap2p,
                                      reinterpret_cast<mp2p_b200_pair_pt2pl*>(out.paired_pt2pl.data() + b2l), cap2l, 0, &n2p, &n2l));
        out.paired_pt2pt.resize(b2p + n2p), out.paired_pt2pl.resize(b2l + n2l);
        if (!allowMatchAlreadyMatchedGlobalPoints_)  // :291-295
            for (size_t i = b2p; i < out.paired_pt2pt.size(); i++) lbf.mark_as_set(out.paired_pt2pt[i].localIdx);
        size_t i = 0;  // :262 (output in ascending local index: parallel walk)
        for (size_t k = b2l; k < out.paired_pt2pl.size(); k++)
        {
            const auto& r = out.paired_pt2pl[k].pt_local;
            while (i < lx.size() && !(lx[i] == r.x && ly[i] == r.y && lz[i] == r.z && !lbf[i])) i++;
            if (i < lx.size()) lbf.mark_as_set(i++);
        }
    }
};
IMPLEMENTS_MRPT_OBJECT(Matcher_Adaptive_B200, Matcher, mp2p_icp)

/** Drop-in for Matcher_Point2Line (mp2p_icp/src/Matcher_Point2Line.cpp:35-163). */
class Matcher_Point2Line_B200 : public Matcher_Points_Base
{
    DEFINE_MRPT_OBJECT(Matcher_Point2Line_B200, mp2p_icp)
   public:
    int  device = 0;              //!< YAML `device`: the GPU this object works on
    bool cacheLocalCloud = true;  //!< YAML: keep the local layer on the device while its fingerprint is unchanged
    void initialize(const mrpt::containers::yaml& params) override
    {
        Matcher_Points_Base::initialize(params);
        MCP_LOAD_OPT(params, device);
        MCP_LOAD_OPT(params, cacheLocalCloud);
        MCP_LOAD_REQ(params, distanceThreshold);
        MCP_LOAD_REQ(params, knn);
        MCP_LOAD_REQ(params, lineEigenThreshold);
        MCP_LOAD_REQ(params, minimumLinePoints);
        ASSERT_GE_(minimumLinePoints, 2UL);
    }
    double   distanceThreshold  = 0.50;
    uint32_t knn                = 4;
    uint32_t minimumLinePoints  = 4;
    double   lineEigenThreshold = 0.01;

   private:
    void implMatchOneLayer(const mrpt::maps::CMetricMap& pcGlobal, const mrpt::maps::CPointsMap& pcLocal,
                           const mrpt::poses::CPose3D& localPose, MatchState& ms, const layer_name_t&,
                           const layer_name_t& localName, Pairings& out) const override
    {
        using namespace b200_detail;
        mp2p_b200_map* gmap = device_map(pcGlobal, device);
        const auto&    lx   = pcLocal.getPointsBufferRef_x();
        const auto&    ly   = pcLocal.getPointsBufferRef_y();
        const auto&    lz   = pcLocal.getPointsBufferRef_z();
        double         T[12];
        pose12(localPose, T);
        mp2p_b200_pt2ln_params p{distanceThreshold, knn, minimumLinePoints, lineEigenThreshold,
                                 allowMatchAlreadyMatchedPoints_, bounding_box_intersection_check_epsilon_};
        auto&        lbf    = ms.localPairedBitField.point_layers[localName];
        const auto   lbits  = to_bits(lbf, lx.size());
        const size_t before = out.paired_pt2ln.size();
        out.paired_pt2ln.resize(before + lx.size());
        static_assert(sizeof(mp2p_icp::point_line_pair_t) == sizeof(mp2p_b200_pair_pt2ln));
        uint64_t     cnt = 0, pot = 0;
        const float* resident = device_cloud(pcLocal, device, cacheLocalCloud);
        check(mp2p_b200_match_pt2ln(ctx(device), gmap, resident ? resident : lx.data(), resident ? nullptr : ly.data(),
                                    resident ? nullptr : lz.data(), lx.size(),
                                    resident ? MP2P_B200_LOCAL_CLOUD : MP2P_B200_LOCAL_HOST, T, &p, bits_or_null(lbits),
                                    reinterpret_cast<mp2p_b200_pair_pt2ln*>(out.paired_pt2ln.data() + before), lx.size(), 0,
                                    &cnt, &pot));
        out.paired_pt2ln.resize(before + cnt);
        out.potential_pairings += pot;
        // :159 — mark the local points (output is in ascending local index: parallel walk)
        size_t i = 0;
        for (size_t k = before; k < out.paired_pt2ln.size(); k++)
        {
            const auto& r = out.paired_pt2ln[k].pt_local;
            while (i < lx.size() && !(double(lx[i]) == r.x && double(ly[i]) == r.y && double(lz[i]) == r.z && !lbf[i])) i++;
            if (i < lx.size()) lbf.mark_as_set(i++);
        }
    }
};
IMPLEMENTS_MRPT_OBJECT(Matcher_Point2Line_B200, Matcher, mp2p_icp)

/** Drop-in for QualityEvaluator_PairedRatio (mp2p_icp/src/QualityEvaluator_PairedRatio.cpp:27-73):
 *  the extra matcher pass of the non-reuse mode runs on the GPU. The reference holds its matcher by
 *  value (QualityEvaluator_PairedRatio.h:62), so the swap needs this class, named in the YAML's
 *  `quality:` list (ICP.cpp:589-606). */
class QualityEvaluator_PairedRatio_B200 : public QualityEvaluator
{
    DEFINE_MRPT_OBJECT(QualityEvaluator_PairedRatio_B200, mp2p_icp)
   public:
    int  device = 0;              //!< YAML `device`: the GPU this object works on
    bool cacheLocalCloud = true;  //!< YAML: keep the local layer on the device while its fingerprint is unchanged
    void initialize(const mrpt::containers::yaml& params) override
    {
        MCP_LOAD_OPT(params, device);
        MCP_LOAD_OPT(params, reuse_icp_pairings);
        MCP_LOAD_OPT(params, absolute_minimum_pairing_ratio);
        if (!reuse_icp_pairings)
        {
            mrpt::containers::yaml p = params;
            if (!p.has("allowMatchAlreadyMatchedGlobalPoints")) p["allowMatchAlreadyMatchedGlobalPoints"] = true;
            matcher_.initialize(p);
        }
    }
    Result evaluate(const metric_map_t& pcGlobal, const metric_map_t& pcLocal, const mrpt::poses::CPose3D& localPose,
                    const Pairings& pairingsFromICP) const override
    {
        const Pairings* pairings = &pairingsFromICP;
        Pairings        newPairings;
        if (!reuse_icp_pairings)
        {
            MatchState ms(pcGlobal, pcLocal);
            matcher_.match(pcGlobal, pcLocal, localPose, {}, ms, newPairings);
            pairings = &newPairings;
        }
        const auto nEffectiveLocalPoints = pairings->potential_pairings;
        Result     r;
        r.quality      = nEffectiveLocalPoints ? pairings->size() / double(nEffectiveLocalPoints) : .0;
        r.hard_discard = r.quality < absolute_minimum_pairing_ratio;
        return r;
    }
    void attachToParameterSource(ParameterSource& source) override
    {
        source.attach(*this);
        source.attach(matcher_);
    }

   private:
    Matcher_Points_DistanceThreshold_B200 matcher_;
    bool                                  reuse_icp_pairings             = true;
    double                                absolute_minimum_pairing_ratio = 0.20;
};
IMPLEMENTS_MRPT_OBJECT(QualityEvaluator_PairedRatio_B200, QualityEvaluator, mp2p_icp)

}  // namespace mp2p_icp

namespace mp2p_icp_filters
{
/** Drop-in for FilterDecimateVoxels (mp2p_icp_filters/src/FilterDecimateVoxels.cpp:109-378): same
 *  YAML parameters (initialize() is the reference's own), the voxel grid on the GPU. The points of all input layers
 *  go through ONE grid, as upstream. Output ORDER: ascending (cx, cy, cz) — the reference's order only with
 *  use_tsl_robin_map = false; with the default hash map its order is implementation-defined, the point SET is equal.
 *  DecimateMethod::RandomPoint (an unseeded generator upstream) goes to the reference implementation. */
class FilterDecimateVoxels_B200 : public FilterDecimateVoxels
{
    DEFINE_MRPT_OBJECT(FilterDecimateVoxels_B200, mp2p_icp_filters)
   public:
    int  device = 0;  //!< YAML `device`
    void initialize(const mrpt::containers::yaml& c) override
    {
        FilterDecimateVoxels::initialize(c);
        MCP_LOAD_OPT(c, device);
    }
    void filter(mp2p_icp::metric_map_t& inOut) const override
    {
        using namespace mp2p_icp::b200_detail;
        if (params_.decimate_method == DecimateMethod::RandomPoint) return FilterDecimateVoxels::filter(inOut);
        checkAllParametersAreRealized();
        std::vector<const mrpt::maps::CPointsMap*> in;  // :116-141
        for (const auto& name : params_.input_pointcloud_layer)
        {
            auto it = inOut.layers.find(name);
            if (it == inOut.layers.end())
            {
                if (params_.error_on_missing_input_layer) THROW_EXCEPTION_FMT("Input layer '%s' not found on input map.", name.c_str());
                continue;
            }
            const auto* pc = mp2p_icp::MapToPointsMap(*it->second);
            if (!pc) THROW_EXCEPTION_FMT("Layer '%s' must be of point cloud type.", name.c_str());
            in.push_back(pc);
        }
        ASSERT_(!in.empty());
        ASSERT_(!params_.output_pointcloud_layer.empty());
        auto outPc = GetOrCreatePointLayer(inOut, params_.output_pointcloud_layer, false, in.at(0)->GetRuntimeClass()->className);
        // small layers are copied whole (:156-189), the others go through the grid
        std::vector<float> cx, cy, cz;  // only used when several layers have to be put behind each other
        const float *      x = nullptr, *y = nullptr, *z = nullptr;
        size_t             n = 0, n_grid_layers = 0;
        for (const auto* pc : in)
        {
            const auto &xs = pc->getPointsBufferRef_x(), &ys = pc->getPointsBufferRef_y(), &zs = pc->getPointsBufferRef_z();
            if (params_.minimum_input_points_to_filter > 0 && xs.size() <= params_.minimum_input_points_to_filter)
            {
                for (size_t i = 0; i < xs.size(); i++)
                    if (params_.flatten_to.has_value())
                        outPc->insertPointFast(xs[i], ys[i], static_cast<float>(*params_.flatten_to));
                    else
                        outPc->insertPointFrom(*pc, i);
                continue;
            }
            if (n_grid_layers++ == 0)
                x = xs.data(), y = ys.data(), z = zs.data(), n = xs.size();
            else
            {
                if (cx.empty()) cx.assign(x, x + n), cy.assign(y, y + n), cz.assign(z, z + n);
                cx.insert(cx.end(), xs.begin(), xs.end()), cy.insert(cy.end(), ys.begin(), ys.end()), cz.insert(cz.end(), zs.begin(), zs.end());
                x = cx.data(), y = cy.data(), z = cz.data(), n = cx.size();
            }
        }
        if (!n) return;
        mp2p_b200_decimate_params p{params_.voxel_filter_resolution, static_cast<int32_t>(params_.decimate_method),
                                    params_.flatten_to.has_value() ? 1 : 0,
                                    params_.flatten_to.has_value() ? static_cast<float>(*params_.flatten_to) : 0.f};
        std::vector<float> ox(n), oy(n), oz(n);
        uint64_t           cnt = 0;
        check(mp2p_b200_filter_decimate_voxels(ctx(device), x, y, z, n, 0, &p, ox.data(), oy.data(), oz.data(), nullptr, n, 0, &cnt));
        outPc->reserve(outPc->size() + cnt);
        for (uint64_t i = 0; i < cnt; i++) outPc->insertPointFast(ox[i], oy[i], oz[i]);
        outPc->mark_as_modified();
    }
};
IMPLEMENTS_MRPT_OBJECT(FilterDecimateVoxels_B200, FilterDecimateVoxels, mp2p_icp_filters)
}  // namespace mp2p_icp_filters

// at global scope, like the reference's own (mp2p_icp/src/register.cpp:43-69)
MRPT_INITIALIZER(register_mp2p_icp_b200)
{
    using mrpt::rtti::registerClass;
    registerClass(CLASS_ID(mp2p_icp::Matcher_Points_DistanceThreshold_B200));
    registerClass(CLASS_ID(mp2p_icp::Matcher_Points_InlierRatio_B200));
    registerClass(CLASS_ID(mp2p_icp::Matcher_Point2Line_B200));
    registerClass(CLASS_ID(mp2p_icp::Matcher_Adaptive_B200));
    registerClass(CLASS_ID(mp2p_icp::Solver_Horn_B200));
    registerClass(CLASS_ID(mp2p_icp::Matcher_Point2Plane_B200));
    registerClass(CLASS_ID(mp2p_icp::Solver_GaussNewton_B200));
    registerClass(CLASS_ID(mp2p_icp::QualityEvaluator_PairedRatio_B200));
    registerClass(CLASS_ID(mp2p_icp_filters::FilterDecimateVoxels_B200));
}

#endif  // MP2P_B200_WITH_MRPT
