// SHAPE STUBS — test infrastructure, not the reference. Declarations with the names, virtual signatures and
// member layouts of the mp2p_icp classes that mp2p_icp_b200/host/mrpt_plugin.cpp derives from or touches,
// restated from the reference headers cited per item (paths relative to the reference checkout). Used only
// to type-check the plugin where MRPT / mp2p_icp are not installed (tests/test_plugin_shape.py).
#pragma once
#include <mrpt_shape.h>

#include <any>
#include <set>

namespace mp2p_icp
{
using layer_name_t = std::string;  // mp2p_icp_map/include/mp2p_icp/layer_name_t.h

// mp2p_icp_map/include/mp2p_icp/NearestPlaneCapable.h:39-51 (not used by the plugin; kept for MapToNP)
class NearestPlaneCapable;

// mp2p_icp_map/include/mp2p_icp/metricmap.h:64-91, 273-292
class metric_map_t : public mrpt::serialization::CSerializable
{
   public:
    std::map<layer_name_t, mrpt::maps::CMetricMap::Ptr> layers;
};
const mrpt::maps::CPointsMap*              MapToPointsMap(const mrpt::maps::CMetricMap& map);
const mrpt::maps::NearestNeighborsCapable* MapToNN(const mrpt::maps::CMetricMap& map, bool throwIfNotImplemented);

// mp2p_icp_map/include/mp2p_icp/pointcloud_bitfield.h:46-133
struct pointcloud_bitfield_t
{
    struct DenseOrSparseBitField
    {
        void               assign(size_t numElements, bool dense);
        [[nodiscard]] bool operator[](const size_t id) const;
        void               mark_as_set(const size_t id);

       private:
        std::optional<std::vector<bool>> dense_;
        std::set<uint64_t>               sparse_;
    };
    std::map<layer_name_t, DenseOrSparseBitField> point_layers;
    std::vector<bool>                             lines, planes;
    void                                          initialize_from(const metric_map_t& pc);
};

// mp2p_icp_map/include/mp2p_icp/Parameterizable.h:60-200
class Parameterizable;
class ParameterSource
{
   public:
    void attach(Parameterizable& obj);
    void updateVariable(const std::string& variable, double value);
    void realize();
};
class Parameterizable
{
   public:
    virtual ~Parameterizable() = default;
    virtual void attachToParameterSource(ParameterSource& source) { source.attach(*this); }
    void         checkAllParametersAreRealized() const;
    void         unrealizeParameters();

   protected:
    // only these three target types exist upstream (Parameterizable.h:155-158)
    void parseAndDeclareParameter(const std::string& value, double& target);
    void parseAndDeclareParameter(const std::string& value, float& target);
    void parseAndDeclareParameter(const std::string& value, uint32_t& target);
};
#define DECLARE_PARAMETER_IN_OPT(__yaml, __variable, __object)    \
    __object.mp2p_icp::Parameterizable::parseAndDeclareParameter( \
        (__yaml).getOrDefault(#__variable, std::to_string(__variable)), __variable);
#define DECLARE_PARAMETER_OPT(__yaml, __variable) DECLARE_PARAMETER_IN_OPT(__yaml, __variable, (*this))
#define DECLARE_PARAMETER_IN_REQ(__yaml, __variable, __object)                                                      \
    if (!(__yaml).has(#__variable))                                                                                 \
        throw std::invalid_argument(mrpt::format("Required parameter `%s` not an existing key in dictionary.", #__variable)); \
    (__object).mp2p_icp::Parameterizable::parseAndDeclareParameter((__yaml)[#__variable].as<std::string>(), __variable);
#define DECLARE_PARAMETER_REQ(__yaml, __variable) DECLARE_PARAMETER_IN_REQ(__yaml, __variable, (*this))

// mp2p_icp_map/include/mp2p_icp/plane_patch.h:30-34, point_plane_pair_t.h:34-38
struct plane_patch_t
{
    mrpt::math::TPlane   plane;
    mrpt::math::TPoint3D centroid;
};
struct point_plane_pair_t
{
    plane_patch_t         pl_global;
    mrpt::math::TPoint3Df pt_local;
};
using MatchedPointPlaneList = std::vector<point_plane_pair_t>;

// mp2p_icp/include/mp2p_icp/Pairings.h:34-133
struct matched_plane_t
{
    plane_patch_t p_global, p_local;
};
struct matched_line_t
{
    mrpt::math::TLine3D ln_global, ln_local;
};
struct point_line_pair_t
{
    mrpt::math::TLine3D  ln_global;
    mrpt::math::TPoint3D pt_local;
};
struct Pairings
{
    virtual ~Pairings() = default;
    mrpt::tfest::TMatchingPairList              paired_pt2pt;
    std::vector<point_line_pair_t>              paired_pt2ln;
    MatchedPointPlaneList                       paired_pt2pl;
    std::vector<matched_line_t>                 paired_ln2ln;
    std::vector<matched_plane_t>                paired_pl2pl;
    uint64_t                                    potential_pairings = 0;
    std::vector<std::pair<std::size_t, double>> point_weights;
    virtual bool                                empty() const;
    virtual size_t                              size() const;
};
struct OutlierIndices
{
    std::vector<std::size_t> point2point, line2line, plane2plane;
};

// mp2p_icp/include/mp2p_icp/robust_kernels.h:29-40, PairWeights.h:28-38, WeightParameters.h:36-63, OptimalTF_Result.h:28-35
enum class RobustKernel : uint8_t
{
    None = 0,
    GemanMcClure,
    Cauchy,
};
struct PairWeights
{
    double pt2pt = 1.0, pt2ln = 1.0, pt2pl = 1.0, ln2ln = 1.0, pl2pl = 1.0;
};
struct WeightParameters
{
    bool                                use_scale_outlier_detector = false;
    double                              scale_outlier_threshold{1.20};
    PairWeights                         pair_weights;
    RobustKernel                        robust_kernel = RobustKernel::None;
    std::optional<mrpt::poses::CPose3D> currentEstimateForRobust;
    double                              robust_kernel_param = 1.0;
};
struct OptimalTF_Result
{
    mrpt::poses::CPose3D optimalPose;
    double               optimalScale = 1.0;
    OutlierIndices       outliers;
};

// mp2p_icp/include/mp2p_icp/Matcher.h:36-108
struct MatchContext
{
    uint32_t icpIteration = 0;
};
struct MatchState
{
    MatchState(const metric_map_t& pcGlobal, const metric_map_t& pcLocal);
    pointcloud_bitfield_t localPairedBitField, globalPairedBitField;
};
class Matcher : public mrpt::system::COutputLogger, public mrpt::rtti::CObject, public mp2p_icp::Parameterizable
{
    DEFINE_VIRTUAL_MRPT_OBJECT(Matcher, mp2p_icp)
   public:
    virtual void initialize(const mrpt::containers::yaml& params);
    virtual bool match(const metric_map_t& pcGlobal, const metric_map_t& pcLocal, const mrpt::poses::CPose3D& localPose,
                       const MatchContext& mc, MatchState& ms, Pairings& out) const;
    uint32_t     runFromIteration = 0, runUpToIteration = 0;
    bool         enabled = true;

   protected:
    virtual bool impl_match(const metric_map_t& pcGlobal, const metric_map_t& pcLocal, const mrpt::poses::CPose3D& localPose,
                            const MatchContext& mc, MatchState& ms, Pairings& out) const = 0;
};

// mp2p_icp/include/mp2p_icp/Matcher_Points_Base.h:40-128
class Matcher_Points_Base : public Matcher
{
   public:
    Matcher_Points_Base() = default;
    std::map<std::string, std::map<std::string, double>> weight_pt2pt_layers;
    uint64_t                   maxLocalPointsPerLayer_ = 0, localPointsSampleSeed_ = 0;
    bool                       allowMatchAlreadyMatchedPoints_       = false;
    bool                       allowMatchAlreadyMatchedGlobalPoints_ = false;
    std::optional<std::size_t> kdtree_leaf_max_points_;
    double                     bounding_box_intersection_check_epsilon_ = 0.20;
    void                       initialize(const mrpt::containers::yaml& params) override;

   protected:
    bool impl_match(const metric_map_t& pcGlobal, const metric_map_t& pcLocal, const mrpt::poses::CPose3D& localPose,
                    const MatchContext& mc, MatchState& ms, Pairings& out) const override final;

   private:
    virtual void implMatchOneLayer(const mrpt::maps::CMetricMap& pcGlobal, const mrpt::maps::CPointsMap& pcLocal,
                                   const mrpt::poses::CPose3D& localPose, MatchState& ms, const layer_name_t& globalName,
                                   const layer_name_t& localName, Pairings& out) const = 0;
};

// mp2p_icp/include/mp2p_icp/Solver.h:43-101
class Solver;
struct SolverContext
{
    std::optional<mrpt::poses::CPose3D>               guessRelativePose, currentCorrectionFromInitialGuess, lastIcpStepIncrement;
    std::optional<mrpt::poses::CPose3DPDFGaussianInf> prior;
    mutable std::map<const Solver*, std::map<std::string, std::any>> perSolverPersistentData;
    std::optional<uint32_t>                                          icpIteration;
};
class Solver : public mrpt::system::COutputLogger, public mrpt::rtti::CObject, public mp2p_icp::Parameterizable
{
    DEFINE_VIRTUAL_MRPT_OBJECT(Solver, mp2p_icp)
   public:
    virtual void initialize(const mrpt::containers::yaml& params);
    virtual bool optimal_pose(const Pairings& pairings, OptimalTF_Result& out, const SolverContext& sc) const;
    uint32_t     runFromIteration = 0, runUpToIteration = 0;
    double       runUntilTranslationCorrectionSmallerThan = 0;
    bool         enabled                                  = true;

   protected:
    virtual bool impl_optimal_pose(const Pairings& pairings, OptimalTF_Result& out, const SolverContext& sc) const = 0;
};
// mp2p_icp/include/mp2p_icp/Solver_Horn.h:28-45
class Solver_Horn : public Solver
{
    DEFINE_MRPT_OBJECT(Solver_Horn, mp2p_icp)
   public:
    WeightParameters pairingsWeightParameters;
    void             initialize(const mrpt::containers::yaml& params) override;

   protected:
    bool impl_optimal_pose(const Pairings& pairings, OptimalTF_Result& out, const SolverContext& sc) const override;
};
// mp2p_icp/include/mp2p_icp/Solver_GaussNewton.h:32-52
class Solver_GaussNewton : public Solver
{
    DEFINE_MRPT_OBJECT(Solver_GaussNewton, mp2p_icp)
   public:
    uint32_t     maxIterations = 5;
    PairWeights  pairWeights;
    RobustKernel robustKernel      = RobustKernel::None;
    double       robustKernelParam = 1.0;
    bool         innerLoopVerbose  = false;
    void         initialize(const mrpt::containers::yaml& params) override;

   protected:
    bool impl_optimal_pose(const Pairings& pairings, OptimalTF_Result& out, const SolverContext& sc) const override;
};

// mp2p_icp/include/mp2p_icp/QualityEvaluator.h:30-60
class QualityEvaluator : public mrpt::system::COutputLogger, public mrpt::rtti::CObject, public mp2p_icp::Parameterizable
{
    DEFINE_VIRTUAL_MRPT_OBJECT(QualityEvaluator, mp2p_icp)
   public:
    struct Result
    {
        double quality      = .0;
        bool   hard_discard = false;
    };
    virtual void   initialize(const mrpt::containers::yaml& params)                                                    = 0;
    virtual Result evaluate(const metric_map_t& pcGlobal, const metric_map_t& pcLocal, const mrpt::poses::CPose3D& localPose,
                            const Pairings& pairingsFromICP) const = 0;
};

// mp2p_icp_filters/include/mp2p_icp_filters/FilterBase.h (the filter plugin class derives from it)
}  // namespace mp2p_icp
