// SHAPE STUBS — test infrastructure, not the reference (see mp2p_icp_shape.h). mp2p_icp_filters declarations the
// plugin's FilterDecimateVoxels_B200 derives from or touches, restated from the cited reference headers.
#pragma once
#include <mp2p_icp_shape.h>

namespace mp2p_icp_filters
{
// mp2p_icp_filters/include/mp2p_icp_filters/FilterBase.h:44-74
class FilterBase : public mrpt::rtti::CObject, public mrpt::system::COutputLogger, public mp2p_icp::Parameterizable
{
    DEFINE_VIRTUAL_MRPT_OBJECT(FilterBase, mp2p_icp_filters)
   public:
    FilterBase();
    virtual ~FilterBase();
    virtual void initialize(const mrpt::containers::yaml& cfg_block) = 0;
    virtual void filter(mp2p_icp::metric_map_t& inOut) const        = 0;
};
// mp2p_icp_filters/include/mp2p_icp_filters/GetOrCreatePointLayer.h:33-35
[[nodiscard]] mrpt::maps::CPointsMap::Ptr GetOrCreatePointLayer(mp2p_icp::metric_map_t& m, const std::string& layerName,
                                                                bool               allowEmptyName = true,
                                                                const std::string& classForLayerCreation = "mrpt::maps::CSimplePointsMap");
// mp2p_icp_filters/include/mp2p_icp_filters/FilterDecimateVoxels.h:34-120
enum class DecimateMethod : uint8_t
{
    FirstPoint = 0,
    ClosestToAverage,
    VoxelAverage,
    RandomPoint
};
class FilterDecimateVoxels : public mp2p_icp_filters::FilterBase
{
    DEFINE_MRPT_OBJECT(FilterDecimateVoxels, mp2p_icp_filters)
   public:
    FilterDecimateVoxels();
    void initialize(const mrpt::containers::yaml& c) override;
    void filter(mp2p_icp::metric_map_t& inOut) const override;
    struct Parameters
    {
        std::vector<std::string> input_pointcloud_layer = {"raw"};
        bool                     error_on_missing_input_layer = true;
        std::string              output_pointcloud_layer;
        float                    voxel_filter_resolution = 1.0f;
        bool                     use_tsl_robin_map       = true;
        uint32_t                 minimum_input_points_to_filter = 0;
        std::optional<double>    flatten_to;
        DecimateMethod           decimate_method = DecimateMethod::FirstPoint;
    };
    Parameters params_;
};
}  // namespace mp2p_icp_filters
