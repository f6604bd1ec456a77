// SHAPE STUBS — test infrastructure, not MRPT. The smallest set of declarations with the names, member
// signatures and layouts of the MRPT 2.x types that mp2p_icp_b200/host/mrpt_plugin.cpp touches, so that
// the plugin (which can only be BUILT where MRPT >= 2.11.5 exists — not in this image) is at least
// type-checked against the interface it claims to implement: `g++ -fsyntax-only -DMP2P_B200_WITH_MRPT
// -Itests/stubs mrpt_plugin.cpp` (tests/test_plugin_shape.py, __graft_entry__.build()).
// Written from the uses in the reference tree (file:line cited per item), not from MRPT sources.
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

namespace mrpt
{
template <class T>
using aligned_std_vector = std::vector<T>;  // Matcher_Points_Base.h:105 (x_locals, y_locals, z_locals)
template <class T>
inline T square(const T x)
{
    return x * x;
}
std::string format(const char* fmt, ...);  // Parameterizable.h:193 (DECLARE_PARAMETER_IN_REQ)

namespace math
{
struct TPoint3Df  // Matcher_Points_Base.h:99-100; TMatchingPair members
{
    float x = 0, y = 0, z = 0;
};
struct TPoint3D  // Pairings.h:66 (point_line_pair_t::pt_local)
{
    double x = 0, y = 0, z = 0;
};
struct TPlane  // plane_patch.h:32
{
    double coefs[4] = {0, 0, 0, 0};
};
struct TLine3D  // Pairings.h:65: {pBase, director[3]}
{
    TPoint3D              pBase;
    std::array<double, 3> director{{0, 0, 0}};
};
struct CMatrixDouble44  // optimal_tf_horn.cpp / CPose3D(const CMatrixDouble44&)
{
    double               m[4][4] = {};
    static CMatrixDouble44 Identity()
    {
        CMatrixDouble44 r;
        for (int i = 0; i < 4; i++) r.m[i][i] = 1.0;
        return r;
    }
    double& operator()(int r, int c) { return m[r][c]; }
    double  operator()(int r, int c) const { return m[r][c]; }
};
struct CMatrixDouble33
{
    double m[3][3] = {};
    double operator()(int r, int c) const { return m[r][c]; }
};
// Matcher_Adaptive.cpp:190-199
template <class T, class V>
void linspace(T first, T last, size_t count, V& out);
void confidenceIntervalsFromHistogram(const std::vector<double>& xs, const std::vector<double>& vals, double& out_lower,
                                      double& out_upper, double confidenceInterval);
}  // namespace math

namespace poses
{
class CPose3D  // Matcher.h:98 (localPose), Solver.h:46 (guessRelativePose), OptimalTF_Result.h:31
{
   public:
    CPose3D() = default;
    explicit CPose3D(const math::CMatrixDouble44&) {}
    const math::CMatrixDouble33& getRotationMatrix() const { return R_; }
    std::array<double, 3>        m_coords{{0, 0, 0}};

   private:
    math::CMatrixDouble33 R_;
};
class CPose3DPDFGaussianInf  // Solver.h:55 (SolverContext::prior)
{
};
}  // namespace poses

namespace containers
{
class yaml  // Matcher.h:91 (initialize), Parameterizable.h:180-196
{
   public:
    struct node
    {
        template <class T>
        T as() const;
        template <class T>
        node& operator=(const T&);
    };
    bool        has(const std::string& key) const;
    node        operator[](const char* key);
    const node  operator[](const char* key) const;
    template <class T>
    T getOrDefault(const std::string& key, const T& def) const;
};
}  // namespace containers

namespace rtti
{
struct TRuntimeClassId
{
};
class CObject  // Matcher.h:84
{
   public:
    using Ptr = std::shared_ptr<CObject>;
    virtual ~CObject() = default;
};
void registerClass(const TRuntimeClassId* cls);  // register.cpp:45
}  // namespace rtti

namespace system
{
class COutputLogger  // Matcher.h:83
{
};
}  // namespace system

namespace serialization
{
class CSerializable : public rtti::CObject
{
};
}  // namespace serialization

namespace tfest
{
#pragma pack(push, 1)
struct TMatchingPair  // Matcher_Points_DistanceThreshold.cpp:106-113 (fields), SURVEY appendix A: 36 bytes
{
    uint32_t        globalIdx = 0, localIdx = 0;
    math::TPoint3Df global, local;
    float           errorSquareAfterTransformation = 0;
};
#pragma pack(pop)
using TMatchingPairList = std::vector<TMatchingPair>;  // Pairings.h:95
}  // namespace tfest

namespace maps
{
class NearestNeighborsCapable  // pointcloud_bitfield.h:108-113
{
   public:
    virtual ~NearestNeighborsCapable() = default;
    virtual size_t nn_index_count() const        = 0;
    virtual bool   nn_has_indices_or_ids() const = 0;
};
class CMetricMap : public serialization::CSerializable  // Matcher_Points_Base.h:126
{
   public:
    using Ptr = std::shared_ptr<CMetricMap>;
    struct ClassInfo
    {
        const char* className = "";
    };
    const ClassInfo* GetRuntimeClass() const { return &info_; }  // FilterDecimateVoxels.cpp:150

   private:
    ClassInfo info_;
};
class CPointsMap : public CMetricMap, public NearestNeighborsCapable  // Matcher_Points_Base.cpp:201-203
{
   public:
    const mrpt::aligned_std_vector<float>& getPointsBufferRef_x() const { return x_; }
    const mrpt::aligned_std_vector<float>& getPointsBufferRef_y() const { return y_; }
    const mrpt::aligned_std_vector<float>& getPointsBufferRef_z() const { return z_; }
    size_t                                 size() const { return x_.size(); }
    using Ptr = std::shared_ptr<CPointsMap>;
    void reserve(size_t n) { x_.reserve(n), y_.reserve(n), z_.reserve(n); }                         // FilterDecimateVoxels.cpp:152
    void insertPointFast(float x, float y, float z) { x_.push_back(x), y_.push_back(y), z_.push_back(z); }  // :175
    void insertPointFrom(const CPointsMap& o, size_t i) { insertPointFast(o.x_[i], o.y_[i], o.z_[i]); }     // :179
    void mark_as_modified() {}                                                                      // Matcher_Points_Base.cpp:112
    size_t                                 nn_index_count() const override { return x_.size(); }
    bool                                   nn_has_indices_or_ids() const override { return true; }

   private:
    mrpt::aligned_std_vector<float> x_, y_, z_;
};
}  // namespace maps
}  // namespace mrpt

// ---- macros (mrpt/core/exceptions.h, mrpt/rtti/CObject.h, mrpt/core/initializer.h, mrpt/containers/yaml.h) ----
#define THROW_EXCEPTION_FMT(fmt_, ...)                        \
    do                                                        \
    {                                                         \
        char b_[512];                                         \
        std::snprintf(b_, sizeof(b_), fmt_, __VA_ARGS__);     \
        throw std::runtime_error(b_);                         \
    } while (0)
#define ASSERT_(c_)                                           \
    do                                                        \
    {                                                         \
        if (!(c_)) throw std::runtime_error("ASSERT_ " #c_);  \
    } while (0)
#define ASSERTMSG_(c_, m_)                       \
    do                                           \
    {                                            \
        if (!(c_)) throw std::runtime_error(m_); \
    } while (0)
#define ASSERT_GT_(a_, b_) ASSERT_((a_) > (b_))
#define ASSERT_GE_(a_, b_) ASSERT_((a_) >= (b_))
#define ASSERT_LT_(a_, b_) ASSERT_((a_) < (b_))
// DEFINE_VIRTUAL_MRPT_OBJECT / DEFINE_MRPT_OBJECT(Class, NS): Matcher.h:86, Matcher_Points_DistanceThreshold.h:40
#define DEFINE_VIRTUAL_MRPT_OBJECT(class_, ns_)                                \
   public:                                                                     \
    using Ptr = std::shared_ptr<class_>;                                       \
    static const mrpt::rtti::TRuntimeClassId* _GetBaseClass();                 \
    static const mrpt::rtti::TRuntimeClassId& GetRuntimeClassIdStatic();       \
                                                                               \
   private:
#define DEFINE_MRPT_OBJECT(class_, ns_)                                        \
   public:                                                                     \
    using Ptr = std::shared_ptr<class_>;                                       \
    static const mrpt::rtti::TRuntimeClassId& GetRuntimeClassIdStatic();       \
    static std::shared_ptr<mrpt::rtti::CObject> CreateObject();                \
                                                                               \
   private:
// IMPLEMENTS_MRPT_OBJECT(Class, Base, NS): Matcher_Points_DistanceThreshold.cpp:30 — at namespace scope
#define IMPLEMENTS_MRPT_OBJECT(class_, base_, ns_)                                                             \
    const mrpt::rtti::TRuntimeClassId& ns_::class_::GetRuntimeClassIdStatic()                                  \
    {                                                                                                          \
        static_assert(std::is_base_of<ns_::base_, ns_::class_>::value, #class_ " must derive from " #base_);   \
        static mrpt::rtti::TRuntimeClassId id;                                                                 \
        return id;                                                                                             \
    }                                                                                                          \
    std::shared_ptr<mrpt::rtti::CObject> ns_::class_::CreateObject() { return std::make_shared<ns_::class_>(); }
#define DEFINE_SERIALIZABLE(class_, ns_) DEFINE_MRPT_OBJECT(class_, ns_)
#define CLASS_ID(T_) (&T_::GetRuntimeClassIdStatic())
// MRPT_INITIALIZER(f): register.cpp:43 — a function run at load time, GLOBAL scope
#define MRPT_INITIALIZER(f_)                  \
    static void f_();                         \
    namespace                                 \
    {                                         \
    struct f_##_runner                        \
    {                                         \
        f_##_runner() { f_(); }               \
    } f_##_instance;                          \
    }                                         \
    static void f_()
// MCP_LOAD_REQ / MCP_LOAD_OPT(yaml, var): Matcher_Adaptive.cpp:36-49
#define MCP_LOAD_REQ(y_, v_)                                                                  \
    do                                                                                        \
    {                                                                                         \
        if (!(y_).has(#v_)) throw std::invalid_argument("Required parameter `" #v_ "` missing"); \
        v_ = (y_)[#v_].as<decltype(v_)>();                                                    \
    } while (0)
#define MCP_LOAD_OPT(y_, v_) v_ = (y_).getOrDefault(#v_, v_)
