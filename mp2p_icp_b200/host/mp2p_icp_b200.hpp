// MRPT-free C++ host mirror of the reference's Matcher / Solver plugin interface for the hot path.
//
// Same class names, parameter names, argument meaning and error behaviour as the reference
// (exceptions for bad parameters, `false` for "did not run"), over minimal stand-ins for the MRPT /
// mp2p_icp_map types the path touches. Everything computes through the C ABI (mp2p_b200.h): no CPU
// implementation of the path lives here. With MRPT available the real plugin classes are in
// mrpt_plugin.cpp; this header is what can be built and tested without MRPT.
//
// Reference interfaces mirrored (paths relative to the reference checkout):
//   Matcher, MatchContext, MatchState, run_matchers   mp2p_icp/include/mp2p_icp/Matcher.h:36-109, src/Matcher.cpp:35-88
//   Matcher_Points_Base                              mp2p_icp/src/Matcher_Points_Base.cpp:30-181
//   Matcher_Points_DistanceThreshold                 mp2p_icp/src/Matcher_Points_DistanceThreshold.cpp:39-269
//   Matcher_Point2Plane                              mp2p_icp/src/Matcher_Point2Plane.cpp:35-114
//   Solver, SolverContext                            mp2p_icp/include/mp2p_icp/Solver.h:43-102, src/Solver.cpp:28-64
//   Solver_Horn / Solver_GaussNewton                 mp2p_icp/src/Solver_Horn.cpp:33-61, Solver_GaussNewton.cpp:29-67
//   Pairings                                         mp2p_icp/include/mp2p_icp/Pairings.h:84-194, src/Pairings.cpp:123-147
//   Parameterizable, ParameterSource                 mp2p_icp_map/include/mp2p_icp/Parameterizable.h, src/Parameterizable.cpp
//   Matcher_Points_InlierRatio                       mp2p_icp/src/Matcher_Points_InlierRatio.cpp:35-143
//   Matcher_Point2Line                               mp2p_icp/src/Matcher_Point2Line.cpp:35-163
//   Matcher_Adaptive                                 mp2p_icp/src/Matcher_Adaptive.cpp:32-314
//   QualityEvaluator, QualityEvaluator_PairedRatio   mp2p_icp/include/mp2p_icp/QualityEvaluator.h, src/QualityEvaluator_PairedRatio.cpp:27-73
//   ICP::align loop (the caller)                     mp2p_icp/src/ICP.cpp:108-338, evaluate_quality :608-634
#pragma once
#include <cctype>
#include <cmath>
#include <deque>
#include <cstdint>
#include <cstring>
#include <atomic>
#include <map>
#include <memory>
#include <optional>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "mp2p_b200.h"

namespace mp2p_icp_b200
{
using layer_name_t = std::string;

inline void check(int rc, const char* what)
{
    if (rc != MP2P_B200_OK) throw std::runtime_error(std::string(what) + ": " + mp2p_b200_last_error());
}

// ---- mrpt::poses::CPose3D stand-in: 3x4 row-major [R|t] -----------------------------------
struct CPose3D
{
    double m[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    CPose3D()    = default;
    CPose3D(double x, double y, double z, double yaw, double pitch, double roll)
    {
        const double cy = std::cos(yaw), sy = std::sin(yaw), cp = std::cos(pitch), sp = std::sin(pitch),
                     cr = std::cos(roll), sr = std::sin(roll);
        const double v[12] = {cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr, x,
                              sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr, y,
                              -sp,     cp * sr,                cp * cr,                z};
        std::memcpy(m, v, sizeof(m));
    }
    static CPose3D Identity() { return {}; }
    CPose3D        operator+(const CPose3D& b) const  // composition a (+) b
    {
        CPose3D o;
        for (int r = 0; r < 3; r++)
        {
            for (int c = 0; c < 3; c++)
                o.m[4 * r + c] = m[4 * r] * b.m[c] + m[4 * r + 1] * b.m[4 + c] + m[4 * r + 2] * b.m[8 + c];
            o.m[4 * r + 3] = m[4 * r] * b.m[3] + m[4 * r + 1] * b.m[7] + m[4 * r + 2] * b.m[11] + m[4 * r + 3];
        }
        return o;
    }
    CPose3D inverse() const
    {
        CPose3D o;
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) o.m[4 * r + c] = m[4 * c + r];
        for (int r = 0; r < 3; r++) o.m[4 * r + 3] = -(m[r] * m[3] + m[4 + r] * m[7] + m[8 + r] * m[11]);
        return o;
    }
    CPose3D operator-(const CPose3D& b) const { return b.inverse() + *this; }  // a (-) b = b^-1 (+) a
    void    inverseComposePoint(double gx, double gy, double gz, double& lx, double& ly, double& lz) const
    {
        const double dx = gx - m[3], dy = gy - m[7], dz = gz - m[11];
        lx = m[0] * dx + m[4] * dy + m[8] * dz;
        ly = m[1] * dx + m[5] * dy + m[9] * dz;
        lz = m[2] * dx + m[6] * dy + m[10] * dz;
    }
    /** SE(3) log as (v, w): only the norms are needed by the ICP termination rule (ICP.cpp:194-229). */
    void log_norms(double& n_xyz, double& n_rot) const
    {
        const double tr = m[0] + m[5] + m[10];
        double       c  = 0.5 * (tr - 1.0);
        c               = c > 1 ? 1 : (c < -1 ? -1 : c);
        const double th = std::acos(c);
        n_rot           = th;
        // v = V^-1 t ;  |v| via V^-1 = I - 1/2 W + D W^2
        double w[3] = {m[9] - m[6], m[2] - m[8], m[4] - m[1]};
        double k    = th < 1e-7 ? 0.5 * (1.0 + th * th / 6.0) : th / (2.0 * std::sin(th));
        if (M_PI - th < 1e-6) k = 0;  // not reached by ICP increments
        for (double& x : w) x *= k;
        const double th2 = th * th;
        double       D;
        if (th < 1e-6)
            D = 1.0 / 12.0 + th2 / 720.0;
        else
        {
            const double A = std::sin(th) / th, B = (1.0 - std::cos(th)) / th2;
            D              = (1.0 - A / (2.0 * B)) / th2;
        }
        const double t[3] = {m[3], m[7], m[11]};
        const double wxt[3] = {w[1] * t[2] - w[2] * t[1], w[2] * t[0] - w[0] * t[2], w[0] * t[1] - w[1] * t[0]};
        const double wxwxt[3] = {w[1] * wxt[2] - w[2] * wxt[1], w[2] * wxt[0] - w[0] * wxt[2], w[0] * wxt[1] - w[1] * wxt[0]};
        double       v2 = 0;
        for (int i = 0; i < 3; i++)
        {
            const double v = t[i] - 0.5 * wxt[i] + D * wxwxt[i];
            v2 += v * v;
        }
        n_xyz = std::sqrt(v2);
    }
};

// ---- mrpt::maps::CPointsMap stand-in: three SoA float buffers + a modification stamp ---------
class CPointsMap
{
   public:
    using Ptr = std::shared_ptr<CPointsMap>;
    static Ptr Create() { return std::make_shared<CPointsMap>(); }
    void       insertPoint(float x, float y, float z)
    {
        xs_.push_back(x), ys_.push_back(y), zs_.push_back(z);
        mark_as_modified();
    }
    void                      insertPointFast(float x, float y, float z) { insertPoint(x, y, z); }
    size_t                    size() const { return xs_.size(); }
    bool                      empty() const { return xs_.empty(); }
    const std::vector<float>& getPointsBufferRef_x() const { return xs_; }
    const std::vector<float>& getPointsBufferRef_y() const { return ys_; }
    const std::vector<float>& getPointsBufferRef_z() const { return zs_; }
    // Stamps are unique across ALL objects and modifications (process-wide counter), so a device
    // cache keyed on (address, stamp) can never mistake a new map that reuses a freed address for
    // the one it cached. Copies get a fresh stamp as well.
    void                      mark_as_modified() { stamp_ = next_stamp(); }  // invalidates the NN index
    uint64_t                  stamp() const { return stamp_; }
    CPointsMap() = default;
    CPointsMap(const CPointsMap& o) : xs_(o.xs_), ys_(o.ys_), zs_(o.zs_) {}
    CPointsMap& operator=(const CPointsMap& o)
    {
        xs_ = o.xs_, ys_ = o.ys_, zs_ = o.zs_, stamp_ = next_stamp();
        return *this;
    }
    void                      reserve(size_t n) { xs_.reserve(n), ys_.reserve(n), zs_.reserve(n); }

   private:
    static uint64_t next_stamp()
    {
        static std::atomic<uint64_t> counter{0};
        return ++counter;
    }
    std::vector<float> xs_, ys_, zs_;
    uint64_t           stamp_ = next_stamp();
};

/** mp2p_icp::metric_map_t: named point layers (metricmap.h:64-151). */
struct metric_map_t
{
    static constexpr const char*               PT_LAYER_RAW = "raw";
    std::map<layer_name_t, CPointsMap::Ptr>    layers;
};

/** mp2p_icp::Pairings restricted to the pairing kinds of the hot path (Pairings.h:84-194). */
struct Pairings
{
    std::vector<mp2p_b200_pair_pt2pt>           paired_pt2pt;
    std::vector<mp2p_b200_pair_pt2pl>           paired_pt2pl;
    std::vector<mp2p_b200_pair_pt2ln>           paired_pt2ln;
    std::vector<std::pair<std::size_t, double>> point_weights;
    uint64_t                                    potential_pairings = 0;
    bool        empty() const { return paired_pt2pt.empty() && paired_pt2pl.empty() && paired_pt2ln.empty(); }
    std::size_t size() const { return paired_pt2pt.size() + paired_pt2pl.size() + paired_pt2ln.size(); }
    /** Pairings::push_back(const Pairings&): appends the lists and potential_pairings but NOT
     *  point_weights (Pairings.cpp:123-131) — kept as is. */
    void push_back(const Pairings& o)
    {
        paired_pt2pt.insert(paired_pt2pt.end(), o.paired_pt2pt.begin(), o.paired_pt2pt.end());
        paired_pt2pl.insert(paired_pt2pl.end(), o.paired_pt2pl.begin(), o.paired_pt2pl.end());
        paired_pt2ln.insert(paired_pt2ln.end(), o.paired_pt2ln.begin(), o.paired_pt2ln.end());
        potential_pairings += o.potential_pairings;
    }
};

struct MatchContext
{
    uint32_t icpIteration = 0;
};

/** MatchState: one "already paired" flag per point and layer (Matcher.h:36-70, pointcloud_bitfield.h). */
struct MatchState
{
    MatchState(const metric_map_t& pcGlobal, const metric_map_t& pcLocal)
    {
        for (const auto& kv : pcGlobal.layers) globalPaired[kv.first].assign((kv.second->size() + 31) / 32, 0u);
        for (const auto& kv : pcLocal.layers) localPaired[kv.first].assign((kv.second->size() + 31) / 32, 0u);
    }
    std::map<layer_name_t, std::vector<uint32_t>> globalPaired, localPaired;
    static void mark(std::vector<uint32_t>& w, size_t i) { w[i >> 5] |= 1u << (i & 31); }
};

/** Flat stand-in for mrpt::containers::yaml maps: key -> scalar text. */
class ParameterMap
{
   public:
    ParameterMap() = default;
    ParameterMap(std::initializer_list<std::pair<const std::string, std::string>> il) : kv_(il) {}
    template <class T>
    void set(const std::string& k, const T& v)
    {
        std::ostringstream s;
        s.precision(17);
        s << v;
        kv_[k] = s.str();
    }
    bool has(const std::string& k) const { return kv_.count(k) != 0; }
    template <class T>
    T get(const std::string& k) const
    {
        std::istringstream s(kv_.at(k));
        T                  v{};
        if (kv_.at(k) == "true") return static_cast<T>(1);
        if (kv_.at(k) == "false") return static_cast<T>(0);
        s >> v;
        return v;
    }
    template <class T>
    T getOrDefault(const std::string& k, const T& d) const
    {
        return has(k) ? get<T>(k) : d;
    }
    /** DECLARE_PARAMETER_REQ: std::invalid_argument if missing (Parameterizable.h:176-181). */
    template <class T>
    T required(const std::string& k) const
    {
        if (!has(k)) throw std::invalid_argument("Required parameter `" + k + "` not an existing key");
        return get<T>(k);
    }
    std::string getString(const std::string& k, const std::string& d) const { return has(k) ? kv_.at(k) : d; }

   private:
    std::map<std::string, std::string> kv_;
};

// ---- formula-capable parameters (mp2p_icp_map/include/mp2p_icp/Parameterizable.h, src/Parameterizable.cpp) ----
/** Arithmetic over named variables: numbers, identifiers, + - * / ^, unary minus, parentheses, and the
 *  functions abs, sqrt, min, max — the subset of MRPT's expression language (mrpt::expr, exprtk) that the
 *  reference's pipelines use (`threshold: "MATCH_THRESHOLD*2.0"`, tests/test-mp2p_matcher_pt2pt_parameterizable.cpp).
 *  Throws std::runtime_error on syntax errors and unknown variables. */
class Expression
{
   public:
    static double eval(const std::string& text, const std::map<std::string, double>& vars)
    {
        Expression e{text, vars, 0};
        const double v = e.sum();
        e.skip();
        if (e.pos_ != text.size()) throw std::runtime_error("expression: unexpected `" + text.substr(e.pos_) + "` in `" + text + "`");
        return v;
    }

   private:
    const std::string&                   s_;
    const std::map<std::string, double>& vars_;
    size_t                               pos_;
    void skip()
    {
        while (pos_ < s_.size() && std::isspace((unsigned char)s_[pos_])) pos_++;
    }
    bool eat(char c)
    {
        skip();
        if (pos_ < s_.size() && s_[pos_] == c) return pos_++, true;
        return false;
    }
    double sum()
    {
        double v = product();
        for (;;)
        {
            if (eat('+'))
                v += product();
            else if (eat('-'))
                v -= product();
            else
                return v;
        }
    }
    double product()
    {
        double v = power();
        for (;;)
        {
            if (eat('*'))
                v *= power();
            else if (eat('/'))
                v /= power();
            else
                return v;
        }
    }
    double power()
    {
        const double b = unary();
        return eat('^') ? std::pow(b, power()) : b;
    }
    double unary()
    {
        if (eat('-')) return -unary();
        if (eat('+')) return unary();
        return atom();
    }
    double atom()
    {
        skip();
        if (eat('('))
        {
            const double v = sum();
            if (!eat(')')) throw std::runtime_error("expression: missing `)` in `" + s_ + "`");
            return v;
        }
        if (pos_ < s_.size() && (std::isdigit((unsigned char)s_[pos_]) || s_[pos_] == '.'))
        {
            size_t       used = 0;
            const double v    = std::stod(s_.substr(pos_), &used);
            pos_ += used;
            return v;
        }
        if (pos_ < s_.size() && (std::isalpha((unsigned char)s_[pos_]) || s_[pos_] == '_'))
        {
            const size_t b = pos_;
            while (pos_ < s_.size() && (std::isalnum((unsigned char)s_[pos_]) || s_[pos_] == '_')) pos_++;
            const std::string name = s_.substr(b, pos_ - b);
            if (name == "true") return 1.0;
            if (name == "false") return 0.0;
            if (eat('('))
            {
                const double a = sum();
                double       r = 0;
                if (name == "abs" || name == "sqrt")
                    r = name == "abs" ? std::fabs(a) : std::sqrt(a);
                else if (name == "min" || name == "max")
                {
                    if (!eat(',')) throw std::runtime_error("expression: `" + name + "` takes two arguments");
                    const double c = sum();
                    r              = name == "min" ? std::min(a, c) : std::max(a, c);
                }
                else
                    throw std::runtime_error("expression: unknown function `" + name + "`");
                if (!eat(')')) throw std::runtime_error("expression: missing `)` in `" + s_ + "`");
                return r;
            }
            const auto it = vars_.find(name);
            if (it == vars_.end()) throw std::runtime_error("expression: unknown variable `" + name + "`");
            return it->second;
        }
        throw std::runtime_error("expression: cannot parse `" + s_ + "`");
    }
    Expression(const std::string& s, const std::map<std::string, double>& v, size_t p) : s_(s), vars_(v), pos_(p) {}
};

class ParameterSource;

/** Parameterizable (Parameterizable.h:110-190): parameters declared from YAML text that may be formulas over
 *  variables; constant ones are evaluated at declaration, the others whenever the attached ParameterSource
 *  realizes (Parameterizable.cpp:107-140, :46-105). */
class Parameterizable
{
   public:
    struct Declared
    {
        std::string expression;
        double*     target_d = nullptr;
        uint32_t*   target_u = nullptr;
        bool        is_constant = false, has_been_evaluated = false;
        void        store(double v) const
        {
            if (target_d) *target_d = v;
            if (target_u) *target_u = static_cast<uint32_t>(v);
        }
    };
    /** checkAllParametersAreRealized (Parameterizable.cpp:142-154) */
    void checkAllParametersAreRealized() const
    {
        for (const auto& d : declared_)
            if (!d.has_been_evaluated)
                throw std::runtime_error("Parameter `" + d.expression + "` was not realized: attach the object to a ParameterSource and call realize()");
    }
    std::deque<Declared>& declaredParameters() { return declared_; }
    void                  unrealizeParameters() { declared_.clear(); }

   protected:
    /** DECLARE_PARAMETER_REQ / _OPT (Parameterizable.h:176-190) */
    template <class T>
    void declareParameter(const ParameterMap& p, const std::string& name, T& target, bool required)
    {
        if (!p.has(name))
        {
            if (required) throw std::invalid_argument("Required parameter `" + name + "` not an existing key");
            return;
        }
        Declared& d  = declared_.emplace_back();
        d.expression = p.getString(name, "");
        if constexpr (std::is_same_v<T, double>)
            d.target_d = &target;
        else
            d.target_u = &target;
        try
        {
            d.store(Expression::eval(d.expression, {}));
            d.is_constant = d.has_been_evaluated = true;
        }
        catch (const std::exception&)
        {
            // needs variables that are not defined yet: evaluated by ParameterSource::realize()
        }
    }

   private:
    std::deque<Declared> declared_;
};

/** ParameterSource (Parameterizable.h:46-96, Parameterizable.cpp:22-105) */
class ParameterSource
{
   public:
    void attach(Parameterizable& obj)
    {
        for (auto& d : obj.declaredParameters()) attached_.push_back(&d);
    }
    void updateVariable(const std::string& variable, double value) { variables_[variable] = value; }
    void realize()
    {
        for (auto* d : attached_)
        {
            if (d->is_constant) continue;
            d->store(Expression::eval(d->expression, variables_));  // throws on variables that are still unknown
            d->has_been_evaluated = true;
        }
    }
    std::map<std::string, double> getVariableValues() const { return variables_; }

   private:
    std::vector<Parameterizable::Declared*> attached_;
    std::map<std::string, double>           variables_;
};

// ---- device context + map cache ------------------------------------------------------------
class Device
{
   public:
    /** One Device (context + caches) per GPU, created on first use. */
    static Device& instance(int device = 0)
    {
        static std::map<int, std::unique_ptr<Device>> all;
        auto&                                         d = all[device];
        if (!d) d.reset(new Device(device));
        return *d;
    }
    mp2p_b200_ctx* ctx() { return ctx_; }
    /** nn_prepare_for_3d_queries(): index of a global layer, rebuilt when the layer was modified. */
    mp2p_b200_map* map_for(const CPointsMap& layer)
    {
        Entry& e = cache_[&layer];
        if (!e.map || e.stamp != layer.stamp() || e.n != layer.size())
        {
            if (e.map) mp2p_b200_map_destroy(e.map);
            e.map = nullptr;
            check(mp2p_b200_map_create(ctx_, layer.getPointsBufferRef_x().data(), layer.getPointsBufferRef_y().data(),
                                       layer.getPointsBufferRef_z().data(), layer.size(), 0, &e.map),
                  "mp2p_b200_map_create");
            e.stamp = layer.stamp(), e.n = layer.size();
        }
        return e.map;
    }
    /** Device-resident copy of a LOCAL layer (mp2p_b200_cloud: caller order + Morton-sorted copy),
     *  cached like the index of a global layer: ICP::align() matches the same, unmodified local
     *  layer at every iteration (ICP.cpp:123-308), so it crosses PCIe once per align(). */
    const float* cloud_for(const CPointsMap& layer)
    {
        Entry& e = cache_[&layer];
        if (!e.cloud || e.cloud_stamp != layer.stamp() || e.cloud_n != layer.size())
        {
            if (e.cloud) mp2p_b200_cloud_destroy(e.cloud);
            e.cloud = nullptr;
            check(mp2p_b200_cloud_create(ctx_, layer.getPointsBufferRef_x().data(), layer.getPointsBufferRef_y().data(),
                                         layer.getPointsBufferRef_z().data(), layer.size(), 0, &e.cloud),
                  "mp2p_b200_cloud_create");
            e.cloud_stamp = layer.stamp(), e.cloud_n = layer.size();
        }
        return reinterpret_cast<const float*>(e.cloud);  // passed as `lx` with MP2P_B200_LOCAL_CLOUD
    }
    /** Witness of the pairings a matcher call just returned to the host: the count and 8 sample
     *  records. A solver that is handed a list with the same count and the same samples takes it
     *  for that output (nothing modifies the Pairings between run_matchers and run_solvers,
     *  ICP.cpp:143-170) and lets the library read its device-resident copy
     *  (MP2P_B200_PAIRS_LAST_MATCH) instead of uploading the records again. */
    template <class Rec>
    void note_match_output(const Rec* recs, size_t n)
    {
        Witness& w = sizeof(Rec) == sizeof(mp2p_b200_pair_pt2pt) ? wit2p_ : wit2l_;
        w.n        = n;
        for (size_t k = 0; k < 8 && n; k++) std::memcpy(w.sample[k], &recs[k * (n - 1) / 7], sizeof(Rec));
    }
    template <class Rec>
    bool is_last_match_output(const Rec* recs, size_t n) const
    {
        const Witness& w = sizeof(Rec) == sizeof(mp2p_b200_pair_pt2pt) ? wit2p_ : wit2l_;
        if (!n || w.n != n) return false;
        for (size_t k = 0; k < 8; k++)
            if (std::memcmp(w.sample[k], &recs[k * (n - 1) / 7], sizeof(Rec)) != 0) return false;
        return true;
    }
    void forget(const CPointsMap& layer)
    {
        auto it = cache_.find(&layer);
        if (it == cache_.end()) return;
        if (it->second.map) mp2p_b200_map_destroy(it->second.map);
        if (it->second.cloud) mp2p_b200_cloud_destroy(it->second.cloud);
        cache_.erase(it);
    }
    ~Device()
    {
        for (auto& kv : cache_)
        {
            if (kv.second.map) mp2p_b200_map_destroy(kv.second.map);
            if (kv.second.cloud) mp2p_b200_cloud_destroy(kv.second.cloud);
        }
        mp2p_b200_ctx_destroy(ctx_);
    }

   private:
    explicit Device(int device) { check(mp2p_b200_ctx_create(device, nullptr, &ctx_), "mp2p_b200_ctx_create"); }
    struct Entry
    {
        mp2p_b200_map*   map   = nullptr;
        uint64_t         stamp = 0;
        size_t           n     = 0;
        mp2p_b200_cloud* cloud = nullptr;
        uint64_t         cloud_stamp = 0;
        size_t           cloud_n     = 0;
    };
    struct Witness
    {
        size_t        n = ~size_t(0);
        unsigned char sample[8][sizeof(mp2p_b200_pair_pt2pl)];
    };
    mp2p_b200_ctx*                      ctx_ = nullptr;
    std::map<const CPointsMap*, Entry>  cache_;
    Witness                             wit2p_, wit2l_;
};

// ---- Matcher hierarchy -----------------------------------------------------------------------
class Matcher : public Parameterizable
{
   public:
    using Ptr          = std::shared_ptr<Matcher>;
    virtual ~Matcher() = default;
    virtual void initialize(const ParameterMap& params)  // Matcher.cpp:28-33
    {
        runFromIteration = params.getOrDefault<uint32_t>("runFromIteration", runFromIteration);
        runUpToIteration = params.getOrDefault<uint32_t>("runUpToIteration", runUpToIteration);
        enabled          = params.getOrDefault<int>("enabled", enabled ? 1 : 0) != 0;
    }
    /** Matcher.cpp:35-44; returns false if the matcher did not run. */
    bool match(const metric_map_t& pcGlobal, const metric_map_t& pcLocal, const CPose3D& localPose,
               const MatchContext& mc, MatchState& ms, Pairings& out) const
    {
        if (!enabled) return false;
        if (mc.icpIteration < runFromIteration) return false;
        if (runUpToIteration > 0 && mc.icpIteration > runUpToIteration) return false;
        return impl_match(pcGlobal, pcLocal, localPose, mc, ms, out);
    }
    uint32_t runFromIteration = 0, runUpToIteration = 0;
    bool     enabled = true;

   protected:
    virtual bool impl_match(const metric_map_t& pcGlobal, const metric_map_t& pcLocal, const CPose3D& localPose,
                            const MatchContext& mc, MatchState& ms, Pairings& out) const = 0;
};
using matcher_list_t = std::vector<Matcher::Ptr>;

class Matcher_Points_Base : public Matcher
{
   public:
    /** w["globalLayer"]["localLayer"] = weight (Matcher_Points_Base.h:45-53) */
    std::map<std::string, std::map<std::string, double>> weight_pt2pt_layers;
    bool   allowMatchAlreadyMatchedPoints_       = false;
    bool   allowMatchAlreadyMatchedGlobalPoints_ = false;
    double bounding_box_intersection_check_epsilon_ = 0.20;

    void initialize(const ParameterMap& params) override  // Matcher_Points_Base.cpp:132-181
    {
        Matcher::initialize(params);
        if (params.getOrDefault<uint64_t>("maxLocalPointsPerLayer", 0) != 0)
            throw std::invalid_argument("maxLocalPointsPerLayer != 0 is not offered (reference bug Q1/Q2, see DESIGN.md)");
        allowMatchAlreadyMatchedPoints_ = params.getOrDefault<int>("allowMatchAlreadyMatchedPoints", 0) != 0;
        allowMatchAlreadyMatchedGlobalPoints_ = params.getOrDefault<int>("allowMatchAlreadyMatchedGlobalPoints", 0) != 0;
        bounding_box_intersection_check_epsilon_ =
            params.getOrDefault<double>("bounding_box_intersection_check_epsilon", bounding_box_intersection_check_epsilon_);
    }

   protected:
    bool impl_match(const metric_map_t& pcGlobal, const metric_map_t& pcLocal, const CPose3D& localPose,
                    const MatchContext&, MatchState& ms, Pairings& out) const final  // Matcher_Points_Base.cpp:30-130
    {
        out = Pairings();
        for (const auto& glKV : pcGlobal.layers)
        {
            std::map<std::string, std::optional<double>> localLayers;
            if (!weight_pt2pt_layers.empty())
            {
                const auto it = weight_pt2pt_layers.find(glKV.first);
                if (it == weight_pt2pt_layers.end()) continue;
                for (const auto& kv : it->second) localLayers[kv.first] = kv.second;
            }
            else
                localLayers[glKV.first] = {};
            for (const auto& lw : localLayers)
            {
                const auto itLocal = pcLocal.layers.find(lw.first);
                if (itLocal == pcLocal.layers.end())
                {
                    if (!lw.second.has_value()) continue;
                    throw std::runtime_error("Local pointcloud layer '" + lw.first + "' not found matching global layer '" + glKV.first + "'");
                }
                if (!glKV.second || !itLocal->second) throw std::runtime_error("null layer");
                const size_t nBefore = out.paired_pt2pt.size();
                implMatchOneLayer(*glKV.second, *itLocal->second, localPose, ms, glKV.first, lw.first, out);
                const size_t nAfter = out.paired_pt2pt.size();
                if (lw.second.has_value() && nAfter != nBefore) out.point_weights.emplace_back(nAfter - nBefore, *lw.second);
            }
        }
        return true;
    }

   private:
    virtual void implMatchOneLayer(const CPointsMap& pcGlobal, const CPointsMap& pcLocal, const CPose3D& localPose,
                                   MatchState& ms, const layer_name_t& globalName, const layer_name_t& localName,
                                   Pairings& out) const = 0;
};

class Matcher_Points_DistanceThreshold : public Matcher_Points_Base
{
   public:
    double   threshold           = 0.50;
    double   thresholdAngularDeg = 0.50;
    uint32_t pairingsPerPoint    = 1;
    void     initialize(const ParameterMap& params) override  // …DistanceThreshold.cpp:39-46
    {
        Matcher_Points_Base::initialize(params);
        unrealizeParameters();
        declareParameter(params, "threshold", threshold, true);  // DECLARE_PARAMETER_REQ: may be formulas
        declareParameter(params, "thresholdAngularDeg", thresholdAngularDeg, true);
        declareParameter(params, "pairingsPerPoint", pairingsPerPoint, false);
    }

   private:
    void implMatchOneLayer(const CPointsMap& pcGlobal, const CPointsMap& pcLocal, const CPose3D& localPose,
                           MatchState& ms, const layer_name_t& globalName, const layer_name_t& localName,
                           Pairings& out) const override
    {
        checkAllParametersAreRealized();                                                                   // :56
        if (!(pairingsPerPoint >= 1)) throw std::runtime_error("Assert failed: pairingsPerPoint >= 1");  // :57-59
        if (!(threshold > .0)) throw std::runtime_error("Assert failed: threshold > 0");
        if (!(thresholdAngularDeg >= .0)) throw std::runtime_error("Assert failed: thresholdAngularDeg >= 0");
        Device&                dev = Device::instance();
        mp2p_b200_pt2pt_params p{threshold, thresholdAngularDeg, pairingsPerPoint, allowMatchAlreadyMatchedPoints_,
                                 allowMatchAlreadyMatchedGlobalPoints_, bounding_box_intersection_check_epsilon_};
        auto&        lbits  = ms.localPaired.at(localName);
        auto&        gbits  = ms.globalPaired.at(globalName);
        const size_t before = out.paired_pt2pt.size(), cap = pcLocal.size() * pairingsPerPoint;
        out.paired_pt2pt.resize(before + cap);
        uint64_t cnt = 0, pot = 0;
        check(mp2p_b200_match_pt2pt(dev.ctx(), dev.map_for(pcGlobal), dev.cloud_for(pcLocal), nullptr, nullptr,
                                    pcLocal.size(), MP2P_B200_LOCAL_CLOUD, localPose.m, &p, lbits.data(), gbits.data(),
                                    out.paired_pt2pt.data() + before, cap, 0, &cnt, &pot),
              "mp2p_b200_match_pt2pt");
        out.paired_pt2pt.resize(before + cnt);
        out.potential_pairings += pot;
        dev.note_match_output(out.paired_pt2pt.data() + before, before == 0 ? cnt : 0);
        if (!allowMatchAlreadyMatchedGlobalPoints_)  // lambdaAddPair :116-120
            for (size_t i = before; i < out.paired_pt2pt.size(); i++)
            {
                MatchState::mark(lbits, out.paired_pt2pt[i].localIdx);
                MatchState::mark(gbits, out.paired_pt2pt[i].globalIdx);
            }
    }
};

/** Matcher_Points_InlierRatio (mp2p_icp/include/mp2p_icp/Matcher_Points_InlierRatio.h,
 *  mp2p_icp/src/Matcher_Points_InlierRatio.cpp:35-143) over mp2p_b200_match_inlier_ratio. */
class Matcher_Points_InlierRatio : public Matcher_Points_Base
{
   public:
    double inliersRatio = 0.80;
    void   initialize(const ParameterMap& params) override  // :35-39
    {
        Matcher_Points_Base::initialize(params);
        inliersRatio = params.required<double>("inliersRatio");
    }

   private:
    void implMatchOneLayer(const CPointsMap& pcGlobal, const CPointsMap& pcLocal, const CPose3D& localPose,
                           MatchState& ms, const layer_name_t& globalName, const layer_name_t& localName,
                           Pairings& out) const override
    {
        if (!(inliersRatio > 0.0)) throw std::runtime_error("Assert failed: inliersRatio > 0");  // :49-50
        if (!(inliersRatio < 1.0)) throw std::runtime_error("Assert failed: inliersRatio < 1");
        Device&                       dev = Device::instance();
        mp2p_b200_inlier_ratio_params p{inliersRatio, allowMatchAlreadyMatchedPoints_, allowMatchAlreadyMatchedGlobalPoints_,
                                        bounding_box_intersection_check_epsilon_};
        auto&        lbits  = ms.localPaired.at(localName);
        auto&        gbits  = ms.globalPaired.at(globalName);
        const size_t before = out.paired_pt2pt.size(), cap = pcLocal.size();
        out.paired_pt2pt.resize(before + cap);
        uint64_t cnt = 0, pot = 0;
        check(mp2p_b200_match_inlier_ratio(dev.ctx(), dev.map_for(pcGlobal), dev.cloud_for(pcLocal), nullptr, nullptr,
                                           pcLocal.size(), MP2P_B200_LOCAL_CLOUD, localPose.m, &p, lbits.data(), gbits.data(),
                                           out.paired_pt2pt.data() + before, cap, 0, &cnt, &pot),
              "mp2p_b200_match_inlier_ratio");
        out.paired_pt2pt.resize(before + cnt);
        out.potential_pairings += pot;
        dev.note_match_output(out.paired_pt2pt.data() + before, before == 0 ? cnt : 0);
        for (size_t i = before; i < out.paired_pt2pt.size(); i++)  // :133-135 (unconditional in this matcher)
        {
            MatchState::mark(lbits, out.paired_pt2pt[i].localIdx);
            MatchState::mark(gbits, out.paired_pt2pt[i].globalIdx);
        }
    }
};

class Matcher_Point2Plane : public Matcher_Points_Base
{
   public:
    double   distanceThreshold = 0.50, searchRadius = 1.0, planeEigenThreshold = 0.01;
    uint32_t knn = 5, minimumPlanePoints = 5;
    void     initialize(const ParameterMap& params) override  // Matcher_Point2Plane.cpp:35-39 (+ plane-fit params)
    {
        Matcher_Points_Base::initialize(params);
        distanceThreshold   = params.required<double>("distanceThreshold");
        searchRadius        = params.getOrDefault<double>("searchRadius", searchRadius);
        knn                 = params.getOrDefault<uint32_t>("knn", knn);
        minimumPlanePoints  = static_cast<uint32_t>(params.getOrDefault<double>("minimumPlanePoints", minimumPlanePoints));
        planeEigenThreshold = params.getOrDefault<double>("planeEigenThreshold", planeEigenThreshold);
    }

   private:
    void implMatchOneLayer(const CPointsMap& pcGlobal, const CPointsMap& pcLocal, const CPose3D& localPose,
                           MatchState& ms, const layer_name_t&, const layer_name_t& localName, Pairings& out) const override
    {
        Device&                dev = Device::instance();
        mp2p_b200_pt2pl_params p{distanceThreshold, searchRadius, knn, minimumPlanePoints, planeEigenThreshold,
                                 allowMatchAlreadyMatchedPoints_, bounding_box_intersection_check_epsilon_};
        auto&        lbits  = ms.localPaired.at(localName);
        const size_t before = out.paired_pt2pl.size(), cap = pcLocal.size();
        out.paired_pt2pl.resize(before + cap);
        uint64_t cnt = 0, pot = 0;
        check(mp2p_b200_match_pt2pl(dev.ctx(), dev.map_for(pcGlobal), dev.cloud_for(pcLocal), nullptr, nullptr,
                                    pcLocal.size(), MP2P_B200_LOCAL_CLOUD, localPose.m, &p, lbits.data(), out.paired_pt2pl.data() + before,
                                    cap, 0, &cnt, &pot),
              "mp2p_b200_match_pt2pl");
        out.paired_pt2pl.resize(before + cnt);
        out.potential_pairings += pot;
        dev.note_match_output(out.paired_pt2pl.data() + before, before == 0 ? cnt : 0);
        // Matcher_Point2Plane.cpp:109 — the local point is marked; which one it was is recoverable
        // from pt_local only through the coordinates, so re-identify by a parallel walk (the output
        // is in ascending local index and each local point pairs at most once).
        const auto &lx = pcLocal.getPointsBufferRef_x(), &ly = pcLocal.getPointsBufferRef_y(),
                   &lz = pcLocal.getPointsBufferRef_z();
        size_t i = 0;
        for (size_t k = before; k < out.paired_pt2pl.size(); k++)
        {
            const auto& r = out.paired_pt2pl[k];
            while (i < lx.size() && !(lx[i] == r.local_x && ly[i] == r.local_y && lz[i] == r.local_z &&
                                      !((lbits[i >> 5] >> (i & 31)) & 1u)))
                i++;
            if (i < lx.size()) MatchState::mark(lbits, i++);
        }
    }
};

/** Matcher_Point2Line (mp2p_icp/include/mp2p_icp/Matcher_Point2Line.h:38-70,
 *  mp2p_icp/src/Matcher_Point2Line.cpp:35-163) over mp2p_b200_match_pt2ln. */
class Matcher_Point2Line : public Matcher_Points_Base
{
   public:
    double   distanceThreshold  = 0.50;
    uint32_t knn                = 4;
    uint32_t minimumLinePoints  = 4;
    double   lineEigenThreshold = 0.01;
    void     initialize(const ParameterMap& params) override  // :35-45
    {
        Matcher_Points_Base::initialize(params);
        distanceThreshold  = params.required<double>("distanceThreshold");
        knn                = params.required<uint32_t>("knn");
        lineEigenThreshold = params.required<double>("lineEigenThreshold");
        minimumLinePoints  = params.required<uint32_t>("minimumLinePoints");
        if (!(minimumLinePoints >= 2)) throw std::runtime_error("Assert failed: minimumLinePoints >= 2");
    }

   private:
    void implMatchOneLayer(const CPointsMap& pcGlobal, const CPointsMap& pcLocal, const CPose3D& localPose,
                           MatchState& ms, const layer_name_t&, const layer_name_t& localName, Pairings& out) const override
    {
        Device&                dev = Device::instance();
        mp2p_b200_pt2ln_params p{distanceThreshold, knn, minimumLinePoints, lineEigenThreshold,
                                 allowMatchAlreadyMatchedPoints_, bounding_box_intersection_check_epsilon_};
        auto&        lbits  = ms.localPaired.at(localName);
        const size_t before = out.paired_pt2ln.size(), cap = pcLocal.size();
        out.paired_pt2ln.resize(before + cap);
        uint64_t cnt = 0, pot = 0;
        check(mp2p_b200_match_pt2ln(dev.ctx(), dev.map_for(pcGlobal), dev.cloud_for(pcLocal), nullptr, nullptr, pcLocal.size(),
                                    MP2P_B200_LOCAL_CLOUD, localPose.m, &p, lbits.data(), out.paired_pt2ln.data() + before, cap, 0,
                                    &cnt, &pot),
              "mp2p_b200_match_pt2ln");
        out.paired_pt2ln.resize(before + cnt);
        out.potential_pairings += pot;
        // :159 — the local point is marked; re-identified by a parallel walk (ascending local index)
        const auto &lx = pcLocal.getPointsBufferRef_x(), &ly = pcLocal.getPointsBufferRef_y(), &lz = pcLocal.getPointsBufferRef_z();
        size_t      i  = 0;
        for (size_t k = before; k < out.paired_pt2ln.size(); k++)
        {
            const auto& r = out.paired_pt2ln[k];
            while (i < lx.size() && !((double)lx[i] == r.local[0] && (double)ly[i] == r.local[1] && (double)lz[i] == r.local[2] &&
                                      !((lbits[i >> 5] >> (i & 31)) & 1u)))
                i++;
            if (i < lx.size()) MatchState::mark(lbits, i++);
        }
    }
};

/** Matcher_Adaptive (mp2p_icp/include/mp2p_icp/Matcher_Adaptive.h:39-98, mp2p_icp/src/Matcher_Adaptive.cpp:32-314)
 *  over mp2p_b200_match_adaptive. The histogram -> confidence-interval step uses the library's
 *  restatement of the two MRPT helpers (parity unpinned); the MRPT plugin calls MRPT. */
class Matcher_Adaptive : public Matcher_Points_Base
{
   public:
    double   confidenceInterval        = 0.80;
    double   firstToSecondDistanceMax  = 1.2;
    double   absoluteMaxSearchDistance = 5.0;
    bool     enableDetectPlanes        = false;
    uint32_t maxPt2PtCorrespondences   = 1;
    uint32_t planeSearchPoints         = 8;
    uint32_t planeMinimumFoundPoints   = 4;
    double   planeMinimumDistance      = 0.10;
    double   planeEigenThreshold       = 0.01;
    double   minimumCorrDist           = 0.1;
    void     initialize(const ParameterMap& params) override  // :32-57
    {
        Matcher_Points_Base::initialize(params);
        confidenceInterval        = params.required<double>("confidenceInterval");
        firstToSecondDistanceMax  = params.required<double>("firstToSecondDistanceMax");
        absoluteMaxSearchDistance = params.required<double>("absoluteMaxSearchDistance");
        minimumCorrDist           = params.getOrDefault<double>("minimumCorrDist", minimumCorrDist);
        enableDetectPlanes        = params.required<int>("enableDetectPlanes") != 0;
        planeSearchPoints         = params.getOrDefault<uint32_t>("planeSearchPoints", planeSearchPoints);
        planeMinimumFoundPoints   = params.getOrDefault<uint32_t>("planeMinimumFoundPoints", planeMinimumFoundPoints);
        planeEigenThreshold       = params.getOrDefault<double>("planeEigenThreshold", planeEigenThreshold);
        maxPt2PtCorrespondences   = params.getOrDefault<uint32_t>("maxPt2PtCorrespondences", maxPt2PtCorrespondences);
        planeMinimumDistance      = params.getOrDefault<double>("planeMinimumDistance", planeMinimumDistance);
        if (!(confidenceInterval < 1.0)) throw std::runtime_error("Assert failed: confidenceInterval < 1.0");
        if (!(confidenceInterval > 0.0)) throw std::runtime_error("Assert failed: confidenceInterval > 0.0");
        if (!(planeSearchPoints >= planeMinimumFoundPoints)) throw std::runtime_error("Assert failed: planeSearchPoints >= planeMinimumFoundPoints");
        if (!(planeMinimumFoundPoints >= 3)) throw std::runtime_error("Assert failed: planeMinimumFoundPoints >= 3");
        if (!(planeEigenThreshold > 0.0)) throw std::runtime_error("Assert failed: planeEigenThreshold > 0.0");
    }

   private:
    void implMatchOneLayer(const CPointsMap& pcGlobal, const CPointsMap& pcLocal, const CPose3D& localPose,
                           MatchState& ms, const layer_name_t& globalName, const layer_name_t& localName, Pairings& out) const override
    {
        Device&                   dev = Device::instance();
        mp2p_b200_adaptive_params p{confidenceInterval, firstToSecondDistanceMax, absoluteMaxSearchDistance, minimumCorrDist,
                                    enableDetectPlanes, planeSearchPoints, planeMinimumFoundPoints, maxPt2PtCorrespondences,
                                    planeEigenThreshold, planeMinimumDistance, allowMatchAlreadyMatchedPoints_,
                                    allowMatchAlreadyMatchedGlobalPoints_, bounding_box_intersection_check_epsilon_};
        auto&        lbits = ms.localPaired.at(localName);
        auto&        gbits = ms.globalPaired.at(globalName);
        const size_t b2p = out.paired_pt2pt.size(), b2l = out.paired_pt2pl.size();
        const size_t cap2p = pcLocal.size() * maxPt2PtCorrespondences, cap2l = pcLocal.size();
        out.paired_pt2pt.resize(b2p + cap2p), out.paired_pt2pl.resize(b2l + cap2l);
        uint64_t n2p = 0, n2l = 0, pot = 0;
        check(mp2p_b200_match_adaptive(dev.ctx(), dev.map_for(pcGlobal), dev.cloud_for(pcLocal), nullptr, nullptr, pcLocal.size(),
                                       MP2P_B200_LOCAL_CLOUD, localPose.m, &p, lbits.data(), gbits.data(), out.paired_pt2pt.data() + b2p,
                                       cap2p, out.paired_pt2pl.data() + b2l, cap2l, 0, &n2p, &n2l, nullptr, &pot),
              "mp2p_b200_match_adaptive");
        out.paired_pt2pt.resize(b2p + n2p), out.paired_pt2pl.resize(b2l + n2l);
        out.potential_pairings += pot;
        if (!allowMatchAlreadyMatchedGlobalPoints_)  // :291-295 (the local bit only; global points are never marked, :303-311)
            for (size_t i = b2p; i < out.paired_pt2pt.size(); i++) MatchState::mark(lbits, out.paired_pt2pt[i].localIdx);
        // :262 — local points that got a plane; re-identified by a parallel walk (ascending local index)
        const auto &lx = pcLocal.getPointsBufferRef_x(), &ly = pcLocal.getPointsBufferRef_y(), &lz = pcLocal.getPointsBufferRef_z();
        size_t      i  = 0;
        for (size_t k = b2l; k < out.paired_pt2pl.size(); k++)
        {
            const auto& r = out.paired_pt2pl[k];
            while (i < lx.size() && !(lx[i] == r.local_x && ly[i] == r.local_y && lz[i] == r.local_z && !((lbits[i >> 5] >> (i & 31)) & 1u)))
                i++;
            if (i < lx.size()) MatchState::mark(lbits, i++);
        }
    }
};

/** run_matchers (Matcher.cpp:46-88) */
inline Pairings run_matchers(const matcher_list_t& matchers, const metric_map_t& pcGlobal, const metric_map_t& pcLocal,
                             const CPose3D& local_wrt_global, const MatchContext& mc, MatchState* userMS = nullptr)
{
    Pairings                  pairings;
    std::optional<MatchState> localMS;
    MatchState*               ms = userMS;
    if (!ms)
    {
        localMS.emplace(pcGlobal, pcLocal);
        ms = &*localMS;
    }
    for (const auto& m : matchers)
    {
        if (!m) throw std::runtime_error("null matcher");
        Pairings pc;
        m->match(pcGlobal, pcLocal, local_wrt_global, mc, *ms, pc);
        pairings.push_back(pc);
    }
    return pairings;
}

// ---- Solver hierarchy -------------------------------------------------------------------------
struct OptimalTF_Result
{
    CPose3D optimalPose;
};
struct SolverContext
{
    std::optional<uint32_t> icpIteration;
    std::optional<CPose3D>  guessRelativePose;
    std::optional<CPose3D>  lastIcpStepIncrement;
    mutable std::map<const void*, bool> perSolverFinished;  // perSolverPersistentData["finished"]
};

class Solver
{
   public:
    using Ptr         = std::shared_ptr<Solver>;
    virtual ~Solver() = default;
    virtual void initialize(const ParameterMap& p)  // Solver.cpp:28-34
    {
        runFromIteration = p.getOrDefault<uint32_t>("runFromIteration", runFromIteration);
        runUpToIteration = p.getOrDefault<uint32_t>("runUpToIteration", runUpToIteration);
        enabled          = p.getOrDefault<int>("enabled", enabled ? 1 : 0) != 0;
        runUntilTranslationCorrectionSmallerThan =
            p.getOrDefault<double>("runUntilTranslationCorrectionSmallerThan", runUntilTranslationCorrectionSmallerThan);
    }
    bool optimal_pose(const Pairings& pairings, OptimalTF_Result& out, const SolverContext& sc) const  // Solver.cpp:36-64
    {
        if (!enabled) return false;
        if (sc.icpIteration && *sc.icpIteration < runFromIteration) return false;
        if (sc.icpIteration && runUpToIteration > 0 && *sc.icpIteration > runUpToIteration) return false;
        if (runUntilTranslationCorrectionSmallerThan > 0)
        {
            if (sc.perSolverFinished.count(this)) return false;
            if (sc.lastIcpStepIncrement)
            {
                const auto&  d = sc.lastIcpStepIncrement->m;
                const double n = std::sqrt(d[3] * d[3] + d[7] * d[7] + d[11] * d[11]);
                if (n < runUntilTranslationCorrectionSmallerThan)
                {
                    sc.perSolverFinished[this] = true;
                    return false;
                }
            }
        }
        return impl_optimal_pose(pairings, out, sc);
    }
    uint32_t runFromIteration = 0, runUpToIteration = 0;
    bool     enabled = true;
    double   runUntilTranslationCorrectionSmallerThan = 0;

   protected:
    virtual bool impl_optimal_pose(const Pairings& pairings, OptimalTF_Result& out, const SolverContext& sc) const = 0;
};
using solver_list_t = std::vector<Solver::Ptr>;

inline int robust_kernel_from_string(const std::string& s)
{
    if (s == "None" || s == "RobustKernel::None") return 0;
    if (s == "GemanMcClure" || s == "RobustKernel::GemanMcClure") return 1;
    if (s == "Cauchy" || s == "RobustKernel::Cauchy") return 2;
    throw std::invalid_argument("Unknown kernel type");  // robust_kernels.h:91
}

class Solver_Horn : public Solver
{
   public:
    mp2p_b200_horn_params pairingsWeightParameters{0, 1.20, 1.0, 0, 1.0, {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0}};
    void                  initialize(const ParameterMap& p) override  // Solver_Horn.cpp:33-39
    {
        Solver::initialize(p);
        auto& w                      = pairingsWeightParameters;
        w.use_scale_outlier_detector = p.getOrDefault<int>("use_scale_outlier_detector", w.use_scale_outlier_detector);
        w.scale_outlier_threshold    = p.getOrDefault<double>("scale_outlier_threshold", w.scale_outlier_threshold);
        w.robust_kernel              = robust_kernel_from_string(p.getString("robust_kernel", "None"));
        w.robust_kernel_param        = p.getOrDefault<double>("robust_kernel_param", w.robust_kernel_param);
        assumeUnmodifiedPairings     = p.getOrDefault<int>("assumeUnmodifiedPairings", 0) != 0;
    }
    /** Opt-in (default off): a list whose count and 8 sampled records equal the last matcher output is
     *  solved from the copy that call left on the device instead of being uploaded again. Only for
     *  pipelines in which nothing edits the Pairings between run_matchers and run_solvers. */
    bool assumeUnmodifiedPairings = false;

   protected:
    bool impl_optimal_pose(const Pairings& pairings, OptimalTF_Result& out, const SolverContext& sc) const override
    {
        out = OptimalTF_Result();
        mp2p_b200_horn_params prm = pairingsWeightParameters;
        if (prm.robust_kernel != 0)
        {
            if (!sc.guessRelativePose) throw std::runtime_error("Assert failed: currentEstimateForRobust.has_value()");
            std::memcpy(prm.currentEstimateForRobust, sc.guessRelativePose->m, sizeof(prm.currentEstimateForRobust));
        }
        std::vector<uint64_t> wc;
        std::vector<double>   wv;
        for (const auto& b : pairings.point_weights) wc.push_back(b.first), wv.push_back(b.second);
        int32_t solved = 0;
        Device&   dev    = Device::instance();
        if (!pairings.paired_pt2ln.empty())
            throw std::runtime_error("Solver_Horn over point-to-line pairings (pt2ln_pl_to_pt2pt.cpp:88-108, TLine3D::closestPointTo) "
                                     "is not offered: use Solver_GaussNewton");
        if (!pairings.paired_pt2pl.empty())
        {
            // Solver_Horn.cpp:51-55: pt2pl pairings are projected to pt2pt by pt2ln_pl_to_pt2pt and the
            // solve runs over the projected list ONLY (the reference's `out` starts empty there)
            if (!sc.guessRelativePose) throw std::runtime_error("Assert failed: sc.guessRelativePose.has_value()");
            const auto& l2l = pairings.paired_pt2pl;
            check(mp2p_b200_solve_horn_pt2pl(dev.ctx(), l2l.data(), l2l.size(),
                                             (assumeUnmodifiedPairings && dev.is_last_match_output(l2l.data(), l2l.size())) ? MP2P_B200_PAIRS_LAST_MATCH : 0,
                                             sc.guessRelativePose->m, &prm, out.optimalPose.m, &solved),
                  "mp2p_b200_solve_horn_pt2pl");
            return solved != 0;
        }
        const int origin = (assumeUnmodifiedPairings && dev.is_last_match_output(pairings.paired_pt2pt.data(), pairings.paired_pt2pt.size()))
                               ? MP2P_B200_PAIRS_LAST_MATCH
                               : 0;
        check(mp2p_b200_solve_horn(dev.ctx(), pairings.paired_pt2pt.data(), pairings.paired_pt2pt.size(), origin,
                                   &prm, wc.data(), wv.data(), wc.size(), out.optimalPose.m, &solved),
              "mp2p_b200_solve_horn");
        return solved != 0;
    }
};

class Solver_GaussNewton : public Solver
{
   public:
    uint32_t    maxIterations = 5;
    std::string robustKernel  = "None";
    double      robustKernelParam = 1.0, w_pt2pt = 1.0, w_pt2pl = 1.0, w_pt2ln = 1.0;
    void        initialize(const ParameterMap& p) override  // Solver_GaussNewton.cpp:29-40
    {
        Solver::initialize(p);
        maxIterations     = p.required<uint32_t>("maxIterations");
        robustKernel      = p.getString("robustKernel", robustKernel);
        robustKernelParam = p.getOrDefault<double>("robustKernelParam", robustKernelParam);
        robust_kernel_from_string(robustKernel);
        assumeUnmodifiedPairings = p.getOrDefault<int>("assumeUnmodifiedPairings", 0) != 0;
    }
    bool assumeUnmodifiedPairings = false;  //!< opt-in, see Solver_Horn

   protected:
    bool impl_optimal_pose(const Pairings& pairings, OptimalTF_Result& out, const SolverContext& sc) const override
    {
        out = OptimalTF_Result();
        if (!sc.guessRelativePose) throw std::runtime_error("Assert failed: sc.guessRelativePose.has_value()");  // :57
        mp2p_b200_gn_params prm{maxIterations, 1e-7, 0.0, w_pt2pt, w_pt2pl, robust_kernel_from_string(robustKernel), robustKernelParam};
        uint32_t            iters  = 0;
        int32_t             solved = 0;
        Device&    dev   = Device::instance();
        const auto &l2p = pairings.paired_pt2pt;
        const auto &l2l = pairings.paired_pt2pl;
        // Pairings::point_weights re-weight the pt2pt term block by block (optimal_tf_gauss_newton.cpp:118-128);
        // the device accumulation carries one uniform pt2pt weight: refuse instead of returning another pose
        // (the MRPT plugin hands such calls to the reference's own solver)
        if (!pairings.point_weights.empty() && !l2p.empty())
            throw std::runtime_error("Solver_GaussNewton over weighted point layers (Pairings::point_weights) is not offered on the device");
        // every non-empty list must be the witnessed output of the last matcher call of its kind
        if (!pairings.paired_pt2ln.empty())  // point-to-line term, optimal_tf_gauss_newton.cpp:182-203
        {
            check(mp2p_b200_solve_gauss_newton_ex(dev.ctx(), l2p.data(), l2p.size(), l2l.data(), l2l.size(),
                                                  pairings.paired_pt2ln.data(), pairings.paired_pt2ln.size(), 0, &prm, w_pt2ln,
                                                  sc.guessRelativePose->m, out.optimalPose.m, &iters, &solved),
                  "mp2p_b200_solve_gauss_newton_ex");
            return solved != 0;
        }
        const bool last  = assumeUnmodifiedPairings && (l2p.empty() || dev.is_last_match_output(l2p.data(), l2p.size())) &&
                          (l2l.empty() || dev.is_last_match_output(l2l.data(), l2l.size())) && !pairings.empty();
        check(mp2p_b200_solve_gauss_newton(dev.ctx(), pairings.paired_pt2pt.data(), pairings.paired_pt2pt.size(),
                                           pairings.paired_pt2pl.data(), pairings.paired_pt2pl.size(),
                                           last ? MP2P_B200_PAIRS_LAST_MATCH : 0, &prm,
                                           sc.guessRelativePose->m, out.optimalPose.m, &iters, &solved),
              "mp2p_b200_solve_gauss_newton");
        return solved != 0;
    }
};

// ---- the caller: ICP::align (ICP.cpp:108-308), restated so whole alignments can be run --------
// ------------------------------------------------------------------------------------------------
/** QualityEvaluator (mp2p_icp/include/mp2p_icp/QualityEvaluator.h). */
class QualityEvaluator
{
   public:
    using Ptr                   = std::shared_ptr<QualityEvaluator>;
    virtual ~QualityEvaluator() = default;
    struct Result
    {
        double quality      = 0;      // [0,1]
        bool   hard_discard = false;  // ICP::evaluate_quality then reports 0 (ICP.cpp:622-626)
    };
    virtual void   initialize(const ParameterMap& params) = 0;
    virtual Result evaluate(const metric_map_t& pcGlobal, const metric_map_t& pcLocal, const CPose3D& localPose,
                            const Pairings& pairingsFromICP) const = 0;
};

/** QualityEvaluator_PairedRatio (mp2p_icp/src/QualityEvaluator_PairedRatio.cpp:27-73): ratio of
 *  pairings over potential pairings, from the last ICP pairings (reuse_icp_pairings, the default) or
 *  from one more pass of a Matcher_Points_DistanceThreshold of its own that may pair a global point
 *  several times (:37-40). That extra pass runs on the device; only its COUNT matters (:66-67). */
class QualityEvaluator_PairedRatio : public QualityEvaluator
{
   public:
    void initialize(const ParameterMap& params) override  // :27-44
    {
        reuse_icp_pairings             = params.getOrDefault<int>("reuse_icp_pairings", reuse_icp_pairings ? 1 : 0) != 0;
        absolute_minimum_pairing_ratio = params.getOrDefault<double>("absolute_minimum_pairing_ratio", absolute_minimum_pairing_ratio);
        if (!reuse_icp_pairings)
        {
            ParameterMap p = params;
            if (!p.has("allowMatchAlreadyMatchedGlobalPoints")) p.set("allowMatchAlreadyMatchedGlobalPoints", 1);
            matcher_.initialize(p);
        }
    }
    Result evaluate(const metric_map_t& pcGlobal, const metric_map_t& pcLocal, const CPose3D& localPose,
                    const Pairings& pairingsFromICP) const override  // :46-73
    {
        const Pairings* pairings = &pairingsFromICP;
        Pairings        newPairings;
        if (!reuse_icp_pairings)
        {
            MatchState ms(pcGlobal, pcLocal);
            matcher_.match(pcGlobal, pcLocal, localPose, {}, ms, newPairings);
            pairings = &newPairings;
        }
        Result r;
        r.quality      = pairings->potential_pairings ? pairings->size() / double(pairings->potential_pairings) : .0;
        r.hard_discard = r.quality < absolute_minimum_pairing_ratio;
        return r;
    }
    bool   reuse_icp_pairings             = true;
    double absolute_minimum_pairing_ratio = 0.20;

   private:
    Matcher_Points_DistanceThreshold matcher_;
};

struct QualityEvaluatorEntry  // ICP.h quality_eval_list_t
{
    QualityEvaluator::Ptr obj;
    double                relativeWeight = 1.0;
};
using quality_eval_list_t = std::vector<QualityEvaluatorEntry>;

enum class IterTermReason
{
    Undefined,
    NoPairings,
    SolverError,
    MaxIterations,
    Stalled,
    QualityCheckpointFailed
};
struct Parameters  // Parameters.h:42-52
{
    uint32_t maxIterations    = 40;
    double   minAbsStep_trans = 5e-4, minAbsStep_rot = 1e-4;
    std::map<uint32_t, double> quality_checkpoints;  // iteration -> minimum quality (Parameters.h:61-73)
};
struct Results
{
    CPose3D        optimal_tf;
    uint32_t       nIterations       = 0;
    IterTermReason terminationReason = IterTermReason::Undefined;
    double         quality           = 0;  // Results.h:41
    Pairings       finalPairings;
};

class ICP
{
   public:
    matcher_list_t&      matchers() { return matchers_; }
    solver_list_t&       solvers() { return solvers_; }
    quality_eval_list_t& quality_evaluators() { return quality_evaluators_; }
    /** ICP::evaluate_quality, ICP.cpp:608-634 */
    static double evaluate_quality(const quality_eval_list_t& evaluators, const metric_map_t& pcGlobal,
                                   const metric_map_t& pcLocal, const CPose3D& localPose, const Pairings& finalPairings)
    {
        if (evaluators.empty()) throw std::runtime_error("Assert failed: !evaluators.empty()");
        double sumW = .0, sumEvals = .0;
        for (const auto& e : evaluators)
        {
            if (!(e.relativeWeight > 0)) throw std::runtime_error("Assert failed: relativeWeight > 0");
            const auto r = e.obj->evaluate(pcGlobal, pcLocal, localPose, finalPairings);
            if (r.hard_discard) return 0;
            sumEvals += e.relativeWeight * r.quality, sumW += e.relativeWeight;
        }
        return sumEvals / sumW;
    }
    void align(const metric_map_t& pcLocal, const metric_map_t& pcGlobal, const CPose3D& initialGuess,
               const Parameters& p, Results& result)
    {
        result           = Results();
        CPose3D current  = initialGuess, prev = initialGuess;
        std::optional<CPose3D> prev2, lastCorrection;
        SolverContext          sc;
        Pairings               pairings;
        result.terminationReason = IterTermReason::MaxIterations;
        for (result.nIterations = 0; result.nIterations < p.maxIterations; result.nIterations++)
        {
            MatchContext mc;
            mc.icpIteration = result.nIterations;
            pairings        = run_matchers(matchers_, pcGlobal, pcLocal, current, mc);  // :143
            if (pairings.empty())
            {
                result.terminationReason = IterTermReason::NoPairings;  // :148
                break;
            }
            sc.icpIteration = result.nIterations;
            sc.guessRelativePose    = current;
            sc.lastIcpStepIncrement = lastCorrection;
            OptimalTF_Result sol;
            bool             solvedOk = false;
            for (const auto& s : solvers_)  // run_solvers, ICP.cpp:469-479
                if (s->optimal_pose(pairings, sol, sc))
                {
                    solvedOk = true;
                    break;
                }
            if (!solvedOk)
            {
                result.terminationReason = IterTermReason::SolverError;
                break;
            }
            current            = sol.optimalPose;
            const CPose3D dSol = current - prev;  // :203
            lastCorrection     = dSol;
            double dxyz, drot;
            dSol.log_norms(dxyz, drot);
            if (prev2)  // :208-215
            {
                double d2x, d2r;
                (current - *prev2).log_norms(d2x, d2r);
                dxyz = std::min(dxyz, d2x), drot = std::min(drot, d2r);
            }
            if (std::abs(dxyz) < p.minAbsStep_trans && std::abs(drot) < p.minAbsStep_rot)  // :228-229
            {
                result.terminationReason = IterTermReason::Stalled;
                break;
            }
            if (auto itQ = p.quality_checkpoints.find(result.nIterations); itQ != p.quality_checkpoints.end())  // :257-280
                if (evaluate_quality(quality_evaluators_, pcGlobal, pcLocal, current, pairings) < itQ->second)
                {
                    result.terminationReason = IterTermReason::QualityCheckpointFailed;
                    break;
                }
            prev2 = prev;
            prev  = current;
        }
        if (!quality_evaluators_.empty())  // :322-324 (the reference asserts a non-empty list)
            result.quality = evaluate_quality(quality_evaluators_, pcGlobal, pcLocal, current, pairings);
        result.optimal_tf    = current;
        result.finalPairings = std::move(pairings);
    }

   private:
    matcher_list_t      matchers_;
    solver_list_t       solvers_;
    quality_eval_list_t quality_evaluators_;
};

}  // namespace mp2p_icp_b200
