// Mirrors the reference's tests/test-mp2p_matcher_pt2pl.cpp (disabled upstream,
// tests/CMakeLists.txt:37): plane x=10 known answer, blob rejection, and the two run_matchers
// pipelines with a pt2pt matcher sharing the MatchState.
#include <cstdio>
#include <iostream>

#include "mp2p_icp_b200.hpp"

using namespace mp2p_icp_b200;

#define ASSERT_(c)                                                                      \
    do                                                                                  \
    {                                                                                   \
        if (!(c))                                                                       \
        {                                                                               \
            std::fprintf(stderr, "%s:%d: assert failed: %s\n", __FILE__, __LINE__, #c); \
            return 1;                                                                   \
        }                                                                               \
    } while (0)
#define ASSERT_NEAR_(a, b, tol) ASSERT_(std::abs((a) - (b)) <= (tol))

static CPointsMap::Ptr generateGlobalPoints()
{
    auto pts = CPointsMap::Create();
    for (int ix = 0; ix < 10; ix++)
        for (int iy = 0; iy < 10; iy++) pts->insertPoint(ix * 0.01f, 5.0f + iy * 0.01f, .0f);
    for (int iy = 0; iy < 10; iy++)
        for (int iz = 0; iz < 10; iz++) pts->insertPoint(10.0f, iy * 0.01f, iz * 0.01f);
    for (int ix = 0; ix < 10; ix++)
        for (int iy = 0; iy < 10; iy++)
            for (int iz = 0; iz < 10; iz++) pts->insertPoint(20.0f + ix * 0.01f, iy * 0.01f, iz * 0.01f);
    return pts;
}
static CPointsMap::Ptr generateLocalPoints()
{
    auto pts = CPointsMap::Create();
    pts->insertPointFast(0.f, 0.f, 0.f);
    pts->insertPointFast(2.f, 0.f, 0.f);
    return pts;
}

int main()
{
    try
    {
        metric_map_t pcGlobal, pcLocal;
        pcGlobal.layers[metric_map_t::PT_LAYER_RAW] = generateGlobalPoints();
        pcLocal.layers[metric_map_t::PT_LAYER_RAW]  = generateLocalPoints();

        auto         m = std::make_shared<Matcher_Point2Plane>();
        ParameterMap p;
        p.set("distanceThreshold", 0.1);
        p.set("searchRadius", 0.1);
        p.set("minimumPlanePoints", 5.0);
        p.set("knn", 5);
        p.set("planeEigenThreshold", 0.1);
        m->initialize(p);
        {
            Pairings   pairs;
            MatchState ms(pcGlobal, pcLocal);
            m->match(pcGlobal, pcLocal, CPose3D(0, 0, 0, 0, 0, 0), {}, ms, pairs);
            ASSERT_(pairs.empty());
        }
        {
            Pairings   pairs;
            MatchState ms(pcGlobal, pcLocal);
            m->match(pcGlobal, pcLocal, CPose3D(0, 5, 0, 0, 0, 0), {}, ms, pairs);
            ASSERT_(pairs.size() == 1U);
            ASSERT_(pairs.paired_pt2pl.size() == 1U);
        }
        {
            Pairings   pairs;
            MatchState ms(pcGlobal, pcLocal);
            m->match(pcGlobal, pcLocal, CPose3D(8.04, 0, 0.0, 0, 0, 0), {}, ms, pairs);
            ASSERT_(pairs.size() == 1U);
            ASSERT_(pairs.paired_pt2pl.size() == 1U);
            const auto& p0 = pairs.paired_pt2pl.at(0);
            ASSERT_NEAR_(p0.local_x, 2.0, 1e-3);
            ASSERT_NEAR_(p0.local_y, 0.0, 1e-3);
            ASSERT_NEAR_(p0.local_z, 0.0, 1e-3);
            ASSERT_NEAR_(p0.centroid[0], 10.0, 0.01);
            ASSERT_NEAR_(p0.centroid[1], 0.0, 0.01);
            ASSERT_NEAR_(p0.centroid[2], 0.0, 0.01);
            ASSERT_NEAR_(p0.plane_coefs[0], 1.0, 1e-3);
            ASSERT_NEAR_(p0.plane_coefs[1], 0.0, 1e-3);
            ASSERT_NEAR_(p0.plane_coefs[2], 0.0, 1e-3);
            ASSERT_NEAR_(p0.plane_coefs[3], -10.0, 1e-3);
        }
        {
            Pairings   pairs;
            MatchState ms(pcGlobal, pcLocal);
            m->match(pcGlobal, pcLocal, CPose3D(18.053, 0.05, 0.03, 0, 0, 0), {}, ms, pairs);
            ASSERT_(pairs.paired_pt2pl.size() == 0U);
        }
        for (int allow = 1; allow >= 0; allow--)
        {
            auto         mPt2Pt = std::make_shared<Matcher_Points_DistanceThreshold>();
            ParameterMap p2;
            p2.set("threshold", 0.1);
            p2.set("thresholdAngularDeg", .0);
            p2.set("allowMatchAlreadyMatchedPoints", allow);
            mPt2Pt->initialize(p2);
            const Pairings pairs = run_matchers({m, mPt2Pt}, pcGlobal, pcLocal, CPose3D(8.04, 0, 0.0, 0, 0, 0), {});
            ASSERT_(pairs.paired_pt2pl.size() == 1U);
            ASSERT_(pairs.size() == (allow ? 2U : 1U));
            ASSERT_(pairs.paired_pt2pt.size() == (allow ? 1U : 0U));
        }
    }
    catch (std::exception& e)
    {
        std::cerr << e.what() << "\n";
        return 1;
    }
    std::puts("test_matcher_pt2pl OK");
    return 0;
}
