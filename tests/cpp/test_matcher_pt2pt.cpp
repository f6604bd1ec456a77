// Mirrors the reference's tests/test-mp2p_matcher_pt2pt.cpp (same clouds, same four poses, same
// expected (localIdx, globalIdx)) against the MRPT-free host mirror of the plugin interface, i.e.
// through the C ABI onto the GPU. Run by tests/test_cpp_host.py (-m gpu).
#include <cstdio>
#include <iostream>

#include "mp2p_icp_b200.hpp"

using namespace mp2p_icp_b200;

#define ASSERT_(c)                                                                          \
    do                                                                                      \
    {                                                                                       \
        if (!(c))                                                                           \
        {                                                                                   \
            std::fprintf(stderr, "%s:%d: assert failed: %s\n", __FILE__, __LINE__, #c);     \
            return 1;                                                                       \
        }                                                                                   \
    } while (0)

static CPointsMap::Ptr generateGlobalPoints()
{
    auto pts = CPointsMap::Create();
    for (int i = 0; i < 10; i++) pts->insertPoint(i * 0.01f, 5.0f, .0f);
    for (int i = 0; i < 10; i++) pts->insertPoint(10.0f, i * 0.01f, 1.0f);
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

        Matcher_Points_DistanceThreshold m;
        ParameterMap                     p;
        p.set("threshold", 1.05);
        p.set("thresholdAngularDeg", .001);
        m.initialize(p);
        ASSERT_(std::abs(m.threshold - 1.05) < 1e-4);
        ASSERT_(std::abs(m.thresholdAngularDeg - .001) < 1e-4);

        {  // identity
            Pairings   pairs;
            MatchState ms(pcGlobal, pcLocal);
            m.match(pcGlobal, pcLocal, CPose3D(0, 0, 0, 0, 0, 0), {}, ms, pairs);
            ASSERT_(pairs.empty());
        }
        {  // pose #1
            Pairings   pairs;
            MatchState ms(pcGlobal, pcLocal);
            m.match(pcGlobal, pcLocal, CPose3D(0, 5, 0, 0, 0, 0), {}, ms, pairs);
            ASSERT_(pairs.size() == 1);
            ASSERT_(pairs.paired_pt2pt.at(0).localIdx == 0);
            ASSERT_(pairs.paired_pt2pt.at(0).globalIdx == 0);
        }
        {  // pose #2
            Pairings   pairs;
            MatchState ms(pcGlobal, pcLocal);
            m.match(pcGlobal, pcLocal, CPose3D(-2, 5, 0, 0, 0, 0), {}, ms, pairs);
            ASSERT_(pairs.size() == 1);
            ASSERT_(pairs.paired_pt2pt.at(0).globalIdx == 0);
            ASSERT_(pairs.paired_pt2pt.at(0).localIdx == 1);
        }
        {  // pose #3
            Pairings   pairs;
            MatchState ms(pcGlobal, pcLocal);
            m.match(pcGlobal, pcLocal, CPose3D(8.5, -1.0, 1, 45.0 * M_PI / 180.0, 0, 0), {}, ms, pairs);
            ASSERT_(pairs.size() == 1);
            ASSERT_(pairs.paired_pt2pt.at(0).localIdx == 1);
            ASSERT_(pairs.paired_pt2pt.at(0).globalIdx == 19);
        }
        {  // error behaviour: missing required parameter -> std::invalid_argument
            Matcher_Points_DistanceThreshold m2;
            ParameterMap                     q;
            q.set("threshold", 1.0);
            bool thrown = false;
            try
            {
                m2.initialize(q);
            }
            catch (const std::invalid_argument&)
            {
                thrown = true;
            }
            ASSERT_(thrown);
        }
        {  // gates of Matcher::match (Matcher.cpp:35-44)
            Matcher_Points_DistanceThreshold m3;
            p.set("runFromIteration", 3);
            m3.initialize(p);
            Pairings     pairs;
            MatchState   ms(pcGlobal, pcLocal);
            MatchContext mc;
            mc.icpIteration = 1;
            ASSERT_(!m3.match(pcGlobal, pcLocal, CPose3D(0, 5, 0, 0, 0, 0), mc, ms, pairs));
            mc.icpIteration = 3;
            ASSERT_(m3.match(pcGlobal, pcLocal, CPose3D(0, 5, 0, 0, 0, 0), mc, ms, pairs));
            ASSERT_(pairs.size() == 1);
        }
    }
    catch (std::exception& e)
    {
        std::cerr << e.what() << "\n";
        return 1;
    }
    std::puts("test_matcher_pt2pt OK");
    return 0;
}
