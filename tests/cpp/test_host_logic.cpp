// CPU-only tests of the host logic above the C ABI (no GPU, no library call): the gating of matchers and
// solvers, run_matchers, Pairings, the ICP::align loop with its termination reasons and the quality
// evaluation — with mock Matcher / Solver / QualityEvaluator classes, as the reference's own loop logic is
// exercised by tests/test-mp2p_icp_algos.cpp through real ones.
//   Matcher.cpp:28-88, Solver.cpp:28-64, Pairings.cpp:123-147, ICP.cpp:108-338 + :608-634
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

// emits `n` pt2pt pairings per call (0 from iteration `empty_from` on) and counts its calls
struct MockMatcher : Matcher
{
    size_t           n          = 5;
    uint32_t         empty_from = 1000000;
    mutable uint32_t calls      = 0;
    mutable std::vector<uint32_t> seen_iterations;

   protected:
    bool impl_match(const metric_map_t&, const metric_map_t&, const CPose3D&, const MatchContext& mc, MatchState&,
                    Pairings& out) const override
    {
        calls++;
        seen_iterations.push_back(mc.icpIteration);
        out = Pairings();
        if (mc.icpIteration < empty_from) out.paired_pt2pt.resize(n);
        out.potential_pairings = 10;
        out.point_weights.emplace_back(n, 2.0);
        return true;
    }
};

// moves the estimate by a step that halves each call (so the loop stalls), or refuses to solve
struct MockSolver : Solver
{
    double           step0 = 0.1;
    bool             fail  = false;
    mutable uint32_t calls = 0;

   protected:
    bool impl_optimal_pose(const Pairings&, OptimalTF_Result& out, const SolverContext& sc) const override
    {
        if (fail) return false;
        const double step = step0 / double(1u << std::min<uint32_t>(calls, 30u));
        calls++;
        out.optimalPose = *sc.guessRelativePose + CPose3D(step, 0, 0, 0, 0, 0);
        return true;
    }
};

struct MockQuality : QualityEvaluator
{
    double q    = 0.5;
    bool   hard = false;
    void   initialize(const ParameterMap&) override {}
    Result evaluate(const metric_map_t&, const metric_map_t&, const CPose3D&, const Pairings&) const override
    {
        Result r;
        r.quality = q, r.hard_discard = hard;
        return r;
    }
};

static metric_map_t one_layer(size_t n)
{
    metric_map_t m;
    auto         pts = CPointsMap::Create();
    for (size_t i = 0; i < n; i++) pts->insertPoint((float)i, 0.f, 0.f);
    m.layers["raw"] = pts;
    return m;
}

int main()
{
    try
    {
        const metric_map_t g = one_layer(40), l = one_layer(33);
        // ---- Matcher gating (Matcher.cpp:35-44) and run_matchers (Matcher.cpp:46-88)
        {
            auto         a = std::make_shared<MockMatcher>(), b = std::make_shared<MockMatcher>(), c = std::make_shared<MockMatcher>();
            ParameterMap pb, pc;
            pb.set("runFromIteration", 2);
            pb.set("runUpToIteration", 3);
            b->initialize(pb);
            pc.set("enabled", 0);
            c->initialize(pc);
            b->n = 7;
            const matcher_list_t ms{a, b, c};
            size_t               expect[] = {5, 5, 12, 12, 5};
            for (uint32_t it = 0; it < 5; it++)
            {
                MatchContext mc;
                mc.icpIteration   = it;
                const Pairings p  = run_matchers(ms, g, l, CPose3D::Identity(), mc);
                ASSERT_(p.paired_pt2pt.size() == expect[it]);
                ASSERT_(p.potential_pairings == (expect[it] == 12 ? 20u : 10u));
                ASSERT_(p.point_weights.empty());  // Pairings::push_back drops the weights (Pairings.cpp:123-131)
                ASSERT_(!p.empty() && p.size() == expect[it]);
            }
            ASSERT_(a->calls == 5 && b->calls == 2 && c->calls == 0);
            // MatchState has one bit per point of every layer
            MatchState st(g, l);
            ASSERT_(st.globalPaired.at("raw").size() == 2 && st.localPaired.at("raw").size() == 2);
            MatchState::mark(st.localPaired.at("raw"), 32);
            ASSERT_(st.localPaired.at("raw")[1] == 1u);
        }
        // ---- Solver gating (Solver.cpp:36-64), incl. runUntilTranslationCorrectionSmallerThan
        {
            MockSolver   s;
            ParameterMap p;
            p.set("runFromIteration", 1);
            p.set("runUpToIteration", 2);
            s.initialize(p);
            Pairings         pr;
            OptimalTF_Result out;
            SolverContext    sc;
            sc.guessRelativePose = CPose3D::Identity();
            bool ran[4];
            for (uint32_t it = 0; it < 4; it++)
            {
                sc.icpIteration = it;
                ran[it]         = s.optimal_pose(pr, out, sc);
            }
            ASSERT_(!ran[0] && ran[1] && ran[2] && !ran[3]);
            MockSolver   s2;
            ParameterMap p2;
            p2.set("runUntilTranslationCorrectionSmallerThan", 0.01);
            s2.initialize(p2);
            SolverContext sc2;
            sc2.guessRelativePose    = CPose3D::Identity();
            sc2.lastIcpStepIncrement = CPose3D(0.5, 0, 0, 0, 0, 0);
            ASSERT_(s2.optimal_pose(pr, out, sc2));
            sc2.lastIcpStepIncrement = CPose3D(0.001, 0, 0, 0, 0, 0);
            ASSERT_(!s2.optimal_pose(pr, out, sc2));  // finished ...
            sc2.lastIcpStepIncrement = CPose3D(0.5, 0, 0, 0, 0, 0);
            ASSERT_(!s2.optimal_pose(pr, out, sc2));  // ... and stays finished for this align()
            MockSolver off;
            ParameterMap po;
            po.set("enabled", 0);
            off.initialize(po);
            ASSERT_(!off.optimal_pose(pr, out, sc));
        }
        // ---- ICP::align termination reasons (ICP.cpp:143-308)
        {
            Parameters prm;
            prm.maxIterations = 50;
            Results res;
            {  // Stalled: the step halves every iteration; 1- and 2-step increments drop below 5e-4
                ICP  icp;
                auto m = std::make_shared<MockMatcher>();
                auto s = std::make_shared<MockSolver>();
                icp.matchers().push_back(m), icp.solvers().push_back(s);
                icp.align(l, g, CPose3D::Identity(), prm, res);
                ASSERT_(res.terminationReason == IterTermReason::Stalled);
                ASSERT_(res.nIterations == s->calls - 1 && s->calls == m->calls && s->calls >= 8 && s->calls <= 10);
                ASSERT_(std::abs(res.optimal_tf.m[3] - 0.2) < 1e-3);  // sum of the halving steps
                ASSERT_(res.finalPairings.paired_pt2pt.size() == 5 && res.quality == 0.0);
                for (uint32_t k = 0; k < m->seen_iterations.size(); k++) ASSERT_(m->seen_iterations[k] == k);
            }
            {  // MaxIterations
                ICP  icp;
                auto s   = std::make_shared<MockSolver>();
                s->step0 = 1e6;  // never small
                icp.matchers().push_back(std::make_shared<MockMatcher>()), icp.solvers().push_back(s);
                Parameters p3 = prm;
                p3.maxIterations = 7;
                icp.align(l, g, CPose3D::Identity(), p3, res);
                ASSERT_(res.terminationReason == IterTermReason::MaxIterations && res.nIterations == 7 && s->calls == 7);
            }
            {  // NoPairings at iteration 3 (ICP.cpp:146-150)
                ICP  icp;
                auto m        = std::make_shared<MockMatcher>();
                m->empty_from = 3;
                auto s        = std::make_shared<MockSolver>();
                s->step0      = 1e6;
                icp.matchers().push_back(m), icp.solvers().push_back(s);
                icp.align(l, g, CPose3D::Identity(), prm, res);
                ASSERT_(res.terminationReason == IterTermReason::NoPairings && res.nIterations == 3 && s->calls == 3);
            }
            {  // SolverError: the first solver refuses, the second is gated out (run_solvers, ICP.cpp:469-479)
                ICP  icp;
                auto s1  = std::make_shared<MockSolver>();
                s1->fail = true;
                auto         s2 = std::make_shared<MockSolver>();
                ParameterMap p2;
                p2.set("runFromIteration", 5);
                s2->initialize(p2);
                icp.matchers().push_back(std::make_shared<MockMatcher>());
                icp.solvers().push_back(s1), icp.solvers().push_back(s2);
                icp.align(l, g, CPose3D::Identity(), prm, res);
                ASSERT_(res.terminationReason == IterTermReason::SolverError && res.nIterations == 0);
            }
            {  // the first solver that solves wins; quality = weighted mean, 0 on a hard discard (ICP.cpp:608-634)
                ICP  icp;
                auto s1  = std::make_shared<MockSolver>();
                s1->fail = true;
                auto s2  = std::make_shared<MockSolver>();
                icp.matchers().push_back(std::make_shared<MockMatcher>());
                icp.solvers().push_back(s1), icp.solvers().push_back(s2);
                auto q1 = std::make_shared<MockQuality>(), q2 = std::make_shared<MockQuality>();
                q1->q = 0.2, q2->q = 0.8;
                icp.quality_evaluators().push_back({q1, 1.0}), icp.quality_evaluators().push_back({q2, 3.0});
                icp.align(l, g, CPose3D::Identity(), prm, res);
                ASSERT_(res.terminationReason == IterTermReason::Stalled && s2->calls > 0);
                ASSERT_(std::abs(res.quality - (0.2 + 3 * 0.8) / 4) < 1e-15);
                q2->hard = true, s2->calls = 0;
                icp.align(l, g, CPose3D::Identity(), prm, res);
                ASSERT_(res.quality == 0.0);
                q2->hard = false;
                Parameters pq = prm;  // quality checkpoint at iteration 2 (ICP.cpp:257-280)
                pq.quality_checkpoints[2] = 0.9;
                s2->calls                 = 0;
                icp.align(l, g, CPose3D::Identity(), pq, res);
                ASSERT_(res.terminationReason == IterTermReason::QualityCheckpointFailed && res.nIterations == 2);
                pq.quality_checkpoints[2] = 0.5;
                s2->calls                 = 0;
                icp.align(l, g, CPose3D::Identity(), pq, res);
                ASSERT_(res.terminationReason == IterTermReason::Stalled);
            }
        }
        // ---- QualityEvaluator_PairedRatio, reuse mode (QualityEvaluator_PairedRatio.cpp:46-73)
        {
            QualityEvaluator_PairedRatio q;
            q.initialize(ParameterMap());
            Pairings p;
            p.paired_pt2pt.resize(30), p.paired_pt2pl.resize(10), p.potential_pairings = 100;
            auto r = q.evaluate(g, l, CPose3D::Identity(), p);
            ASSERT_(std::abs(r.quality - 0.4) < 1e-15 && !r.hard_discard);
            p.paired_pt2pt.resize(5);
            r = q.evaluate(g, l, CPose3D::Identity(), p);
            ASSERT_(std::abs(r.quality - 0.15) < 1e-15 && r.hard_discard);  // below absolute_minimum_pairing_ratio 0.20
            p.potential_pairings = 0;
            ASSERT_(q.evaluate(g, l, CPose3D::Identity(), p).quality == 0.0);
        }
        // ---- parameters: required / optional / errors (MCP_LOAD_REQ / MCP_LOAD_OPT behaviour)
        {
            ParameterMap p;
            p.set("threshold", 1.5);
            Matcher_Points_DistanceThreshold m;
            bool                             thrown = false;
            try
            {
                m.initialize(p);  // thresholdAngularDeg is required (Matcher_Points_DistanceThreshold.cpp:44)
            }
            catch (const std::invalid_argument&)
            {
                thrown = true;
            }
            ASSERT_(thrown);
            p.set("thresholdAngularDeg", 0.0);
            p.set("pairingsPerPoint", 3);
            m.initialize(p);
            ASSERT_(m.threshold == 1.5 && m.pairingsPerPoint == 3);
            thrown = false;
            try
            {
                robust_kernel_from_string("Huber");
            }
            catch (const std::invalid_argument&)
            {
                thrown = true;
            }
            ASSERT_(thrown && robust_kernel_from_string("RobustKernel::GemanMcClure") == 1);
        }
        // ---- formula parameters: tests/test-mp2p_matcher_pt2pt_parameterizable.cpp:28-53
        {
            Matcher_Points_DistanceThreshold m;
            ParameterMap                     p;
            p.set("threshold", "MATCH_THRESHOLD*2.0");  // defined as an expression
            p.set("thresholdAngularDeg", .0);
            m.initialize(p);
            bool thrown = false;
            try
            {
                m.checkAllParametersAreRealized();  // `threshold` still waits for its variable
            }
            catch (const std::runtime_error&)
            {
                thrown = true;
            }
            ASSERT_(thrown);
            ParameterSource globalParams;
            globalParams.attach(m);
            globalParams.updateVariable("MATCH_THRESHOLD", 1.5);
            globalParams.realize();
            ASSERT_(std::abs(m.threshold - 3.0) < 1e-4 && std::abs(m.thresholdAngularDeg) < 1e-4);
            m.checkAllParametersAreRealized();
            globalParams.updateVariable("MATCH_THRESHOLD", 0.5);
            globalParams.realize();
            ASSERT_(std::abs(m.threshold - 1.0) < 1e-4 && std::abs(m.thresholdAngularDeg) < 1e-4);
            // plain numbers stay what the flat parser made of them (constants, realized at declaration)
            Matcher_Points_DistanceThreshold m2;
            ParameterMap                     p2;
            p2.set("threshold", 0.40 * 0.1552);
            p2.set("thresholdAngularDeg", 0);
            p2.set("pairingsPerPoint", 3);
            m2.initialize(p2);
            m2.checkAllParametersAreRealized();
            ASSERT_(m2.threshold == 0.40 * 0.1552 && m2.thresholdAngularDeg == 0.0 && m2.pairingsPerPoint == 3);
            // the expression language itself
            const std::map<std::string, double> v{{"A", 2.0}, {"b_1", -3.0}};
            ASSERT_(Expression::eval("1e-3", v) == 1e-3 && Expression::eval(" ( A + 1 ) * -b_1 ", v) == 9.0);
            ASSERT_(Expression::eval("max(A, abs(b_1)) / 2 ^ 2", v) == 0.75 && Expression::eval("sqrt(A*8)", v) == 4.0);
            for (const char* bad : {"A +", "unknown*2", "(A", "2 $ 3", "foo(1)"})
            {
                thrown = false;
                try
                {
                    Expression::eval(bad, v);
                }
                catch (const std::runtime_error&)
                {
                    thrown = true;
                }
                ASSERT_(thrown);
            }
        }
    }
    catch (std::exception& e)
    {
        std::cerr << e.what() << "\n";
        return 1;
    }
    std::puts("test_host_logic OK");
    return 0;
}
