// Mirrors the reference's tests/test-mp2p_optimize_pt2pl.cpp (Solver_GaussNewton on 3 pt2pl + 1 pt2pt
// pairings recovers 15 ground-truth poses to 1e-3), tests/test-mp2p_optimize_pt2ln.cpp (three
// point-to-line pairings on the axes, same poses) and the protocol of tests/test-mp2p_icp_algos.cpp
// (full ICP::align on a decimated cloud, |log(GT - est)| < 0.1) through the host mirror.
#include <cstdio>
#include <fstream>
#include <iostream>
#include <random>

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

static mp2p_b200_pair_pt2pl plane_pair(const CPose3D& gt, double nx, double ny, double nz, double gx, double gy, double gz)
{
    mp2p_b200_pair_pt2pl pp{};
    pp.plane_coefs[0] = nx, pp.plane_coefs[1] = ny, pp.plane_coefs[2] = nz, pp.plane_coefs[3] = 0;  // FromPointAndNormal({0,0,0}, n)
    double lx, ly, lz;
    gt.inverseComposePoint(gx, gy, gz, lx, ly, lz);
    pp.local_x = (float)lx, pp.local_y = (float)ly, pp.local_z = (float)lz;
    return pp;
}

static int test_opt_pt2pl(const CPose3D& groundTruth, const Solver& solver)
{
    Pairings p;
    p.paired_pt2pl.push_back(plane_pair(groundTruth, 0, 0, 1, 0.5, 0, 0));
    p.paired_pt2pl.push_back(plane_pair(groundTruth, 1, 0, 0, 0, 0.8, 0));
    p.paired_pt2pl.push_back(plane_pair(groundTruth, 0, 1, 0, 0, 0, 0.3));
    {
        mp2p_b200_pair_pt2pt pp{};
        double               lx, ly, lz;
        groundTruth.inverseComposePoint(0, 0, 0, lx, ly, lz);
        pp.local_x = (float)lx, pp.local_y = (float)ly, pp.local_z = (float)lz;
        p.paired_pt2pt.push_back(pp);
    }
    OptimalTF_Result result;
    SolverContext    sc;
    sc.guessRelativePose = CPose3D::Identity();
    ASSERT_(solver.optimal_pose(p, result, sc));
    double dxyz, drot;
    (result.optimalPose - groundTruth).log_norms(dxyz, drot);
    ASSERT_(std::sqrt(dxyz * dxyz + drot * drot) < 1e-3);
    return 0;
}

// tests/test-mp2p_optimize_pt2ln.cpp:25-78
static int test_opt_pt2ln(const CPose3D& groundTruth, const Solver& solver)
{
    Pairings     p;
    const double axis[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, pt[3][3] = {{0.5, 0, 0}, {0, 0.4, 0}, {0, 0, 0.2}};
    for (int k = 0; k < 3; k++)
    {
        mp2p_b200_pair_pt2ln pp{};  // TLine3D::FromTwoPoints({0,0,0}, axis)
        for (int d = 0; d < 3; d++) pp.director[d] = axis[k][d];
        groundTruth.inverseComposePoint(pt[k][0], pt[k][1], pt[k][2], pp.local[0], pp.local[1], pp.local[2]);
        p.paired_pt2ln.push_back(pp);
    }
    OptimalTF_Result result;
    SolverContext    sc;
    sc.guessRelativePose = CPose3D::Identity();
    ASSERT_(solver.optimal_pose(p, result, sc));
    double dxyz, drot;
    (result.optimalPose - groundTruth).log_norms(dxyz, drot);
    ASSERT_(std::sqrt(dxyz * dxyz + drot * drot) < 1e-3);
    return 0;
}

int main(int argc, char** argv)
{
    try
    {
        const double       D = M_PI / 180.0;
        Solver_GaussNewton solverGN;
        {
            ParameterMap sp;
            sp.set("maxIterations", 25);
            solverGN.initialize(sp);
        }
        const CPose3D gts[] = {CPose3D(0, 0, 0, 0, 0, 0),      CPose3D(1, 0, 0, 0, 0, 0),     CPose3D(0, 1, 0, 0, 0, 0),
                               CPose3D(0, 0, 1, 0, 0, 0),      CPose3D(-2, 0, 0, 0, 0, 0),    CPose3D(0, -3, 0, 0, 0, 0),
                               CPose3D(0, 0, -4, 0, 0, 0),     CPose3D(0, 0, 0, 20 * D, 0, 0), CPose3D(0, 0, 0, -20 * D, 0, 0),
                               CPose3D(0, 0, 0, 0, 10 * D, 0), CPose3D(0, 0, 0, 0, -10 * D, 0), CPose3D(0, 0, 0, 0, 0, 15 * D),
                               CPose3D(0, 0, 0, 0, 0, -15 * D), CPose3D(1, 2, 3, 0, 0, 0),     CPose3D(1, 2, 3, -10 * D, 5 * D, 30 * D)};
        for (const auto& gt : gts)
            if (test_opt_pt2pl(gt, solverGN)) return 1;
        for (const auto& gt : gts)
            if (test_opt_pt2ln(gt, solverGN)) return 1;
        {  // Solver: missing required parameter
            Solver_GaussNewton s2;
            bool               thrown = false;
            try
            {
                s2.initialize(ParameterMap());
            }
            catch (const std::invalid_argument&)
            {
                thrown = true;
            }
            ASSERT_(thrown);
        }

        // ---- ICP::align protocol (tests/test-mp2p_icp_algos.cpp:85-223) on a cloud given as "x y z" text
        if (argc > 1)
        {
            auto          pts = CPointsMap::Create();
            std::ifstream f(argv[1]);
            float         x, y, z;
            size_t        k = 0;
            float         mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
            while (f >> x >> y >> z)
                if (k++ % 10 == 0)
                {
                    pts->insertPoint(x, y, z);
                    const float v[3] = {x, y, z};
                    for (int d = 0; d < 3; d++) mn[d] = std::min(mn[d], v[d]), mx[d] = std::max(mx[d], v[d]);
                }
            ASSERT_(pts->size() > 500);
            const double sz[3]   = {mx[0] - mn[0], mx[1] - mn[1], mx[2] - mn[2]};
            const double max_dim = std::max(sz[0], std::max(sz[1], sz[2]));
            std::mt19937_64                        rng(1234);
            std::uniform_real_distribution<double> U(-1.0, 1.0);
            // matcherKind 0: Matcher_Points_DistanceThreshold, 1: Matcher_Points_InlierRatio (defaults),
            // tests/test-mp2p_icp_algos.cpp:250-262
            // 2: the schedule of demos/icp-settings-kitti.yaml:39-59 — Matcher_Points_DistanceThreshold for iterations
            //    0-5, Matcher_Adaptive from iteration 6 on (confidenceInterval 0.75, firstToSecondDistanceMax 1.2)
            for (int matcherKind = 0; matcherKind < 3; matcherKind++)
            for (int solverKind = 0; solverKind < 2; solverKind++)
                for (int rep = 0; rep < 3; rep++)
                {
                    const CPose3D gt(0.15 * sz[0] * U(rng), 0.15 * sz[1] * U(rng), 0.15 * sz[2] * U(rng), 10 * D * U(rng),
                                     10 * D * U(rng), 10 * D * U(rng));
                    auto reg = CPointsMap::Create();  // changeCoordinatesReference(pts, -gt)
                    for (size_t i = 0; i < pts->size(); i++)
                    {
                        double lx, ly, lz;
                        gt.inverseComposePoint(pts->getPointsBufferRef_x()[i], pts->getPointsBufferRef_y()[i],
                                               pts->getPointsBufferRef_z()[i], lx, ly, lz);
                        reg->insertPoint((float)lx, (float)ly, (float)lz);
                    }
                    metric_map_t pc_ref, pc_mod;
                    pc_ref.layers["raw"] = pts;
                    pc_mod.layers["raw"] = reg;
                    ICP  icp;
                    if (matcherKind == 0)
                    {
                        auto         matcher = std::make_shared<Matcher_Points_DistanceThreshold>();
                        ParameterMap ps;
                        ps.set("threshold", 0.40 * max_dim);
                        ps.set("thresholdAngularDeg", 0);
                        matcher->initialize(ps);
                        icp.matchers().push_back(matcher);
                    }
                    else if (matcherKind == 1)
                        icp.matchers().push_back(std::make_shared<Matcher_Points_InlierRatio>());
                    else
                    {
                        auto         m1 = std::make_shared<Matcher_Points_DistanceThreshold>();
                        ParameterMap p1;
                        p1.set("threshold", 0.40 * max_dim);
                        p1.set("thresholdAngularDeg", 0);
                        p1.set("runFromIteration", 0);
                        p1.set("runUpToIteration", 5);
                        m1->initialize(p1);
                        icp.matchers().push_back(m1);
                        auto         m2 = std::make_shared<Matcher_Adaptive>();
                        ParameterMap p2;
                        p2.set("confidenceInterval", 0.75);
                        p2.set("firstToSecondDistanceMax", 1.2);
                        p2.set("absoluteMaxSearchDistance", 0.40 * max_dim);
                        p2.set("minimumCorrDist", 0.02 * max_dim);
                        p2.set("enableDetectPlanes", 0);
                        p2.set("runFromIteration", 6);
                        p2.set("runUpToIteration", 0);
                        m2->initialize(p2);
                        icp.matchers().push_back(m2);
                    }
                    if (solverKind == 0)
                        icp.solvers().push_back(std::make_shared<Solver_Horn>());
                    else
                    {
                        auto         s = std::make_shared<Solver_GaussNewton>();
                        ParameterMap sp;
                        sp.set("maxIterations", 6);
                        s->initialize(sp);
                        icp.solvers().push_back(s);
                    }
                    // QualityEvaluator_PairedRatio: from the last pairings (default) and from its own
                    // matcher pass (QualityEvaluator_PairedRatio.cpp:27-73)
                    auto qReuse = std::make_shared<QualityEvaluator_PairedRatio>();
                    qReuse->initialize(ParameterMap());
                    auto         qOwn = std::make_shared<QualityEvaluator_PairedRatio>();
                    ParameterMap qp;
                    qp.set("reuse_icp_pairings", 0);
                    qp.set("threshold", 0.05 * max_dim);
                    qp.set("thresholdAngularDeg", 0);
                    qOwn->initialize(qp);
                    icp.quality_evaluators().push_back({qReuse, 1.0});
                    icp.quality_evaluators().push_back({qOwn, 3.0});
                    Parameters icp_params;
                    icp_params.maxIterations = 100;
                    Results res;
                    icp.align(pc_mod, pc_ref, CPose3D::Identity(), icp_params, res);
                    {
                        const auto   r1 = qReuse->evaluate(pc_ref, pc_mod, res.optimal_tf, res.finalPairings);
                        const auto   r2 = qOwn->evaluate(pc_ref, pc_mod, res.optimal_tf, res.finalPairings);
                        const double expect1 = res.finalPairings.size() / double(res.finalPairings.potential_pairings);
                        ASSERT_(std::abs(r1.quality - expect1) < 1e-15 && !r1.hard_discard);
                        // a converged alignment of a cloud with itself pairs (nearly) every point within 5 % of the size
                        ASSERT_(r2.quality > 0.95 && r2.quality <= 1.0);
                        ASSERT_(std::abs(res.quality - (1.0 * r1.quality + 3.0 * r2.quality) / 4.0) < 1e-12);
                        // a checkpoint nobody can pass aborts the loop at that iteration (ICP.cpp:257-280)
                        Parameters pq = icp_params;
                        pq.quality_checkpoints[1] = 2.0;
                        Results rq;
                        icp.align(pc_mod, pc_ref, CPose3D::Identity(), pq, rq);
                        ASSERT_(rq.terminationReason == IterTermReason::QualityCheckpointFailed && rq.nIterations == 1);
                    }
                    double dxyz, drot;
                    (gt - res.optimal_tf).log_norms(dxyz, drot);
                    std::printf("matcher=%d solver=%d rep=%d iters=%u err=%.3e\n", matcherKind, solverKind, rep, res.nIterations,
                                std::sqrt(dxyz * dxyz + drot * drot));
                    ASSERT_(std::sqrt(dxyz * dxyz + drot * drot) < 0.1);
                    ASSERT_(!res.finalPairings.empty());
                }
        }
    }
    catch (std::exception& e)
    {
        std::cerr << e.what() << "\n";
        return 1;
    }
    std::puts("test_optimize_and_align OK");
    return 0;
}
