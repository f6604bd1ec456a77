// ORACLE — TEST INFRASTRUCTURE ONLY (see se3.hpp header).
//
// Exact k-nearest-neighbour search over three SoA float arrays, restating what the reference
// obtains from MRPT's `NearestNeighborsCapable` on a `CPointsMap` (a nanoflann
// KDTreeSingleIndexAdaptor<L2_Simple_Adaptor<float>,...,3>; EXTERNAL to /root/reference, version
// "MRPT >= 2.11.5", no lockfile). Reference call sites:
//   mp2p_icp/src/Matcher_Points_DistanceThreshold.cpp:92      nn_prepare_for_3d_queries()
//   mp2p_icp/src/Matcher_Points_DistanceThreshold.cpp:161-163 nn_single_search()
//   mp2p_icp/src/Matcher_Points_DistanceThreshold.cpp:174-177 nn_radius_search()
//   mp2p_icp/src/Matcher_Points_DistanceThreshold.cpp:246-248 nn_multiple_search()
//   mp2p_icp/src/Matcher_Points_Base.cpp:105-114              leaf_max_size override
//
// Published algorithm restated here (nanoflann): points stay in their SoA arrays, the tree
// permutes an index vector; leaves hold <= leaf_max (default 10) indices; distance is
//     d2 = ((dx*dx) + dy*dy) + dz*dz          accumulated in FLOAT, dims x,y,z, no FMA
// and a candidate enters the result set only on a strict improvement. nanoflann's winner among
// exact-distance ties depends on traversal order (unpinned upstream); this
// oracle pins the rule "lowest point index wins" by ordering candidates on the pair (d2, index).
// The pruning bound is the same per-dimension accumulated cut distance nanoflann uses, evaluated in
// the same float order as the point distance, hence a rigorous lower bound of the computed d2
// (rounding is monotonic), so the search is EXACT w.r.t. the float metric; tests cross-check it
// against brute force.
#pragma once
#include <algorithm>
#include <cstdint>
#include <limits>
#include <numeric>
#include <vector>

namespace orc
{
struct KDTree
{
    const float *x = nullptr, *y = nullptr, *z = nullptr;
    size_t       n = 0;
    int          leaf_max = 10;

    struct Node
    {
        // leaf: left == right == -1, [lo,hi) into vind
        int32_t  left = -1, right = -1;
        uint32_t lo = 0, hi = 0;
        int      dim = 0;
        float    divlow = 0, divhigh = 0;
    };
    std::vector<Node>     nodes;
    std::vector<uint32_t> vind;
    float                 bbmin[3], bbmax[3];

    float coord(uint32_t i, int d) const { return d == 0 ? x[i] : (d == 1 ? y[i] : z[i]); }

    void build(const float* xs, const float* ys, const float* zs, size_t count, int leafMax)
    {
        x = xs, y = ys, z = zs, n = count, leaf_max = leafMax > 0 ? leafMax : 10;
        vind.resize(n);
        std::iota(vind.begin(), vind.end(), 0u);
        nodes.clear();
        nodes.reserve(2 * (n / leaf_max + 1));
        for (int d = 0; d < 3; d++)
        {
            bbmin[d] = std::numeric_limits<float>::max();
            bbmax[d] = -std::numeric_limits<float>::max();
        }
        for (size_t i = 0; i < n; i++)
            for (int d = 0; d < 3; d++)
            {
                bbmin[d] = std::min(bbmin[d], coord(i, d));
                bbmax[d] = std::max(bbmax[d], coord(i, d));
            }
        if (n) build_rec(0, static_cast<uint32_t>(n));
    }

    int32_t build_rec(uint32_t lo, uint32_t hi)
    {
        const int32_t id = static_cast<int32_t>(nodes.size());
        nodes.emplace_back();
        if (hi - lo <= static_cast<uint32_t>(leaf_max))
        {
            nodes[id].lo = lo, nodes[id].hi = hi;
            return id;
        }
        // widest dimension of THIS subset, median split
        float mn[3], mx[3];
        for (int d = 0; d < 3; d++) mn[d] = std::numeric_limits<float>::max(), mx[d] = -mn[d];
        for (uint32_t k = lo; k < hi; k++)
            for (int d = 0; d < 3; d++)
            {
                const float v = coord(vind[k], d);
                mn[d] = std::min(mn[d], v), mx[d] = std::max(mx[d], v);
            }
        int dim = 0;
        if (mx[1] - mn[1] > mx[dim] - mn[dim]) dim = 1;
        if (mx[2] - mn[2] > mx[dim] - mn[dim]) dim = 2;
        const uint32_t mid = lo + (hi - lo) / 2;
        std::nth_element(vind.begin() + lo, vind.begin() + mid, vind.begin() + hi,
                         [&](uint32_t a, uint32_t b) { return coord(a, dim) < coord(b, dim); });
        float lowmax = -std::numeric_limits<float>::max(), highmin = -lowmax;
        for (uint32_t k = lo; k < mid; k++) lowmax = std::max(lowmax, coord(vind[k], dim));
        for (uint32_t k = mid; k < hi; k++) highmin = std::min(highmin, coord(vind[k], dim));
        const int32_t l = build_rec(lo, mid);
        const int32_t r = build_rec(mid, hi);
        Node&         nd = nodes[id];
        nd.left = l, nd.right = r, nd.dim = dim, nd.divlow = lowmax, nd.divhigh = highmin;
        return id;
    }

    // Result set ordered by (d2, idx); keeps the K best with d2 < radius2 (strict).
    struct ResultSet
    {
        int       K;
        float     radius2;
        int       count = 0;
        float*    d2;
        uint32_t* idx;
        ResultSet(int k, float r2, float* d, uint32_t* i) : K(k), radius2(r2), d2(d), idx(i) {}
        float worst() const { return count < K ? radius2 : d2[K - 1]; }
        static bool less(float da, uint32_t ia, float db, uint32_t ib)
        {
            return da < db || (da == db && ia < ib);
        }
        void add(float d, uint32_t i)
        {
            if (!(d < radius2)) return;
            if (count == K && !less(d, i, d2[K - 1], idx[K - 1])) return;
            int pos = (count < K) ? count++ : K - 1;
            while (pos > 0 && less(d, i, d2[pos - 1], idx[pos - 1]))
            {
                d2[pos] = d2[pos - 1], idx[pos] = idx[pos - 1];
                pos--;
            }
            d2[pos] = d, idx[pos] = i;
        }
    };

    static inline float dist2(float qx, float qy, float qz, float px, float py, float pz)
    {
        const float dx = qx - px, dy = qy - py, dz = qz - pz;
        float       d  = dx * dx;  // no FMA: this translation unit is built with -ffp-contract=off
        d              = d + dy * dy;
        d              = d + dz * dz;
        return d;
    }

    void search_rec(int32_t id, const float q[3], float cut[3], ResultSet& rs) const
    {
        const Node& nd = nodes[id];
        if (nd.left < 0)
        {
            for (uint32_t k = nd.lo; k < nd.hi; k++)
            {
                const uint32_t i = vind[k];
                rs.add(dist2(q[0], q[1], q[2], x[i], y[i], z[i]), i);
            }
            return;
        }
        const int   d     = nd.dim;
        const float val   = q[d];
        const float diff1 = val - nd.divlow, diff2 = val - nd.divhigh;
        int32_t     best, other;
        float       cutd;
        if (diff1 + diff2 < 0)
        {
            best = nd.left, other = nd.right;
            cutd = diff2 * diff2;  // distance to the nearest point coordinate of the far side
        }
        else
        {
            best = nd.right, other = nd.left;
            cutd = diff1 * diff1;
        }
        search_rec(best, q, cut, rs);
        const float saved = cut[d];
        cut[d]            = cutd;
        float lb          = cut[0];
        lb                = lb + cut[1];
        lb                = lb + cut[2];
        // `<=`: an equal-distance point with a lower index must still be visited.
        if (lb <= rs.worst()) search_rec(other, q, cut, rs);
        cut[d] = saved;
    }

    // returns number found (<= K), ascending (d2, idx)
    int knn(const float q[3], int K, float radius2, uint32_t* idx, float* d2) const
    {
        ResultSet rs(K, radius2, d2, idx);
        if (!n) return 0;
        float cut[3];
        for (int d = 0; d < 3; d++)
        {
            float c = 0;
            if (q[d] < bbmin[d]) c = (q[d] - bbmin[d]) * (q[d] - bbmin[d]);
            if (q[d] > bbmax[d]) c = (q[d] - bbmax[d]) * (q[d] - bbmax[d]);
            cut[d] = c;
        }
        search_rec(0, q, cut, rs);
        return rs.count;
    }

    int knn_bruteforce(const float q[3], int K, float radius2, uint32_t* idx, float* d2) const
    {
        ResultSet rs(K, radius2, d2, idx);
        for (size_t i = 0; i < n; i++)
            rs.add(dist2(q[0], q[1], q[2], x[i], y[i], z[i]), static_cast<uint32_t>(i));
        return rs.count;
    }
};
}  // namespace orc
