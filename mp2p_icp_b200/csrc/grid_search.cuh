// Device-side exact k-NN over the multi-resolution hashed voxel index (product code).
//
// Index layout (built in index.cu): map points are sorted by the 63-bit Morton code of their
// finest-level voxel coordinates (2^21 voxels along the longest bbox axis), so the voxel of ANY
// level L (cell = s0 * 2^L) is one contiguous run of the sorted array. For each level from the
// chosen finest one upwards there is an open-addressing hash table  (cx,cy,cz) -> (start,count).
//
// Search for one query: visit the 3x3x3 voxel block around the query at the finest level, keep
// the K best candidates ordered by the pair (d2, original index); the block guarantees that every
// unvisited point is farther than `m` (distance from the query to the block's faces). If the K-th
// best d2 <= m^2 the result is exact and the search stops, otherwise it is repeated one level up
// (voxels twice as large) — the top level is a single voxel, so termination is unconditional.
// Voxels whose box lower bound exceeds the current K-th distance are skipped without a lookup.
//
// Exactness w.r.t. the reference metric (nanoflann L2_Simple on float, see oracle/kdtree.hpp):
// d2 = ((dx*dx)+dy*dy)+dz*dz with non-fused float ops; bounds are made conservative by 4 finest
// quanta (covers the float rounding of the voxel coordinate function, which is monotonic) and a
// 1e-6 relative margin; candidates compare on (d2, index) so exact ties resolve to the lowest
// original index, the rule the oracle pins.
#pragma once
#include "common.cuh"

namespace mp2p
{
__device__ __forceinline__ unsigned long long morton_expand21(uint32_t v)
{
    unsigned long long x = v & 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}
__device__ __forceinline__ unsigned long long morton63(uint32_t x, uint32_t y, uint32_t z)
{
    return morton_expand21(x) | (morton_expand21(y) << 1) | (morton_expand21(z) << 2);
}

// voxel coordinate function (float, monotonic in p): u = (p - o) * inv_s0
__device__ __forceinline__ float grid_u(float p, float o, float inv_s0)
{
    return __fmul_rn(__fsub_rn(p, o), inv_s0);
}

__device__ __forceinline__ unsigned long long cell_key(uint32_t cx, uint32_t cy, uint32_t cz)
{
    return (unsigned long long)cx | ((unsigned long long)cy << 21) | ((unsigned long long)cz << 42);
}
__device__ __forceinline__ uint32_t cell_hash(unsigned long long key, uint32_t shift)
{
    return (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> shift);
}

__device__ __forceinline__ bool grid_lookup(const GridView& g, int rl, uint32_t cx, uint32_t cy,
                                            uint32_t cz, uint32_t& start, uint32_t& count)
{
    const unsigned long long key   = cell_key(cx, cy, cz);
    const uint32_t           shift = g.level_shift[rl];
    const uint32_t           mask  = (1u << (64 - shift)) - 1u;
    const CellEntry*         t     = g.table + g.level_off[rl];
    uint32_t                 h     = cell_hash(key, shift);
    while (true)
    {
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(t + h));
        const unsigned long long k = (unsigned long long)raw.x | ((unsigned long long)raw.y << 32);
        if (k == key)
        {
            start = raw.z;
            count = raw.w;
            return true;
        }
        if (k == kEmptyKey) return false;
        h = (h + 1) & mask;
    }
}

// reference float metric, never fused
__device__ __forceinline__ float dist2_ref(float qx, float qy, float qz, float px, float py, float pz)
{
    const float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
    float       d  = __fmul_rn(dx, dx);
    d              = __fadd_rn(d, __fmul_rn(dy, dy));
    d              = __fadd_rn(d, __fmul_rn(dz, dz));
    return d;
}

// 27 neighbour offsets ordered centre, 6 faces, 12 edges, 8 corners: bits [1:0]=dx+1, [3:2]=dy+1,
// [5:4]=dz+1
__constant__ uint8_t kNeighbourOrder[27] = {
    0x15,                                                                    // (0,0,0)
    0x14, 0x16, 0x11, 0x19, 0x05, 0x25,                                      // faces
    0x10, 0x12, 0x18, 0x1a, 0x04, 0x06, 0x24, 0x26, 0x01, 0x09, 0x21, 0x29,  // edges
    0x00, 0x02, 0x08, 0x0a, 0x20, 0x22, 0x28, 0x2a};                         // corners

struct SearchCounters
{
    uint32_t probes = 0, cands = 0, levels = 0;
};

// ---- sub-warp cooperative search --------------------------------------------------------------
// A query is processed by a GROUP of G consecutive lanes, G = 8, 16 or 32 (the smallest >= k: four,
// two or one query per warp). The group members hold the same query and
//   * read a voxel's points G at a time (one coalesced request per step);
//   * keep the K best of the current level DISTRIBUTED over the group, one 64-bit key per lane,
//     ascending by lane (lane r holds the r-th smallest): the K-th distance is one shuffle away,
//     exact at all times, and a new candidate enters with one shuffle-up (group_insert) — no
//     per-lane lists, no merge pass, two registers of state.
__device__ __forceinline__ unsigned long long point_key(float qx, float qy, float qz, const float4 p)
{
    const float d2 = dist2_ref(qx, qy, qz, p.x, p.y, p.z);
    return ((unsigned long long)__float_as_uint(d2) << 32) | (uint32_t)__float_as_int(p.w);
}

// position of offset (dz,dy,dx in 0..2, index dz*9+dy*3+dx) inside kNeighbourOrder
__host__ __device__ constexpr int neighbour_rank(int idx)
{
    constexpr int order[27] = {0x15, 0x14, 0x16, 0x11, 0x19, 0x05, 0x25, 0x10, 0x12, 0x18, 0x1a, 0x04, 0x06, 0x24,
                               0x26, 0x01, 0x09, 0x21, 0x29, 0x00, 0x02, 0x08, 0x0a, 0x20, 0x22, 0x28, 0x2a};
    const int     code      = (idx % 3) | (((idx / 3) % 3) << 2) | ((idx / 9) << 4);
    for (int i = 0; i < 27; i++)
        if (order[i] == code) return i;
    return 0;
}

// Exact k nearest neighbours (k <= G) of (qx,qy,qz) within radius2 (strict <), by (d2, index).
// MUST be called by all 32 lanes of the warp, converged (`enabled` = false for lanes without a
// query): the groups of a warp run ONE control flow — every loop below is warp-uniform and the
// bodies are predicated per group — so the 32/G queries of a warp share every instruction issue
// and their memory requests go out together. (Letting each group run its own loops serialises the
// groups: measured 0.9k warp-instructions per QUERY, profiles/r01_c3_search_*.txt.)
// `sub` = lane index inside the group. On return lane r of the group holds the r-th best key in
// `mine` (ascending); keys >= (radius2 bits << 32) are "not found".
// `rl_start`: relative level to start from — the finest level whose voxels hold about 0.75 k points
// on average (start_level(), host side), so that the centre voxel alone usually settles the k-th
// distance and the neighbours can be pruned; any start level is correct.
constexpr uint32_t kLongRun = 96;                 // points from which a run is scanned by the whole warp
template <int G>
__device__ __forceinline__ void knn_search_v1(const GridView& g, bool enabled, float qx, float qy, float qz,
                                              float radius2, int K, int rl_start, unsigned long long& mine, int sub,
                                              SearchCounters& sc)
{
    constexpr unsigned       FULL     = 0xffffffffu;
    const int                lane     = threadIdx.x & 31;
    const unsigned           gmask    = (G == 32 ? FULL : ((1u << (G & 31)) - 1u)) << (lane - sub);
    const int                kth_lane = (lane - sub) + K - 1;  // warp lane holding the K-th best
    const unsigned long long sentinel = (unsigned long long)__float_as_uint(radius2) << 32;
    mine                              = sentinel;
    bool live = enabled && radius2 > 0.f;  // group-uniform
    if (live)
    {
        // reject queries farther than the radius from the map bbox (conservative: strictly greater)
        const float ex = fmaxf(fmaxf(g.bbmin[0] - qx, qx - g.bbmax[0]), 0.f);
        const float ey = fmaxf(fmaxf(g.bbmin[1] - qy, qy - g.bbmax[1]), 0.f);
        const float ez = fmaxf(fmaxf(g.bbmin[2] - qz, qz - g.bbmax[2]), 0.f);
        if ((ex * ex + ey * ey + ez * ez) * 0.999999f > radius2) live = false;
    }

    const float lim = 4194304.f;  // 2^22
    const float ux  = fminf(fmaxf(grid_u(qx, g.ox, g.inv_s0), -lim), lim);
    const float uy  = fminf(fmaxf(grid_u(qy, g.oy, g.inv_s0), -lim), lim);
    const float uz  = fminf(fmaxf(grid_u(qz, g.oz, g.inv_s0), -lim), lim);
    const int   Ix = (int)floorf(ux), Iy = (int)floorf(uy), Iz = (int)floorf(uz);
    const float q2 = g.s0_lo * g.s0_lo * 0.999999f;  // quanta^2 -> metres^2, rounded down

    float kth = radius2;  // upper bound of the K-th best distance found so far (all levels)

    // Offer the group's run of `count` consecutive points (count = 0: nothing for this group), G per
    // step; all groups of the warp step together. The K best stay distributed over the group, one
    // key per lane, ascending: the K-th key is one shuffle away, a passing candidate enters with one
    // shuffle-up — the lanes with cc < mine form a suffix of the group, its first lane takes cc, the
    // others take their left neighbour's key.
    auto scan_run = [&](const float4* __restrict__ run, uint32_t count)
    {
        // A LONG run (a coarse voxel of a query that had to climb) is scanned by the whole warp for
        // its owner group — 32 points per step, kAhead steps of loads in flight — instead of by the
        // G lanes of the group while the other groups of the warp wait: such queries are rare but a
        // single one otherwise outlives the rest of the launch.
        if (G < 32)
        {
            unsigned big = __ballot_sync(FULL, count >= kLongRun && sub == 0);
            while (big)  // warp-uniform
            {
                const int o = __ffs(big) - 1;  // first lane of the owner group
                big &= big - 1;
                const float4* __restrict__ orun = reinterpret_cast<const float4*>(
                    __shfl_sync(FULL, reinterpret_cast<unsigned long long>(run), o));
                const uint32_t ocount = __shfl_sync(FULL, count, o);
                const float    oqx = __shfl_sync(FULL, qx, o), oqy = __shfl_sync(FULL, qy, o), oqz = __shfl_sync(FULL, qz, o);
                const float    okth  = __shfl_sync(FULL, kth, o);
                const int      okl   = o + K - 1;                   // lane holding the owner's K-th key
                const bool     owner = (unsigned)(lane - o) < (unsigned)G;
                constexpr int  kAheadW = 4;
                for (uint32_t j0 = 0; j0 < ocount; j0 += kAheadW * 32)
                {
                    float4 p[kAheadW];
#pragma unroll
                    for (int u = 0; u < kAheadW; u++)
                    {
                        const uint32_t j = j0 + u * 32 + lane;
                        p[u]             = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (j < ocount) p[u] = __ldg(orun + j);
                    }
#pragma unroll
                    for (int u = 0; u < kAheadW; u++)
                    {
                        if (j0 + u * 32 >= ocount) break;  // warp-uniform
                        const uint32_t           j    = j0 + u * 32 + lane;
                        const bool               in   = j < ocount;
                        const unsigned long long c    = in ? point_key(oqx, oqy, oqz, p[u]) : ~0ull;
                        const unsigned long long kkey = __shfl_sync(FULL, mine, okl);
                        const bool pass = in && c < kkey && __uint_as_float((uint32_t)(c >> 32)) <= okth;
                        unsigned   pm   = __ballot_sync(FULL, pass);
                        while (pm)  // warp-uniform
                        {
                            const int src = __ffs(pm) - 1;
                            pm &= pm - 1;
                            const unsigned long long cc = __shfl_sync(FULL, c, src);
                            const unsigned long long up = __shfl_up_sync(FULL, mine, 1, G);
                            if (owner && cc < mine) mine = (sub == 0 || !(cc < up)) ? cc : up;
                        }
                    }
                }
                const unsigned long long kkey = __shfl_sync(FULL, mine, okl);
                if (owner) kth = fminf(kth, __uint_as_float((uint32_t)(kkey >> 32))), count = 0;
            }
        }
        const uint32_t steps = __reduce_max_sync(FULL, count);
        constexpr int  kAhead = 4;  // steps whose loads are issued together: a long run (a coarse
                                    // voxel of a query that had to climb) is latency-bound otherwise
        for (uint32_t j0 = 0; j0 < steps; j0 += kAhead * G)
        {
            float4 p[kAhead];
#pragma unroll
            for (int u = 0; u < kAhead; u++)
            {
                const uint32_t j = j0 + u * G + sub;
                p[u]             = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j < count) p[u] = __ldg(run + j);
            }
#pragma unroll
            for (int u = 0; u < kAhead; u++)
            {
                if (j0 + u * G >= steps) break;  // warp-uniform
                const uint32_t           j    = j0 + u * G + sub;
                const bool               in   = j < count;
                const unsigned long long c    = in ? point_key(qx, qy, qz, p[u]) : ~0ull;
                const unsigned long long kkey = __shfl_sync(FULL, mine, kth_lane);
                const bool pass = in && c < kkey && __uint_as_float((uint32_t)(c >> 32)) <= kth;
                unsigned   pm   = __ballot_sync(FULL, pass) & gmask;
                while (__any_sync(FULL, pm != 0))
                {
                    const bool ins = pm != 0;
                    const int  src = ins ? __ffs(pm) - 1 : lane;
                    pm &= pm - 1;
                    const unsigned long long cc = __shfl_sync(FULL, c, src);
                    const unsigned long long up = __shfl_up_sync(FULL, mine, 1, G);  // left neighbour's key
                    if (ins && cc < mine) mine = (sub == 0 || !(cc < up)) ? cc : up;
                }
            }
        }
        const unsigned long long kkey = __shfl_sync(FULL, mine, kth_lane);
        if (count) kth = fminf(kth, __uint_as_float((uint32_t)(kkey >> 32)));
    };

    for (int rl = rl_start; rl < g.n_levels; rl++)
    {
        if (!__any_sync(FULL, live)) break;
        const int L = g.level_first + rl;
        if (live) mine = sentinel;  // the list is rebuilt at every level: no key is ever offered twice

        if (L == kGridBits)
        {
            // top level: the single voxel holds every point; a query outside the grid (possible only
            // with a radius larger than its distance to the bbox) must still see all of them
            if (live && sub == 0) sc.probes++, sc.cands += g.n_points, sc.levels++;
            scan_run(g.pts, live ? g.n_points : 0u);
            break;
        }

        const int   cmax = ((1 << kGridBits) - 1) >> L;
        const float s    = (float)(1 << L);  // voxel edge in finest quanta
        const int   cx = Ix >> L, cy = Iy >> L, cz = Iz >> L;
        const float fx = ux - (float)cx * s, fy = uy - (float)cy * s, fz = uz - (float)cz * s;
        // per-axis gap (in quanta, made conservative) to the -1 / +1 neighbour slabs
        const float gxl = fmaxf(fx - 4.f, 0.f), gxh = fmaxf(s - fx - 4.f, 0.f);
        const float gyl = fmaxf(fy - 4.f, 0.f), gyh = fmaxf(s - fy - 4.f, 0.f);
        const float gzl = fmaxf(fz - 4.f, 0.f), gzh = fmaxf(s - fz - 4.f, 0.f);

        // The 3x3x3 block of this level, centre first then faces/edges/corners. Per round a group
        // takes up to G voxels of its to-do mask, ONE PER LANE: each lane bounds and hash-probes its
        // own voxel (the probes of a round are in flight together — a dependent chain of one probe
        // per voxel is what a query that climbs through sparse space would otherwise pay), then the
        // runs found are offered one after the other, each re-checked against the bound the
        // previous ones tightened. After the centre, the survivors are collected once in the mask
        // (hierarchically: a slab or a row that is too far drops all its voxels at once).
        const float ax[3] = {gxl * gxl * q2, 0.f, gxh * gxh * q2};
        const float ay[3] = {gyl * gyl * q2, 0.f, gyh * gyh * q2};
        const float az[3] = {gzl * gzl * q2, 0.f, gzh * gzh * q2};
        uint32_t    todo  = live ? 1u : 0u;  // bit i <-> kNeighbourOrder[i]; start with the centre
        bool        first = live;
        while (__any_sync(FULL, todo != 0))
        {
            // lane `sub` takes the sub-th pending voxel
            const int      n_take = min(__popc(todo), G);
            uint32_t       start = 0, count = 0;
            float          lb    = 0.f;
            uint32_t       taken = 0;  // bit of my voxel
            if (sub < n_take)
            {
                const int nb = (int)__fns(todo, 0, sub + 1);
                taken        = 1u << nb;
                const uint32_t code = kNeighbourOrder[nb];
                const int      dx = (int)(code & 3u) - 1, dy = (int)((code >> 2) & 3u) - 1,
                          dz = (int)((code >> 4) & 3u) - 1;
                lb = (dx < 0 ? ax[0] : (dx > 0 ? ax[2] : 0.f)) + (dy < 0 ? ay[0] : (dy > 0 ? ay[2] : 0.f)) +
                     (dz < 0 ? az[0] : (dz > 0 ? az[2] : 0.f));
                const int nx = cx + dx, ny = cy + dy, nz = cz + dz;
                // strict `>`: an equal-distance lower index must still be seen
                if (!(lb > kth) && (unsigned)nx <= (unsigned)cmax && (unsigned)ny <= (unsigned)cmax &&
                    (unsigned)nz <= (unsigned)cmax)
                {
                    sc.probes++;
                    if (!grid_lookup(g, rl, (uint32_t)nx, (uint32_t)ny, (uint32_t)nz, start, count)) count = 0;
                }
            }
            // drop the voxels taken this round from the group's mask (OR over the group's lanes)
            {
                uint32_t t = taken;
#pragma unroll
                for (int o = G / 2; o > 0; o >>= 1) t |= __shfl_xor_sync(FULL, t, o);
                todo &= ~t;
            }
            // offer the runs in voxel order; a run whose bound fell behind is skipped
            const int rounds = __reduce_max_sync(FULL, (unsigned)n_take);
            for (int t = 0; t < rounds; t++)
            {
                const int      src = (lane - sub) + t;
                const uint32_t rs = __shfl_sync(FULL, start, src), rc = __shfl_sync(FULL, count, src);
                const float    rlb = __shfl_sync(FULL, lb, src);
                const uint32_t c   = (t < n_take && !(rlb > kth)) ? rc : 0u;
                if (sub == 0) sc.cands += c;
                scan_run(g.pts + rs, c);
            }
            if (first)
            {
                first = false;
#pragma unroll
                for (int dz3 = 0; dz3 < 3; dz3++)
                {
                    if (az[dz3] > kth) continue;
#pragma unroll
                    for (int dy3 = 0; dy3 < 3; dy3++)
                    {
                        const float t = az[dz3] + ay[dy3];
                        if (t > kth) continue;
#pragma unroll
                        for (int dx3 = 0; dx3 < 3; dx3++)
                        {
                            if (dx3 == 1 && dy3 == 1 && dz3 == 1) continue;
                            if (t + ax[dx3] > kth) continue;
                            todo |= 1u << neighbour_rank(dz3 * 9 + dy3 * 3 + dx3);
                        }
                    }
                }
            }
        }
        // `mine` now holds the exact K best of this level's block; everything outside the 3x3x3
        // block is at least m quanta away
        const float mx = s + fminf(fx, s - fx), my = s + fminf(fy, s - fy), mz = s + fminf(fz, s - fz);
        const float m  = fmaxf(fminf(mx, fminf(my, mz)) - 4.f, 0.f);
        if (live)
        {
            if (sub == 0) sc.levels++;
            if (kth <= m * m * q2) live = false;
        }
    }
}

// ---- dense-list search with pruned descent (round 2) ------------------------------------------
// Same index, same exactness argument, same group layout (G lanes per query, the K best keys
// distributed over the group in ascending order) as knn_search_v1. What changes:
//  (1) DENSE LISTS. v1 offered every voxel's run on its own (a warp-uniform call per run, ~9 per
//      level, each with its fixed cost and at least one scan step even when three of the four groups
//      of the warp had nothing to offer). Here the voxels of a ROUND — the centre voxel; then the
//      neighbours the K-th distance found so far cannot exclude, the 6 faces before the 20 edges /
//      corners; then whatever the descent stack holds — are taken by the lanes of the group by
//      static assignment (voxel n -> lane n % G, no bit-select), bounded, hash-probed with all probes
//      of a lane in flight, and the runs found go into the group's RUN TABLE in shared memory
//      (ballot-ranked positions, one in-group prefix scan: end[r] = inclusive prefix of the run
//      lengths, adj[r] = start - exclusive prefix, so candidate v of the concatenated list sits at
//      pts[adj[r] + v]). The list is scanned densely: step s offers candidates [s*G, s*G + G)
//      whatever run they belong to (each lane walks the table with a cursor), four steps of loads in
//      flight. A long list is scanned by the whole warp for its owner group, 32 candidates a step.
//  (2) PRUNED DESCENT. A query that has to climb (a far return the pose error moved off its surface:
//      7 % of a C3 scan) met voxels of hundreds to thousands of points in v1 and read all of them —
//      half of all candidates of the launch, and whole CTAs of such queries formed its tail (ncu:
//      one SM busy for 400 k cycles, the average 284 k). Here a voxel holding more than kDescend
//      points is not scanned: its 8 children (one table level down — every level's voxel is the
//      union of its children, and their runs are sub-runs of its run) go on the group's stack, are
//      bounded against the current K-th distance, probed, and scanned or split again. Depth first,
//      so the first leaves tighten the bound for everything still on the stack. The candidates of a
//      climbing query scale with the surface inside its search sphere, not with the coarse block.
// Exactness: within one level every voxel of the 3x3x3 block is scanned, or excluded by a lower
// bound that exceeds an upper bound of the K-th distance, or replaced by its children, which
// partition it; no point is offered twice (the list is rebuilt per level, as before).
constexpr uint32_t kDescend  = 64;  // points above which a voxel is split instead of scanned
constexpr int      kStackCap = 64;  // descent stack entries per group (overflow: the voxel is scanned)
struct RunTable  // one per query group, shared memory
{
    uint32_t           end[36];
    uint32_t           adj[36];
    unsigned long long stk_key[kStackCap];  // cell_key of a voxel waiting to be visited
    uint8_t            stk_rl[kStackCap];   // its relative level
};
template <int G, int NT = (int)kQueryTile>
struct KnnShared
{
    RunTable tb[NT / G];  // the groups of warp w own tb[w * 32 / G ...]
    uint8_t  nb[32];      // kNeighbourOrder (divergent lookups: shared memory, not the constant bank)
};
// every thread of the CTA, followed by a __syncthreads() before the first search
template <int G, int NT>
__device__ __forceinline__ void knn_shared_init(KnnShared<G, NT>& ks)
{
    if (threadIdx.x < 27) ks.nb[threadIdx.x] = kNeighbourOrder[threadIdx.x];
}

template <int G>
__device__ __forceinline__ void knn_search(const GridView& g, bool enabled, float qx, float qy, float qz,
                                           float radius2, int K, int rl_start, unsigned long long& mine, int sub,
                                           SearchCounters& sc, RunTable* tables, const uint8_t* nb_order, int n_phases)
{
    constexpr unsigned       FULL     = 0xffffffffu;
    constexpr int            T        = (26 + G - 1) / G;  // voxels a lane may have to take per round
    const int                lane     = threadIdx.x & 31;
    const int                gbase    = lane - sub;
    const unsigned           gmask    = (G == 32 ? FULL : ((1u << (G & 31)) - 1u)) << gbase;
    const unsigned           below    = gmask & ((1u << lane) - 1u);
    const int                kth_lane = gbase + K - 1;  // warp lane holding the K-th best
    RunTable* const          tb_warp  = tables + (threadIdx.x >> 5) * (32 / G);  // the tables of this warp's groups
    RunTable&                tb       = tb_warp[gbase / G];
    const unsigned long long sentinel = (unsigned long long)__float_as_uint(radius2) << 32;
    mine                              = sentinel;
    bool live = enabled && radius2 > 0.f;  // group-uniform
    if (live)
    {
        // reject queries farther than the radius from the map bbox (conservative: strictly greater)
        const float ex = fmaxf(fmaxf(g.bbmin[0] - qx, qx - g.bbmax[0]), 0.f);
        const float ey = fmaxf(fmaxf(g.bbmin[1] - qy, qy - g.bbmax[1]), 0.f);
        const float ez = fmaxf(fmaxf(g.bbmin[2] - qz, qz - g.bbmax[2]), 0.f);
        if ((ex * ex + ey * ey + ez * ez) * 0.999999f > radius2) live = false;
    }
    const float lim = 4194304.f;  // 2^22
    const float ux  = fminf(fmaxf(grid_u(qx, g.ox, g.inv_s0), -lim), lim);
    const float uy  = fminf(fmaxf(grid_u(qy, g.oy, g.inv_s0), -lim), lim);
    const float uz  = fminf(fmaxf(grid_u(qz, g.oz, g.inv_s0), -lim), lim);
    const int   Ix = (int)floorf(ux), Iy = (int)floorf(uy), Iz = (int)floorf(uz);
    const float q2 = g.s0_lo * g.s0_lo * 0.999999f;  // quanta^2 -> metres^2, rounded down
    float       kth = radius2;  // upper bound of the K-th best distance found so far (all levels)
    int         sp  = 0;        // descent stack height (group-uniform)

    // one offer of a candidate key to the owner group's list (all 32 lanes; `ins` = this lane's group
    // takes part): lanes with cc < mine form a suffix of the group, its first lane takes cc, the others
    // take their left neighbour's key
#define MP2P_KNN_INSERT(ins_, cc_)                                                     \
    {                                                                                  \
        const unsigned long long up_ = __shfl_up_sync(FULL, mine, 1, G);               \
        if ((ins_) && (cc_) < mine) mine = (sub == 0 || !((cc_) < up_)) ? (cc_) : up_; \
    }

    // lower bound (metres^2, conservative by 4 quanta per axis) of the distance from the query to the
    // voxel (vx,vy,vz) of absolute level L; for the neighbours of the query's own voxel this is the
    // per-axis gap formula of v1 (fx = ux - cx * s is the gap to the -1 slab, s - fx to the +1 slab)
    auto voxel_bound = [&](int L, int vx, int vy, int vz) -> float
    {
        const float s  = (float)(1 << L);
        const float lx = (float)vx * s, ly = (float)vy * s, lz = (float)vz * s;
        const float gx = fmaxf(fmaxf(lx - ux, ux - (lx + s)) - 4.f, 0.f);
        const float gy = fmaxf(fmaxf(ly - uy, uy - (ly + s)) - 4.f, 0.f);
        const float gz = fmaxf(fmaxf(lz - uz, uz - (lz + s)) - 4.f, 0.f);
        return gx * gx * q2 + gy * gy * q2 + gz * gz * q2;
    };

    // scan of the round's list (table already in shared memory); total = its length for this group
    auto dense_scan = [&](uint32_t total)
    {
        if (G < 32)
        {
            unsigned big = __ballot_sync(FULL, total >= kLongRun && sub == 0);
            while (big)  // warp-uniform: the whole warp scans the list of the group starting at lane o
            {
                const int o = __ffs(big) - 1;
                big &= big - 1;
                const RunTable& ot     = tb_warp[o / G];
                const uint32_t  ototal = __shfl_sync(FULL, total, o);
                const float     oqx = __shfl_sync(FULL, qx, o), oqy = __shfl_sync(FULL, qy, o), oqz = __shfl_sync(FULL, qz, o);
                const float     okth  = __shfl_sync(FULL, kth, o);
                const int       okl   = o + K - 1;
                const bool      owner = (unsigned)(lane - o) < (unsigned)G;
                constexpr int   kAheadW = 4;
                uint32_t        r = 0, re = ot.end[0], ra = ot.adj[0];
                for (uint32_t j0 = 0; j0 < ototal; j0 += kAheadW * 32)
                {
                    float4 p[kAheadW];
#pragma unroll
                    for (int u = 0; u < kAheadW; u++)
                    {
                        const uint32_t v = j0 + u * 32 + lane;
                        p[u]             = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (v < ototal)
                        {
                            while (v >= re) r++, re = ot.end[r], ra = ot.adj[r];
                            p[u] = __ldg(g.pts + (uint32_t)(ra + v));
                        }
                    }
#pragma unroll
                    for (int u = 0; u < kAheadW; u++)
                    {
                        if (j0 + u * 32 >= ototal) break;  // warp-uniform
                        const bool               in   = j0 + u * 32 + lane < ototal;
                        const unsigned long long c    = in ? point_key(oqx, oqy, oqz, p[u]) : ~0ull;
                        const unsigned long long kkey = __shfl_sync(FULL, mine, okl);
                        const bool pass = in && c < kkey && __uint_as_float((uint32_t)(c >> 32)) <= okth;
                        unsigned   pm   = __ballot_sync(FULL, pass);
                        while (pm)  // warp-uniform
                        {
                            const int src = __ffs(pm) - 1;
                            pm &= pm - 1;
                            const unsigned long long cc = __shfl_sync(FULL, c, src);
                            MP2P_KNN_INSERT(owner, cc)
                        }
                    }
                }
                const unsigned long long kkey = __shfl_sync(FULL, mine, okl);
                if (owner) kth = fminf(kth, __uint_as_float((uint32_t)(kkey >> 32))), total = 0;
            }
        }
        const uint32_t steps = __reduce_max_sync(FULL, total);
        constexpr int  kAhead = 4;  // steps whose loads are issued together
        uint32_t       r = 0, re = tb.end[0], ra = tb.adj[0];  // stale values if total == 0: never used then
        for (uint32_t j0 = 0; j0 < steps; j0 += kAhead * G)
        {
            float4 p[kAhead];
#pragma unroll
            for (int u = 0; u < kAhead; u++)
            {
                const uint32_t v = j0 + u * G + sub;
                p[u]             = make_float4(0.f, 0.f, 0.f, 0.f);
                if (v < total)
                {
                    while (v >= re) r++, re = tb.end[r], ra = tb.adj[r];
                    p[u] = __ldg(g.pts + (uint32_t)(ra + v));
                }
            }
#pragma unroll
            for (int u = 0; u < kAhead; u++)
            {
                if (j0 + u * G >= steps) break;  // warp-uniform
                const bool               in   = j0 + u * G + sub < total;
                const unsigned long long c    = in ? point_key(qx, qy, qz, p[u]) : ~0ull;
                const unsigned long long kkey = __shfl_sync(FULL, mine, kth_lane);
                const bool pass = in && c < kkey && __uint_as_float((uint32_t)(c >> 32)) <= kth;
                unsigned   pm   = __ballot_sync(FULL, pass) & gmask;
                while (__any_sync(FULL, pm != 0))
                {
                    const bool ins = pm != 0;
                    const int  src = ins ? __ffs(pm) - 1 : lane;
                    pm &= pm - 1;
                    const unsigned long long cc = __shfl_sync(FULL, c, src);
                    MP2P_KNN_INSERT(ins, cc)
                }
            }
        }
        const unsigned long long kkey = __shfl_sync(FULL, mine, kth_lane);
        if (total) kth = fminf(kth, __uint_as_float((uint32_t)(kkey >> 32)));
    };

    // One ROUND: every lane takes up to T voxels — from the 3x3x3 neighbourhood of the level's centre
    // voxel (from_stack = false: neighbour numbers [lo, hi) of kNeighbourOrder) or from the top of the
    // group's descent stack — bounds them against kth, probes the survivors, splits the large ones
    // (8 children onto the stack) and lists the others in the run table. Returns the list's length.
    // (rl, cx, cy, cz = the level being searched and its centre voxel; only used if !from_stack.)
    auto round = [&](bool from_stack, int lo, int hi, bool want, int rl, int cx, int cy, int cz) -> uint32_t
    {
        uint32_t           st[T], cn[T], h[T];
        unsigned long long ckey[T];
        uint4              raw[T];
        int                erl[T];
        bool               need[T];
        const int          n_pop = from_stack ? min(sp, G * T) : 0;
#pragma unroll
        for (int t = 0; t < T; t++)
        {
            const int n = sub + G * t;
            st[t] = cn[t] = h[t] = 0u, ckey[t] = 0ull, need[t] = false, erl[t] = rl, raw[t] = make_uint4(0u, 0u, 0u, 0u);
            int  vx = 0, vy = 0, vz = 0;
            bool have = false;
            if (from_stack)
            {
                if (n < n_pop)
                {
                    const unsigned long long e = tb.stk_key[sp - 1 - n];
                    erl[t]                     = tb.stk_rl[sp - 1 - n];
                    vx = (int)(e & 0x1fffffu), vy = (int)((e >> 21) & 0x1fffffu), vz = (int)((e >> 42) & 0x1fffffu);
                    have = true;
                }
            }
            else if (want && lo + n < hi)
            {
                const uint32_t code = nb_order[lo + n];
                vx = cx + (int)(code & 3u) - 1, vy = cy + (int)((code >> 2) & 3u) - 1, vz = cz + (int)((code >> 4) & 3u) - 1;
                const int cmax = ((1 << kGridBits) - 1) >> (g.level_first + rl);
                have = (unsigned)vx <= (unsigned)cmax && (unsigned)vy <= (unsigned)cmax && (unsigned)vz <= (unsigned)cmax;
            }
            // strict `>`: an equal-distance lower index must still be seen
            if (have && !(voxel_bound(g.level_first + erl[t], vx, vy, vz) > kth))
            {
                need[t] = true;
                ckey[t] = cell_key((uint32_t)vx, (uint32_t)vy, (uint32_t)vz);
                h[t]    = cell_hash(ckey[t], g.level_shift[erl[t]]);
                raw[t]  = __ldg(reinterpret_cast<const uint4*>(g.table + g.level_off[erl[t]] + h[t]));  // the lane's probes go out together
            }
        }
        sp -= n_pop;
        __syncwarp();  // the popped entries are in registers before anybody pushes over them
#pragma unroll
        for (int t = 0; t < T; t++)
        {
            if (!need[t]) continue;
            sc.probes++;
            const uint32_t   hmask = (1u << (64 - g.level_shift[erl[t]])) - 1u;
            const CellEntry* tab   = g.table + g.level_off[erl[t]];
            while (true)  // linear probing continues on a collision
            {
                const unsigned long long k = (unsigned long long)raw[t].x | ((unsigned long long)raw[t].y << 32);
                if (k == ckey[t])
                {
                    st[t] = raw[t].z, cn[t] = raw[t].w;
                    break;
                }
                if (k == kEmptyKey) break;
                h[t]   = (h[t] + 1) & hmask;
                raw[t] = __ldg(reinterpret_cast<const uint4*>(tab + h[t]));
            }
        }
        // large voxels are split (while the stack has room), the others listed: positions by ballot rank
        uint32_t n_runs = 0, pos[T];
        int      sp_new = sp;
        bool     split[T];
#pragma unroll
        for (int t = 0; t < T; t++)
        {
            const bool     large = cn[t] > kDescend && erl[t] > 0;
            const unsigned bl    = __ballot_sync(FULL, large) & gmask;
            const int      at    = sp_new + 8 * __popc(bl & below);
            split[t]             = large && at + 8 <= kStackCap;
            // (lanes are served in order, so the ones that fit form a prefix of the large ones)
            const unsigned bs = __ballot_sync(FULL, split[t]) & gmask;
            if (split[t])
            {
                const unsigned long long e  = ckey[t];
                const unsigned long long x2 = (e & 0x1fffffull) << 1, y2 = ((e >> 21) & 0x1fffffull) << 1, z2 = ((e >> 42) & 0x1fffffull) << 1;
#pragma unroll
                for (int c = 0; c < 8; c++)
                {
                    tb.stk_key[at + c] = (x2 | (unsigned long long)(c & 1)) | ((y2 | (unsigned long long)((c >> 1) & 1)) << 21) |
                                         ((z2 | (unsigned long long)(c >> 2)) << 42);
                    tb.stk_rl[at + c] = (uint8_t)(erl[t] - 1);
                }
            }
            sp_new += 8 * __popc(bs);
            const bool     listed = cn[t] != 0u && !split[t];
            const unsigned b      = __ballot_sync(FULL, listed) & gmask;
            pos[t]                = n_runs + __popc(b & below);
            n_runs += __popc(b);
            if (listed) sc.cands += cn[t];
        }
        sp = sp_new;
        __syncwarp();
#pragma unroll
        for (int t = 0; t < T; t++)
            if (cn[t] != 0u && !split[t]) tb.end[pos[t]] = cn[t], tb.adj[pos[t]] = st[t];
        __syncwarp();
        // lengths -> inclusive prefix, starts -> start - exclusive prefix; G entries per pass
        uint32_t       carry  = 0;
        const uint32_t passes = __reduce_max_sync(FULL, (n_runs + G - 1) / G);
        for (uint32_t c = 0; c < passes; c++)
        {
            const uint32_t e    = c * G + sub;
            const uint32_t x    = e < n_runs ? tb.end[e] : 0u;
            uint32_t       incl = x;
#pragma unroll
            for (int o = 1; o < G; o <<= 1)
            {
                const uint32_t y = __shfl_up_sync(FULL, incl, o, G);
                if (sub >= o) incl += y;
            }
            incl += carry;
            if (e < n_runs) tb.end[e] = incl, tb.adj[e] -= incl - x;
            carry = __shfl_sync(FULL, incl, gbase + G - 1);
        }
        __syncwarp();
        return carry;
    };

    // visits everything the rounds so far put on the stacks, depth first
    auto drain = [&]()
    {
        while (__any_sync(FULL, sp > 0))
        {
            const uint32_t total = round(true, 0, 0, false, 0, 0, 0, 0);
            dense_scan(total);
        }
    };

    for (int rl = rl_start; rl < g.n_levels; rl++)
    {
        if (!__any_sync(FULL, live)) break;
        const int  L   = g.level_first + rl;
        const bool top = L == kGridBits;  // the single voxel holding every point
        if (live) mine = sentinel;        // the list is rebuilt at every level: no key is ever offered twice

        const int   cmax = ((1 << kGridBits) - 1) >> L;
        const float s    = (float)(1 << L);  // voxel edge in finest quanta
        const int   cx = Ix >> L, cy = Iy >> L, cz = Iz >> L;
        const float fx = ux - (float)cx * s, fy = uy - (float)cy * s, fz = uz - (float)cz * s;
        // smallest bound any face voxel / any edge or corner voxel can have (per-axis gaps as in voxel_bound)
        const float gxl = fmaxf(fx - 4.f, 0.f), gxh = fmaxf(s - fx - 4.f, 0.f);
        const float gyl = fmaxf(fy - 4.f, 0.f), gyh = fmaxf(s - fy - 4.f, 0.f);
        const float gzl = fmaxf(fz - 4.f, 0.f), gzh = fmaxf(s - fz - 4.f, 0.f);
        const float mnx = fminf(gxl * gxl * q2, gxh * gxh * q2), mny = fminf(gyl * gyl * q2, gyh * gyh * q2),
                    mnz = fminf(gzl * gzl * q2, gzh * gzh * q2);
        const float min_face = fminf(mnx, fminf(mny, mnz));
        const float min_edge = fminf(mnx + mny, fminf(mnx + mnz, mny + mnz)) * 0.999999f;

        // ---- centre voxel: every lane of the group makes the same probe (one broadcast request)
        {
            uint32_t start = 0, count = 0;
            if (top)
            {
                if (live) count = g.n_points;
                if (live && sub == 0) sc.probes++;
            }
            else if (live && (unsigned)cx <= (unsigned)cmax && (unsigned)cy <= (unsigned)cmax && (unsigned)cz <= (unsigned)cmax)
            {
                if (!grid_lookup(g, rl, (uint32_t)cx, (uint32_t)cy, (uint32_t)cz, start, count)) count = 0;
                if (sub == 0) sc.probes++;
            }
            const bool split = count > kDescend && rl > 0;  // (sp == 0 here: the stack has room)
            __syncwarp();
            if (split)
            {
                if (sub < 8)
                {
                    const unsigned long long x2 = (unsigned long long)(top ? 0 : cx) << 1, y2 = (unsigned long long)(top ? 0 : cy) << 1,
                                             z2 = (unsigned long long)(top ? 0 : cz) << 1;
                    tb.stk_key[sub] = (x2 | (unsigned long long)(sub & 1)) | ((y2 | (unsigned long long)((sub >> 1) & 1)) << 21) |
                                      ((z2 | (unsigned long long)(sub >> 2)) << 42);
                    tb.stk_rl[sub]  = (uint8_t)(rl - 1);
                }
                sp = 8, count = 0;
            }
            else if (sub == 0)
            {
                tb.end[0] = count, tb.adj[0] = start;
                sc.cands += count;
            }
            __syncwarp();
            dense_scan(count);
            drain();
        }
        // ---- the neighbours the bound cannot exclude: faces, then edges and corners (or all at once)
        if (!top)
        {
#pragma unroll 1
            for (int phase = 1; phase < n_phases; phase++)
            {
                const int   lo = phase == 1 ? 1 : 7, hi = (phase == 1 && n_phases == 3) ? 7 : 27;
                const float mn = phase == 1 ? min_face : min_edge;
                const bool  want = live && !(mn > kth);
                if (!__any_sync(FULL, want)) continue;  // nothing of this phase can matter to any group
                const uint32_t total = round(false, lo, hi, want, rl, cx, cy, cz);
                dense_scan(total);
                drain();
            }
        }
        if (top)
        {
            if (live && sub == 0) sc.levels++;
            break;
        }
        // `mine` now holds the exact K best of this level's block; everything outside the 3x3x3
        // block is at least m quanta away
        const float mx = s + fminf(fx, s - fx), my = s + fminf(fy, s - fy), mz = s + fminf(fz, s - fz);
        const float m  = fmaxf(fminf(mx, fminf(my, mz)) - 4.f, 0.f);
        if (live)
        {
            if (sub == 0) sc.levels++;
            if (kth <= m * m * q2) live = false;
        }
    }
#undef MP2P_KNN_INSERT
}


// warp-aggregated accumulation of the per-thread counters into stats[0..3] (measurement hook)
__device__ __forceinline__ void flush_search_stats(const SearchCounters& sc, uint32_t n_valid,
                                                   unsigned long long* stats)
{
    if (!stats) return;
    uint32_t a = sc.probes, b = sc.cands, c = n_valid, d = sc.levels > 1 ? 1u : 0u;
    const unsigned mask = __activemask();
    a = __reduce_add_sync(mask, a), b = __reduce_add_sync(mask, b);
    c = __reduce_add_sync(mask, c), d = __reduce_add_sync(mask, d);
    const uint32_t mc = __reduce_max_sync(mask, sc.cands), mp = __reduce_max_sync(mask, sc.probes);
    const uint32_t ml = __reduce_max_sync(mask, sc.levels);
    if ((threadIdx.x & 31) == (__ffs(mask) - 1))
    {
        atomicAdd(stats + 0, (unsigned long long)a), atomicAdd(stats + 1, (unsigned long long)b);
        atomicAdd(stats + 2, (unsigned long long)c), atomicAdd(stats + 3, (unsigned long long)d);
        atomicMax(stats + 4, (unsigned long long)mc), atomicMax(stats + 5, (unsigned long long)mp);
        atomicMax(stats + 6, (unsigned long long)ml);
        if (mc > 2000u) atomicAdd(stats + 7, 1ull);  // warps holding a query with > 2000 candidates
    }
}

}  // namespace mp2p
