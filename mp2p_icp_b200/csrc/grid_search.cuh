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
                                            uint32_t cz, uint32_t& start, uint32_t& count, uint32_t* slot = nullptr)
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
            if (slot) *slot = h;
            return true;
        }
        if (k == kEmptyKey) return false;
        h = (h + 1) & mask;
    }
}

// Lower bound (metres^2, conservative by the same 4 quanta per axis as the cube bounds below) of the distance from
// a query at (ux,uy,uz) finest quanta to the TIGHT box of the points of voxel (vx,vy,vz), absolute level L, found in
// slot `slot` of table rl (GridView::box; layout in index.cu). q2 = quanta^2 -> metres^2, rounded down.
constexpr uint32_t kBoxMin = 4;  // voxels with fewer points are scanned without looking at their box
__device__ __forceinline__ float box_bound(const GridView& g, int rl, uint32_t slot, int L, int vx, int vy, int vz, float ux,
                                           float uy, float uz, float q2)
{
    const uint2 b    = __ldg(g.box + g.level_off[rl] + slot);
    const float unit = (float)(1 << max(L - 8, 0)), s = (float)(1 << L);  // (all products below are integers < 2^24: exact)
    const float ox = (float)vx * s, oy = (float)vy * s, oz = (float)vz * s;
    const float lx = ox + (float)(b.x & 255u) * unit, hx = ox + (float)((b.y & 255u) + 1u) * unit;
    const float ly = oy + (float)((b.x >> 8) & 255u) * unit, hy = oy + (float)(((b.y >> 8) & 255u) + 1u) * unit;
    const float lz = oz + (float)((b.x >> 16) & 255u) * unit, hz = oz + (float)(((b.y >> 16) & 255u) + 1u) * unit;
    const float gx = fmaxf(fmaxf(lx - ux, ux - hx) - 4.f, 0.f);
    const float gy = fmaxf(fmaxf(ly - uy, uy - hy) - 4.f, 0.f);
    const float gz = fmaxf(fmaxf(lz - uz, uz - hz) - 4.f, 0.f);
    return gx * gx * q2 + gy * gy * q2 + gz * gz * q2;
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
    uint32_t rounds = 0, steps = 0, inserts = 0;  // warp-uniform loop trip counts (trace hook)
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

// ---- stack-driven search with pruned descent (round 2) ----------------------------------------
// Same index, same exactness argument, same group layout (G lanes per query, the K best keys
// distributed over the group in ascending order) as knn_search_v1. What changes:
//  (1) ONE MECHANISM. Every voxel that may matter — the centre voxel of a level, then its face
//      neighbours, then its edge / corner neighbours, then the children of a voxel that was split —
//      is an entry (cell key, level) on the group's STACK in shared memory. A ROUND pops up to G
//      entries, one per lane: the lane bounds its voxel against the K-th distance found so far,
//      hash-probes it if it survives (the probes of a round are in flight together), and either
//      lists the run found in the group's RUN TABLE or splits it. The table (ballot-ranked
//      positions, one in-group prefix scan: end[r] = inclusive prefix of the run lengths, adj[r] =
//      start - exclusive prefix, so candidate v of the round's concatenated list sits at
//      pts[adj[r] + v]) is then scanned DENSELY: step s offers candidates [s*G, s*G + G) whatever
//      run they belong to, four steps of loads in flight. v1 offered every run on its own (~9
//      warp-uniform calls per level, each with its fixed cost and at least one scan step even when
//      three of the four groups of the warp had nothing to offer). The kernel holds one copy of the
//      round and one of the scan: its code is half of v1's (the first cut of this design inlined
//      them per phase, 108 KB of SASS, and stalled on instruction fetch).
//  (2) PRUNED DESCENT. A query that has to climb (a far return the pose error moved off its surface:
//      7 % of a C3 scan) met voxels of hundreds to thousands of points in v1 and read all of them —
//      half of all candidates of the launch, and whole CTAs of such queries formed its tail (ncu:
//      one SM busy for 400 k cycles, the average 284 k). Here a voxel holding more than kDescend
//      points is not scanned: its 8 children (one table level down — a voxel is the union of its
//      children and their runs are sub-runs of its run) go on the stack. Depth first, so the first
//      leaves tighten the bound for everything still waiting.
// Exactness: within one level every voxel of the 3x3x3 block is scanned, or excluded by a lower
// bound that exceeds an upper bound of the K-th distance, or replaced by its children, which
// partition it; no point is offered twice (the list is rebuilt per level, as before).
constexpr uint32_t kDescend  = 64;  // points above which a voxel is split instead of scanned
constexpr uint32_t kLongList = 64;  // candidates from which a round's list is scanned by the whole warp
constexpr int      kStackCap = 64;  // stack entries per group (a voxel that cannot be split for lack of room is scanned)
struct RunTable  // one per query group, shared memory
{
    uint32_t           end[32];
    uint32_t           adj[32];
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
    const int                lane     = threadIdx.x & 31;
    const int                gbase    = lane - sub;
    const unsigned           gmask    = (G == 32 ? FULL : ((1u << (G & 31)) - 1u)) << gbase;
    const unsigned           below    = gmask & ((1u << lane) - 1u);
    const int                kth_lane = gbase + K - 1;  // warp lane holding the K-th best
    RunTable&                tb       = tables[(threadIdx.x >> 5) * (32 / G) + gbase / G];
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
    int         sp  = 0;        // stack height (group-uniform)
    int         next_rl = rl_start;  // the next level this group searches (levels in between are skipped)

    // lower bound (metres^2, conservative by 4 quanta per axis) of the distance from the query to the
    // voxel (vx,vy,vz) of absolute level L; for the neighbours of the query's own voxel this is the
    // per-axis gap formula of v1 (ux - cx * s is the gap to the -1 slab, (cx + 1) * s - ux to the +1 slab)
    auto voxel_bound = [&](int L, int vx, int vy, int vz) -> float
    {
        const float s  = (float)(1 << L);
        const float lx = (float)vx * s, ly = (float)vy * s, lz = (float)vz * s;
        const float gx = fmaxf(fmaxf(lx - ux, ux - (lx + s)) - 4.f, 0.f);
        const float gy = fmaxf(fmaxf(ly - uy, uy - (ly + s)) - 4.f, 0.f);
        const float gz = fmaxf(fmaxf(lz - uz, uz - (lz + s)) - 4.f, 0.f);
        return gx * gx * q2 + gy * gy * q2 + gz * gz * q2;
    };

    for (int rl = rl_start; rl < g.n_levels; rl++)
    {
        if (!__any_sync(FULL, live)) break;
        const int  L    = g.level_first + rl;
        const bool top  = L == kGridBits;         // the single voxel holding every point
        const bool here = live && rl >= next_rl;  // this group searches this level
        if (here) mine = sentinel;                // the list is rebuilt at every level: no key is ever offered twice
        const int   cmax = ((1 << kGridBits) - 1) >> L;
        const float s    = (float)(1 << L);  // voxel edge in finest quanta
        const int   cx = top ? 0 : (Ix >> L), cy = top ? 0 : (Iy >> L), cz = top ? 0 : (Iz >> L);
        const int   nph = top ? 1 : n_phases;
        // smallest bound any face voxel / any edge or corner voxel of the block can have (the per-axis gaps of
        // voxel_bound): a phase none of whose voxels can matter to any group of the warp is skipped outright
        float min_face, min_edge;
        {
            const float fx = ux - (float)cx * s, fy = uy - (float)cy * s, fz = uz - (float)cz * s;
            const float ax = fmaxf(fminf(fx, s - fx) - 4.f, 0.f), ay = fmaxf(fminf(fy, s - fy) - 4.f, 0.f), az = fmaxf(fminf(fz, s - fz) - 4.f, 0.f);
            const float mnx = ax * ax * q2, mny = ay * ay * q2, mnz = az * az * q2;
            min_face = fminf(mnx, fminf(mny, mnz)) * 0.999999f;
            min_edge = fminf(mnx + mny, fminf(mnx + mnz, mny + mnz)) * 0.999999f;
        }

#pragma unroll 1
        for (int phase = 0; phase < nph; phase++)
        {
            // ---- this phase's voxels onto the stack: the centre | the 6 faces | the 20 edges and corners
            // (n_phases == 2: all 26 neighbours at once), each bounded against kth first
            {
                const int lo = phase == 0 ? 0 : (phase == 1 ? 1 : 7), hi = phase == 0 ? 1 : ((phase == 1 && n_phases == 3) ? 7 : 27);
                if (phase > 0 && !__any_sync(FULL, here && !((phase == 1 ? min_face : min_edge) > kth))) continue;
                __syncwarp();
#pragma unroll 1
                for (int n0 = lo; n0 < hi; n0 += G)  // (warp-uniform; at most ceil(20 / G) passes)
                {
                    const int n  = n0 + sub;
                    bool      ok = false;
                    int       vx = 0, vy = 0, vz = 0;
                    if (here && n < hi)
                    {
                        const uint32_t code = nb_order[n];
                        vx = cx + (int)(code & 3u) - 1, vy = cy + (int)((code >> 2) & 3u) - 1, vz = cz + (int)((code >> 4) & 3u) - 1;
                        ok = (unsigned)vx <= (unsigned)cmax && (unsigned)vy <= (unsigned)cmax && (unsigned)vz <= (unsigned)cmax &&
                             !(voxel_bound(L, vx, vy, vz) > kth);  // strict `>`: an equal-distance lower index must still be seen
                    }
                    const unsigned b = __ballot_sync(FULL, ok) & gmask;
                    if (ok)
                    {
                        // (pushed in reverse so that the nearer voxels of kNeighbourOrder pop first)
                        const int at   = sp + __popc(b) - 1 - __popc(b & below);
                        tb.stk_key[at] = cell_key((uint32_t)vx, (uint32_t)vy, (uint32_t)vz);
                        tb.stk_rl[at]  = (uint8_t)rl;
                    }
                    sp += __popc(b);
                }
                __syncwarp();
            }
            // ---- rounds until every stack of the warp is empty
#pragma unroll 1
            while (__any_sync(FULL, sp > 0))
            {
                // pop: lane `sub` takes the sub-th entry from the top, bounds it again (kth may have
                // tightened since it was pushed) and probes it
                const int          n_pop = min(sp, G);
                sc.rounds++;
                uint32_t           start = 0, count = 0;
                int                erl   = 0;
                unsigned long long ckey  = 0;
                if (sub < n_pop)
                {
                    ckey = tb.stk_key[sp - 1 - sub];
                    erl  = tb.stk_rl[sp - 1 - sub];
                    const int vx = (int)(ckey & 0x1fffffu), vy = (int)((ckey >> 21) & 0x1fffffu), vz = (int)((ckey >> 42) & 0x1fffffu);
                    if (!(voxel_bound(g.level_first + erl, vx, vy, vz) > kth))
                    {
                        sc.probes++;
                        uint32_t slot;
                        if (!grid_lookup(g, erl, (uint32_t)vx, (uint32_t)vy, (uint32_t)vz, start, count, &slot))
                            count = 0;
                        else if (g.box && count >= kBoxMin && box_bound(g, erl, slot, g.level_first + erl, vx, vy, vz, ux, uy, uz, q2) > kth)
                            count = 0;  // the voxel's points sit in a corner of it the K-th distance does not reach
                    }
                }
                sp -= n_pop;
                __syncwarp();  // the popped entries are in registers before anybody pushes over them
                // split the large voxels (while the stack has room), list the others. The children of a voxel
                // are bounded BEFORE they take a stack slot — lane c of the group looks at child c — so that
                // only those the K-th distance cannot exclude cost a round later (three of four fail for a
                // query that climbed through empty space)
                const bool large = count > kDescend && erl > 0;
                bool       split = false;
                unsigned   bl    = __ballot_sync(FULL, large);
                while (bl)  // warp-uniform: one large voxel of one group per pass
                {
                    const int o = __ffs(bl) - 1;  // the lane holding it
                    bl &= bl - 1;
                    const bool               mine_grp = (unsigned)(o - gbase) < (unsigned)G;
                    const unsigned long long pk  = __shfl_sync(FULL, ckey, o);
                    const int                prl = __shfl_sync(FULL, erl, o) - 1;
                    const bool               room = sp + 8 <= kStackCap;  // (group-uniform)
                    bool                     ok   = false;
                    unsigned long long       ck   = 0;
                    if (mine_grp && room && sub < 8)
                    {
                        const int vx = (int)((pk & 0x1fffffull) << 1) | (sub & 1), vy = (int)(((pk >> 21) & 0x1fffffull) << 1) | ((sub >> 1) & 1),
                                  vz = (int)(((pk >> 42) & 0x1fffffull) << 1) | (sub >> 2);
                        ok = !(voxel_bound(g.level_first + prl, vx, vy, vz) > kth);
                        ck = cell_key((uint32_t)vx, (uint32_t)vy, (uint32_t)vz);
                    }
                    const unsigned bc = __ballot_sync(FULL, ok) & gmask;
                    if (ok)
                    {
                        const int at   = sp + __popc(bc & below);
                        tb.stk_key[at] = ck;
                        tb.stk_rl[at]  = (uint8_t)prl;
                    }
                    if (mine_grp && room)
                    {
                        sp += __popc(bc);
                        if (lane == o) split = true;
                    }
                }
                const bool     listed = count != 0u && !split;
                const unsigned b      = __ballot_sync(FULL, listed) & gmask;
                const uint32_t n_runs = __popc(b);
                if (listed) sc.cands += count;
                // run table: lengths -> inclusive prefix, starts -> start - exclusive prefix (n_runs <= G: one pass)
                uint32_t total;
                {
                    if (listed)
                    {
                        const int at2 = __popc(b & below);
                        tb.end[at2] = count, tb.adj[at2] = start;
                    }
                    __syncwarp();
                    const bool     mine_e = (uint32_t)sub < n_runs;  // lane `sub` finishes entry `sub`
                    const uint32_t len    = mine_e ? tb.end[sub] : 0u;
                    uint32_t       incl   = len;
#pragma unroll
                    for (int o = 1; o < G; o <<= 1)
                    {
                        const uint32_t y = __shfl_up_sync(FULL, incl, o, G);
                        if (sub >= o) incl += y;
                    }
                    if (mine_e) tb.end[sub] = incl, tb.adj[sub] -= incl - len;
                    total = __shfl_sync(FULL, incl, gbase + G - 1);
                }
                __syncwarp();

                // ---- a LONG list (a query that climbed: its groups' lists are long in different rounds, so
                // stepping them side by side would cost the warp the SUM of their lengths) is scanned by the
                // whole warp for its owner group, 32 candidates per step, through the owner's table
                if (G < 32)
                {
                    unsigned big = __ballot_sync(FULL, total >= kLongList && sub == 0);
                    while (big)  // warp-uniform
                    {
                        const int o = __ffs(big) - 1;  // first lane of the owner group
                        big &= big - 1;
                        const RunTable& ot     = tables[(threadIdx.x >> 5) * (32 / G) + o / G];
                        const uint32_t  ototal = __shfl_sync(FULL, total, o);
                        const float     oqx = __shfl_sync(FULL, qx, o), oqy = __shfl_sync(FULL, qy, o), oqz = __shfl_sync(FULL, qz, o);
                        const float     okth  = __shfl_sync(FULL, kth, o);
                        const int       okl   = o + K - 1;  // lane holding the owner's K-th key
                        const bool      owner = (unsigned)(lane - o) < (unsigned)G;
                        uint32_t        r = 0, re = ot.end[0], ra = ot.adj[0];
#pragma unroll 1
                        for (uint32_t j0 = 0; j0 < ototal; j0 += 64)
                        {
                            float4 p[2];
#pragma unroll
                            for (int u = 0; u < 2; u++)
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
                            for (int u = 0; u < 2; u++)
                            {
                                if (j0 + u * 32 >= ototal) break;  // warp-uniform
                                sc.steps++;
                                const bool               in   = j0 + u * 32 + lane < ototal;
                                const unsigned long long c    = in ? point_key(oqx, oqy, oqz, p[u]) : ~0ull;
                                const unsigned long long kkey = __shfl_sync(FULL, mine, okl);
                                const bool pass = in && c < kkey && __uint_as_float((uint32_t)(c >> 32)) <= okth;
                                unsigned   pm   = __ballot_sync(FULL, pass);
                                while (pm)  // warp-uniform
                                {
                                    const int src = __ffs(pm) - 1;
                                    pm &= pm - 1;
                                    sc.inserts++;
                                    const unsigned long long cc = __shfl_sync(FULL, c, src);
                                    const unsigned long long up = __shfl_up_sync(FULL, mine, 1, G);
                                    if (owner && cc < mine) mine = (sub == 0 || !(cc < up)) ? cc : up;
                                }
                            }
                        }
                        const unsigned long long kkey = __shfl_sync(FULL, mine, okl);
                        if (owner) kth = fminf(kth, __uint_as_float((uint32_t)(kkey >> 32))), total = 0;
                    }
                }
                // ---- dense scan of the round's list, G candidates per step, four steps of loads in flight
                const uint32_t steps = __reduce_max_sync(FULL, total);
                constexpr int  kAhead = 4;
                uint32_t       r = 0, re = tb.end[0], ra = tb.adj[0];  // stale values if total == 0: never used then
#pragma unroll 1
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
                        sc.steps++;
                        const bool               in   = j0 + u * G + sub < total;
                        const unsigned long long c    = in ? point_key(qx, qy, qz, p[u]) : ~0ull;
                        const unsigned long long kkey = __shfl_sync(FULL, mine, kth_lane);
                        const bool pass = in && c < kkey && __uint_as_float((uint32_t)(c >> 32)) <= kth;
                        unsigned   pm   = __ballot_sync(FULL, pass) & gmask;
                        while (__any_sync(FULL, pm != 0))
                        {
                            // one offer to the group's list: lanes with cc < mine form a suffix of the group, its
                            // first lane takes cc, the others take their left neighbour's key
                            const bool ins = pm != 0;
                            const int  src = ins ? __ffs(pm) - 1 : lane;
                            pm &= pm - 1;
                            sc.inserts++;
                            const unsigned long long cc = __shfl_sync(FULL, c, src);
                            const unsigned long long up = __shfl_up_sync(FULL, mine, 1, G);
                            if (ins && cc < mine) mine = (sub == 0 || !(cc < up)) ? cc : up;
                        }
                    }
                }
                const unsigned long long kkey = __shfl_sync(FULL, mine, kth_lane);
                if (total) kth = fminf(kth, __uint_as_float((uint32_t)(kkey >> 32)));
            }
        }
        if (top)
        {
            if (here && sub == 0) sc.levels++;
            break;
        }
        // `mine` now holds the exact K best of this level's block; everything outside the 3x3x3
        // block is at least m quanta away
        const float fx = ux - (float)cx * s, fy = uy - (float)cy * s, fz = uz - (float)cz * s;
        const float mx = s + fminf(fx, s - fx), my = s + fminf(fy, s - fy), mz = s + fminf(fz, s - fz);
        const float m  = fmaxf(fminf(mx, fminf(my, mz)) - 4.f, 0.f);
        if (here)
        {
            if (sub == 0) sc.levels++;
            if (kth <= m * m * q2)
                live = false;
            else
            {
                // not settled: go straight to the first level whose block is certain to settle the bound found
                // (its margin is at least one voxel edge) instead of trying every level in between
                next_rl = rl + 1;
                while (next_rl < g.n_levels - 1)
                {
                    const float e = fmaxf((float)(1 << (g.level_first + next_rl)) - 4.f, 0.f);
                    if (e * e * q2 >= kth) break;
                    next_rl++;
                }
            }
        }
    }
}

// ---- one thread per query (k <= kThreadKMax) ---------------------------------------------------
// The group searches above spread ONE query's candidates over G lanes and pay for it at every step: the
// K best keys live in G different lanes, so a candidate enters the list through a ballot, two shuffles of
// the key and a shuffle-up, one candidate at a time per group (per-line profile r02_ncu_lines_c3_knn.txt:
// half of the kernel's instructions are that bookkeeping, a quarter is the stack / run table it shares
// through shared memory). Here a query is ONE thread:
//   * the K best keys sit in the thread's registers, descending (best[0] = the K-th best, the only key a
//     candidate is compared with); a candidate that passes sinks through a branch-free chain of KM - 1
//     compare-exchanges. 32 queries take a candidate each per step and nothing is exchanged;
//   * queries are walked in Morton order, so the lanes of a warp stand in the same or adjacent voxels:
//     their hash probes and point reads mostly hit the same addresses (one L1 transaction) and their
//     loops have nearly the same trip counts;
//   * per level the same three phases as above (centre | faces | edges + corners), the bound of the 26
//     neighbours from three per-axis gaps (two adds per voxel); the runs found in a phase go into the
//     thread's run list (shared memory, column per thread: no bank conflict whatever row a lane is in)
//     and are scanned densely, four loads in flight; a voxel above kDescend points is split depth-first
//     WITHOUT a stack (the walk derives the next sibling / the parent from the voxel coordinates), every
//     leaf bounded first and scanned at once so that the bound tightens on the way.
// Exactness: as knn_search — every voxel of the level's block is scanned, or excluded by a conservative
// lower bound above an upper bound of the K-th distance, or replaced by its children.
constexpr int kThreadKMax = 20;
constexpr int kRunCap     = 8;  // runs a thread lists before it scans them
constexpr int kDfsFlush   = 4;  // inside a split voxel: leaves listed before they are scanned (nearest first)
template <int NT>
struct KnnThreadShared
{
    uint32_t start[kRunCap][NT];
    uint32_t count[kRunCap][NT];
    float    bound[kRunCap][NT];  // lower bound of the distance to the run's points (its voxel's tight box)
};

// Work budget of a thread: a query that has cost more probes / candidates than this without settling hands itself
// over (its position goes on the deferred list, served afterwards by the warp-per-query group search) — a lone
// lane walking 700 candidates through a split voxel holds its warp for longer than the rest of the launch takes.
// list == NULL: no deferral. A full list (cap entries) means the thread carries on by itself.
struct DeferList
{
    uint32_t* list;
    uint32_t* count;
    uint32_t  cap, max_probes, max_cands;
};

// returns true if the query was deferred (best[] is meaningless then)
template <int KM, int NT>
__device__ __forceinline__ bool knn_search_thread(const GridView& g, bool enabled, float qx, float qy, float qz, float radius2,
                                                  int K, int rl_start, unsigned long long (&best)[KM], SearchCounters& sc,
                                                  KnnThreadShared<NT>& ks, const DeferList& df, uint32_t qpos)
{
    const int                tid      = threadIdx.x;
    const unsigned long long sentinel = (unsigned long long)__float_as_uint(radius2) << 32;
#pragma unroll
    for (int j = 0; j < KM; j++) best[j] = j < K ? sentinel : 0ull;  // (unused slots hold the smallest key: nothing sinks past them)
    bool live = enabled && radius2 > 0.f;
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
    uint32_t    n_runs = 0;

    // The walk is a state machine with ONE producer loop and ONE scan in the code (a scan inlined at every
    // place a list can fill up made the first group search of this round stall on instruction fetch), and
    // the scan — where the instructions go — is where the lanes of a warp meet again after producing.
    // (Every lane of the warp runs the outer loop and takes part in its votes, finished lanes with nothing to
    // do: left to themselves the lanes never reconverged after their first different probe count and the
    // kernel ran them one after the other, 20x slower.)
    constexpr unsigned FULL = 0xffffffffu;
    int      rl    = rl_start;
    int      phase = 0;         // 0 = the centre voxel, 1 = the 6 face neighbours, 2 = the 20 edge / corner neighbours
    uint32_t mask  = 1u << 13;  // voxels of the phase still to visit: bit dz*9 + dy*3 + dx (each 0..2)
    bool     dfs   = false;     // inside a voxel that was split: (erl, vx, vy, vz) is the next voxel of the walk
    int      erl = 0, vx = 0, vy = 0, vz = 0;
    bool     deferred = false, may_defer = df.list != nullptr;
    // the child of voxel (px,py,pz), relative level prl, the query is in or nearest to: bit 0 = upper half in x, ...
    auto near_child = [&](int prl, int px, int py, int pz) -> int
    {
        const float h = (float)(1 << (g.level_first + prl - 1));  // child edge, finest quanta
        return (int)(ux >= (float)(2 * px + 1) * h) | ((int)(uy >= (float)(2 * py + 1) * h) << 1) | ((int)(uz >= (float)(2 * pz + 1) * h) << 2);
    };
#pragma unroll 1
    while (__any_sync(FULL, live))
    {
        const int  L    = g.level_first + rl;
        const bool top  = L == kGridBits;  // the single voxel holding every point
        const int  cmax = ((1 << kGridBits) - 1) >> L;
        const int  cx = top ? 0 : (Ix >> L), cy = top ? 0 : (Iy >> L), cz = top ? 0 : (Iz >> L);
        bool       phase_done = false;
        // ---- produce: list runs until the phase is through, the list is full or a split voxel's leaves are due
#pragma unroll 1
        while (live)
        {
            if (may_defer && (sc.probes > df.max_probes || sc.cands > df.max_cands))
            {
                const uint32_t at = atomicAdd(df.count, 1u);
                may_defer         = false;
                if (at < df.cap)
                {
                    df.list[at] = qpos;
                    deferred = true, live = false, n_runs = 0;
                    break;
                }
            }
            uint32_t start = 0, count = 0;
            int      prl, px, py, pz;  // the voxel probed in this step
            if (dfs)
                prl = erl, px = vx, py = vy, pz = vz;
            else
            {
                if (!mask)
                {
                    phase_done = true;
                    break;
                }
                if (n_runs == (uint32_t)kRunCap) break;
                const int bit = __ffs(mask) - 1;
                mask &= mask - 1;
                prl = rl, px = cx + bit % 3 - 1, py = cy + (bit % 9) / 3 - 1, pz = cz + bit / 9 - 1;
            }
            bool  found = false;
            float bnd;
            {
                // lower bound (metres^2, conservative by 4 quanta per axis) of the distance to the voxel; for the
                // block's own voxels it was tested when the mask was made, against a K-th distance that may have
                // tightened since
                const float e  = (float)(1 << (g.level_first + prl));
                const float lx = (float)px * e, ly = (float)py * e, lz = (float)pz * e;
                const float gx = fmaxf(fmaxf(lx - ux, ux - (lx + e)) - 4.f, 0.f);
                const float gy = fmaxf(fmaxf(ly - uy, uy - (ly + e)) - 4.f, 0.f);
                const float gz = fmaxf(fmaxf(lz - uz, uz - (lz + e)) - 4.f, 0.f);
                bnd            = gx * gx * q2 + gy * gy * q2 + gz * gz * q2;
                const unsigned pmax = (unsigned)(((1 << kGridBits) - 1) >> (g.level_first + prl));
                if ((unsigned)px <= pmax && (unsigned)py <= pmax && (unsigned)pz <= pmax && !(bnd > kth))  // strict `>`: an equal-distance lower index must still be seen
                {
                    sc.probes++;
                    uint32_t slot;
                    found = grid_lookup(g, prl, (uint32_t)px, (uint32_t)py, (uint32_t)pz, start, count, &slot);
                    if (found && g.box && count >= kBoxMin)
                    {
                        // (the voxel's points may sit in a corner of it the K-th distance does not reach)
                        bnd = box_bound(g, prl, slot, g.level_first + prl, px, py, pz, ux, uy, uz, q2);
                        if (bnd > kth) found = false;
                    }
                }
            }
            const bool split = found && count > kDescend && prl > 0;
            if (found && !split)
            {
                ks.start[n_runs][tid] = start, ks.count[n_runs][tid] = count, ks.bound[n_runs][tid] = bnd;
                n_runs++;
            }
            if (split)
            {
                // into the child nearest to the query. No stack: from a finished voxel the walk goes to the next
                // sibling (the i-th visited child is nearest ^ i), from the last one up to the parent's next sibling
                const int c = near_child(prl, px, py, pz);
                dfs = true, erl = prl - 1, vx = (px << 1) | (c & 1), vy = (py << 1) | ((c >> 1) & 1), vz = (pz << 1) | (c >> 2);
                if (n_runs) break;  // what is listed so far first: it tightens the bound the children are held against
                continue;
            }
            if (dfs)
            {
                int i;
                for (;;)
                {
                    const int c = near_child(erl + 1, vx >> 1, vy >> 1, vz >> 1);
                    i           = ((vx & 1) | ((vy & 1) << 1) | ((vz & 1) << 2)) ^ c;
                    if (i < 7)
                    {
                        const int b = (i + 1) ^ c;
                        vx = (vx & ~1) | (b & 1), vy = (vy & ~1) | ((b >> 1) & 1), vz = (vz & ~1) | (b >> 2);
                        break;
                    }
                    vx >>= 1, vy >>= 1, vz >>= 1, erl++;
                    if (erl == rl)
                    {
                        dfs = false;
                        break;
                    }
                }
                if (n_runs >= (uint32_t)(dfs ? kDfsFlush : 1)) break;  // leaves are scanned a few at a time: depth first, the bound tightens on the way
            }
        }
        // ---- scan the listed runs, nearest first, for as long as the K-th distance reaches them; four loads in
        // flight per lane (warp-uniform loop: a lane that is through idles)
        __syncwarp();
        {
            uint32_t todo = (1u << n_runs) - 1u, p = 0, e = 0;
#pragma unroll 1
            for (;;)
            {
                if (p >= e && todo)
                {
                    const float kcur = fminf(kth, __uint_as_float((uint32_t)(best[0] >> 32)));
                    int         pick = -1;
                    float       bmin = 3.4e38f;
#pragma unroll
                    for (int r = 0; r < kRunCap; r++)
                        if ((todo >> r) & 1u)
                        {
                            const float b = ks.bound[r][tid];
                            if (b < bmin) bmin = b, pick = r;
                        }
                    if (bmin > kcur)
                        todo = 0;  // nothing left that could hold a better point
                    else
                    {
                        todo &= ~(1u << pick), p = ks.start[pick][tid], e = p + ks.count[pick][tid];
                        sc.cands += e - p;  // (runs listed but never scanned are not candidates)
                    }
                }
                const bool busy = p < e;
                if (!__any_sync(FULL, busy)) break;
                constexpr int kAhead = 4;
                float4        pt[kAhead];
#pragma unroll
                for (int u = 0; u < kAhead; u++) pt[u] = (busy && p + u < e) ? __ldg(g.pts + p + u) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int u = 0; u < kAhead; u++)
                {
                    const unsigned long long c = (busy && p + u < e) ? point_key(qx, qy, qz, pt[u]) : ~0ull;
                    if (c < best[0] && __uint_as_float((uint32_t)(c >> 32)) <= kth)
                    {
                        // the old K-th best drops out, c sinks to its place (branch-free compare-exchange chain)
                        unsigned long long x = c;
#pragma unroll
                        for (int j = 1; j < KM; j++)
                        {
                            const bool               lt = x < best[j];
                            const unsigned long long hi = lt ? best[j] : x;
                            x                           = lt ? x : best[j];
                            best[j - 1]                 = hi;
                        }
                        best[KM - 1] = x;
                    }
                }
                if (busy) p += kAhead;
            }
            if (n_runs) kth = fminf(kth, __uint_as_float((uint32_t)(best[0] >> 32)));
            n_runs = 0;
        }
        __syncwarp();
        if (!phase_done) continue;  // (lanes that are through: phase_done is false, they only take part in the votes)
        if (phase < 2 && !top)
        {
            // ---- next phase: the neighbours the bound does not exclude, from the squared conservative gaps
            // (metres^2) to the -1 / 0 / +1 slabs per axis
            phase++;
            const float s  = (float)(1 << L);
            const float fx = ux - (float)cx * s, fy = uy - (float)cy * s, fz = uz - (float)cz * s;
            const float glx = fmaxf(fx - 4.f, 0.f), ghx = fmaxf(s - fx - 4.f, 0.f);
            const float gly = fmaxf(fy - 4.f, 0.f), ghy = fmaxf(s - fy - 4.f, 0.f);
            const float glz = fmaxf(fz - 4.f, 0.f), ghz = fmaxf(s - fz - 4.f, 0.f);
            const float ax[3] = {glx * glx * q2, 0.f, ghx * ghx * q2};
            const float ay[3] = {gly * gly * q2, 0.f, ghy * ghy * q2};
            const float az[3] = {glz * glz * q2, 0.f, ghz * ghz * q2};
            const bool  okx[3] = {cx >= 1 && cx - 1 <= cmax, (unsigned)cx <= (unsigned)cmax, cx + 1 >= 0 && cx + 1 <= cmax};
            const bool  oky[3] = {cy >= 1 && cy - 1 <= cmax, (unsigned)cy <= (unsigned)cmax, cy + 1 >= 0 && cy + 1 <= cmax};
            const bool  okz[3] = {cz >= 1 && cz - 1 <= cmax, (unsigned)cz <= (unsigned)cmax, cz + 1 >= 0 && cz + 1 <= cmax};
            uint32_t    m = 0;
#pragma unroll
            for (int dz = 0; dz < 3; dz++)
#pragma unroll
                for (int dy = 0; dy < 3; dy++)
#pragma unroll
                    for (int dx = 0; dx < 3; dx++)
                    {
                        if (dx == 1 && dy == 1 && dz == 1) continue;
                        if (okx[dx] && oky[dy] && okz[dz] && !(az[dz] + ay[dy] + ax[dx] > kth)) m |= 1u << (dz * 9 + dy * 3 + dx);
                    }
            constexpr uint32_t kFaces = (1u << 4) | (1u << 10) | (1u << 12) | (1u << 14) | (1u << 16) | (1u << 22);
            mask = m & (phase == 1 ? kFaces : ~kFaces);
            continue;
        }
        // ---- the level is through: best[] holds the exact K best of its block; everything outside the 3x3x3
        // block is at least m quanta away
        sc.levels++;
        if (top)
        {
            live = false;
            continue;
        }
        const float s  = (float)(1 << L);
        const float fx = ux - (float)cx * s, fy = uy - (float)cy * s, fz = uz - (float)cz * s;
        const float mx = s + fminf(fx, s - fx), my = s + fminf(fy, s - fy), mz = s + fminf(fz, s - fz);
        const float mq = fmaxf(fminf(mx, fminf(my, mz)) - 4.f, 0.f);
        if (kth <= mq * mq * q2)
        {
            live = false;
            continue;
        }
        // not settled: go straight to the first level whose block is certain to settle the bound found (its
        // margin is at least one voxel edge) instead of trying every level in between
        rl++;
        while (rl < g.n_levels - 1)
        {
            const float e = fmaxf((float)(1 << (g.level_first + rl)) - 4.f, 0.f);
            if (e * e * q2 >= kth) break;
            rl++;
        }
        phase = 0, mask = 1u << 13;
#pragma unroll
        for (int j = 0; j < KM; j++)
            if (j < K) best[j] = sentinel;  // the list is rebuilt at every level: no key is ever offered twice
    }
    return deferred;
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
