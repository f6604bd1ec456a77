// Index build: replaces nn_prepare_for_3d_queries() (the nanoflann KD-tree build inside MRPT,
// call site mp2p_icp/src/Matcher_Points_DistanceThreshold.cpp:92) with a multi-resolution hashed
// voxel index over Morton-sorted points. Built once per map modification, amortised over all ICP
// iterations of an align() exactly like the reference's KD-tree.
//
// Steps (all on the context stream):
//   k_bbox          min/max of the layer                      N*12 B read
//   k_morton_keys   63-bit Morton key of the finest voxel     N*12 B read, N*12 B write
//   radix sort      (key,idx) pairs — hand-written stable LSD radix sort, 8 passes of 8 bits
//                   (radix_sort.cuh); no library kernel anywhere in this library
//   k_gather        sorted float4 {x,y,z,idx} + original-order float4
//   k_level_hist    for every sorted position, the coarsest level at which it opens a new voxel
//   (host)          choose the finest level with >= kTargetOccupancy points per occupied voxel
//   k_insert_cells  each voxel opener finds its run end (galloping) and inserts (key,start,count)
//                   into that level's open-addressing table
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "grid_search.cuh"
#include "radix_sort.cuh"

namespace mp2p
{
namespace
{
// finest level = the finest one with at least this many points per occupied voxel (measurement
// knob: MP2P_TARGET_OCC overrides the default)
float target_occupancy()
{
    static const float v = [] {
        const char* e = getenv("MP2P_TARGET_OCC");
        const float f = e ? (float)atof(e) : 0.f;
        return f > 0.5f ? f : 2.5f;
    }();
    return v;
}

__device__ __forceinline__ uint32_t f2ord(float f)
{
    const uint32_t u = __float_as_uint(f);
    return u ^ ((uint32_t)((int32_t)u >> 31) | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(uint32_t o)
{
    const uint32_t u = o ^ (((o >> 31) - 1u) | 0x80000000u);
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

// bbox[0..2] = min (ordered uints), bbox[3..5] = max
__global__ void __launch_bounds__(256) k_bbox(const float* __restrict__ x, const float* __restrict__ y,
                                              const float* __restrict__ z, uint32_t n,
                                              uint32_t* __restrict__ bbox)
{
    float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const float a = x[i], b = y[i], c = z[i];
        mn[0] = fminf(mn[0], a), mx[0] = fmaxf(mx[0], a);
        mn[1] = fminf(mn[1], b), mx[1] = fmaxf(mx[1], b);
        mn[2] = fminf(mn[2], c), mx[2] = fmaxf(mx[2], c);
    }
#pragma unroll
    for (int d = 0; d < 3; d++)
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
            mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
            mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
        }
    }
    if ((threadIdx.x & 31) == 0)
    {
#pragma unroll
        for (int d = 0; d < 3; d++)
        {
            atomicMin(bbox + d, f2ord(mn[d]));
            atomicMax(bbox + 3 + d, f2ord(mx[d]));
        }
    }
}

__global__ void __launch_bounds__(256)
    k_morton_keys(const float* __restrict__ x, const float* __restrict__ y,
                  const float* __restrict__ z, uint32_t n, float ox, float oy, float oz, float inv_s0,
                  unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int      cmax = (1 << kGridBits) - 1;
    const int      ix = min(max((int)floorf(grid_u(x[i], ox, inv_s0)), 0), cmax);
    const int      iy = min(max((int)floorf(grid_u(y[i], oy, inv_s0)), 0), cmax);
    const int      iz = min(max((int)floorf(grid_u(z[i], oz, inv_s0)), 0), cmax);
    keys[i]           = morton63((uint32_t)ix, (uint32_t)iy, (uint32_t)iz);
    vals[i]           = i;
}

__global__ void __launch_bounds__(256)
    k_gather(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z,
             const uint32_t* __restrict__ vals, uint32_t n, float4* __restrict__ pts,
             float4* __restrict__ pts_orig)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t i = vals[j];
    pts[j]           = make_float4(x[i], y[i], z[i], __int_as_float((int)i));
    pts_orig[j]      = make_float4(x[j], y[j], z[j], 0.f);
}

// top level (coarsest L) at which sorted position j opens a new voxel: voxel keys at level L are
// (morton >> 3L); position j opens voxels at all levels 0..floor(hb/3), hb = highest differing bit
// against its predecessor. -1: same finest voxel as predecessor.
__device__ __forceinline__ int opener_top_level(const unsigned long long* keys, uint32_t j)
{
    if (j == 0) return kMaxLevels - 1;
    const unsigned long long d = keys[j] ^ keys[j - 1];
    if (d == 0) return -1;
    return (63 - __clzll((long long)d)) / 3;
}

__global__ void __launch_bounds__(256)
    k_level_hist(const unsigned long long* __restrict__ keys, uint32_t n, uint32_t* __restrict__ hist)
{
    __shared__ uint32_t sh[kMaxLevels];
    if (threadIdx.x < kMaxLevels) sh[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n)
    {
        const int t = opener_top_level(keys, j);
        if (t >= 0) atomicAdd(&sh[t], 1u);
    }
    __syncthreads();
    if (threadIdx.x < kMaxLevels && sh[threadIdx.x]) atomicAdd(hist + threadIdx.x, sh[threadIdx.x]);
}

struct LevelLayout
{
    int      level_first, n_levels;
    uint32_t level_off[kMaxLevels], level_shift[kMaxLevels];
};

__device__ __forceinline__ uint32_t compact3(unsigned long long x)
{
    x &= 0x1249249249249249ull;
    x = (x ^ (x >> 2)) & 0x10c30c30c30c30c3ull;
    x = (x ^ (x >> 4)) & 0x100f00f00f00f00full;
    x = (x ^ (x >> 8)) & 0x1f0000ff0000ffull;
    x = (x ^ (x >> 16)) & 0x1f00000000ffffull;
    x = (x ^ (x >> 32)) & 0x1fffffull;
    return (uint32_t)x;
}

__global__ void __launch_bounds__(256)
    k_insert_cells(const unsigned long long* __restrict__ keys, uint32_t n, LevelLayout lay,
                   CellEntry* __restrict__ table)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int top = opener_top_level(keys, j);
    if (top < lay.level_first) return;
    const unsigned long long mk = keys[j];
    const int last = min(top, lay.level_first + lay.n_levels - 1);
    uint32_t  end  = j + 1;  // runs are nested: the end at level L is a lower bound for level L+1
    for (int L = lay.level_first; L <= last; L++)
    {
        const int                sh   = 3 * L;
        const unsigned long long cell = (sh >= 63) ? 0ull : (mk >> sh);
        // gallop to the first position whose voxel key at this level differs
        uint32_t step = 1, lo = end;  // invariant: [j, lo) in the voxel
        while (lo < n)
        {
            const uint32_t           probe = min(lo + step - 1, n - 1);
            const unsigned long long c     = (sh >= 63) ? 0ull : (keys[probe] >> sh);
            if (c != cell)
            {
                uint32_t hi = probe;  // keys[hi] outside; binary search in [lo, hi]
                while (lo < hi)
                {
                    const uint32_t           mid = (lo + hi) >> 1;
                    const unsigned long long cm  = (sh >= 63) ? 0ull : (keys[mid] >> sh);
                    if (cm == cell)
                        lo = mid + 1;
                    else
                        hi = mid;
                }
                break;
            }
            lo = probe + 1;
            step <<= 1;
        }
        end = lo;
        // voxel coordinates at this level
        const uint32_t cx = compact3(mk) >> L, cy = compact3(mk >> 1) >> L, cz = compact3(mk >> 2) >> L;
        const unsigned long long key = cell_key(cx, cy, cz);
        const int                rl  = L - lay.level_first;
        const uint32_t           shift = lay.level_shift[rl];
        const uint32_t           mask  = (1u << (64 - shift)) - 1u;
        CellEntry*               t     = table + lay.level_off[rl];
        uint32_t                 h     = cell_hash(key, shift);
        while (true)
        {
            const unsigned long long prev = atomicCAS(&t[h].key, kEmptyKey, key);
            if (prev == kEmptyKey)
            {
                t[h].start = j;
                t[h].count = end - j;
                break;
            }
            h = (h + 1) & mask;
        }
    }
}

// ---- tight boxes (GridView::box) --------------------------------------------------------------
// Box of the voxel in table slot `slot` of level L (edge 2^L finest quanta), per axis the smallest and the
// largest finest-voxel coordinate of its points relative to the voxel, in units of 2^max(L - 8, 0) quanta:
// word x = lo_x | lo_y << 8 | lo_z << 16, word y = hi_x | hi_y << 8 | hi_z << 16 (hi inclusive: the box ends at
// (hi + 1) units). Finest table: from the points' Morton keys. Every level above: the union of the children's
// boxes (eight lookups one table down), rounded outwards.
constexpr uint32_t kBoxScanMax = 4096;  // a finest-level voxel holding more points (duplicates) keeps the full cube
__global__ void __launch_bounds__(256)
    k_box_finest(const unsigned long long* __restrict__ keys, const CellEntry* __restrict__ table, uint32_t n_slots, int L,
                 uint2* __restrict__ box)
{
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n_slots) return;
    const CellEntry e = table[slot];
    if (e.key == kEmptyKey) return;
    const int sh = max(L - 8, 0);
    uint32_t  lo[3] = {255u, 255u, 255u}, hi[3] = {0u, 0u, 0u};
    if (e.count > kBoxScanMax)
        for (int d = 0; d < 3; d++) lo[d] = 0u, hi[d] = (uint32_t)min((1 << L) - 1, 255);
    else
    {
        const uint32_t o[3] = {(uint32_t)(e.key & 0x1fffffu) << L, (uint32_t)((e.key >> 21) & 0x1fffffu) << L, (uint32_t)((e.key >> 42) & 0x1fffffu) << L};
        for (uint32_t j = e.start; j < e.start + e.count; j++)
        {
            const unsigned long long mk = keys[j];
            for (int d = 0; d < 3; d++)
            {
                const uint32_t q = (compact3(mk >> d) - o[d]) >> sh;
                lo[d] = min(lo[d], q), hi[d] = max(hi[d], q);
            }
        }
    }
    box[slot] = make_uint2(lo[0] | (lo[1] << 8) | (lo[2] << 16), hi[0] | (hi[1] << 8) | (hi[2] << 16));
}
__global__ void __launch_bounds__(256)
    k_box_merge(const CellEntry* __restrict__ table, LevelLayout lay, int rl, uint32_t n_slots, uint2* __restrict__ box)
{
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n_slots) return;
    const CellEntry e = table[lay.level_off[rl] + slot];
    if (e.key == kEmptyKey) return;
    const int      L   = lay.level_first + rl;
    const int      shp = max(L - 8, 0), shc = max(L - 1 - 8, 0);
    const uint32_t v[3] = {(uint32_t)(e.key & 0x1fffffu), (uint32_t)((e.key >> 21) & 0x1fffffu), (uint32_t)((e.key >> 42) & 0x1fffffu)};
    uint32_t       lo[3] = {255u, 255u, 255u}, hi[3] = {0u, 0u, 0u};
    const uint32_t   shift = lay.level_shift[rl - 1], mask = (1u << (64 - shift)) - 1u;
    const CellEntry* t     = table + lay.level_off[rl - 1];
    for (uint32_t c = 0; c < 8; c++)
    {
        const uint32_t           b[3] = {c & 1u, (c >> 1) & 1u, c >> 2};
        const unsigned long long key  = cell_key(2 * v[0] + b[0], 2 * v[1] + b[1], 2 * v[2] + b[2]);
        uint32_t                 h    = cell_hash(key, shift);
        while (true)
        {
            const unsigned long long k = t[h].key;
            if (k == key)
            {
                const uint2 cb = box[lay.level_off[rl - 1] + h];
                for (int d = 0; d < 3; d++)
                {
                    const uint32_t off = b[d] << (L - 1);  // the child's origin inside this voxel, finest quanta
                    const uint32_t l   = off + (((cb.x >> (8 * d)) & 255u) << shc);
                    const uint32_t u   = off + ((((cb.y >> (8 * d)) & 255u) + 1u) << shc) - 1u;
                    lo[d] = min(lo[d], l >> shp), hi[d] = max(hi[d], u >> shp);
                }
                break;
            }
            if (k == kEmptyKey) break;
            h = (h + 1) & mask;
        }
    }
    box[lay.level_off[rl] + slot] = make_uint2(lo[0] | (lo[1] << 8) | (lo[2] << 16), hi[0] | (hi[1] << 8) | (hi[2] << 16));
}

__global__ void k_fill_u64(unsigned long long* p, size_t n, unsigned long long v)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = v;
}
// resident local cloud: coarse (30-bit) Morton keys of the cloud's own bounding box
__global__ void __launch_bounds__(256)
    k_cloud_keys(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z, uint32_t n,
                 float ox, float oy, float oz, float inv_s0, unsigned long long* __restrict__ keys,
                 uint32_t* __restrict__ vals)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cmax = (1 << kGridBits) - 1;
    const int ix = min(max((int)floorf(grid_u(x[i], ox, inv_s0)), 0), cmax);
    const int iy = min(max((int)floorf(grid_u(y[i], oy, inv_s0)), 0), cmax);
    const int iz = min(max((int)floorf(grid_u(z[i], oz, inv_s0)), 0), cmax);
    keys[i]      = morton63((uint32_t)ix, (uint32_t)iy, (uint32_t)iz) >> 33;  // 10 bits per axis
    vals[i]      = i;
}

__global__ void __launch_bounds__(256)
    k_cloud_gather(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z,
                   const uint32_t* __restrict__ vals, uint32_t n, float* __restrict__ sx, float* __restrict__ sy,
                   float* __restrict__ sz)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t i = vals[j];
    sx[j] = x[i], sy[j] = y[i], sz[j] = z[i];
}
}  // namespace

// Upload (or copy) a local cloud and sort a second copy along a Morton curve. The order only
// affects which queries share a warp — results are written back under the caller's indices — so
// any finite quantisation is valid; non-finite points simply land in the first/last cell.
int build_cloud(mp2p_b200_ctx* ctx, mp2p_b200_cloud* cloud, const float* x, const float* y, const float* z,
                uint64_t n64, int on_device)
{
    if (n64 >= (1ull << 31))
    {
        set_error("local cloud too large: %llu points (max 2^31-1)", (unsigned long long)n64);
        return MP2P_B200_ERR_ARG;
    }
    const uint32_t n  = (uint32_t)n64;
    cudaStream_t   st = ctx->stream;
    cloud->ctx = ctx, cloud->n = n;
    if (n == 0) return 0;
    MP2P_CUDA_TRY(cudaEventRecord(ctx->ev0, st));
    // padded like the staging buffers of stage_local (whole query tiles are always readable)
    const size_t bytes = ((size_t)n + kQueryTile) / kQueryTile * kQueryTile * sizeof(float);
    for (DevBuf* b : {&cloud->d_x, &cloud->d_y, &cloud->d_z, &cloud->d_sx, &cloud->d_sy, &cloud->d_sz})
    {
        MP2P_TRY(b->ensure(bytes));
        MP2P_CUDA_TRY(cudaMemsetAsync(b->p, 0, bytes, st));
    }
    MP2P_TRY(cloud->d_perm.ensure((size_t)n * 4));
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    MP2P_CUDA_TRY(cudaMemcpyAsync(cloud->d_x.p, x, (size_t)n * 4, kind, st));
    MP2P_CUDA_TRY(cudaMemcpyAsync(cloud->d_y.p, y, (size_t)n * 4, kind, st));
    MP2P_CUDA_TRY(cudaMemcpyAsync(cloud->d_z.p, z, (size_t)n * 4, kind, st));
    const float *dx = cloud->d_x.as<float>(), *dy = cloud->d_y.as<float>(), *dz = cloud->d_z.as<float>();

    DevBuf small, k0, k1, v1, tmp;
    auto   cleanup = [&]() { small.release(), k0.release(), k1.release(), v1.release(), tmp.release(); };
    MP2P_TRY(small.ensure(64));
    uint32_t* d_bbox = small.as<uint32_t>();
    {
        const uint32_t init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0, 0, 0};
        MP2P_CUDA_TRY(cudaMemcpyAsync(d_bbox, init, sizeof(init), cudaMemcpyHostToDevice, st));
        k_bbox<<<(int)std::min<uint32_t>((n + 255) / 256, 148 * 8), 256, 0, st>>>(dx, dy, dz, n, d_bbox);
        count_launch(ctx);
    }
    uint32_t h_bbox[6];
    MP2P_CUDA_TRY(cudaMemcpyAsync(h_bbox, d_bbox, sizeof(h_bbox), cudaMemcpyDeviceToHost, st));
    MP2P_CUDA_TRY(cudaStreamSynchronize(st));
    float  bmin[3];
    double extent = 0;
    for (int d = 0; d < 3; d++)
    {
        bmin[d]        = ord2f(h_bbox[d]);
        const float mx = ord2f(h_bbox[3 + d]);
        if (!std::isfinite(bmin[d])) bmin[d] = -3.0e38f;
        if (std::isfinite(mx)) extent = std::max(extent, (double)mx - (double)bmin[d]);
    }
    if (!(extent >= 1e-6)) extent = 1e-6;
    if (!(extent < 1e38)) extent = 1e38;
    const float inv_s0 = (float)((double)(1u << kGridBits) / (extent * (1.0 + 1e-4)));

    MP2P_TRY(k0.ensure(n * 8ull));
    MP2P_TRY(k1.ensure(n * 8ull));
    MP2P_TRY(v1.ensure(n * 4ull));
    k_cloud_keys<<<(n + 255) / 256, 256, 0, st>>>(dx, dy, dz, n, bmin[0], bmin[1], bmin[2], inv_s0,
                                                  k0.as<unsigned long long>(), cloud->d_perm.as<uint32_t>());
    count_launch(ctx);
    {
        const uint32_t n_tiles = (n + rs::kTile - 1) / rs::kTile;
        MP2P_TRY(tmp.ensure(((size_t)256 * n_tiles + 256) * sizeof(uint32_t)));
        MP2P_TRY(rs::sort_pairs(ctx, k0.as<unsigned long long>(), cloud->d_perm.as<uint32_t>(),
                                k1.as<unsigned long long>(), v1.as<uint32_t>(), n, 30, tmp.as<uint32_t>()));
    }
    k_cloud_gather<<<(n + 255) / 256, 256, 0, st>>>(dx, dy, dz, cloud->d_perm.as<uint32_t>(), n,
                                                    cloud->d_sx.as<float>(), cloud->d_sy.as<float>(),
                                                    cloud->d_sz.as<float>());
    count_launch(ctx);
    MP2P_CUDA_TRY(cudaEventRecord(ctx->ev1, st));
    MP2P_CUDA_TRY(cudaStreamSynchronize(st));
    MP2P_CUDA_TRY(cudaGetLastError());
    cudaEventElapsedTime(&cloud->build_ms, ctx->ev0, ctx->ev1);
    cleanup();
    return 0;
}

int build_index(mp2p_b200_ctx* ctx, mp2p_b200_map* map, const float* x, const float* y,
                const float* z, uint64_t n64, int on_device)
{
    if (n64 >= (1ull << 31))
    {
        set_error("map layer too large: %llu points (max 2^31-1)", (unsigned long long)n64);
        return MP2P_B200_ERR_ARG;
    }
    const uint32_t n  = (uint32_t)n64;
    cudaStream_t   st = ctx->stream;
    map->ctx          = ctx;
    map->info         = mp2p_b200_map_info{};
    map->info.n_points = n;
    map->view          = GridView{};
    map->view.n_points = n;
    if (n == 0)
    {
        for (int d = 0; d < 3; d++) map->view.bbmin[d] = 3.4e38f, map->view.bbmax[d] = -3.4e38f;
        return 0;
    }

    MP2P_CUDA_TRY(cudaEventRecord(ctx->ev0, st));

    DevBuf       sx, sy, sz;  // staging SoA if the caller's pointers are host memory
    const float *dx = x, *dy = y, *dz = z;
    if (!on_device)
    {
        MP2P_TRY(sx.ensure(n * 4ull));
        MP2P_TRY(sy.ensure(n * 4ull));
        MP2P_TRY(sz.ensure(n * 4ull));
        MP2P_CUDA_TRY(cudaMemcpyAsync(sx.p, x, n * 4ull, cudaMemcpyHostToDevice, st));
        MP2P_CUDA_TRY(cudaMemcpyAsync(sy.p, y, n * 4ull, cudaMemcpyHostToDevice, st));
        MP2P_CUDA_TRY(cudaMemcpyAsync(sz.p, z, n * 4ull, cudaMemcpyHostToDevice, st));
        dx = sx.as<float>(), dy = sy.as<float>(), dz = sz.as<float>();
    }
    auto cleanup = [&]() { sx.release(), sy.release(), sz.release(); };

    // ---- bbox
    DevBuf    small;
    MP2P_TRY(small.ensure(256));
    uint32_t* d_bbox = small.as<uint32_t>();
    uint32_t* d_hist = d_bbox + 8;
    {
        const uint32_t init[8] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0, 0, 0, 0, 0};
        MP2P_CUDA_TRY(cudaMemcpyAsync(d_bbox, init, sizeof(init), cudaMemcpyHostToDevice, st));
        MP2P_CUDA_TRY(cudaMemsetAsync(d_hist, 0, kMaxLevels * 4, st));
        const int blocks = (int)std::min<uint32_t>((n + 255) / 256, 148 * 8);
        k_bbox<<<blocks, 256, 0, st>>>(dx, dy, dz, n, d_bbox);
        count_launch(ctx);
    }
    uint32_t h_bbox[6];
    MP2P_CUDA_TRY(cudaMemcpyAsync(h_bbox, d_bbox, sizeof(h_bbox), cudaMemcpyDeviceToHost, st));
    MP2P_CUDA_TRY(cudaStreamSynchronize(st));
    float bmin[3], bmax[3];
    for (int d = 0; d < 3; d++) bmin[d] = ord2f(h_bbox[d]), bmax[d] = ord2f(h_bbox[3 + d]);
    for (int d = 0; d < 3; d++)
        if (!std::isfinite(bmin[d]) || !std::isfinite(bmax[d]))
        {
            cleanup();
            small.release();
            set_error("map layer contains non-finite coordinates");
            return MP2P_B200_ERR_ARG;
        }
    double extent = 0;
    for (int d = 0; d < 3; d++) extent = std::max(extent, (double)bmax[d] - (double)bmin[d]);
    if (extent < 1e-6) extent = 1e-6;
    const double s0     = extent * (1.0 + 1e-4) / (double)(1u << kGridBits);
    const float  inv_s0 = (float)(1.0 / s0);
    const double s0_eff = 1.0 / (double)inv_s0;  // the quantum the float function really uses
    float        s0_lo  = (float)(s0_eff * (1.0 - 1e-6));
    if ((double)s0_lo > s0_eff * (1.0 - 5e-7)) s0_lo = std::nextafter(s0_lo, 0.f);

    GridView& v = map->view;
    v.ox = bmin[0], v.oy = bmin[1], v.oz = bmin[2];
    v.inv_s0 = inv_s0, v.s0_lo = s0_lo;
    for (int d = 0; d < 3; d++) v.bbmin[d] = bmin[d], v.bbmax[d] = bmax[d];

    // ---- Morton keys + sort
    DevBuf k0, k1, v0, v1, tmp;
    MP2P_TRY(k0.ensure(n * 8ull));
    MP2P_TRY(k1.ensure(n * 8ull));
    MP2P_TRY(v0.ensure(n * 4ull));
    MP2P_TRY(v1.ensure(n * 4ull));
    auto cleanup2 = [&]() { cleanup(), k0.release(), k1.release(), v0.release(), v1.release(), tmp.release(), small.release(); };
    k_morton_keys<<<(n + 255) / 256, 256, 0, st>>>(dx, dy, dz, n, v.ox, v.oy, v.oz, inv_s0,
                                                   k0.as<unsigned long long>(), v0.as<uint32_t>());
    count_launch(ctx);
    {
        const uint32_t n_tiles = (n + rs::kTile - 1) / rs::kTile;
        MP2P_TRY(tmp.ensure(((size_t)256 * n_tiles + 256) * sizeof(uint32_t)));
        MP2P_TRY(rs::sort_pairs(ctx, k0.as<unsigned long long>(), v0.as<uint32_t>(), k1.as<unsigned long long>(),
                                v1.as<uint32_t>(), n, 63, tmp.as<uint32_t>()));
    }
    const unsigned long long* d_keys = k0.as<unsigned long long>();
    const uint32_t*           d_vals = v0.as<uint32_t>();

    // ---- gather
    MP2P_TRY(map->d_pts.ensure(n * 16ull));
    MP2P_TRY(map->d_pts_orig.ensure(n * 16ull));
    k_gather<<<(n + 255) / 256, 256, 0, st>>>(dx, dy, dz, d_vals, n, map->d_pts.as<float4>(),
                                              map->d_pts_orig.as<float4>());
    count_launch(ctx);
    v.pts = map->d_pts.as<float4>(), v.pts_orig = map->d_pts_orig.as<float4>();

    // ---- per-level voxel counts -> finest useful level -> table layout
    k_level_hist<<<(n + 255) / 256, 256, 0, st>>>(d_keys, n, d_hist);
    count_launch(ctx);
    uint32_t h_hist[kMaxLevels];
    MP2P_CUDA_TRY(cudaMemcpyAsync(h_hist, d_hist, sizeof(h_hist), cudaMemcpyDeviceToHost, st));
    MP2P_CUDA_TRY(cudaStreamSynchronize(st));
    uint64_t cells[kMaxLevels];  // occupied voxels per level
    {
        uint64_t acc = 0;
        for (int L = kMaxLevels - 1; L >= 0; L--) acc += h_hist[L], cells[L] = acc;
    }
    int Lf = kGridBits;  // the single-voxel level always qualifies
    for (int L = 0; L <= kGridBits; L++)
        if ((double)n / (double)cells[L] >= target_occupancy())
        {
            Lf = L;
            break;
        }
    LevelLayout lay{};
    lay.level_first = Lf, lay.n_levels = kGridBits - Lf + 1;
    uint64_t total = 0;
    for (int rl = 0; rl < lay.n_levels; rl++)
    {
        const uint64_t c    = cells[Lf + rl];
        uint32_t       lg   = 1;
        while ((1ull << lg) < 2 * c) lg++;
        lay.level_off[rl]   = (uint32_t)total;
        lay.level_shift[rl] = 64 - lg;
        total += 1ull << lg;
    }
    if (total >= (1ull << 32))
    {
        cleanup2();
        set_error("index too large");
        return MP2P_B200_ERR_NOMEM;
    }
    MP2P_TRY(map->d_table.ensure(total * sizeof(CellEntry)));
    MP2P_CUDA_TRY(cudaMemsetAsync(map->d_table.p, 0xff, total * sizeof(CellEntry), st));
    k_insert_cells<<<(n + 255) / 256, 256, 0, st>>>(d_keys, n, lay, map->d_table.as<CellEntry>());
    count_launch(ctx);
    v.table = map->d_table.as<CellEntry>(), v.level_first = Lf, v.n_levels = lay.n_levels;
    for (int rl = 0; rl < lay.n_levels; rl++)
    {
        v.level_off[rl] = lay.level_off[rl], v.level_shift[rl] = lay.level_shift[rl];
        v.level_occupancy[rl] = (float)((double)n / (double)cells[Lf + rl]);
    }

    // ---- tight boxes, finest table first, then level by level upwards. Optional ($MP2P_INDEX_BOX=1): on C3 they
    // cut the candidates of the k > 1 search by 23 % and its time by 4 % (the search is bound by its per-round
    // costs, not by the candidates), 0.5 % of an iteration, for 8 more bytes per table slot and 11 % more build time
    static const bool with_boxes = [] {
        const char* e = getenv("MP2P_INDEX_BOX");
        return e && atoi(e) == 1;
    }();
    v.box = nullptr;
    if (with_boxes)
    {
        MP2P_TRY(map->d_box.ensure(total * sizeof(uint2)));
        uint2* box = map->d_box.as<uint2>();
        for (int rl = 0; rl < lay.n_levels; rl++)
        {
            const uint32_t n_slots = 1u << (64 - lay.level_shift[rl]);
            if (rl == 0)
                k_box_finest<<<(n_slots + 255) / 256, 256, 0, st>>>(d_keys, map->d_table.as<CellEntry>(), n_slots, Lf, box);
            else
                k_box_merge<<<(n_slots + 255) / 256, 256, 0, st>>>(map->d_table.as<CellEntry>(), lay, rl, n_slots, box);
            count_launch(ctx);
        }
        v.box = box;
    }

    // ---- first-claim words (see match.cu): one u64 per map point, all ones = "never claimed"
    MP2P_TRY(map->d_claim.ensure(n * 8ull));
    k_fill_u64<<<148 * 4, 256, 0, st>>>(map->d_claim.as<unsigned long long>(), n, ~0ull);
    count_launch(ctx);
    map->epoch = 0;

    MP2P_CUDA_TRY(cudaEventRecord(ctx->ev1, st));
    MP2P_CUDA_TRY(cudaStreamSynchronize(st));
    MP2P_CUDA_TRY(cudaGetLastError());
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    cleanup2();

    mp2p_b200_map_info& inf = map->info;
    for (int d = 0; d < 3; d++) inf.bbox_min[d] = bmin[d], inf.bbox_max[d] = bmax[d];
    inf.finest_cell_size = (float)(s0_eff * (double)(1u << Lf));
    inf.n_levels         = (uint32_t)lay.n_levels;
    inf.n_finest_cells   = cells[Lf];
    inf.index_bytes      = n * 40ull + total * (sizeof(CellEntry) + (v.box ? sizeof(uint2) : 0));
    inf.build_ms         = ms;
    return 0;
}
}  // namespace mp2p
