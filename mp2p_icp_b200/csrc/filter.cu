// FilterDecimateVoxels on the device (product code).
//
// Replaces mp2p_icp_filters::FilterDecimateVoxels::filter over ONE input layer
// (mp2p_icp_filters/src/FilterDecimateVoxels.cpp:109-378) — the step right before ICP::align() in the
// reference's pipelines (demos/icp-settings-kitti.yaml:76-82) — so that the decimated local layer can stay
// on the device from the filter to the matchers (mp2p_b200_cloud_create_decimated).
//
// The reference hashes every point into a voxel map (PointCloudToVoxelGridSingle.h:52-110: index per axis
// = int32(coordinate / resolution), float division, truncation toward zero) and walks the map. Here:
//   k_fd_minmax   voxel index of every point, extrema per axis (warp redux + 6 atomics per warp)
//   k_fd_keys     sort key = the three indices, offset by their minima, packed x | y | z (x most
//                 significant): ascending keys = ascending (cx, cy, cz), the order of the reference's
//                 std::map walk (use_tsl_robin_map = false); value = point index
//   rs::sort_pairs  the index build's stable LSD radix sort, ceil(bits / 8) passes: the members of a voxel
//                 stay in ascending point index = the reference's insertion order
//   k_fd_flags    run heads (one per occupied voxel; with flatten_to one per (cx, cy) column: the first
//                 voxel of the column in this order emits, FilterDecimateVoxels.cpp:210-224,335-349)
//   scan          exclusive prefix of the flags over tiles (two small kernels)
//   k_fd_emit     one thread per emitting voxel: FirstPoint = its first member; VoxelAverage = float sums
//                 in member order times float(1 / n) (:265-277,300-304); ClosestToAverage = the member
//                 closest to that mean, first on ties (:279-298). The sums are SEQUENTIAL float additions
//                 in the reference's order — the price of bit-identical averages is one thread per voxel.
// Output order: ascending (cx, cy, cz). That is the reference's order where its container has one (std::map);
// with the default tsl::robin_map the reference's order is implementation-defined and only the SET is pinned.
#include "common.cuh"
#include "radix_sort.cuh"

namespace mp2p
{
namespace
{
constexpr int kFdThreads = 256;

__device__ __forceinline__ int voxel_index(float c, float resolution)
{
    return __float2int_rz(__fdiv_rn(c, resolution));  // static_cast<int32_t>(c / resolution)
}

__global__ void __launch_bounds__(kFdThreads)
    k_fd_minmax(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z, uint32_t n, float res,
                int* __restrict__ mm /* min xyz, max xyz */)
{
    const uint32_t i  = blockIdx.x * kFdThreads + threadIdx.x;
    const bool     in = i < n;
    int            lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
    if (in)
    {
        lo[0] = hi[0] = voxel_index(x[i], res);
        lo[1] = hi[1] = voxel_index(y[i], res);
        lo[2] = hi[2] = voxel_index(z[i], res);
    }
#pragma unroll
    for (int d = 0; d < 3; d++)
    {
        const int a = __reduce_min_sync(0xffffffffu, lo[d]), b = __reduce_max_sync(0xffffffffu, hi[d]);
        if ((threadIdx.x & 31) == 0) atomicMin(mm + d, a), atomicMax(mm + 3 + d, b);
    }
}

struct FdPack
{
    int      min[3];
    uint32_t by, bz;  // bits of the y and z fields (x is the most significant field)
};

__global__ void __launch_bounds__(kFdThreads)
    k_fd_keys(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z, uint32_t n, float res,
              FdPack p, unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals)
{
    const uint32_t i = blockIdx.x * kFdThreads + threadIdx.x;
    if (i >= n) return;
    const unsigned long long cx = (unsigned long long)(uint32_t)(voxel_index(x[i], res) - p.min[0]);
    const unsigned long long cy = (unsigned long long)(uint32_t)(voxel_index(y[i], res) - p.min[1]);
    const unsigned long long cz = (unsigned long long)(uint32_t)(voxel_index(z[i], res) - p.min[2]);
    keys[i] = (cx << (p.by + p.bz)) | (cy << p.bz) | cz;
    vals[i] = i;
}

// flags[j] = 1 if sorted position j emits a point; tile_sum[b] = emitters of tile b
__global__ void __launch_bounds__(kFdThreads)
    k_fd_flags(const unsigned long long* __restrict__ keys, uint32_t n, uint32_t column_shift /* bz if flatten, else 0 */,
               uint8_t* __restrict__ flags, uint32_t* __restrict__ tile_sum)
{
    __shared__ uint32_t wsum[kFdThreads / 32];
    const uint32_t      j = blockIdx.x * kFdThreads + threadIdx.x;
    bool                f = false;
    if (j < n) f = j == 0 || (keys[j] >> column_shift) != (keys[j - 1] >> column_shift);
    if (j < n) flags[j] = f ? 1 : 0;
    const unsigned b = __ballot_sync(0xffffffffu, f);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = __popc(b);
    __syncthreads();
    if (threadIdx.x == 0)
    {
        uint32_t s = 0;
        for (int w = 0; w < kFdThreads / 32; w++) s += wsum[w];
        tile_sum[blockIdx.x] = s;
    }
}

// in-place exclusive scan of tile_sum[0..n_tiles) by ONE CTA; total -> *count
__global__ void __launch_bounds__(kFdThreads) k_fd_scan_tiles(uint32_t* __restrict__ tile_sum, uint32_t n_tiles, unsigned long long* __restrict__ count)
{
    __shared__ uint32_t wsum[kFdThreads / 32];
    __shared__ uint32_t carry;
    const int           lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t b = 0; b < n_tiles; b += kFdThreads)
    {
        const uint32_t t = b + threadIdx.x;
        const uint32_t v = t < n_tiles ? tile_sum[t] : 0u;
        uint32_t       s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += y;
        }
        if (lane == 31) wsum[warp] = s;
        __syncthreads();
        uint32_t off = carry;
        for (int w = 0; w < warp; w++) off += wsum[w];
        if (t < n_tiles) tile_sum[t] = off + s - v;
        __syncthreads();
        if (threadIdx.x == kFdThreads - 1) carry = off + s;
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = carry;
}

__global__ void __launch_bounds__(kFdThreads)
    k_fd_emit(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z,
              const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ vals, uint32_t n,
              const uint8_t* __restrict__ flags, const uint32_t* __restrict__ tile_off, int method, int has_flatten,
              float flatten_to, uint64_t capacity, float* __restrict__ ox, float* __restrict__ oy, float* __restrict__ oz,
              long long* __restrict__ osrc)
{
    __shared__ uint32_t wsum[kFdThreads / 32];
    const uint32_t      j    = blockIdx.x * kFdThreads + threadIdx.x;
    const int           lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool          f    = j < n && flags[j];
    const unsigned      b    = __ballot_sync(0xffffffffu, f);
    if (lane == 0) wsum[warp] = __popc(b);
    __syncthreads();
    if (!f) return;
    uint32_t pos = tile_off[blockIdx.x] + __popc(b & ((1u << lane) - 1u));
    for (int w = 0; w < warp; w++) pos += wsum[w];
    if (pos >= capacity) return;

    const unsigned long long key = keys[j];
    long long                src = -1;
    float                    px = 0.f, py = 0.f, pz = 0.f;
    if (method == 0)
        src = vals[j];
    else
    {
        // the voxel's members: sorted positions [j, e) with the same key, ascending point index
        float    mx = 0.f, my = 0.f, mz = 0.f;
        uint32_t e = j;
        for (; e < n && keys[e] == key; e++)
        {
            const uint32_t i = vals[e];
            mx = __fadd_rn(mx, x[i]), my = __fadd_rn(my, y[i]), mz = __fadd_rn(mz, z[i]);
        }
        const float inv_n = __fdiv_rn(1.0f, (float)(e - j));
        mx = __fmul_rn(mx, inv_n), my = __fmul_rn(my, inv_n), mz = __fmul_rn(mz, inv_n);
        if (method == 1)
        {
            float best = 0.f;
            for (uint32_t t = j; t < e; t++)
            {
                const uint32_t i  = vals[t];
                const float    dx = __fsub_rn(x[i], mx), dy = __fsub_rn(y[i], my), dz = __fsub_rn(z[i], mz);
                const float    d  = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                if (t == j || d < best) best = d, src = i;
            }
        }
        else
            px = mx, py = my, pz = mz;
    }
    if (src >= 0) px = x[src], py = y[src], pz = z[src];
    if (has_flatten) pz = flatten_to;
    ox[pos] = px, oy[pos] = py, oz[pos] = pz;
    if (osrc) osrc[pos] = src;
}

int bits_for(uint32_t range)  // bits needed for values 0..range
{
    int b = 0;
    while (b < 32 && (range >> b) != 0u) b++;
    return b;
}
}  // namespace

// x, y, z: DEVICE arrays of n points. Outputs: device arrays of `capacity` entries (d_osrc may be NULL);
// *d_count_out = device address of the number of points produced (a u64 the caller may copy back).
int run_decimate_voxels(mp2p_b200_ctx* ctx, const float* dx, const float* dy, const float* dz, uint64_t n,
                        const mp2p_b200_decimate_params* prm, float* d_ox, float* d_oy, float* d_oz, long long* d_osrc,
                        uint64_t capacity, uint64_t* h_count)
{
    *h_count = 0;
    if (n == 0) return 0;
    if (n >= 0xFFFFFFFFull)
    {
        set_error("decimate_voxels: n must be < 2^32 - 1");
        return MP2P_B200_ERR_ARG;
    }
    if (!(prm->voxel_filter_resolution > 0.f) || prm->decimate_method < 0 || prm->decimate_method > 2)
    {
        set_error("decimate_voxels: voxel_filter_resolution must be > 0 and decimate_method one of FirstPoint (0), "
                  "ClosestToAverage (1), VoxelAverage (2); RandomPoint draws from an unseeded generator upstream and is not offered");
        return MP2P_B200_ERR_ARG;
    }
    cudaStream_t   st    = ctx->stream;
    const uint32_t nn    = (uint32_t)n;
    const uint32_t tiles = (nn + kFdThreads - 1) / kFdThreads;
    const float    res   = prm->voxel_filter_resolution;
    // scratch: [mm 6 ints | count u64] + keys a/b + vals a/b + flags + tile sums + radix scratch
    const uint32_t rs_tiles = (nn + rs::kTile - 1) / rs::kTile;
    MP2P_TRY(ctx->d_fd_small.ensure(64));
    MP2P_TRY(ctx->d_fd_keys.ensure((size_t)nn * 16));
    MP2P_TRY(ctx->d_fd_vals.ensure((size_t)nn * 8));
    MP2P_TRY(ctx->d_fd_flags.ensure((size_t)nn + (size_t)tiles * 4 + 16));
    MP2P_TRY(ctx->d_fd_rs.ensure(((size_t)256 * rs_tiles + 256) * 4));
    int*                mm    = ctx->d_fd_small.as<int>();
    unsigned long long* count = reinterpret_cast<unsigned long long*>(ctx->d_fd_small.as<char>() + 32);
    const int           init[6] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN};
    MP2P_CUDA_TRY(cudaMemcpyAsync(mm, init, sizeof(init), cudaMemcpyHostToDevice, st));
    k_fd_minmax<<<tiles, kFdThreads, 0, st>>>(dx, dy, dz, nn, res, mm);
    count_launch(ctx);
    int h_mm[6];
    MP2P_CUDA_TRY(cudaMemcpyAsync(h_mm, mm, sizeof(h_mm), cudaMemcpyDeviceToHost, st));
    MP2P_CUDA_TRY(cudaStreamSynchronize(st));
    FdPack p{};
    int    bits[3];
    for (int d = 0; d < 3; d++)
    {
        p.min[d] = h_mm[d];
        bits[d]  = bits_for((uint32_t)((long long)h_mm[3 + d] - (long long)h_mm[d]));
    }
    p.by = (uint32_t)bits[1], p.bz = (uint32_t)bits[2];
    const int key_bits = bits[0] + bits[1] + bits[2];
    if (key_bits > 64)
    {
        set_error("decimate_voxels: the voxel indices span %d + %d + %d bits (> 64): resolution too fine for the cloud's extent", bits[0],
                  bits[1], bits[2]);
        return MP2P_B200_ERR_ARG;
    }
    unsigned long long* ka = ctx->d_fd_keys.as<unsigned long long>();
    unsigned long long* kb = ka + nn;
    uint32_t*           va = ctx->d_fd_vals.as<uint32_t>();
    uint32_t*           vb = va + nn;
    k_fd_keys<<<tiles, kFdThreads, 0, st>>>(dx, dy, dz, nn, res, p, ka, va);
    count_launch(ctx);
    if (key_bits > 0) MP2P_TRY(rs::sort_pairs(ctx, ka, va, kb, vb, nn, key_bits, ctx->d_fd_rs.as<uint32_t>()));
    uint8_t*  flags    = ctx->d_fd_flags.as<uint8_t>();
    uint32_t* tile_sum = reinterpret_cast<uint32_t*>(ctx->d_fd_flags.as<char>() + (((size_t)nn + 15) & ~(size_t)15));
    k_fd_flags<<<tiles, kFdThreads, 0, st>>>(ka, nn, prm->has_flatten_to ? p.bz : 0u, flags, tile_sum);
    k_fd_scan_tiles<<<1, kFdThreads, 0, st>>>(tile_sum, tiles, count);
    k_fd_emit<<<tiles, kFdThreads, 0, st>>>(dx, dy, dz, ka, va, nn, flags, tile_sum, prm->decimate_method, prm->has_flatten_to,
                                            prm->flatten_to, capacity, d_ox, d_oy, d_oz, d_osrc);
    count_launch(ctx, 3);
    unsigned long long h = 0;
    MP2P_CUDA_TRY(cudaMemcpyAsync(&h, count, 8, cudaMemcpyDeviceToHost, st));
    MP2P_CUDA_TRY(cudaStreamSynchronize(st));
    MP2P_CUDA_TRY(cudaGetLastError());
    *h_count = h;
    if (h > capacity)
    {
        set_error("decimate_voxels: %llu occupied voxels but capacity is %llu", h, (unsigned long long)capacity);
        return MP2P_B200_ERR_CAPACITY;
    }
    return 0;
}
}  // namespace mp2p
