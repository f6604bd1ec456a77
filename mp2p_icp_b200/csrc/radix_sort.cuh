// Hand-written LSD radix sort of (64-bit key, 32-bit value) pairs for the index build
// (Morton key -> original point index). 8 bits per pass, stable. Per pass:
//   k_rs_hist     per tile of 2048 pairs: 256-bin digit histogram          reads  8 B/pair
//   k_rs_scan_a   per digit: exclusive scan of that digit's counts over the tiles (+ digit total)
//   k_rs_scan_b   exclusive scan of the 256 digit totals
//   k_rs_scatter  per tile: stable local ranks (warp match_any + per-warp digit counters),
//                 scatter to  base[digit] + prefix[digit][tile] + local rank   reads 12, writes 12 B/pair
// Bytes per pair and pass: 8 + 12 + 12 = 32  =>  B_sort = passes * N * 32.
#pragma once
#include "common.cuh"

namespace mp2p
{
namespace rs
{
constexpr int kThreads = 256;
constexpr int kItems   = 8;
constexpr int kTile    = kThreads * kItems;  // 2048 pairs per CTA
constexpr int kWarps   = kThreads / 32;

static __global__ void __launch_bounds__(kThreads)
    k_rs_hist(const unsigned long long* __restrict__ keys, uint32_t n, int shift,
              uint32_t* __restrict__ hist /*[256][n_tiles]*/, uint32_t n_tiles)
{
    __shared__ uint32_t sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * kTile;
#pragma unroll
    for (int r = 0; r < kItems; r++)
    {
        const uint32_t i = base + r * kThreads + threadIdx.x;
        if (i < n) atomicAdd(&sh[(uint32_t)(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * n_tiles + blockIdx.x] = sh[threadIdx.x];
}

// one CTA per digit: in-place exclusive scan over the tiles, total[digit] = sum
static __global__ void __launch_bounds__(kThreads)
    k_rs_scan_a(uint32_t* __restrict__ hist, uint32_t n_tiles, uint32_t* __restrict__ total)
{
    __shared__ uint32_t wsum[kWarps];
    __shared__ uint32_t carry;
    uint32_t*           row  = hist + (size_t)blockIdx.x * n_tiles;
    const int           lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t b = 0; b < n_tiles; b += kThreads)
    {
        const uint32_t t = b + threadIdx.x;
        const uint32_t v = t < n_tiles ? row[t] : 0u;
        uint32_t       x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        uint32_t off = carry;
        for (int w = 0; w < warp; w++) off += wsum[w];
        if (t < n_tiles) row[t] = off + x - v;
        __syncthreads();
        if (threadIdx.x == kThreads - 1) carry = off + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) total[blockIdx.x] = carry;
}

static __global__ void __launch_bounds__(256) k_rs_scan_b(uint32_t* __restrict__ total /*in: counts, out: exclusive*/)
{
    __shared__ uint32_t wsum[8];
    const int           lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t      v    = total[threadIdx.x];
    uint32_t            x    = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    uint32_t off = 0;
    for (int w = 0; w < warp; w++) off += wsum[w];
    total[threadIdx.x] = off + x - v;
}

static __global__ void __launch_bounds__(kThreads)
    k_rs_scatter(const unsigned long long* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                 unsigned long long* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t n, int shift,
                 const uint32_t* __restrict__ prefix /*[256][n_tiles]*/, const uint32_t* __restrict__ dbase /*[256]*/,
                 uint32_t n_tiles)
{
    __shared__ uint32_t wcount[kWarps][256];  // per-warp digit counts of the current round
    __shared__ uint32_t dbase_s[256];         // running output position per digit for this tile
    const int           lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    dbase_s[threadIdx.x] = dbase[threadIdx.x] + prefix[(size_t)threadIdx.x * n_tiles + blockIdx.x];
    const uint32_t base  = blockIdx.x * kTile;
    for (int r = 0; r < kItems; r++)
    {
        for (int w = 0; w < kWarps; w++) wcount[w][threadIdx.x] = 0;  // 256 threads clear 8x256 counters
        __syncthreads();
        const uint32_t           i     = base + r * kThreads + threadIdx.x;
        const bool               in    = i < n;
        const unsigned long long k     = in ? keys_in[i] : 0ull;
        const uint32_t           v     = in ? vals_in[i] : 0u;
        const uint32_t           d     = (uint32_t)(k >> shift) & 255u;
        const unsigned           alive = __ballot_sync(0xffffffffu, in);
        uint32_t                 rank_in_warp = 0;
        if (in)
        {
            const unsigned peers = __match_any_sync(alive, d);
            rank_in_warp         = __popc(peers & ((1u << lane) - 1u));
            if (rank_in_warp == 0) wcount[warp][d] = __popc(peers);  // one writer per (warp, digit)
        }
        __syncthreads();
        if (in)
        {
            uint32_t before = 0;
            for (int w = 0; w < warp; w++) before += wcount[w][d];
            const uint32_t pos = dbase_s[d] + before + rank_in_warp;
            keys_out[pos]      = k;
            vals_out[pos]      = v;
        }
        __syncthreads();
        {  // advance the per-digit running positions by this round's totals
            uint32_t tot = 0;
            for (int w = 0; w < kWarps; w++) tot += wcount[w][threadIdx.x];
            dbase_s[threadIdx.x] += tot;
        }
        __syncthreads();
    }
}

// Sorts in place logically: result ends in (keys_a, vals_a) after an even number of passes.
// scratch: hist[256*n_tiles] + total[256] (uint32).
inline int sort_pairs(mp2p_b200_ctx* ctx, unsigned long long* keys_a, uint32_t* vals_a, unsigned long long* keys_b,
                      uint32_t* vals_b, uint32_t n, int key_bits, uint32_t* scratch)
{
    const uint32_t n_tiles = (n + kTile - 1) / kTile;
    uint32_t*      hist    = scratch;
    uint32_t*      total   = scratch + (size_t)256 * n_tiles;
    const int      passes  = (key_bits + 7) / 8;
    unsigned long long *kin = keys_a, *kout = keys_b;
    uint32_t *          vin = vals_a, *vout = vals_b;
    for (int p = 0; p < passes; p++)
    {
        const int shift = 8 * p;
        k_rs_hist<<<n_tiles, kThreads, 0, ctx->stream>>>(kin, n, shift, hist, n_tiles);
        k_rs_scan_a<<<256, kThreads, 0, ctx->stream>>>(hist, n_tiles, total);
        k_rs_scan_b<<<1, 256, 0, ctx->stream>>>(total);
        k_rs_scatter<<<n_tiles, kThreads, 0, ctx->stream>>>(kin, vin, kout, vout, n, shift, hist, total, n_tiles);
        count_launch(ctx, 4);
        unsigned long long* tk = kin;
        kin                    = kout;
        kout                   = tk;
        uint32_t* tv = vin;
        vin          = vout;
        vout         = tv;
    }
    if (kin != keys_a)  // odd number of passes: bring the result home
    {
        MP2P_CUDA_TRY(cudaMemcpyAsync(keys_a, kin, (size_t)n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        MP2P_CUDA_TRY(cudaMemcpyAsync(vals_a, vin, (size_t)n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return 0;
}
}  // namespace rs
}  // namespace mp2p
