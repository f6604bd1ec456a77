// Deterministic grid-wide reduction of NV doubles per thread into a 32-double packet, in ONE launch.
#pragma once
#include "common.cuh"

namespace mp2p
{
constexpr int    kReduceThreads  = 256;
constexpr int    kReduceWarps    = kReduceThreads / 32;

// (compile-time recursion: every index below is a constant, the values stay in registers)
template <int P>
__device__ __forceinline__ void warp_plain_steps(double (&v)[P])
{
#pragma unroll
    for (int off = 16; off >= P; off >>= 1)
#pragma unroll
        for (int i = 0; i < P; i++) v[i] += __shfl_xor_sync(0xffffffffu, v[i], off);
}
template <int P, int OFF>
__device__ __forceinline__ void warp_halving_steps(double (&v)[P], int lane)
{
    if constexpr (OFF >= 1)
    {
        const bool up = (lane & OFF) != 0;
#pragma unroll
        for (int i = 0; i < OFF; i++)  // 2 * OFF values are still held
        {
            const double send = up ? v[i] : v[i + OFF];
            const double keep = up ? v[i + OFF] : v[i];
            v[i]              = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
        }
        warp_halving_steps<P, OFF / 2>(v, lane);
    }
}

// Block reduction + grid fold without a second launch: every CTA stores its NV partial sums in row
// `slot` (a CTA-unique index in [0, n_slots)) and takes a ticket; the LAST CTA to arrive sums the
// rows of all CTAs in a FIXED order (8 chunks of rows in parallel, then the 8 chunk sums in order)
// into the packet and re-arms the ticket for the next launch. The order depends only on n_slots,
// so results are run-to-run bit-stable. Must be called by all kReduceThreads threads of the CTA.
// Returns true (to every thread of the CTA) in the CTA that folded the packet, false elsewhere.
template <int NV>
__device__ __forceinline__ bool block_reduce_to_packet(double (&acc)[NV], double* __restrict__ partials,
                                                       unsigned int* __restrict__ ticket,
                                                       double* __restrict__ packet, unsigned slot,
                                                       unsigned n_slots)
{
    static_assert(NV <= 32, "packet holds 32 doubles");
    __shared__ double   sh[kReduceWarps][32];
    __shared__ unsigned is_last;
    const int           lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // Warp sums of the NV values by RECURSIVE HALVING instead of NV separate shuffle trees: at every
    // step the lanes of a pair trade half of the values they still hold (the lane whose bit is clear
    // keeps the lower half and receives its partner's, and vice versa), so the P = 2^k >= NV values
    // cost P - 1 exchanges in total — 31 for the 29 Gauss-Newton sums, where the trees cost 29 x 5 = 145
    // (59 % of that kernel's instructions, profiles/r01_v20_ncu_lines_c3_gn.txt). Lane l ends with the
    // sum of value l & (P - 1). The order of additions is fixed, so results stay run-to-run identical.
    {
        constexpr int P = NV <= 1 ? 1 : (NV <= 2 ? 2 : (NV <= 4 ? 4 : (NV <= 8 ? 8 : (NV <= 16 ? 16 : 32))));
        double        v[P];
#pragma unroll
        for (int i = 0; i < P; i++) v[i] = i < NV ? acc[i] : 0.0;
        if (P <= 16) warp_plain_steps<P>(v);  // more lanes than values: plain exchanges first
        warp_halving_steps<P, P / 2>(v, lane);
        if (lane < P) sh[warp][lane] = v[0];
    }
    __syncthreads();
    if (threadIdx.x < NV)
    {
        double s = 0;
#pragma unroll
        for (int w = 0; w < kReduceWarps; w++) s += sh[w][threadIdx.x];
        partials[(size_t)slot * NV + threadIdx.x] = s;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == n_slots - 1) ? 1u : 0u;
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    // thread (warp = chunk, lane = value): chunk c sums rows [c*per, (c+1)*per) in order
    const unsigned per = (n_slots + kReduceWarps - 1) / kReduceWarps;
    double         s   = 0;
    if (lane < NV)
    {
        // 8 independent partial sums keep 8 loads in flight (the rows sit in L2: a dependent chain
        // of ~0.4 us loads would cost more than the whole streaming pass); combined in fixed order
        const unsigned b0 = warp * per, b1 = min(b0 + per, n_slots);
        double         p[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        unsigned       b = b0;
        for (; b + 8 <= b1; b += 8)
        {
            double v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = __ldcg(partials + (size_t)(b + k) * NV + lane);
#pragma unroll
            for (int k = 0; k < 8; k++) p[k] += v[k];
        }
#pragma unroll
        for (int k = 0; k < 8; k++)  // remainder, static indices (a dynamic p[k] would go to local memory)
            if (b + k < b1) p[k] += __ldcg(partials + (size_t)(b + k) * NV + lane);
        s = ((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]));
    }
    __syncthreads();
    sh[warp][lane] = s;
    __syncthreads();
    if (threadIdx.x < 32)
    {
        double t = 0;
        if (threadIdx.x < NV)
#pragma unroll
            for (int w = 0; w < kReduceWarps; w++) t += sh[w][threadIdx.x];
        packet[threadIdx.x] = t;
    }
    if (threadIdx.x == 0) *ticket = 0u;
    return true;
}

// solve.cu: scratch rows + ticket of a context
int solve_scratch(mp2p_b200_ctx* ctx, size_t rows, unsigned int** ticket, double** partials);
}  // namespace mp2p
