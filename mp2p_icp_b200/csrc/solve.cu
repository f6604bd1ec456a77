// Solver kernels (product code): the O(N_pairs) accumulations of Solver_Horn / Solver_GaussNewton.
//
//   k_gn_accumulate   pt2pt loop (optimal_tf_gauss_newton.cpp:149-180) and pt2pl loop (:267-286):
//                     residual + Jacobian (errorTerms.cpp:36-66,115-161) in the closed form
//                     Ji = A [R | -R [l]x]  (A = I for pt2pt, -n n^T/|n|^2 for pt2pl), robust weight
//                     (robust_kernels.h:57-94), H += w Ji^T Ji (upper triangle, 21), g += w Ji^T e (6).
//   k_horn_sums       eval_centroids_robust (Pairings.cpp:68-110).
//   k_horn_moments    visit_correspondences (visit_correspondences.h:100-212) + S accumulation
//                     (optimal_tf_horn.cpp:101-117).
//
// Reduction skeleton (shared by all three): every thread keeps NV double accumulators over a
// grid-stride range, warp-shuffle tree -> one partial per warp in shared memory -> one partial per
// CTA in global memory -> the LAST CTA to arrive (ticket) folds the CTA partials in FIXED order into a
// 32-double packet (reduce.cuh) — one launch. Grid size is a pure function of N, so results are
// run-to-run bit-stable. The packet is what a multi-GPU caller all-reduces.
//
// Pair records are AoS (36 / 72 bytes). A warp stages 32 consecutive records with fully coalesced
// 128-byte loads into shared memory and each lane then reads its own record (stride 9 / 18 words).
#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "host_math.hpp"
#include "reduce.cuh"
#include "peer.cuh"

namespace mp2p
{
namespace
{
constexpr int kSolveThreads = kReduceThreads;
constexpr int kWarps        = kReduceWarps;

// Stage 32 records of WORDS 4-byte words each into this warp's shared buffer, coalesced.
template <int WORDS>
__device__ __forceinline__ void warp_stage_records(const uint32_t* __restrict__ src_words,
                                                   uint64_t first_rec, uint64_t n_rec, uint32_t* sh)
{
    const int      lane  = threadIdx.x & 31;
    const uint64_t w0    = first_rec * WORDS;
    const uint64_t w_end = n_rec * WORDS;
#pragma unroll
    for (int k = 0; k < WORDS; k++)
    {
        const uint64_t w = w0 + (uint64_t)k * 32 + lane;
        sh[k * 32 + lane] = (w < w_end) ? __ldg(src_words + w) : 0u;
    }
    __syncwarp();
}

__device__ __forceinline__ double robust_weight(int kernel, double param, double e2)
{
    // robust_kernels.h:57-94 — argument is the SQUARED error
    if (kernel == 1)
    {
        const double d = e2 + param;  // GemanMcClure: c^2/(e^2+c)^2
        return (param * param) / (d * d);
    }
    if (kernel == 2) return (param * param) / (e2 + param * param);  // Cauchy
    return 1.0;
}

struct GNArgs
{
    uint64_t n2p, n2l;
    double   w2p, w2l;
    int      kernel;
    double   kparam;
    // point-to-line pairings (optimal_tf_gauss_newton.cpp:182-203), 72-byte records
    uint64_t        n2ln;
    double          w2ln;
    const uint32_t* p2ln;
};

constexpr int kGNV = 29;  // 21 H + 6 g + err + count

// acc += w * (a a^T upper, a r) for a 1x6 row a and scalar residual r
__device__ __forceinline__ void add_row(double (&acc)[kGNV], const double (&a)[6], double r, double w)
{
    int idx = 0;
#pragma unroll
    for (int i = 0; i < 6; i++)
    {
        const double wa = w * a[i];
#pragma unroll
        for (int j = i; j < 6; j++) acc[idx++] += wa * a[j];
    }
#pragma unroll
    for (int i = 0; i < 6; i++) acc[21 + i] += w * a[i] * r;
}

// One Gauss-Newton update (single thread): delta = -H^{-1} g by LDL^T, pose <- pose (+) exp(delta),
// convergence flags (optimal_tf_gauss_newton.cpp:344-365). state[0] = done, state[1] = updates.
__device__ __forceinline__ void gn_step_device(const double* packet, double minDelta, double maxCost, double* pose,
                                               uint32_t* state)
{
    if (state[0]) return;
    if (sqrt(packet[27]) <= maxCost)
    {
        state[0] = 1;
        return;
    }
    double H[36], g[6], delta[6];
    int    idx = 0;
    for (int i = 0; i < 6; i++)
        for (int j = i; j < 6; j++) H[6 * i + j] = H[6 * j + i] = packet[idx++];
    for (int i = 0; i < 6; i++) g[i] = -packet[21 + i];
    if (!hm::ldlt_solve6_nopivot(H, g, delta)) hm::ldlt_solve6(H, g, delta);
    hm::Pose34 P;
    for (int k = 0; k < 12; k++) P.m[k] = pose[k];
    const hm::Pose34 Pn = hm::compose(P, hm::se3_exp(delta));
    for (int k = 0; k < 12; k++) pose[k] = Pn.m[k];
    state[1] += 1;
    double nrm = 0;
    for (int k = 0; k < 6; k++) nrm += delta[k] * delta[k];
    if (sqrt(nrm) < minDelta) state[0] = 1;
}

// `step_state` != NULL: the CTA that folds the packet also applies the Gauss-Newton update to `pose`
// (every other CTA has read the pose long before: they all passed the ticket) — the single-GPU inner
// loop is then ONE launch per iteration. A multi-GPU caller all-reduces the packet first and uses
// k_gn_step.
// WITH_LINES = false compiles the point-to-line loop out (the pt2pt / pt2pl hot path keeps its registers)
// The pair loops of one Gauss-Newton accumulation at `pose` into the caller's accumulators (shared by the
// one-launch-per-iteration kernel and the whole-inner-loop kernel). `sh` = this warp's staging buffer.
template <bool WITH_LINES>
__device__ __forceinline__ void gn_accumulate_body(const uint32_t* __restrict__ p2p, const uint32_t* __restrict__ p2l, const GNArgs& a,
                                                   const double* pose, uint32_t* sh, double (&acc)[kGNV])
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double              R[9], t[3];
#pragma unroll
    for (int r = 0; r < 3; r++)
    {
        R[3 * r] = pose[4 * r], R[3 * r + 1] = pose[4 * r + 1], R[3 * r + 2] = pose[4 * r + 2];
        t[r] = pose[4 * r + 3];
    }
#pragma unroll
    for (int v = 0; v < kGNV; v++) acc[v] = 0;

    const uint64_t warp_global = (uint64_t)blockIdx.x * kWarps + warp;
    const uint64_t warp_stride = (uint64_t)gridDim.x * kWarps;

    // ---- point-to-point: e = R l + t - g ; J = [R | -R [l]x]
    for (uint64_t base = warp_global * 32; base < a.n2p; base += warp_stride * 32)
    {
        warp_stage_records<9>(p2p, base, a.n2p, sh);
        if (base + lane < a.n2p)
        {
            const uint32_t* rec = sh + lane * 9;
            const double    gx = __uint_as_float(rec[2]), gy = __uint_as_float(rec[3]), gz = __uint_as_float(rec[4]);
            const double    lx = __uint_as_float(rec[5]), ly = __uint_as_float(rec[6]), lz = __uint_as_float(rec[7]);
            double          e[3];
#pragma unroll
            for (int r = 0; r < 3; r++) e[r] = R[3 * r] * lx + R[3 * r + 1] * ly + R[3 * r + 2] * lz + t[r];
            e[0] -= gx, e[1] -= gy, e[2] -= gz;
            const double e2 = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
            double       w  = a.w2p;
            if (a.kernel) w *= robust_weight(a.kernel, a.kparam, e2);
            acc[27] += w * e2;
            acc[28] += 1.0;
#pragma unroll
            for (int r = 0; r < 3; r++)
            {
                const double R0 = R[3 * r], R1 = R[3 * r + 1], R2 = R[3 * r + 2];
                const double row[6] = {R0, R1, R2, R2 * ly - R1 * lz, R0 * lz - R2 * lx, R1 * lx - R0 * ly};
                add_row(acc, row, e[r], w);
            }
        }
        __syncwarp();
    }
    // ---- point-to-plane: scalar residual r = (n.g + d)/|n|, row a = n^T/|n| [R | -R [l]x];
    //      identical to the reference's 3-vector form since (n n^T/|n|^2)^2 = n n^T/|n|^2.
    for (uint64_t base = warp_global * 32; base < a.n2l; base += warp_stride * 32)
    {
        warp_stage_records<18>(p2l, base, a.n2l, sh);
        if (base + lane < a.n2l)
        {
            const uint32_t* rec = sh + lane * 18;
            double          c[4];
#pragma unroll
            for (int k = 0; k < 4; k++)
                c[k] = __longlong_as_double(((long long)rec[2 * k + 1] << 32) | (long long)rec[2 * k]);
            const double lx = __uint_as_float(rec[14]), ly = __uint_as_float(rec[15]), lz = __uint_as_float(rec[16]);
            double       g[3];
#pragma unroll
            for (int r = 0; r < 3; r++) g[r] = R[3 * r] * lx + R[3 * r + 1] * ly + R[3 * r + 2] * lz + t[r];
            const double mod_n = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
            const double ev    = c[0] * g[0] + c[1] * g[1] + c[2] * g[2] + c[3];
            const double e2    = ev * ev / mod_n;  // |e|^2 of the reference's 3-vector error
            double       w     = a.w2l;
            if (a.kernel) w *= robust_weight(a.kernel, a.kparam, e2);
            acc[27] += w * e2;
            acc[28] += 1.0;
            const double inv_n = 1.0 / sqrt(mod_n);
            const double n0 = c[0] * inv_n, n1 = c[1] * inv_n, n2 = c[2] * inv_n;
            // nR = n^T R
            const double q0 = n0 * R[0] + n1 * R[3] + n2 * R[6];
            const double q1 = n0 * R[1] + n1 * R[4] + n2 * R[7];
            const double q2 = n0 * R[2] + n1 * R[5] + n2 * R[8];
            const double row[6] = {q0, q1, q2, q2 * ly - q1 * lz, q0 * lz - q2 * lx, q1 * lx - q0 * ly};
            add_row(acc, row, ev * inv_n, w);
        }
        __syncwarp();
    }
    // ---- point-to-line (errorTerms.cpp:67-113): q = T(+)l - pBase, e = q - u (u.q),
    //      J = M [R | -R [l]x] with M = I - u u^T formed as written (:92-97: no |u| = 1 assumption);
    //      the cost term is weight^2 |e|^2 there (optimal_tf_gauss_newton.cpp:197)
    for (uint64_t base = warp_global * 32; WITH_LINES && base < a.n2ln; base += warp_stride * 32)
    {
        warp_stage_records<18>(a.p2ln, base, a.n2ln, sh);
        if (base + lane < a.n2ln)
        {
            const uint32_t* rec = sh + lane * 18;
            double          v[9];
#pragma unroll
            for (int k = 0; k < 9; k++) v[k] = __longlong_as_double(((long long)rec[2 * k + 1] << 32) | (long long)rec[2 * k]);
            const double lx = v[6], ly = v[7], lz = v[8];
            double       q[3];
#pragma unroll
            for (int r = 0; r < 3; r++) q[r] = R[3 * r] * lx + R[3 * r + 1] * ly + R[3 * r + 2] * lz + t[r] - v[r];
            const double ux = v[3], uy = v[4], uz = v[5];
            const double uq = ux * q[0] + uy * q[1] + uz * q[2];
            const double e[3] = {q[0] - ux * uq, q[1] - uy * uq, q[2] - uz * uq};
            const double e2 = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
            double       w  = a.w2ln;
            if (a.kernel) w *= robust_weight(a.kernel, a.kparam, e2);
            acc[27] += w * w * e2;
            acc[28] += 1.0;
            const double u[3] = {ux, uy, uz};
#pragma unroll
            for (int r = 0; r < 3; r++)
            {
                // row r of M [R | -R [l]x] = sum_c M[r][c] * B[c][:]
                double row[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
                for (int c = 0; c < 3; c++)
                {
                    const double m  = (r == c ? 1.0 : 0.0) - u[r] * u[c];
                    const double R0 = R[3 * c], R1 = R[3 * c + 1], R2 = R[3 * c + 2];
                    row[0] += m * R0, row[1] += m * R1, row[2] += m * R2;
                    row[3] += m * (R2 * ly - R1 * lz), row[4] += m * (R0 * lz - R2 * lx), row[5] += m * (R1 * lx - R0 * ly);
                }
                add_row(acc, row, e[r], w);
            }
        }
        __syncwarp();
    }
}

template <bool WITH_LINES>
__global__ void __launch_bounds__(kSolveThreads)
    k_gn_accumulate(const uint32_t* __restrict__ p2p, const uint32_t* __restrict__ p2l, GNArgs a,
                    double* pose, double* __restrict__ partials,
                    unsigned int* __restrict__ ticket, double* __restrict__ packet,
                    const unsigned long long* __restrict__ d_n2p, const unsigned long long* __restrict__ d_n2l,
                    const uint32_t* d_done, uint32_t* step_state, double minDelta, double maxCost)
{
    if (d_done && *d_done) return;  // the device-side GN loop already converged
    if (d_n2p) a.n2p = *d_n2p;
    if (d_n2l) a.n2l = *d_n2l;
    __shared__ uint32_t stage[kWarps][32 * 18];
    double              acc[kGNV];
    gn_accumulate_body<WITH_LINES>(p2p, p2l, a, pose, stage[threadIdx.x >> 5], acc);
    const bool folded = block_reduce_to_packet<kGNV>(acc, partials, ticket, packet, blockIdx.x, gridDim.x);
    if (folded && step_state && threadIdx.x < 32)
    {
        __syncwarp();  // the packet was written by this warp's lanes
        if (threadIdx.x == 0) gn_step_device(packet, minDelta, maxCost, pose, step_state);
    }
}

// ------------------------------------------------------------------------------------------
// The WHOLE inner loop of optimal_tf_gauss_newton (optimal_tf_gauss_newton.cpp:70-366) in ONE cooperative
// launch of co-resident CTAs: per inner iteration  accumulate -> block partials -> the last CTA folds the
// packet (reduce.cuh) [-> SHARDED: all-reduces it with the peers through the NVLink mailboxes, rank order,
// bit-identical everywhere] -> applies the LDL^T step to the pose -> ONE grid barrier -> next iteration.
// Against one launch per iteration this saves the launch gaps and, for a query-sharded run, the two extra
// launches per iteration of the stand-alone all-reduce and step kernels: the 6x6 / 6x1 all-reduce the north
// star names costs one mailbox round trip inside the kernel. After convergence the remaining iterations only
// pass their barrier (the barrier count of a launch must not depend on the data: the arrival counter is
// shared with the next launch).
struct GNCoop
{
    unsigned long long* arrivals;  // monotonically increasing across launches (never reset)
    unsigned long long  target;    // arrivals value that completes this launch's barrier 0 (+ grid per further one)
};
__device__ __forceinline__ void gn_grid_barrier(const GNCoop& cs, unsigned which)
{
    __syncthreads();
    if (threadIdx.x == 0)
    {
        const unsigned long long target = cs.target + (unsigned long long)which * gridDim.x;
        __threadfence();
        atomicAdd(cs.arrivals, 1ull);
        while (*reinterpret_cast<volatile unsigned long long*>(cs.arrivals) < target) __nanosleep(20);
        __threadfence();
    }
    __syncthreads();
}

template <bool SHARDED>
__global__ void __launch_bounds__(kSolveThreads)
    k_gn_loop(const uint32_t* __restrict__ p2p, const uint32_t* __restrict__ p2l, GNArgs a, double* pose, double* __restrict__ partials,
              unsigned int* __restrict__ ticket, double* __restrict__ packet, const unsigned long long* __restrict__ d_n2p,
              const unsigned long long* __restrict__ d_n2l, uint32_t* state, double minDelta, double maxCost, uint32_t max_iter,
              GNCoop cs, PeerLaunch pl)
{
    if (d_n2p) a.n2p = *d_n2p;
    if (d_n2l) a.n2l = *d_n2l;
    __shared__ uint32_t stage[kWarps][32 * 18];
    __shared__ double   s_pose[12];
    __shared__ uint32_t s_done;
    for (uint32_t it = 0; it < max_iter; it++)
    {
        // pose and flag as the folding CTA of the previous iteration left them (behind the barrier)
        if (threadIdx.x < 12) s_pose[threadIdx.x] = __ldcg(pose + threadIdx.x);
        if (threadIdx.x == 12) s_done = __ldcg(state);
        __syncthreads();
        if (!s_done)  // (the same on every CTA — and, sharded, on every rank: the reduced packets are bit-identical)
        {
            double acc[kGNV];
            gn_accumulate_body<false>(p2p, p2l, a, s_pose, stage[threadIdx.x >> 5], acc);
            const bool folded = block_reduce_to_packet<kGNV>(acc, partials, ticket, packet, blockIdx.x, gridDim.x);
            if (folded && threadIdx.x < 32)
            {
                __syncwarp();  // the packet was written by this warp's lanes
                if (SHARDED) peer_allreduce_warp(pl.view, pl.pkt_epoch + it, packet, threadIdx.x), __syncwarp();
                if (threadIdx.x == 0)
                {
                    gn_step_device(packet, minDelta, maxCost, pose, state);
                    __threadfence();
                }
            }
        }
        else if (SHARDED && blockIdx.x == 0 && threadIdx.x < 32)
            // converged: the exchange still takes place (on a packet nobody reads) — the number of exchanges of a
            // call must not depend on the data, a peer on the multi-launch path is waiting for this one
            peer_allreduce_warp(pl.view, pl.pkt_epoch + it, packet, threadIdx.x);
        gn_grid_barrier(cs, it);
    }
}

// ------------------------------------------------------------------------------------------
constexpr int kH1V = 8;  // sum local(3), sum global(3), count (non-outliers), pairs (all)

__global__ void __launch_bounds__(kSolveThreads)
    k_horn_sums(const uint32_t* __restrict__ p2p, uint64_t n, const uint8_t* __restrict__ outlier,
                double* __restrict__ partials, unsigned int* __restrict__ ticket,
                double* __restrict__ packet, const unsigned long long* __restrict__ d_n)
{
    if (d_n) n = *d_n;
    __shared__ uint32_t stage[kWarps][32 * 9];
    const int           lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t*           sh = stage[warp];
    double              acc[kH1V];
#pragma unroll
    for (int v = 0; v < kH1V; v++) acc[v] = 0;
    const uint64_t warp_global = (uint64_t)blockIdx.x * kWarps + warp;
    const uint64_t warp_stride = (uint64_t)gridDim.x * kWarps;
    for (uint64_t base = warp_global * 32; base < n; base += warp_stride * 32)
    {
        warp_stage_records<9>(p2p, base, n, sh);
        const uint64_t i = base + lane;
        if (i < n) acc[7] += 1.0;
        if (i < n && !(outlier && outlier[i]))
        {
            const uint32_t* rec = sh + lane * 9;
            acc[3] += (double)__uint_as_float(rec[2]), acc[4] += (double)__uint_as_float(rec[3]);
            acc[5] += (double)__uint_as_float(rec[4]);
            acc[0] += (double)__uint_as_float(rec[5]), acc[1] += (double)__uint_as_float(rec[6]);
            acc[2] += (double)__uint_as_float(rec[7]);
            acc[6] += 1.0;
        }
        __syncwarp();
    }
    block_reduce_to_packet<kH1V>(acc, partials, ticket, packet, blockIdx.x, gridDim.x);
}

struct HornArgs
{
    uint64_t n, n_total;
    int      n_total_mode;  // 0: n_total given; 1: = *d_n (single GPU, fused); 2: = sums[7] (all-reduced)
    int      use_scale_outlier;
    double   scale_thr, w_pt2pt;
    int      kernel;
    double   kparam;
    double   Rref[9];  // rotation+translation of currentEstimateForRobust
    double   tref[3];
    uint32_t n_wblocks;
};

constexpr int kH2V = 12;  // S(9), w_sum, new outliers, pairs used

__global__ void __launch_bounds__(kSolveThreads)
    k_horn_moments(const uint32_t* __restrict__ p2p, HornArgs a, const double* __restrict__ sums,
                   const uint64_t* __restrict__ wprefix, const double* __restrict__ wvalue,
                   uint8_t* __restrict__ outlier, uint64_t first_global_index,
                   double* __restrict__ partials, unsigned int* __restrict__ ticket,
                   double* __restrict__ packet, const unsigned long long* __restrict__ d_n)
{
    if (d_n) a.n = *d_n;
    if (a.n_total_mode == 1) a.n_total = a.n;
    if (a.n_total_mode == 2) a.n_total = (uint64_t)sums[7];
    __shared__ uint32_t stage[kWarps][32 * 9];
    const int           lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t*           sh = stage[warp];
    // centroids (Pairings.cpp:78,104-106): sums * 1/(n - outliers)
    const double wc   = 1.0 / sums[6];
    const double cl[3] = {sums[0] * wc, sums[1] * wc, sums[2] * wc};
    const double cg[3] = {sums[3] * wc, sums[4] * wc, sums[5] * wc};
    // visit_correspondences.h:85-87: waPoints = wPt / (wPt * nPt2Pt)
    const double waPoints = a.w_pt2pt / (a.w_pt2pt * (double)a.n_total);
    double       acc[kH2V];
#pragma unroll
    for (int v = 0; v < kH2V; v++) acc[v] = 0;
    const uint64_t warp_global = (uint64_t)blockIdx.x * kWarps + warp;
    const uint64_t warp_stride = (uint64_t)gridDim.x * kWarps;
    for (uint64_t base = warp_global * 32; base < a.n; base += warp_stride * 32)
    {
        warp_stage_records<9>(p2p, base, a.n, sh);
        const uint64_t i = base + lane;
        if (i < a.n && !(outlier && outlier[i]))
        {
            const uint32_t* rec = sh + lane * 9;
            const double    bi[3] = {(double)__uint_as_float(rec[2]) - cg[0], (double)__uint_as_float(rec[3]) - cg[1],
                                     (double)__uint_as_float(rec[4]) - cg[2]};
            const double    ri[3] = {(double)__uint_as_float(rec[5]) - cl[0], (double)__uint_as_float(rec[6]) - cl[1],
                                     (double)__uint_as_float(rec[7]) - cl[2]};
            const double bn = sqrt(bi[0] * bi[0] + bi[1] * bi[1] + bi[2] * bi[2]);
            const double rn = sqrt(ri[0] * ri[0] + ri[1] * ri[1] + ri[2] * ri[2]);
            bool         use = !(bn < 1e-4 || rn < 1e-4);  // visit_correspondences.h:141-146
            if (use && a.use_scale_outlier)                  // :158-169
            {
                const double mism = fmax(bn, rn) / fmin(bn, rn);
                if (mism > a.scale_thr)
                {
                    use = false;
                    if (outlier) outlier[i] = 1;
                    acc[10] += 1.0;
                }
            }
            if (use)
            {
                double wi = waPoints;
                if (a.n_wblocks)  // Pairings::point_weights run-length blocks (:127-133)
                {
                    const uint64_t gi = first_global_index + i;
                    uint32_t       lo = 0, hi = a.n_wblocks - 1;
                    while (lo < hi)
                    {
                        const uint32_t mid = (lo + hi) >> 1;
                        if (gi < wprefix[mid + 1])
                            hi = mid;
                        else
                            lo = mid + 1;
                    }
                    wi *= wvalue[lo];
                }
                if (a.kernel)  // :200-210
                {
                    double r2[3];
#pragma unroll
                    for (int r = 0; r < 3; r++)
                        r2[r] = a.Rref[3 * r] * ri[0] + a.Rref[3 * r + 1] * ri[1] + a.Rref[3 * r + 2] * ri[2] + a.tref[r];
                    const double e2 = (r2[0] - bi[0]) * (r2[0] - bi[0]) + (r2[1] - bi[1]) * (r2[1] - bi[1]) +
                                      (r2[2] - bi[2]) * (r2[2] - bi[2]);
                    wi *= robust_weight(a.kernel, a.kparam, e2);
                }
                acc[9] += wi;
                acc[11] += 1.0;
#pragma unroll
                for (int r = 0; r < 3; r++)
#pragma unroll
                    for (int c = 0; c < 3; c++) acc[3 * r + c] += wi * ri[r] * bi[c];  // S += w r b^T
            }
        }
        __syncwarp();
    }
    block_reduce_to_packet<kH2V>(acc, partials, ticket, packet, blockIdx.x, gridDim.x);
}

}  // namespace
// scratch of the one-launch reductions: [0,64) the (self re-arming) ticket counter, then one row
// of 32 doubles per CTA / tile
int solve_scratch(mp2p_b200_ctx* ctx, size_t rows, unsigned int** ticket, double** partials)
{
    const size_t need = 64 + rows * 32 * sizeof(double);
    if (ctx->d_partials.bytes < need)
    {
        MP2P_TRY(ctx->d_partials.ensure(std::max<size_t>(need, 64 + (size_t)148 * 8 * 32 * sizeof(double))));
        MP2P_CUDA_TRY(cudaMemsetAsync(ctx->d_partials.p, 0, ctx->d_partials.bytes, ctx->stream));
    }
    *ticket   = ctx->d_partials.as<unsigned int>();
    *partials = reinterpret_cast<double*>(ctx->d_partials.as<char>() + 64);
    return 0;
}

namespace
{
int solve_grid(uint64_t n)
{
    // latency-bound at ICP sizes: one record per thread (one staging trip per warp) until the
    // machine is full (8 CTAs per SM), grid-stride beyond that
    const uint64_t blocks = (n + kSolveThreads - 1) / kSolveThreads;
    return (int)std::max<uint64_t>(1, std::min<uint64_t>(blocks, 148 * 8));
}
}  // namespace

int run_gn_accumulate(mp2p_b200_ctx* ctx, const mp2p_b200_pair_pt2pt* d2p, uint64_t n2p,
                      const mp2p_b200_pair_pt2pl* d2l, uint64_t n2l, const mp2p_b200_gn_params* prm,
                      const double* d_pose, double* d_packet, const unsigned long long* d_n2p,
                      const unsigned long long* d_n2l, const uint32_t* d_done, uint32_t* d_step_state,
                      const mp2p_b200_pair_pt2ln* d2ln, uint64_t n2ln, double w_pt2ln)
{
    const int blocks = solve_grid(std::max(std::max(n2p, n2l), n2ln));
    unsigned int* ticket;
    double*       partials;
    MP2P_TRY(solve_scratch(ctx, blocks, &ticket, &partials));
    GNArgs a{n2p, n2l, prm->w_pt2pt, prm->w_pt2pl, prm->kernel, prm->kernelParam, n2ln, w_pt2ln,
             reinterpret_cast<const uint32_t*>(d2ln)};
    prof_begin(ctx, 4);
    if (n2ln)
        k_gn_accumulate<true><<<blocks, kSolveThreads, 0, ctx->stream>>>(
            reinterpret_cast<const uint32_t*>(d2p), reinterpret_cast<const uint32_t*>(d2l), a, const_cast<double*>(d_pose),
            partials, ticket, d_packet, d_n2p, d_n2l, d_done, d_step_state, prm->minDelta, prm->maxCost);
    else
        k_gn_accumulate<false><<<blocks, kSolveThreads, 0, ctx->stream>>>(
            reinterpret_cast<const uint32_t*>(d2p), reinterpret_cast<const uint32_t*>(d2l), a, const_cast<double*>(d_pose),
            partials, ticket, d_packet, d_n2p, d_n2l, d_done, d_step_state, prm->minDelta, prm->maxCost);
    prof_end(ctx, 4);
    count_launch(ctx);
    return 0;
}

__global__ void k_gn_step(const double* __restrict__ packet, double minDelta, double maxCost,
                          double* __restrict__ pose, uint32_t* __restrict__ state)
{
    if (threadIdx.x == 0) gn_step_device(packet, minDelta, maxCost, pose, state);
}

// The inner loop as ONE cooperative launch (k_gn_loop). Returns 1 — nothing enqueued — when that is not possible
// (no cooperative launch on the device, switched off by $MP2P_GN_COOP=0): the caller then enqueues one launch per
// iteration. `peer` != NULL: the packet is all-reduced with the peers inside the kernel (epochs pkt_epoch+1 ...).
int run_gn_coop_loop(mp2p_b200_ctx* ctx, const mp2p_b200_pair_pt2pt* d2p, uint64_t n2p, const mp2p_b200_pair_pt2pl* d2l, uint64_t n2l,
                     const mp2p_b200_gn_params* prm, double* d_pose, uint32_t* d_state, double* d_packet,
                     const unsigned long long* d_n2p, const unsigned long long* d_n2l, mp2p_b200_peer* peer)
{
    if (prm->maxInnerLoopIterations == 0) return 1;
    // co-residency bound per device (contexts on different GPUs of one process must not share it)
    static int occ[64][2];
    static bool have[64] = {};
    const int  dev = ctx->device;
    if (dev < 0 || dev >= 64) return 1;
    if (!have[dev])
    {
        int coop = 0, n_sm = 0, b0 = 0, b1 = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        if (coop)
        {
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b0, k_gn_loop<false>, kSolveThreads, 0);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b1, k_gn_loop<true>, kSolveThreads, 0);
        }
        const char* e = getenv("MP2P_GN_COOP");
        if (e && atoi(e) == 0) b0 = b1 = 0;
        occ[dev][0] = b0 * n_sm, occ[dev][1] = b1 * n_sm;
        have[dev]   = true;
    }
    const int cap = occ[dev][peer ? 1 : 0];
    if (cap <= 0) return 1;
    // (the loop is bound by its grid barrier and the fold of the CTA partials, not by the pass over the records:
    // $MP2P_GN_BLOCKS caps the grid for measurements)
    static const int max_blocks = [] {
        const char* e = getenv("MP2P_GN_BLOCKS");
        return e ? atoi(e) : 0;
    }();
    int blocks = std::min(solve_grid(std::max(n2p, n2l)), cap);
    if (max_blocks > 0) blocks = std::min(blocks, max_blocks);
    unsigned int* ticket;
    double*       partials;
    MP2P_TRY(solve_scratch(ctx, blocks, &ticket, &partials));
    if (!ctx->d_coop.p)
    {
        MP2P_TRY(ctx->d_coop.ensure(64));
        MP2P_CUDA_TRY(cudaMemsetAsync(ctx->d_coop.p, 0, 64, ctx->stream));
        ctx->coop_arrivals = 0, ctx->coop_epoch = 0;
    }
    GNArgs a{n2p, n2l, prm->w_pt2pt, prm->w_pt2pl, prm->kernel, prm->kernelParam, 0, 1.0, nullptr};
    GNCoop cs{ctx->d_coop.as<unsigned long long>(), ctx->coop_arrivals + (unsigned long long)blocks};
    PeerLaunch pl{};
    if (peer) pl.view = peer->view, pl.pkt_epoch = peer->pkt_epoch + 1;
    const uint32_t* p2p   = reinterpret_cast<const uint32_t*>(d2p);
    const uint32_t* p2l   = reinterpret_cast<const uint32_t*>(d2l);
    double          minDelta = prm->minDelta, maxCost = prm->maxCost;
    uint32_t        max_iter = prm->maxInnerLoopIterations;
    MP2P_CUDA_TRY(cudaMemsetAsync(d_state, 0, 8, ctx->stream));
    void* args[] = {&p2p, &p2l, &a, &d_pose, &partials, &ticket, &d_packet, &d_n2p, &d_n2l, &d_state, &minDelta, &maxCost, &max_iter, &cs, &pl};
    void* fn = peer ? reinterpret_cast<void*>(k_gn_loop<true>) : reinterpret_cast<void*>(k_gn_loop<false>);
    prof_begin(ctx, 4);
    const cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(blocks), dim3(kSolveThreads), args, 0, ctx->stream);
    prof_end(ctx, 4);
    if (e != cudaSuccess)
    {
        cudaGetLastError();  // e.g. cudaErrorCooperativeLaunchTooLarge: take the multi-launch path
        return 1;
    }
    ctx->coop_arrivals += (unsigned long long)max_iter * blocks;
    if (peer) peer->pkt_epoch += max_iter;
    count_launch(ctx);
    return 0;
}

int run_gn_device_loop(mp2p_b200_ctx* ctx, const mp2p_b200_pair_pt2pt* d2p, uint64_t n2p,
                       const mp2p_b200_pair_pt2pl* d2l, uint64_t n2l, const mp2p_b200_gn_params* prm,
                       double* d_pose, uint32_t* d_state, double* d_packet,
                       const unsigned long long* d_n2p, const unsigned long long* d_n2l,
                       const mp2p_b200_pair_pt2ln* d2ln, uint64_t n2ln, double w_pt2ln)
{
    if (!n2ln)  // the pt2pt / pt2pl loop in one cooperative launch
    {
        const int rc = run_gn_coop_loop(ctx, d2p, n2p, d2l, n2l, prm, d_pose, d_state, d_packet, d_n2p, d_n2l, nullptr);
        if (rc <= 0) return rc;
    }
    MP2P_CUDA_TRY(cudaMemsetAsync(d_state, 0, 8, ctx->stream));
    for (uint32_t it = 0; it < prm->maxInnerLoopIterations; it++)  // optimal_tf_gauss_newton.cpp:70
    {
        // accumulate + update in one launch (the folding CTA applies the step)
        MP2P_TRY(run_gn_accumulate(ctx, d2p, n2p, d2l, n2l, prm, d_pose, d_packet, d_n2p, d_n2l, d_state, d_state, d2ln,
                                   n2ln, w_pt2ln));
    }
    return 0;
}

int run_gn_step(mp2p_b200_ctx* ctx, const double* d_packet, const mp2p_b200_gn_params* prm, double* d_pose,
                uint32_t* d_state)
{
    k_gn_step<<<1, 32, 0, ctx->stream>>>(d_packet, prm->minDelta, prm->maxCost, d_pose, d_state);
    count_launch(ctx);
    return 0;
}

int run_horn_sums(mp2p_b200_ctx* ctx, const mp2p_b200_pair_pt2pt* d2p, uint64_t n,
                  const uint8_t* d_outlier, double* d_packet, const unsigned long long* d_n)
{
    const int blocks = solve_grid(n);
    unsigned int* ticket;
    double*       partials;
    MP2P_TRY(solve_scratch(ctx, blocks, &ticket, &partials));
    prof_begin(ctx, 2);
    k_horn_sums<<<blocks, kSolveThreads, 0, ctx->stream>>>(reinterpret_cast<const uint32_t*>(d2p), n,
                                                          d_outlier, partials, ticket,
                                                          d_packet, d_n);
    prof_end(ctx, 2);
    count_launch(ctx);
    return 0;
}

int run_horn_moments(mp2p_b200_ctx* ctx, const mp2p_b200_pair_pt2pt* d2p, uint64_t n,
                     const mp2p_b200_horn_params* prm, const double* d_sums_packet,
                     uint64_t n_total_pairs, const uint64_t* d_wcount_prefix, const double* d_wvalue,
                     uint32_t n_wblocks, uint8_t* d_outlier, double* d_packet, const unsigned long long* d_n,
                     int n_total_mode)
{
    const int     blocks = solve_grid(n);
    unsigned int* ticket;
    double*       partials;
    MP2P_TRY(solve_scratch(ctx, blocks, &ticket, &partials));
    HornArgs a{};
    a.n = n, a.n_total = n_total_pairs, a.n_total_mode = n_total_mode;
    a.use_scale_outlier = prm->use_scale_outlier_detector, a.scale_thr = prm->scale_outlier_threshold;
    a.w_pt2pt = prm->w_pt2pt, a.kernel = prm->robust_kernel, a.kparam = prm->robust_kernel_param;
    for (int r = 0; r < 3; r++)
    {
        for (int c = 0; c < 3; c++) a.Rref[3 * r + c] = prm->currentEstimateForRobust[4 * r + c];
        a.tref[r] = prm->currentEstimateForRobust[4 * r + 3];
    }
    a.n_wblocks = n_wblocks;
    prof_begin(ctx, 3);
    k_horn_moments<<<blocks, kSolveThreads, 0, ctx->stream>>>(reinterpret_cast<const uint32_t*>(d2p), a,
                                                             d_sums_packet, d_wcount_prefix, d_wvalue,
                                                             d_outlier, 0, partials, ticket,
                                                             d_packet, d_n);
    prof_end(ctx, 3);
    count_launch(ctx);
    return 0;
}
// ------------------------------------------------------------------------------------------
// pt2ln_pl_to_pt2pt, plane part (pt2ln_pl_to_pt2pt.cpp:25-113): Solver_Horn turns every pt2pl
// pairing into a pt2pt pairing "local point -> its projection on the plane" (Solver_Horn.cpp:51-55)
// and keeps those whose |distance| is at least 25 % of the largest (or the 3 largest).
//   k_pl2pt_project  one thread per pairing: g = T (+) l in double, d = n.g + D, record
//                    {0, 0, float(g - n d), l, 0}, |d| kept aside, atomicMax of its bit pattern
//   k_pl2pt_count / k_pl2pt_scan / k_pl2pt_write   ordered compaction of the records that pass
// The reference emits them by descending |d|; Horn's sums do not depend on the order, this path
// keeps the input order (documented in the header).
// ------------------------------------------------------------------------------------------
namespace
{
constexpr int kConvThreads = 256;
struct PoseArg
{
    double m[12];
};

__global__ void __launch_bounds__(kConvThreads)
    k_pl2pt_project(const mp2p_b200_pair_pt2pl* __restrict__ in, uint64_t n, PoseArg T,
                    mp2p_b200_pair_pt2pt* __restrict__ rec, double* __restrict__ absd, unsigned long long* __restrict__ maxbits)
{
    const uint64_t i = (uint64_t)blockIdx.x * kConvThreads + threadIdx.x;
    double         a = 0.0;
    if (i < n)
    {
        const mp2p_b200_pair_pt2pl p  = in[i];
        const double               lx = p.local_x, ly = p.local_y, lz = p.local_z;
        const double gx = T.m[0] * lx + T.m[1] * ly + T.m[2] * lz + T.m[3];  // composePoint, double
        const double gy = T.m[4] * lx + T.m[5] * ly + T.m[6] * lz + T.m[7];
        const double gz = T.m[8] * lx + T.m[9] * ly + T.m[10] * lz + T.m[11];
        const double d  = p.plane_coefs[0] * gx + p.plane_coefs[1] * gy + p.plane_coefs[2] * gz + p.plane_coefs[3];
        mp2p_b200_pair_pt2pt r;
        r.globalIdx = 0, r.localIdx = 0;  // dummies, as in the reference
        r.global_x = (float)(gx - p.plane_coefs[0] * d), r.global_y = (float)(gy - p.plane_coefs[1] * d);
        r.global_z = (float)(gz - p.plane_coefs[2] * d);
        r.local_x = p.local_x, r.local_y = p.local_y, r.local_z = p.local_z;
        r.errorSquareAfterTransformation = 0.f;
        rec[i]  = r;
        a       = fabs(d);
        absd[i] = a;
    }
    // non-negative doubles order like their bit patterns
    unsigned long long b = (unsigned long long)__double_as_longlong(a);
    for (int o = 16; o > 0; o >>= 1) b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
    if ((threadIdx.x & 31) == 0 && b) atomicMax(maxbits, b);
}

// thr = ratio * max unless thr_override >= 0
__device__ __forceinline__ double conv_threshold(const unsigned long long* maxbits, double thr_override)
{
    return thr_override >= 0.0 ? thr_override : __longlong_as_double((long long)*maxbits) * 0.25;
}

__global__ void __launch_bounds__(kConvThreads)
    k_pl2pt_count(const double* __restrict__ absd, uint64_t n, const unsigned long long* __restrict__ maxbits,
                  double thr_override, uint32_t* __restrict__ counts)
{
    const double   thr  = conv_threshold(maxbits, thr_override);
    const uint64_t i    = (uint64_t)blockIdx.x * kConvThreads + threadIdx.x;
    const int      keep = (i < n && !(absd[i] < thr)) ? 1 : 0;
    const int      c    = __syncthreads_count(keep);
    if (threadIdx.x == 0) counts[blockIdx.x] = (uint32_t)c;
}

// one CTA: exclusive scan of the per-block counts in place, total -> *total
__global__ void __launch_bounds__(1024) k_pl2pt_scan(uint32_t* __restrict__ counts, uint32_t n_blocks, unsigned long long* __restrict__ total)
{
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (uint32_t b0 = 0; b0 < n_blocks; b0 += 1024)
    {
        const uint32_t b = b0 + threadIdx.x;
        const uint32_t v = b < n_blocks ? counts[b] : 0u;
        uint32_t       x = v;
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32)
        {
            uint32_t w = s_warp[threadIdx.x];
            for (int o = 1; o < 32; o <<= 1)
            {
                const uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
                if (threadIdx.x >= o) w += y;
            }
            s_warp[threadIdx.x] = w;  // inclusive over warps
        }
        __syncthreads();
        const uint32_t warp_excl = (threadIdx.x >> 5) ? s_warp[(threadIdx.x >> 5) - 1] : 0u;
        const uint32_t base      = s_base;
        if (b < n_blocks) counts[b] = base + warp_excl + x - v;
        __syncthreads();
        if (threadIdx.x == 0) s_base = base + s_warp[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = s_base;
}

__global__ void __launch_bounds__(kConvThreads)
    k_pl2pt_write(const mp2p_b200_pair_pt2pt* __restrict__ rec, const double* __restrict__ absd, uint64_t n,
                  const unsigned long long* __restrict__ maxbits, double thr_override, const uint32_t* __restrict__ offsets,
                  mp2p_b200_pair_pt2pt* __restrict__ out, uint64_t capacity)
{
    __shared__ uint32_t s_warp[kConvThreads / 32];
    const double        thr  = conv_threshold(maxbits, thr_override);
    const uint64_t      i    = (uint64_t)blockIdx.x * kConvThreads + threadIdx.x;
    const bool          keep = i < n && !(absd[i] < thr);
    const unsigned      m    = __ballot_sync(0xffffffffu, keep);
    const int           lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) s_warp[w] = __popc(m);
    __syncthreads();
    uint32_t off = offsets[blockIdx.x];
    for (int k = 0; k < w; k++) off += s_warp[k];
    off += __popc(m & ((1u << lane) - 1u));
    if (keep && off < capacity) out[off] = rec[i];
}
}  // namespace

// d_in: n pt2pl pairings (device). d_out: room for min(n, capacity) records (device). *h_total (host)
// receives the number of records kept. Synchronises the stream.
int run_pt2pl_to_pt2pt(mp2p_b200_ctx* ctx, const mp2p_b200_pair_pt2pl* d_in, uint64_t n, const double pose[12],
                       mp2p_b200_pair_pt2pt* d_out, uint64_t capacity, uint64_t* h_total)
{
    *h_total = 0;
    if (n == 0) return 0;
    if (n >= 0xFFFFFFFFull)
    {
        set_error("pt2pl_to_pt2pt: n must be < 2^32-1");
        return MP2P_B200_ERR_ARG;
    }
    const uint32_t blocks = (uint32_t)((n + kConvThreads - 1) / kConvThreads);
    // scratch: [maxbits u64][total u64][counts u32 x blocks (padded)][absd f64 x n][records 36 B x n]
    const size_t off_counts = 16, off_absd = (off_counts + (size_t)blocks * 4 + 15) & ~(size_t)15;
    const size_t off_rec = off_absd + n * 8;
    MP2P_TRY(ctx->d_conv.ensure(off_rec + n * sizeof(mp2p_b200_pair_pt2pt)));
    char* base    = ctx->d_conv.as<char>();
    auto* maxbits = reinterpret_cast<unsigned long long*>(base);
    auto* total   = maxbits + 1;
    auto* counts  = reinterpret_cast<uint32_t*>(base + off_counts);
    auto* absd    = reinterpret_cast<double*>(base + off_absd);
    auto* rec     = reinterpret_cast<mp2p_b200_pair_pt2pt*>(base + off_rec);
    cudaStream_t st = ctx->stream;
    MP2P_CUDA_TRY(cudaMemsetAsync(base, 0, 16, st));
    PoseArg T;
    for (int k = 0; k < 12; k++) T.m[k] = pose[k];
    k_pl2pt_project<<<blocks, kConvThreads, 0, st>>>(d_in, n, T, rec, absd, maxbits);
    count_launch(ctx);
    unsigned long long* h = static_cast<unsigned long long*>(ctx->h_pinned);
    double thr_override   = -1.0;
    for (int pass = 0; pass < 2; pass++)
    {
        k_pl2pt_count<<<blocks, kConvThreads, 0, st>>>(absd, n, maxbits, thr_override, counts);
        k_pl2pt_scan<<<1, 1024, 0, st>>>(counts, blocks, total);
        k_pl2pt_write<<<blocks, kConvThreads, 0, st>>>(rec, absd, n, maxbits, thr_override, counts, d_out, capacity);
        count_launch(ctx, 3);
        MP2P_CUDA_TRY(cudaMemcpyAsync(h, total, 8, cudaMemcpyDeviceToHost, st));
        MP2P_CUDA_TRY(cudaStreamSynchronize(st));
        MP2P_CUDA_TRY(cudaGetLastError());
        if (*h >= 3 || *h >= n || pass == 1) break;
        // fewer than 3 pairings reach 25 % of the largest error: the reference keeps the 3 largest
        // (pt2ln_pl_to_pt2pt.cpp:39-41). Rare: settle the threshold on the host.
        std::vector<double> a(n);
        MP2P_CUDA_TRY(cudaMemcpy(a.data(), absd, n * 8, cudaMemcpyDeviceToHost));
        std::nth_element(a.begin(), a.begin() + 2, a.end(), [](double x, double y) { return x > y; });
        thr_override = a[2];
    }
    *h_total = *h;
    if (*h > capacity)
    {
        set_error("pt2pl_to_pt2pt: output capacity %llu too small for %llu pairings", (unsigned long long)capacity,
                  (unsigned long long)*h);
        return MP2P_B200_ERR_CAPACITY;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// covariance() (mp2p_icp/src/covariance.cpp:28-141): the stacked error vector of the final
// pairings — error_point2point / error_point2line / error_point2plane, 3 rows per pairing — differentiated
// numerically w.r.t. (x, y, z, yaw, pitch, roll) by central differences (mrpt::math::estimateJacobian:
// column i = (f(x + h_i e_i) - f(x - h_i e_i)) * 0.5 / h_i), hessian = J^T J. The 12 perturbed poses come
// from the host; one thread per pairing evaluates its 3 x 12 error values exactly as the host loop would,
// forms its 3 x 6 block of J and adds the block's J^T J (upper triangle, 21 values) to its accumulators;
// the grid reduction is the solvers' (reduce.cuh). cov = hessian^-1 on the host.
// ------------------------------------------------------------------------------------------
namespace
{
struct CovPoses
{
    double m[12][12];  // [2 * i] = x + h_i e_i, [2 * i + 1] = x - h_i e_i, row-major 3x4 each
    double inv2h[6];
};
constexpr int kCovV = 21;

__device__ __forceinline__ void cov_point(const double* T, double lx, double ly, double lz, double (&g)[3])
{
#pragma unroll
    for (int r = 0; r < 3; r++) g[r] = T[4 * r] * lx + T[4 * r + 1] * ly + T[4 * r + 2] * lz + T[4 * r + 3];
}
__device__ __forceinline__ void cov_add_block(double (&acc)[kCovV], const double (&J)[3][6])
{
    int idx = 0;
#pragma unroll
    for (int a = 0; a < 6; a++)
#pragma unroll
        for (int b = a; b < 6; b++) acc[idx++] += J[0][a] * J[0][b] + J[1][a] * J[1][b] + J[2][a] * J[2][b];
}

__global__ void __launch_bounds__(kSolveThreads)
    k_cov_accumulate(const uint32_t* __restrict__ p2p, uint64_t n2p, const uint32_t* __restrict__ p2l, uint64_t n2l,
                     const uint32_t* __restrict__ p2ln, uint64_t n2ln, CovPoses P, double* __restrict__ partials,
                     unsigned int* __restrict__ ticket, double* __restrict__ packet)
{
    __shared__ uint32_t stage[kWarps][32 * 18];
    const int           lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t*           sh = stage[warp];
    double              acc[kCovV];
#pragma unroll
    for (int v = 0; v < kCovV; v++) acc[v] = 0;
    const uint64_t warp_global = (uint64_t)blockIdx.x * kWarps + warp;
    const uint64_t warp_stride = (uint64_t)gridDim.x * kWarps;

    for (uint64_t base = warp_global * 32; base < n2p; base += warp_stride * 32)  // covariance.cpp:75-82
    {
        warp_stage_records<9>(p2p, base, n2p, sh);
        if (base + lane < n2p)
        {
            const uint32_t* rec = sh + lane * 9;
            const double    gx = __uint_as_float(rec[2]), gy = __uint_as_float(rec[3]), gz = __uint_as_float(rec[4]);
            const double    lx = __uint_as_float(rec[5]), ly = __uint_as_float(rec[6]), lz = __uint_as_float(rec[7]);
            double          J[3][6];
#pragma unroll
            for (int i = 0; i < 6; i++)
            {
                double gp[3], gm[3];
                cov_point(P.m[2 * i], lx, ly, lz, gp), cov_point(P.m[2 * i + 1], lx, ly, lz, gm);
                J[0][i] = P.inv2h[i] * ((gp[0] - gx) - (gm[0] - gx));
                J[1][i] = P.inv2h[i] * ((gp[1] - gy) - (gm[1] - gy));
                J[2][i] = P.inv2h[i] * ((gp[2] - gz) - (gm[2] - gz));
            }
            cov_add_block(acc, J);
        }
        __syncwarp();
    }
    for (uint64_t base = warp_global * 32; base < n2ln; base += warp_stride * 32)  // :85-93, errorTerms.cpp:67-113
    {
        warp_stage_records<18>(p2ln, base, n2ln, sh);
        if (base + lane < n2ln)
        {
            const uint32_t* rec = sh + lane * 18;
            double          v[9];
#pragma unroll
            for (int k = 0; k < 9; k++) v[k] = __longlong_as_double(((long long)rec[2 * k + 1] << 32) | (long long)rec[2 * k]);
            double J[3][6];
#pragma unroll
            for (int i = 0; i < 6; i++)
            {
                double e[2][3];
#pragma unroll
                for (int sgn = 0; sgn < 2; sgn++)
                {
                    double g[3];
                    cov_point(P.m[2 * i + sgn], v[6], v[7], v[8], g);
                    const double q[3] = {g[0] - v[0], g[1] - v[1], g[2] - v[2]};
                    const double uq   = v[3] * q[0] + v[4] * q[1] + v[5] * q[2];
                    e[sgn][0] = q[0] - v[3] * uq, e[sgn][1] = q[1] - v[4] * uq, e[sgn][2] = q[2] - v[5] * uq;
                }
#pragma unroll
                for (int r = 0; r < 3; r++) J[r][i] = P.inv2h[i] * (e[0][r] - e[1][r]);
            }
            cov_add_block(acc, J);
        }
        __syncwarp();
    }
    for (uint64_t base = warp_global * 32; base < n2l; base += warp_stride * 32)  // :107-115, errorTerms.cpp:115-161
    {
        warp_stage_records<18>(p2l, base, n2l, sh);
        if (base + lane < n2l)
        {
            const uint32_t* rec = sh + lane * 18;
            double          c[4];
#pragma unroll
            for (int k = 0; k < 4; k++) c[k] = __longlong_as_double(((long long)rec[2 * k + 1] << 32) | (long long)rec[2 * k]);
            const double lx = __uint_as_float(rec[14]), ly = __uint_as_float(rec[15]), lz = __uint_as_float(rec[16]);
            const double mod_n = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
            double       J[3][6];
#pragma unroll
            for (int i = 0; i < 6; i++)
            {
                double e[2][3];
#pragma unroll
                for (int sgn = 0; sgn < 2; sgn++)
                {
                    double g[3];
                    cov_point(P.m[2 * i + sgn], lx, ly, lz, g);
                    const double ev = c[0] * g[0] + c[1] * g[1] + c[2] * g[2] + c[3];
#pragma unroll
                    for (int r = 0; r < 3; r++) e[sgn][r] = -(c[r] / mod_n) * ev;
                }
#pragma unroll
                for (int r = 0; r < 3; r++) J[r][i] = P.inv2h[i] * (e[0][r] - e[1][r]);
            }
            cov_add_block(acc, J);
        }
        __syncwarp();
    }
    block_reduce_to_packet<kCovV>(acc, partials, ticket, packet, blockIdx.x, gridDim.x);
}
}  // namespace

// poses[12][12]: the perturbed poses (see CovPoses); d_packet[0..21) <- upper triangle of J^T J, row-major
int run_cov_accumulate(mp2p_b200_ctx* ctx, const mp2p_b200_pair_pt2pt* d2p, uint64_t n2p, const mp2p_b200_pair_pt2pl* d2l,
                       uint64_t n2l, const mp2p_b200_pair_pt2ln* d2ln, uint64_t n2ln, const double poses[12][12],
                       const double inv2h[6], double* d_packet)
{
    const int     blocks = solve_grid(std::max(std::max(n2p, n2l), n2ln));
    unsigned int* ticket;
    double*       partials;
    MP2P_TRY(solve_scratch(ctx, blocks, &ticket, &partials));
    CovPoses P;
    std::memcpy(P.m, poses, sizeof(P.m));
    std::memcpy(P.inv2h, inv2h, sizeof(P.inv2h));
    k_cov_accumulate<<<blocks, kSolveThreads, 0, ctx->stream>>>(reinterpret_cast<const uint32_t*>(d2p), n2p,
                                                                reinterpret_cast<const uint32_t*>(d2l), n2l,
                                                                reinterpret_cast<const uint32_t*>(d2ln), n2ln, P, partials, ticket,
                                                                d_packet);
    count_launch(ctx);
    return 0;
}
}  // namespace mp2p
