// Peer exchange over NVLink: shared device-side pieces (mailbox layout, flags, in-CTA all-reduce).
// See peer.cu for the protocol.
#pragma once
#include "common.cuh"

namespace mp2p
{
constexpr unsigned long long kPeerTimeoutNs = 10ull * 1000ull * 1000ull * 1000ull;
constexpr int                kMaxPeers      = 16;
constexpr size_t             kHdrBytes      = 4096;  // flags: [2 parities][kMaxPeers] u32 for records, then for packets

struct PeerView
{
    char*    box[kMaxPeers];  // mailbox base of every rank (own pointer for the own rank)
    uint32_t rank, world;
    uint64_t rec_words;       // 64-bit words per exchange record
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void wait_flag(const uint32_t* flag, uint32_t epoch)
{
    const unsigned long long t0 = globaltimer_ns();
    while ((int32_t)(ld_acquire_sys(flag) - epoch) < 0)
    {
        __nanosleep(64);
        if (globaltimer_ns() - t0 > kPeerTimeoutNs) __trap();
    }
}

__device__ __forceinline__ uint32_t* rec_flag(char* box, uint32_t parity, uint32_t src) { return reinterpret_cast<uint32_t*>(box) + parity * kMaxPeers + src; }
__device__ __forceinline__ uint32_t* pkt_flag(char* box, uint32_t parity, uint32_t src) { return reinterpret_cast<uint32_t*>(box) + 2 * kMaxPeers + parity * kMaxPeers + src; }
__device__ __forceinline__ double*   pkt_slot(char* box, uint32_t parity, uint32_t src)
{
    return reinterpret_cast<double*>(box + kHdrBytes) + (size_t)(parity * kMaxPeers + src) * MP2P_B200_PACKET_DOUBLES;
}
__host__ __device__ __forceinline__ size_t rec_offset(uint64_t rec_words, uint32_t world, uint32_t parity, uint32_t src)
{
    return kHdrBytes + (size_t)2 * kMaxPeers * MP2P_B200_PACKET_DOUBLES * 8 + ((size_t)parity * world + src) * rec_words * 8;
}

// All-reduce (SUM, rank order) of a 32-double packet by ONE WARP (lane t owns double t), in place.
// `t` = lane index; must be called by all 32 lanes of the warp, converged.
__device__ __forceinline__ void peer_allreduce_warp(const PeerView& pv, uint32_t epoch, double* packet, uint32_t t)
{
    const uint32_t parity = epoch & 1u;
    const double   v      = __ldcg(packet + t);
    for (uint32_t p = 0; p < pv.world; p++) pkt_slot(pv.box[p], parity, pv.rank)[t] = v;
    __threadfence_system();
    __syncwarp();
    if (t < pv.world) st_release_sys(pkt_flag(pv.box[t], parity, pv.rank), epoch);
    if (t < pv.world) wait_flag(pkt_flag(pv.box[pv.rank], parity, t), epoch);
    __syncwarp();
    double s = 0.0;
    for (uint32_t r = 0; r < pv.world; r++) s += *reinterpret_cast<volatile double*>(pkt_slot(pv.box[pv.rank], parity, r) + t);
    packet[t] = s;
}

// what a kernel that exchanges by itself needs: the view plus the epochs the host assigned
struct PeerLaunch
{
    PeerView view;
    uint32_t rec_epoch;   // this launch's record exchange
    uint32_t pkt_epoch;   // this launch's FIRST packet all-reduce (the second one is pkt_epoch + 1)
};
}  // namespace mp2p

// host side of a peer object (peer.cu owns it; match.cu launches kernels that use its view)
struct mp2p_b200_peer
{
    mp2p_b200_ctx* ctx = nullptr;
    mp2p::PeerView view{};
    void*          own = nullptr;   // cudaMalloc'd mailbox of this rank
    void*          opened[mp2p::kMaxPeers] = {};
    size_t         bytes     = 0;
    uint32_t       rec_epoch = 0, pkt_epoch = 0;
    unsigned int*  ticket    = nullptr;
    bool           connected = false;
    // owner-partitioned first claims (mp2p_b200_peer_claims_*): this rank's part, the peers' parts (IPC mappings),
    // the device array of the `world` part pointers, the map size they were made for, the call counter
    void*                 claims_own = nullptr;
    void*                 claims_opened[mp2p::kMaxPeers] = {};
    unsigned long long**  d_claim_parts = nullptr;
    uint64_t              claims_n = 0;
    uint32_t              claim_epoch = 0;
    bool                  claims_connected = false;
};
