// Copies between PAGEABLE host memory and the device (host side of the C-ABI entry points).
//
// The reference hands the plugin std::vector storage (Pairings, CPointsMap buffers): pageable. CUDA moves
// pageable memory through its own staging buffer with ONE thread doing the host-side copy, synchronously:
// ~10 GB/s, 0.45 ms for the 4.5 MB of pairings a C3 iteration returns and 0.45 ms again for handing them to
// the solver — the two transfers were two thirds of the 1.37 ms iteration a caller with pageable buffers saw
// (bench.py, e2e.pageable; 1.04 ms with this path). Here the transfer goes through a pinned bounce buffer of the context in chunks:
//   device -> host: all chunk DMAs are enqueued at once (an event behind each); the caller and ONE helper
//                   thread (more were measured and lose, see run()) wait for the chunk their next slice belongs
//                   to and copy it out while the later chunks are in flight;
//   host -> device: the two fill the bounce buffer slice by slice; whoever completes a chunk enqueues its
//                   DMA, so the first chunk travels while the last is still being filled.
// Pinned / registered / managed memory is recognised (cudaPointerGetAttributes) and copied directly, as are
// small transfers. The workers sleep on a condition variable between transfers (no spinning while idle).
#pragma once
#include <atomic>
#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

namespace mp2p
{
class PageableCopier
{
  public:
    static constexpr size_t kMinBytes  = 256 << 10;  // below this the plain copy wins (wake-up + events)
    static constexpr size_t kSlice     = 128 << 10;  // what one thread copies at a time
    static constexpr int    kMaxChunks = 16;          // DMAs (and events) per transfer

    explicit PageableCopier(int device) : device_(device)
    {
        const char* e = getenv("MP2P_HOST_COPY_THREADS");  // helpers beside the calling thread; 0 = off
        n_workers_    = e ? atoi(e) : 1;
        if (n_workers_ < 0) n_workers_ = 0;
        if (n_workers_ > 15) n_workers_ = 15;
    }
    ~PageableCopier()
    {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
        for (auto& e : ev_)
            if (e) cudaEventDestroy(e);
        if (ev_h2d_) cudaEventDestroy(ev_h2d_);
        if (bounce_[0]) cudaFreeHost(bounce_[0]);
        if (bounce_[1]) cudaFreeHost(bounce_[1]);
    }
    bool enabled() const { return n_workers_ > 0; }

    // true if `p` is ordinary pageable host memory (not pinned, registered, managed or device memory)
    static bool pageable(const void* p)
    {
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, p) != cudaSuccess)
        {
            cudaGetLastError();
            return true;
        }
        return at.type == cudaMemoryTypeUnregistered;
    }

    // device -> pageable host; returns when the bytes are in dst. Everything enqueued on `st` before the call
    // is waited for (as a cudaMemcpyAsync + cudaStreamSynchronize would).
    cudaError_t to_host(void* dst, const void* src_dev, size_t bytes, cudaStream_t st)
    {
        if (cudaError_t e = prepare(0, bytes)) return e;
        Job j{};
        plan(j, bytes);
        j.kind = 0, j.host = static_cast<char*>(dst), j.bounce = bounce_[0], j.stream = st;
        for (int c = 0; c < j.n_chunks; c++)
        {
            const size_t off = (size_t)c * j.chunk, len = std::min(j.chunk, bytes - off);
            if (cudaError_t e = cudaMemcpyAsync(j.bounce + off, static_cast<const char*>(src_dev) + off, len, cudaMemcpyDeviceToHost, st)) return e;
            if (cudaError_t e = cudaEventRecord(ev_[c], st)) return e;
        }
        return run(j);
    }
    // pageable host -> device; returns when src has been read (the DMAs are enqueued on `st`, work enqueued
    // behind them sees the data).
    cudaError_t to_device(void* dst_dev, const void* src, size_t bytes, cudaStream_t st)
    {
        if (cudaError_t e = prepare(1, bytes)) return e;
        if (h2d_pending_)  // the previous transfer's DMAs still read the bounce buffer
        {
            if (cudaError_t e = cudaEventSynchronize(ev_h2d_)) return e;
            h2d_pending_ = false;
        }
        Job j{};
        plan(j, bytes);
        j.kind = 1, j.host = const_cast<char*>(static_cast<const char*>(src)), j.bounce = bounce_[1], j.stream = st;
        j.dev = static_cast<char*>(dst_dev);
        if (cudaError_t e = run(j)) return e;
        h2d_pending_ = true;
        return cudaEventRecord(ev_h2d_, st);
    }

  private:
    struct Job
    {
        int          kind = 0;  // 0: device -> host, 1: host -> device
        char *       host = nullptr, *bounce = nullptr, *dev = nullptr;
        size_t       bytes = 0, chunk = 0;
        int          n_chunks = 0, n_slices = 0, slices_per_chunk = 0;
        cudaStream_t stream = nullptr;
    };
    void plan(Job& j, size_t bytes) const
    {
        j.bytes            = bytes;
        j.slices_per_chunk = (int)std::max<size_t>(1, ((bytes + kMaxChunks - 1) / kMaxChunks + kSlice - 1) / kSlice);
        j.chunk            = (size_t)j.slices_per_chunk * kSlice;
        j.n_chunks         = (int)((bytes + j.chunk - 1) / j.chunk);
        j.n_slices         = (int)((bytes + kSlice - 1) / kSlice);
    }
    cudaError_t prepare(int which, size_t bytes)
    {
        if (!ev_[0])
        {
            for (auto& e : ev_)
                if (cudaError_t r = cudaEventCreateWithFlags(&e, cudaEventDisableTiming)) return r;
            if (cudaError_t r = cudaEventCreateWithFlags(&ev_h2d_, cudaEventDisableTiming)) return r;
        }
        if (bounce_bytes_[which] < bytes)
        {
            if (which == 1 && h2d_pending_) cudaEventSynchronize(ev_h2d_), h2d_pending_ = false;
            if (bounce_[which]) cudaFreeHost(bounce_[which]);
            bounce_[which] = nullptr, bounce_bytes_[which] = 0;
            const size_t want = bytes + bytes / 4 + (1 << 20);
            if (cudaError_t r = cudaHostAlloc(reinterpret_cast<void**>(&bounce_[which]), want, cudaHostAllocDefault)) return r;
            bounce_bytes_[which] = want;
        }
        if (workers_.empty())
            for (int w = 0; w < n_workers_; w++) workers_.emplace_back([this] { worker(); });
        return cudaSuccess;
    }
    // The calling thread publishes the job, works on it like a helper and waits for the last slice. Every
    // thread waits for the event of the chunk its slice belongs to (device -> host) or enqueues the DMA of the
    // chunk it completes (host -> device) by itself.
    // Measured (C3, 2 x 4.55 MB per iteration, 16-vCPU box; iteration time with pageable buffers):
    //   helpers        0 (plain cudaMemcpyAsync)   1        3        7        15
    //   ms             1.33                        1.04     1.14     1.30     1.46
    // More helpers lose to their wake-ups and to one another; a variant in which only the caller talks to
    // CUDA (helpers spin on a landed-chunks counter) was slower still (1.13 / 1.15 / 1.23 with 1 / 2 / 3).
    cudaError_t run(const Job& j)
    {
        {
            // job_ and the counters change only while no helper is inside work(): helpers enter it (busy_++)
            // under the mutex, so with the mutex held and busy_ == 0 none is in and none can get in
            std::unique_lock<std::mutex> lk(m_);
            while (busy_.load(std::memory_order_acquire) != 0)
            {
                lk.unlock();
                std::this_thread::yield();
                lk.lock();
            }
            job_ = j;
            next_.store(0, std::memory_order_relaxed), done_.store(0, std::memory_order_relaxed);
            err_.store(0, std::memory_order_relaxed);
            for (int c = 0; c < j.n_chunks; c++) filled_[c].store(0, std::memory_order_relaxed);
            generation_++;
        }
        cv_.notify_all();
        work();
        while (done_.load(std::memory_order_acquire) < j.n_slices) std::this_thread::yield();
        return (cudaError_t)err_.load();
    }
    void work()
    {
        const Job& j = job_;
        for (;;)
        {
            const int s = next_.fetch_add(1, std::memory_order_relaxed);
            if (s >= j.n_slices) break;
            const size_t off = (size_t)s * kSlice, len = std::min(kSlice, j.bytes - off);
            const int    c   = s / j.slices_per_chunk;
            if (j.kind == 0)
            {
                const cudaError_t e = cudaEventSynchronize(ev_[c]);  // the chunk this slice belongs to has landed
                if (e != cudaSuccess) err_.store((int)e);
                std::memcpy(j.host + off, j.bounce + off, len);
            }
            else
            {
                std::memcpy(j.bounce + off, j.host + off, len);
                const size_t coff = (size_t)c * j.chunk, clen = std::min(j.chunk, j.bytes - coff);
                const int    in_chunk = (int)((clen + kSlice - 1) / kSlice);
                if (filled_[c].fetch_add(1, std::memory_order_acq_rel) + 1 == in_chunk)
                {
                    const cudaError_t e = cudaMemcpyAsync(j.dev + coff, j.bounce + coff, clen, cudaMemcpyHostToDevice, j.stream);
                    if (e != cudaSuccess) err_.store((int)e);
                }
            }
            done_.fetch_add(1, std::memory_order_release);
        }
    }
    void worker()
    {
        cudaSetDevice(device_);
        unsigned long long seen = 0;
        for (;;)
        {
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
                busy_.fetch_add(1, std::memory_order_acq_rel);
            }
            work();
            busy_.fetch_sub(1, std::memory_order_acq_rel);
        }
    }

    int                      device_;
    int                      n_workers_ = 0;
    std::vector<std::thread> workers_;
    std::mutex               m_;
    std::condition_variable  cv_;
    unsigned long long       generation_ = 0;
    bool                     stop_       = false;
    Job                      job_{};
    std::atomic<int>         next_{0}, done_{0}, err_{0}, busy_{0};
    std::atomic<int>         filled_[kMaxChunks];
    cudaEvent_t              ev_[kMaxChunks] = {};
    cudaEvent_t              ev_h2d_         = nullptr;
    bool                     h2d_pending_    = false;
    char*                    bounce_[2]      = {nullptr, nullptr};
    size_t                   bounce_bytes_[2] = {0, 0};
};
}  // namespace mp2p
