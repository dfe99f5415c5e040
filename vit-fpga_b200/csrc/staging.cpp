// staging.cpp -- host-side copy engine behind the host-buffer forward calls (host code only, no CUDA).
//
// net::net_abstract::launch_forward takes a pageable std::vector (include/netAbstract.h:13).  The reference moves it to
// its device with one blocking clEnqueueWriteBuffer per sample (src/netFPGA.cpp:266-273); here a ViT-B batch of 1024
// images is 616 MB per call, and the copy into the page-locked staging slots -- not the GPU -- sets the pace of the call
// unless it runs at memory speed.  So:
//   * one process-wide pool of copy threads (started on first use, shared by every handle and GPU of the process, sized
//     to the cores the process may run on), fed with 1 MiB pieces: no thread creation per call, and several GPUs' host
//     threads share the cores instead of oversubscribing them;
//   * pieces are copied with non-temporal stores (AVX2 `vmovntdq`, 32-byte aligned destination): the staging slot is
//     written once and read only by the DMA engine, so it must neither evict the caller's data from the caches nor pay the
//     read-for-ownership of a cached store (a third less memory traffic per byte staged).
#include "kernels.h"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include <sched.h>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace nc
{

namespace
{

#if defined(__x86_64__)
__attribute__((target("avx2"))) void copy_nt_avx2(char *dst, const char *src, size_t n)
{
    // head: up to the first 32-byte boundary of the destination
    size_t head = (32 - (reinterpret_cast<uintptr_t>(dst) & 31u)) & 31u;
    if (head > n) head = n;
    if (head) memcpy(dst, src, head);
    dst += head, src += head, n -= head;
    size_t i = 0;
    for (; i + 128 <= n; i += 128)
    {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 64));
        const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i), a);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 96), d);
    }
    if (i < n) memcpy(dst + i, src + i, n - i);
    _mm_sfence(); // the non-temporal stores are globally visible before the piece is reported done (the DMA reads them next)
}
#endif

void copy_piece(char *dst, const char *src, size_t n)
{
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2 && n >= 4096)
    {
        copy_nt_avx2(dst, src, n);
        return;
    }
#endif
    memcpy(dst, src, n);
#if defined(__x86_64__)
    _mm_sfence(); // the destination may be write-combined memory: drain this core's WC buffers before the DMA is queued
#endif
}

constexpr size_t PIECE = 1u << 20;

struct Job
{
    char *dst;
    const char *src;
    size_t bytes;
    std::atomic<size_t> next{0};      // next piece to hand out
    std::atomic<size_t> remaining{0}; // pieces not yet finished
    size_t pieces = 0;
};

class CopyPool
{
  public:
    static CopyPool &get()
    {
        static CopyPool pool;
        return pool;
    }

    // Copies [src, src + bytes) to dst with the pool's threads plus the caller; returns when every byte has landed.
    void copy(void *dst, const void *src, size_t bytes)
    {
        if (bytes < 4 * PIECE || workers_.empty())
        {
            copy_piece(static_cast<char *>(dst), static_cast<const char *>(src), bytes);
            return;
        }
        Job job;
        job.dst = static_cast<char *>(dst), job.src = static_cast<const char *>(src), job.bytes = bytes;
        job.pieces = (bytes + PIECE - 1) / PIECE;
        job.remaining.store(job.pieces, std::memory_order_relaxed);
        {
            std::lock_guard<std::mutex> lock(mu_);
            jobs_.push_back(&job);
        }
        cv_.notify_all();
        work_on(job); // the caller copies pieces of its own job too
        // pieces handed to workers may still be in flight
        std::unique_lock<std::mutex> lock(mu_);
        done_cv_.wait(lock, [&] { return job.remaining.load(std::memory_order_acquire) == 0; });
        // the job lives on this stack frame: no pointer to it may stay behind in the queue
        for (auto it = jobs_.begin(); it != jobs_.end(); ++it)
            if (*it == &job)
            {
                jobs_.erase(it);
                break;
            }
    }

    size_t threads() const { return workers_.size() + 1; }

  private:
    CopyPool()
    {
        size_t n = 0;
        if (const char *e = getenv("NETCUDA_COPY_THREADS"))
            n = (size_t)std::max(atoi(e), 1);
        else
        {
            cpu_set_t set;
            CPU_ZERO(&set);
            size_t cores = 0;
            if (sched_getaffinity(0, sizeof(set), &set) == 0) cores = (size_t)CPU_COUNT(&set);
            if (cores == 0) cores = std::thread::hardware_concurrency();
            n = std::min<size_t>(cores ? cores : 1, 32);
        }
        for (size_t i = 1; i < n; i++) workers_.emplace_back([this] { loop(); });
    }
    ~CopyPool()
    {
        {
            std::lock_guard<std::mutex> lock(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &w : workers_) w.join();
    }

    // Copy pieces of `job` until none is left to hand out; the thread that finishes the last piece retires the job.
    void work_on(Job &job)
    {
        for (;;)
        {
            const size_t i = job.next.fetch_add(1, std::memory_order_relaxed);
            if (i >= job.pieces) return;
            const size_t off = i * PIECE, len = std::min(PIECE, job.bytes - off);
            copy_piece(job.dst + off, job.src + off, len);
            if (job.remaining.fetch_sub(1, std::memory_order_acq_rel) == 1)
            {
                std::lock_guard<std::mutex> lock(mu_); // (pairs with the waiter's predicate check: no lost wake-up)
                done_cv_.notify_all();
            }
        }
    }

    void loop()
    {
        for (;;)
        {
            Job *job = nullptr;
            {
                std::unique_lock<std::mutex> lock(mu_);
                cv_.wait(lock, [&] {
                    while (!jobs_.empty() && jobs_.front()->next.load(std::memory_order_relaxed) >= jobs_.front()->pieces) jobs_.pop_front();
                    return stop_ || !jobs_.empty();
                });
                if (stop_) return;
                // round-robin over the open jobs, so that several GPUs' staging copies advance together
                job = jobs_.front();
                if (jobs_.size() > 1)
                {
                    jobs_.pop_front();
                    jobs_.push_back(job);
                }
                // take one piece while holding the lock: the job cannot be retired (and its stack frame left) before this piece is done
                const size_t i = job->next.fetch_add(1, std::memory_order_relaxed);
                if (i >= job->pieces) continue;
                lock.unlock();
                const size_t off = i * PIECE, len = std::min(PIECE, job->bytes - off);
                copy_piece(job->dst + off, job->src + off, len);
                if (job->remaining.fetch_sub(1, std::memory_order_acq_rel) == 1)
                {
                    std::lock_guard<std::mutex> relock(mu_);
                    done_cv_.notify_all();
                }
            }
        }
    }

    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    std::deque<Job *> jobs_;
    std::vector<std::thread> workers_;
    bool stop_ = false;
};

} // namespace

void staging_copy(void *dst, const void *src, size_t bytes) { CopyPool::get().copy(dst, src, bytes); }
int staging_threads() { return (int)CopyPool::get().threads(); }

} // namespace nc
