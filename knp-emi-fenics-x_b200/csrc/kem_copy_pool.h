// kem_copy_pool.h -- persistent worker threads for host-side staging copies.
#pragma once
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

// Pageable caller buffers (what `u.x.array` of a dolfinx Function is) have to pass through
// pinned staging memory; one core moves ~10 GB/s, a PCIe 5 x16 link ~50 GB/s, so the
// staging copies are spread over a few persistent worker threads.
class CopyPool {
public:
    static CopyPool &get()
    {
        // leaked on purpose: destroying a condition variable that detached workers still
        // wait on blocks process exit (pthread_cond_destroy waits for its waiters)
        static CopyPool *pool = new CopyPool;
        return *pool;
    }

    void copy(void *dst, const void *src, size_t bytes) { run_parts(dst, src, 0.0, bytes); }

    // dst[0:n] = value, spread over the pool (the host-side stand-in for the device->host copy
    // of an output column whose value is known without asking the device)
    void fill(double *dst, double value, size_t n) { run_parts(dst, nullptr, value, n * sizeof(double)); }

    int threads() const { return (int)workers_.size() + 1; }

private:
    static void do_part(char *dst, const char *src, double value, size_t bytes)
    {
        if (src) {
            memcpy(dst, src, bytes);
        } else {
            // Non-temporal stores: the array is hundreds of MB and is read next by somebody
            // else (the PDE side), so pulling its lines into the cache first (write-allocate)
            // would double the memory traffic of a fill that already competes with the DMA
            // engines for the host's memory system (eight ranks on one host: 4-5 ms of 39).
            double *d = (double *)dst;
            size_t n = bytes / sizeof(double), i = 0;
#if defined(__SSE2__)
            while (i < n && ((uintptr_t)(d + i) & 15)) d[i++] = value;
            const __m128d v = _mm_set1_pd(value);
            for (; i + 8 <= n; i += 8) {
                _mm_stream_pd(d + i, v);
                _mm_stream_pd(d + i + 2, v);
                _mm_stream_pd(d + i + 4, v);
                _mm_stream_pd(d + i + 6, v);
            }
            _mm_sfence();
#endif
            for (; i < n; ++i) d[i] = value;
        }
    }

    void run_parts(void *dst, const void *src, double value, size_t bytes)
    {
        const size_t min_part = 512u << 10;
        int parts = (int)std::min<size_t>(workers_.size() + 1, (bytes + min_part - 1) / min_part);
        if (parts <= 1) {
            do_part((char *)dst, (const char *)src, value, bytes);
            return;
        }
        std::unique_lock<std::mutex> call_lock(call_mu_);      // one parallel operation at a time
        const size_t per = ((bytes + parts - 1) / parts + 63) / 64 * 64;
        {
            std::lock_guard<std::mutex> lk(mu_);
            dst_ = (char *)dst;
            src_ = (const char *)src;
            value_ = value;
            bytes_ = bytes;
            per_ = per;
            next_ = 1;
            parts_ = parts;
            pending_ = parts - 1;
            ++generation_;
        }
        cv_.notify_all();
        do_part((char *)dst, (const char *)src, value, std::min(per, bytes));   // part 0 on the caller
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [&] { return pending_ == 0; });
    }

    CopyPool()
    {
        // every CPU this process may run on (a rank bound next to its GPU sees only those); the
        // calling thread copies too.  One core moves 4-10 GB/s, the link 50.
        unsigned hw = std::thread::hardware_concurrency();
#if defined(__linux__)
        cpu_set_t set;
        if (sched_getaffinity(0, sizeof set, &set) == 0) hw = (unsigned)CPU_COUNT(&set);
#endif
        // several ranks on one host (torchrun sets LOCAL_WORLD_SIZE) share those CPUs
        if (const char *e = getenv("LOCAL_WORLD_SIZE")) {
            const unsigned ranks = (unsigned)std::max(1, atoi(e));
            hw = std::max(1u, hw / ranks);
        }
        int n = (int)std::min<unsigned>(15u, hw > 1 ? hw - 1 : 0);
        if (const char *e = getenv("KNPEMI_COPY_THREADS")) n = std::max(0, atoi(e) - 1);
        for (int i = 0; i < n; ++i) workers_.emplace_back([this] { run(); });
        for (auto &t : workers_) t.detach();
    }

    void run()
    {
        unsigned long seen = 0;
        for (;;) {
            int part;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return generation_ != seen && next_ < parts_; });
                part = next_++;
                if (next_ >= parts_) seen = generation_;
            }
            const size_t off = (size_t)part * per_;
            if (off < bytes_) do_part(dst_ + off, src_ ? src_ + off : nullptr, value_, std::min(per_, bytes_ - off));
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (--pending_ == 0) done_cv_.notify_one();
            }
        }
    }

    std::vector<std::thread> workers_;
    std::mutex mu_, call_mu_;
    std::condition_variable cv_, done_cv_;
    char *dst_ = nullptr;
    const char *src_ = nullptr;
    double value_ = 0.0;
    size_t bytes_ = 0, per_ = 0;
    int next_ = 0, parts_ = 0, pending_ = 0;
    unsigned long generation_ = 0;
};

