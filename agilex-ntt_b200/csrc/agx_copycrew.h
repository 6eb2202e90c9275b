// agx_copycrew.h -- host-side helper of the host-pointer pipelines (agx_*_host, agx_ref_*): parallel staging copies for
// pageable callers.  Plain C++17, no CUDA (tests/test_abi.py builds and stresses it on the CPU).
#pragma once
#include <algorithm>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace agx {

// Pageable callers pay a host memcpy per chunk and direction (caller's pages <-> pinned staging), and one thread copies at
// 8-10 GB/s -- a fifth of what the PCIe link moves.  CopyCrew splits a chunk's copy into slices run by a few persistent
// helper threads beside the submitting thread.  One crew serves ONE submitting thread at a time (the pipelines keep one
// for staging-in on the calling thread and one for the copy-out helper).  AGX_HOST_COPY_THREADS = threads per direction
// (default: a quarter of the hardware threads, 1..4; 1 = no helpers).
class CopyCrew {
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable go_, done_;
    const std::function<void(size_t)> *job_ = nullptr;
    size_t parts_ = 0, next_ = 0, left_ = 0;
    uint64_t gen_ = 0;
    bool stop_ = false;
    void work() {
        uint64_t seen = 0;
        std::unique_lock<std::mutex> lk(m_);
        for (;;) {
            go_.wait(lk, [&] { return stop_ || gen_ != seen; });
            if (stop_) return;
            seen = gen_;
            while (job_ && next_ < parts_) {
                const size_t i = next_++;
                const std::function<void(size_t)> *f = job_;
                lk.unlock();
                (*f)(i);
                lk.lock();
                if (--left_ == 0) done_.notify_all();
            }
        }
    }
public:
    static int default_threads() {
        if (const char *e = getenv("AGX_HOST_COPY_THREADS")) {
            const long v = atol(e);
            if (v >= 1 && v <= 16) return (int)v;
        }
        const unsigned hw = std::thread::hardware_concurrency();
        return (int)std::min(4u, std::max(1u, hw / 4));
    }
    explicit CopyCrew(int threads = default_threads()) {
        for (int i = 1; i < threads; i++) th_.emplace_back(&CopyCrew::work, this);
    }
    ~CopyCrew() {
        { std::lock_guard<std::mutex> lk(m_); stop_ = true; }
        go_.notify_all();
        for (auto &t : th_) t.join();
    }
    int threads() const { return (int)th_.size() + 1; }
    // fn(0) .. fn(parts-1), each exactly once, on the helpers and the calling thread; returns when all are done
    void run(size_t parts, const std::function<void(size_t)> &fn) {
        if (th_.empty() || parts <= 1) { for (size_t i = 0; i < parts; i++) fn(i); return; }
        std::unique_lock<std::mutex> lk(m_);
        job_ = &fn; parts_ = parts; next_ = 0; left_ = parts; gen_++;
        go_.notify_all();
        while (next_ < parts_) {
            const size_t i = next_++;
            lk.unlock();
            fn(i);
            lk.lock();
            --left_;
        }
        done_.wait(lk, [&] { return left_ == 0; });
        job_ = nullptr;
    }
    void copy(void *dst, const void *src, size_t bytes) {
        constexpr size_t kMinSlice = 1u << 20;                       // below this a slice is not worth a wake-up
        size_t parts = std::min<size_t>((size_t)threads(), bytes / kMinSlice);
        if (parts <= 1) { memcpy(dst, src, bytes); return; }
        const size_t slice = ((bytes + parts - 1) / parts + 4095) & ~(size_t)4095;
        parts = (bytes + slice - 1) / slice;
        run(parts, [&](size_t i) {
            const size_t o = i * slice, n = std::min(slice, bytes - o);
            memcpy(static_cast<char *>(dst) + o, static_cast<const char *>(src) + o, n);
        });
    }
};


}  // namespace agx
