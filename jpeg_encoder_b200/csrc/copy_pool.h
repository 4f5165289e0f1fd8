// Host-only helper of api.cu (no CUDA in here: tests/test_copy_pool.py builds it with g++ and hammers it).
#pragma once
#include <algorithm>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>

namespace jpgb {

// Three helper threads that copy pageable source pixels into the pinned staging buffers next to the calling thread
// (one core moves ~10 GB/s, a fifth of what the link takes). They belong to the context, are started with the first
// pageable upload and sleep on a condition variable in between: handing them a copy costs a wake-up, not a thread start.
class CopyPool {
public:
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> l(mu_);
            stop_ = true;
        }
        cv_work_.notify_all();
        for (auto &t : th_)
            if (t.joinable()) t.join();
    }
    // dst[0, n) = src[0, n) with `threads` (1..4) threads including the caller
    void copy(void *dst, const uint8_t *src, size_t n, unsigned threads) {
        if (threads <= 1 || n == 0) {
            std::memcpy(dst, src, n);
            return;
        }
        if (threads > 4) threads = 4;
        if (!started_) {
            for (unsigned i = 0; i < 3; ++i) th_[i] = std::thread([this, i] { worker(i); });
            started_ = true;
        }
        const size_t part = ((n / threads) + 63) & ~(size_t)63;
        {
            std::lock_guard<std::mutex> l(mu_);
            dst_ = static_cast<uint8_t *>(dst);
            src_ = src;
            n_ = n;
            part_ = part;
            parts_ = threads;
            pending_ = threads - 1;
            ++gen_;
        }
        cv_work_.notify_all();
        std::memcpy(dst, src, std::min(part, n));
        std::unique_lock<std::mutex> l(mu_);
        cv_done_.wait(l, [this] { return pending_ == 0; });
    }

private:
    void worker(unsigned idx) { // takes part idx + 1 of every copy that has that many parts
        uint64_t seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> l(mu_);
            cv_work_.wait(l, [&] { return stop_ || gen_ != seen; });
            if (stop_) return;
            seen = gen_;
            if (idx + 1 >= parts_) continue;
            const size_t lo = (size_t)(idx + 1) * part_, hi = idx + 2 == parts_ ? n_ : std::min(n_, lo + part_);
            uint8_t *d = dst_;
            const uint8_t *sp = src_;
            l.unlock();
            if (lo < hi) std::memcpy(d + lo, sp + lo, hi - lo);
            l.lock();
            if (--pending_ == 0) cv_done_.notify_one();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_work_, cv_done_;
    std::thread th_[3];
    bool started_ = false, stop_ = false;
    uint64_t gen_ = 0;
    unsigned parts_ = 0, pending_ = 0;
    uint8_t *dst_ = nullptr;
    const uint8_t *src_ = nullptr;
    size_t n_ = 0, part_ = 0;
};


} // namespace jpgb
