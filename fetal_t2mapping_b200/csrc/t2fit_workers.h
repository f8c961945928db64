// Host worker pool of the staged host-memory path (gather of the masked rows into pinned staging, unpack / scatter of
// results): plain C++, no CUDA.  fn(part, n_parts) runs on every worker, the caller included.
//
// The staged pipeline hands the pool a task every ~100 us (one per 2.6 MB chunk).  Waking 15 sleeping threads through a
// condition variable costs 30-200 us on the VMs these boxes are, per chunk, and varies from call to call (measured: 1.3-3.0
// ms per c2 volume).  Workers therefore spin on the generation counter for a short while after a task (kSpinUs) before
// they go to sleep: inside one call they are always hot, between calls they cost nothing.
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define T2FIT_CPU_RELAX() _mm_pause()
#else
#define T2FIT_CPU_RELAX() std::this_thread::yield()
#endif

namespace t2fit {

class Workers {
  public:
    explicit Workers(int n) : n_(n < 1 ? 1 : n) {
        for (int t = 1; t < n_; ++t) threads_.emplace_back([this, t] { loop(t); });
    }
    ~Workers() {
        stop_.store(true);
        gen_.fetch_add(1);
        { std::lock_guard<std::mutex> g(m_); }
        cv_.notify_all();
        for (auto& t : threads_) t.join();
    }
    int size() const { return n_; }
    // fn(part, n_parts) on every worker, caller included; returns when all are done.  One caller at a time.
    void run(const std::function<void(int, int)>& fn) {
        fn_ = &fn;
        pending_.store(n_ - 1);
        gen_.fetch_add(1);                       // publishes fn_ and pending_ (sequentially consistent)
        if (sleepers_.load() > 0) {              // a worker that is about to sleep re-checks gen_ under m_
            { std::lock_guard<std::mutex> g(m_); }
            cv_.notify_all();
        }
        fn(0, n_);
        int spins = 0;
        while (pending_.load() != 0) {
            if (++spins < 4096) T2FIT_CPU_RELAX();
            else std::this_thread::yield();
        }
        fn_ = nullptr;
    }

  private:
    static constexpr int kSpinUs = 300;          // how long an idle worker polls before it sleeps
    void loop(int t) {
        uint64_t seen = 0;
        for (;;) {
            uint64_t g = gen_.load();
            if (g == seen) {
                const auto t0 = std::chrono::steady_clock::now();
                int spins = 0;
                while ((g = gen_.load()) == seen) {
                    T2FIT_CPU_RELAX();
                    if ((++spins & 255) == 0 &&
                        std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count() > kSpinUs) {
                        std::unique_lock<std::mutex> l(m_);
                        sleepers_.fetch_add(1);
                        cv_.wait(l, [&] { return gen_.load() != seen; });
                        sleepers_.fetch_sub(1);
                    }
                }
            }
            seen = g;
            if (stop_.load()) return;
            const std::function<void(int, int)>* fn = fn_;
            if (fn) (*fn)(t, n_);
            pending_.fetch_sub(1);
        }
    }
    int n_;
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_;
    const std::function<void(int, int)>* fn_ = nullptr;
    std::atomic<uint64_t> gen_{0};
    std::atomic<int> pending_{0}, sleepers_{0};
    std::atomic<bool> stop_{false};
};

}  // namespace t2fit
