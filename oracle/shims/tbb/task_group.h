// TEST INFRASTRUCTURE — stand-in for the classic-TBB scheduling calls that the reference's
// Render.cpp makes (reference src/Render.cpp:335-358). TBB is not installed in this image and
// carries no arithmetic: this header only decides WHICH thread runs WHICH tile.  Tiles are
// independent and own their PRNG (reference include/cornelis/Tiles.hpp:23-31), so images do not
// depend on the schedule.
//
// Provided: tbb::task_group::run_and_wait, tbb::parallel_for_each, tbb::task::self()
// .cancel_group_execution(), tbb::task_group_status / tbb::canceled.
#pragma once

#include <atomic>
#include <cstdlib>
#include <iterator>
#include <thread>
#include <vector>

namespace tbb {

enum task_group_status { not_complete, complete, canceled };

namespace shim {
// Number of worker threads parallel_for_each uses.  0 = hardware concurrency.  The oracle harness
// sets this before every render; CORNELIS_REF_THREADS overrides the default.
inline std::atomic<int> &requestedThreads() {
    static std::atomic<int> n{0};
    return n;
}
inline std::atomic<bool> &cancelFlag() {
    static std::atomic<bool> f{false};
    return f;
}
inline int workerCount() {
    int n = requestedThreads().load();
    if (n <= 0) {
        if (char const *env = std::getenv("CORNELIS_REF_THREADS"))
            n = std::atoi(env);
    }
    if (n <= 0)
        n = static_cast<int>(std::thread::hardware_concurrency());
    return n > 0 ? n : 1;
}
} // namespace shim

struct task {
    static task &self() {
        static task t;
        return t;
    }
    void cancel_group_execution() { shim::cancelFlag().store(true); }
};

class task_group {
  public:
    template <typename F>
    task_group_status run_and_wait(F &&body) {
        shim::cancelFlag().store(false);
        body();
        return shim::cancelFlag().load() ? canceled : complete;
    }
};

// Dynamic scheduling: every worker pulls the next unclaimed element from one atomic cursor, the
// same work-distribution behaviour tbb::parallel_for_each gives for independent items.
template <typename It, typename F>
void parallel_for_each(It first, It last, F body) {
    auto const count = static_cast<std::size_t>(std::distance(first, last));
    std::atomic<std::size_t> cursor{0};
    auto worker = [&] {
        for (;;) {
            if (shim::cancelFlag().load())
                return;
            std::size_t k = cursor.fetch_add(1);
            if (k >= count)
                return;
            body(*std::next(first, static_cast<std::ptrdiff_t>(k)));
        }
    };
    int const n = shim::workerCount();
    std::vector<std::thread> pool;
    for (int i = 1; i < n; i++)
        pool.emplace_back(worker);
    worker();
    for (auto &t : pool)
        t.join();
}

} // namespace tbb
