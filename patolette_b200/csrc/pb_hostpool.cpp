#include "pb_hostpool.h"

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <stdexcept>
#include <thread>
#include <vector>

namespace {

constexpr int MAX_WORKERS = 15;

struct Pool {
    std::mutex mu;
    std::condition_variable wake, done;
    std::vector<std::thread> workers;
    const std::function<void(int)> *job = nullptr;
    unsigned long generation = 0;
    int active = 0;   // workers that take part in the current job (tid 1..active)
    int pending = 0;  // ... and have not finished it yet
    bool quit = false;
    bool failed = false; // a worker's share of the current job threw
    std::mutex run_mu; // one job at a time

    void worker(int tid) {
        unsigned long seen = 0;
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            wake.wait(lk, [&] { return quit || generation != seen; });
            if (quit) return;
            seen = generation;
            if (tid > active) continue;
            const std::function<void(int)> *fn = job;
            lk.unlock();
            bool threw = false;
            try { (*fn)(tid); } catch (...) { threw = true; } // an exception must not leave a worker thread (std::terminate)
            lk.lock();
            if (threw) failed = true;
            if (--pending == 0) done.notify_one();
        }
    }
    void ensure(int n) {
        while ((int)workers.size() < n) {
            const int tid = (int)workers.size() + 1;
            workers.emplace_back([this, tid] { worker(tid); });
        }
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> lk(mu);
            quit = true;
        }
        wake.notify_all();
        for (auto &t : workers) t.join();
    }
};

Pool &pool() {
    static Pool p;
    return p;
}

} // namespace

int pb_hostpool_default_threads() {
    const unsigned hc = std::thread::hardware_concurrency();
    return (int)std::max(1u, std::min(8u, hc ? hc : 1u));
}

void pb_hostpool_run(int threads, const std::function<void(int)> &fn) {
    threads = std::max(1, std::min(threads, MAX_WORKERS + 1));
    if (threads == 1) {
        fn(0);
        return;
    }
    Pool &p = pool();
    std::lock_guard<std::mutex> one(p.run_mu);
    {
        std::lock_guard<std::mutex> lk(p.mu);
        p.ensure(threads - 1);
        p.job = &fn;
        p.active = threads - 1;
        p.pending = threads - 1;
        p.failed = false;
        p.generation++;
    }
    p.wake.notify_all();
    bool threw = false;
    try { fn(0); } catch (...) { threw = true; } // the workers still have to finish before `fn` goes out of scope
    std::unique_lock<std::mutex> lk(p.mu);
    p.done.wait(lk, [&] { return p.pending == 0; });
    p.job = nullptr;
    if (threw || p.failed) throw std::runtime_error("patolette_b200: a host pool job failed");
}
