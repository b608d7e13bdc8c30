#include "pb_pool.h"

#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <unordered_map>

#include "pb_error.h"

namespace {
struct Block {
    size_t bytes;
    int device;
};
std::mutex g_mu;
std::multimap<std::pair<int, size_t>, void *> g_free; // (device, bytes) -> block
std::unordered_map<void *, Block> g_live;
size_t g_cached = 0;

void flush_locked() {
    for (auto &kv : g_free) cudaFree(kv.second);
    g_free.clear();
    g_cached = 0;
}
} // namespace

void *pb_pool_alloc(size_t bytes) {
    if (bytes == 0) bytes = 1;
    bytes = (bytes + 511) & ~(size_t)511;
    int dev = 0;
    PB_CUDA_OK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_free.lower_bound({dev, bytes});
    if (it != g_free.end() && it->first.first == dev && it->first.second <= bytes + bytes / 8 + 4096) {
        void *p = it->second;
        g_live[p] = Block{it->first.second, dev};
        g_cached -= it->first.second;
        g_free.erase(it);
        return p;
    }
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) { // out of memory: give the cache back and retry once
        cudaGetLastError();
        flush_locked();
        e = cudaMalloc(&p, bytes);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        throw pb_cuda_error(e);
    }
    g_live[p] = Block{bytes, dev};
    return p;
}

void pb_pool_free(void *p) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_live.find(p);
    if (it == g_live.end()) {
        cudaFree(p);
        return;
    }
    g_free.insert({{it->second.device, it->second.bytes}, p});
    g_cached += it->second.bytes;
    g_live.erase(it);
}

void pb_pool_release_all() {
    std::lock_guard<std::mutex> lk(g_mu);
    flush_locked();
}

size_t pb_pool_cached_bytes() {
    std::lock_guard<std::mutex> lk(g_mu);
    return g_cached;
}
