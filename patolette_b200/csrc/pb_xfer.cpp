#include "pb_xfer.h"

#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "pb_error.h"
#include "pb_nccl.h"

namespace {

constexpr size_t CHUNK = 8u << 20;      // bounce buffer size
constexpr size_t DIRECT_BELOW = 4u << 20; // small copies: the driver's own path is fine
constexpr int SLOTS = 2;

struct Lane { // per worker thread: pinned bounce buffers + a stream, created once and kept
    void *slot[SLOTS] = {nullptr, nullptr};
    cudaEvent_t ev[SLOTS] = {nullptr, nullptr};
    cudaStream_t st = nullptr;
    bool ok = false;
    void init() {
        if (ok) return;
        for (int i = 0; i < SLOTS; i++) {
            if (cudaHostAlloc(&slot[i], CHUNK, cudaHostAllocDefault) != cudaSuccess) return;
            if (cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess) return;
        }
        if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) return;
        ok = true;
    }
};

constexpr int MAX_LANES = 8;
Lane g_lanes[16][MAX_LANES]; // [device][worker]

bool is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

int worker_count() {
    unsigned hc = std::thread::hardware_concurrency();
    hc /= (unsigned)std::max(1, pb_nccl_world()); // ranks of an image-sharded job share the host's cores
    int n = hc >= 16 ? 8 : (hc >= 8 ? 4 : 2);
    return std::min(n, MAX_LANES);
}

template <bool H2D>
void striped(char *dev, char *host, size_t bytes) {
    int dev_id = 0;
    PB_CUDA_OK(cudaGetDevice(&dev_id));
    const int T = worker_count();
    // stripes are multiples of CHUNK so every worker streams through whole bounce buffers
    const size_t nchunks = (bytes + CHUNK - 1) / CHUNK;
    std::vector<std::thread> th;
    std::vector<cudaError_t> err(T, cudaSuccess);
    for (int t = 0; t < T; t++) {
        const size_t c0 = nchunks * t / T, c1 = nchunks * (t + 1) / T;
        if (c0 == c1) continue;
        th.emplace_back([=, &err]() {
            cudaSetDevice(dev_id);
            Lane &L = g_lanes[dev_id & 15][t];
            L.init();
            if (!L.ok) { err[t] = cudaErrorMemoryAllocation; return; }
            for (size_t c = c0; c < c1; c++) {
                const int s = (int)((c - c0) % SLOTS);
                const size_t off = c * CHUNK, len = std::min(CHUNK, bytes - off);
                if (H2D) {
                    cudaEventSynchronize(L.ev[s]); // the DMA that last read this slot is done
                    memcpy(L.slot[s], host + off, len);
                    cudaMemcpyAsync(dev + off, L.slot[s], len, cudaMemcpyHostToDevice, L.st);
                    cudaEventRecord(L.ev[s], L.st);
                } else {
                    // keep one DMA in flight while the previous slot is drained into the caller's buffer
                    cudaMemcpyAsync(L.slot[s], dev + off, len, cudaMemcpyDeviceToHost, L.st);
                    cudaEventRecord(L.ev[s], L.st);
                    if (c > c0) {
                        const int ps = (int)((c - 1 - c0) % SLOTS);
                        const size_t poff = (c - 1) * CHUNK;
                        cudaEventSynchronize(L.ev[ps]);
                        memcpy(host + poff, L.slot[ps], std::min(CHUNK, bytes - poff));
                    }
                }
            }
            if (!H2D) {
                const int ps = (int)((c1 - 1 - c0) % SLOTS);
                const size_t poff = (c1 - 1) * CHUNK;
                cudaEventSynchronize(L.ev[ps]);
                memcpy(host + poff, L.slot[ps], std::min(CHUNK, bytes - poff));
            }
            err[t] = cudaStreamSynchronize(L.st);
        });
    }
    for (auto &x : th) x.join();
    for (cudaError_t e : err)
        if (e != cudaSuccess) throw pb_cuda_error(e);
}

} // namespace

bool pb_host_is_pinned(const void *p) { return is_pinned(p); }

void pb_copy_h2d(void *dst_dev, const void *src_host, size_t bytes, cudaStream_t st) {
    if (!bytes) return;
    if (bytes < DIRECT_BELOW || is_pinned(src_host)) {
        PB_CUDA_OK(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, st));
        PB_CUDA_OK(cudaStreamSynchronize(st));
        return;
    }
    PB_CUDA_OK(cudaStreamSynchronize(st)); // the destination may still be in use by earlier work
    striped<true>((char *)dst_dev, (char *)src_host, bytes);
}

void pb_copy_d2h(void *dst_host, const void *src_dev, size_t bytes, cudaStream_t st) {
    if (!bytes) return;
    PB_CUDA_OK(cudaStreamSynchronize(st)); // producers of src_dev run on st
    if (bytes < DIRECT_BELOW || is_pinned(dst_host)) {
        PB_CUDA_OK(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, st));
        PB_CUDA_OK(cudaStreamSynchronize(st));
        return;
    }
    striped<false>((char *)src_dev, (char *)dst_host, bytes);
}
