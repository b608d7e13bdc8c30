// pb_kmeans.cu - weighted KMeans palette refinement.
//
// Reference: lib/src/palette/refine.c:56-221 driving the vendored, patolette-modified
// faiss 1.10.0 (lib/faiss/faiss/Clustering.cpp:267-603, IndexFlatL2 k=1 search through
// utils/distances.cpp:259-343 on the generic build).  Restated, not linked:
//   * f32 samples (n x 3) and centres, f32 weights (refine.c:102-163);
//   * optional subsample of k * max_points_per_centroid points: the first entries of a
//     forward Fisher-Yates permutation driven by std::mt19937(1234)
//     (Clustering.cpp:70-120, utils/random.cpp:184-194).  Only a prefix of the
//     permutation is ever read, so the host computes just that prefix with a sparse
//     swap map - O(subsample), not O(N);
//   * assignment: dis = (|x|^2 + |y|^2) - 2*ip clamped at 0, ip as sgemm_ evaluates a
//     k=3 dot, strict '<' over ascending centroid index - one thread per sample, the
//     centroids in shared memory;
//   * centroid update: per centroid a SEQUENTIAL f32 sum in sample order
//     (Clustering.cpp:178-187): samples are stably sorted by assignment
//     (pb_parallel.cu) and one warp per centroid walks its run;
//   * empty clusters are refilled on the host with faiss' own RNG protocol
//     (Clustering.cpp:216-263).
#include <math.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "pb_common.cuh"
#include "pb_kernels.h"
#include "pb_prof.h"
#include "pb_pipeline.h"
#include "pb_pool.h"

namespace {

// std::mt19937 (faiss::RandomGenerator, utils/random.cpp:35-55)
struct Mt19937 {
    uint32_t s[624];
    int i;
    explicit Mt19937(uint32_t seed) {
        s[0] = seed;
        for (int k = 1; k < 624; k++) s[k] = 1812433253u * (s[k - 1] ^ (s[k - 1] >> 30)) + (uint32_t)k;
        i = 624;
    }
    uint32_t next() {
        if (i >= 624) {
            for (int k = 0; k < 624; k++) {
                uint32_t y = (s[k] & 0x80000000u) | (s[(k + 1) % 624] & 0x7fffffffu);
                s[k] = s[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            i = 0;
        }
        uint32_t y = s[i++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
};

// First `take` entries of faiss::rand_perm(n, seed) (random.cpp:184-194).  Forward Fisher-Yates
// finalises perm[i] at step i and later steps only touch positions > i, so only `take` steps are
// simulated; positions whose value is no longer the identity live in an open-addressing table.
// The result depends on (n, take, seed) only and is cached across calls.
std::vector<uint32_t> rand_perm_prefix(size_t n, size_t take, uint32_t seed) {
    static std::mutex mu;
    static size_t c_n = 0, c_take = 0;
    static uint32_t c_seed = 0;
    static std::vector<uint32_t> c_out;
    std::lock_guard<std::mutex> lk(mu);
    if (c_n == n && c_take == take && c_seed == seed && !c_out.empty()) return c_out;
    std::vector<uint32_t> out(take);
    size_t cap = 1;
    while (cap < take * 4 + 16) cap <<= 1;
    std::vector<uint64_t> tab(cap, ~0ULL); // (position << 32) | value
    auto slot = [&](uint32_t k) { return (size_t)(k * 0x9E3779B1u) & (cap - 1); };
    auto get = [&](uint32_t pos) {
        for (size_t j = slot(pos);; j = (j + 1) & (cap - 1)) {
            const uint64_t e = tab[j];
            if (e == ~0ULL) return pos;
            if ((uint32_t)(e >> 32) == pos) return (uint32_t)e;
        }
    };
    auto put = [&](uint32_t pos, uint32_t v) {
        for (size_t j = slot(pos);; j = (j + 1) & (cap - 1)) {
            const uint64_t e = tab[j];
            if (e == ~0ULL || (uint32_t)(e >> 32) == pos) { tab[j] = ((uint64_t)pos << 32) | v; return; }
        }
    };
    Mt19937 rng(seed);
    for (size_t i = 0; i < take && i + 1 < n; i++) {
        const uint32_t i2 = (uint32_t)(i + (size_t)((uint64_t)rng.next() % (uint64_t)(int)(n - i)));
        const uint32_t vi = get((uint32_t)i), v2 = get(i2);
        out[i] = v2; // perm[i] is final after step i
        put(i2, vi);
    }
    if (take == n && n > 0) out[n - 1] = get((uint32_t)(n - 1));
    c_n = n; c_take = take; c_seed = seed; c_out = out;
    return out;
}

__global__ void k_to_f32(const double *__restrict__ c0, const double *__restrict__ c1,
                         const double *__restrict__ c2, const double *__restrict__ w,
                         const uint32_t *__restrict__ pick, size_t n, float *__restrict__ x0,
                         float *__restrict__ x1, float *__restrict__ x2, float *__restrict__ wf,
                         int *__restrict__ nonfinite) {
    int bad = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t p = pick ? pick[i] : i;
        const float a = (float)c0[p], b = (float)c1[p], c = (float)c2[p]; // refine.c:136-142
        x0[i] = a; x1[i] = b; x2[i] = c;
        if (wf) wf[i] = (float)w[p]; // refine.c:158
        bad |= !(isfinite(a) && isfinite(b) && isfinite(c));
    }
    if (bad) atomicOr(nonfinite, 1);
}

// Clustering.cpp:295-304 scans ALL input samples for NaN/Inf before subsampling.
__global__ void k_scan_finite(const double *__restrict__ c0, const double *__restrict__ c1,
                              const double *__restrict__ c2, size_t n, int *__restrict__ nonfinite) {
    int bad = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        bad |= !(isfinite((float)c0[i]) && isfinite((float)c1[i]) && isfinite((float)c2[i]));
    if (bad) atomicOr(nonfinite, 1);
}

// (y0, y1, y2, |y|^2) per centroid; fvec_norms_L2sqr: ((y0^2 + y1^2) + y2^2), separately rounded
__global__ void k_cen4(const float *__restrict__ cen, int K, float4 *__restrict__ cen4, const int *__restrict__ stop) {
    if (*stop) return;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= K) return;
    const float a = cen[3 * j], b = cen[3 * j + 1], c = cen[3 * j + 2];
    cen4[j] = make_float4(a, b, c, __fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c)));
}

// IN_SMEM: the centroid table sits in shared memory (16 B x K); palettes too large for that are read from
// global memory - every lane asks for the same entry, so the loads are L1-served broadcasts.
template <bool IN_SMEM>
__global__ void __launch_bounds__(256) k_assign(const float *__restrict__ x0, const float *__restrict__ x1,
                                                const float *__restrict__ x2, size_t nx,
                                                const float4 *__restrict__ cen4, int K, bool seq_path,
                                                uint16_t *__restrict__ assign, const int *__restrict__ stop) {
    extern __shared__ float4 s_cen[];
    if (*stop) return; // an earlier iteration produced an empty cluster: the host takes over from there
    if (IN_SMEM) {
        for (int j = threadIdx.x; j < K; j += blockDim.x) s_cen[j] = cen4[j];
        __syncthreads();
    }
    const float4 *tab = IN_SMEM ? s_cen : cen4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nx; i += (size_t)gridDim.x * blockDim.x) {
        const float a = x0[i], b = x1[i], c = x2[i];
        const float xn = __fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c));
        float bd = 0.f;
        int best = 0;
        for (int j = 0; j < K; j++) {
            const float4 y = tab[j];
            float dd;
            if (seq_path) { // fewer than 20 queries: exhaustive_L2sqr_seq -> fvec_L2sqr
                const float d0 = __fsub_rn(a, y.x), d1 = __fsub_rn(b, y.y), d2 = __fsub_rn(c, y.z);
                dd = __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2));
            } else { // distances.cpp:259-343: dis = x_norms + y_norms - 2 * ip, ip from sgemm_ (k = 3)
                const float ip = __fmaf_rn(c, y.z, __fmaf_rn(b, y.y, __fmul_rn(a, y.x)));
                dd = __fsub_rn(__fadd_rn(xn, y.w), __fmul_rn(2.f, ip));
                if (dd < 0) dd = 0;
            }
            if (j == 0 || dd < bd) { bd = dd; best = j; }
        }
        assign[i] = (uint16_t)best;
    }
}

constexpr int KM_TILE = 128;
constexpr int KM_PER = KM_TILE / 32;
constexpr int KM_WARPS = 4;

// (x0, x1, x2, w) of every sample interleaved: the centroid chains GATHER the members of a cluster, and one
// 16-byte sector access per member replaces four 4-byte ones (same idea as the f64 bucket sums, pb_chain.cu)
__global__ void k_pack_samples(const float *__restrict__ x0, const float *__restrict__ x1, const float *__restrict__ x2,
                               const float *__restrict__ wf, size_t nx, float4 *__restrict__ aos) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nx; i += (size_t)gridDim.x * blockDim.x)
        aos[i] = make_float4(x0[i], x1[i], x2[i], wf ? wf[i] : 1.0f);
}

// One warp per centroid: lanes 0..3 chain {hassign, c0, c1, c2} in sample order (Clustering.cpp:178-187).
// Every term is formed by the gathering lanes; the gather of tile t+1 sits in registers during the chain
// over tile t.
template <bool WEIGHTED>
__global__ void __launch_bounds__(KM_WARPS * 32) k_centroid_chains(const float4 *__restrict__ aos,
                                                                   const uint32_t *__restrict__ ord,
                                                                   const uint32_t *__restrict__ class_start, int K,
                                                                   float *__restrict__ out /* K x 4 */,
                                                                   const int *__restrict__ stop) {
    __shared__ float sm_all[KM_WARPS][4][KM_TILE + 1];
    if (*stop) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * KM_WARPS + warp;
    if (c >= K) return;
    float(*sm)[KM_TILE + 1] = sm_all[warp];
    const uint32_t beg = class_start[c], end = class_start[c + 1];
    float acc = 0.f;
    float4 g[KM_PER];
    auto gather = [&](uint32_t i0) {
#pragma unroll
        for (int q = 0; q < KM_PER; q++) {
            const uint32_t i = i0 + q * 32 + lane;
            g[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < end) g[q] = aos[ord[i]];
        }
    };
    if (beg < end) gather(beg);
    for (uint32_t i0 = beg; i0 < end; i0 += KM_TILE) {
        const uint32_t cnt = min((uint32_t)KM_TILE, end - i0);
#pragma unroll
        for (int q = 0; q < KM_PER; q++) {
            const int e = q * 32 + lane;
            const float w = g[q].w;
            // hassign += w ; c[j] += x[j] * w   (unweighted: += 1.0 ; += x[j])
            sm[0][e] = WEIGHTED ? w : 1.0f;
            sm[1][e] = WEIGHTED ? __fmul_rn(g[q].x, w) : g[q].x;
            sm[2][e] = WEIGHTED ? __fmul_rn(g[q].y, w) : g[q].y;
            sm[3][e] = WEIGHTED ? __fmul_rn(g[q].z, w) : g[q].z;
        }
        __syncwarp();
        if (i0 + KM_TILE < end) gather(i0 + KM_TILE);
        if (lane < 4) {
            const float *vp = sm[lane];
#pragma unroll 16
            for (uint32_t e = 0; e < cnt; e++) acc = __fadd_rn(acc, vp[e]);
        }
        __syncwarp();
    }
    if (lane < 4) out[4 * c + lane] = acc;
}

// compute_centroids epilogue (Clustering.cpp:194-203) on the device: centroid = sum * (1 / hassign).
// An empty cluster needs faiss' RNG-driven split_clusters (Clustering.cpp:216-263), which stays on the
// host: the kernel then records the iteration in *stop and leaves the sums untouched for the host.
__global__ void k_finalize_centroids(const float *__restrict__ sums, int K, int iteration, float *__restrict__ cen,
                                     int *__restrict__ stop) {
    __shared__ int s_empty;
    if (*stop) return;
    if (threadIdx.x == 0) s_empty = 0;
    __syncthreads();
    for (int c = threadIdx.x; c < K; c += blockDim.x)
        if (sums[4 * c] == 0.f) s_empty = 1;
    __syncthreads();
    if (s_empty) {
        if (threadIdx.x == 0) *stop = iteration + 1;
        return;
    }
    for (int c = threadIdx.x; c < K; c += blockDim.x) {
        const float norm = __fdiv_rn(1.f, sums[4 * c]);
        for (int j = 0; j < 3; j++) cen[3 * c + j] = __fmul_rn(sums[4 * c + 1 + j], norm);
    }
}

struct DevMem {
    std::vector<void *> ptrs;
    template <typename T>
    T *alloc(size_t count) {
        void *p = pb_pool_alloc((count ? count : 1) * sizeof(T));
        ptrs.push_back(p);
        return (T *)p;
    }
    ~DevMem() {
        for (void *p : ptrs) pb_pool_free(p);
    }
};

} // namespace

// planes: device f64 colours in the quantisation space (original pixel order); d_w: device f64
// weights or nullptr; pal_rm: K x 3 row-major f64 centres, refined in place.
void pb_kmeans_device(const float *d_x0, const float *d_x1, const float *d_x2, const float *d_wf, size_t nx,
                      std::vector<float> &cen, int K, int niter, int sm_count, cudaStream_t st, long *launches) {
    DevMem mem;
    uint16_t *d_assign = mem.alloc<uint16_t>(nx);
    uint32_t *d_ord = mem.alloc<uint32_t>(nx);
    const size_t tiles = pb_scatter_tiles((uint32_t)nx);
    uint32_t *d_tile_hist = mem.alloc<uint32_t>(pb_scatter_table_words(tiles, 1, K));
    uint32_t *d_cstart = mem.alloc<uint32_t>((size_t)K + 1);
    float *d_cen = mem.alloc<float>((size_t)K * 3);
    float *d_sums = mem.alloc<float>((size_t)K * 4);
    float4 *d_aos = mem.alloc<float4>(nx);
    PbSeg *d_seg = mem.alloc<PbSeg>(1);
    PbSeg whole{0u, (uint32_t)nx, 0u, 0u};
    PB_CUDA_OK(cudaMemcpyAsync(d_seg, &whole, sizeof whole, cudaMemcpyHostToDevice, st));
    size_t want = (nx + 255) / 256, cap = (size_t)sm_count * 8;
    const int grid = (int)(want < cap ? (want ? want : 1) : cap);
    const size_t smem_want = (size_t)K * sizeof(float4);
    const bool cen_in_smem = smem_want <= PB_SMEM_PALETTE_LIMIT;
    const size_t smem = cen_in_smem ? smem_want : 0;
    if (smem > 32 * 1024)
        PB_CUDA_OK(cudaFuncSetAttribute(k_assign<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    float4 *d_cen4 = mem.alloc<float4>((size_t)K);
    int *d_stop = mem.alloc<int>(1);
    { PbProfScope _prof("k_pack_samples", st, false);
      k_pack_samples<<<grid, 256, 0, st>>>(d_x0, d_x1, d_x2, d_wf, nx, d_aos); }
    std::vector<float> sums((size_t)K * 4);
    // All iterations are enqueued back to back (Clustering.cpp:442-530); the centroid epilogue runs on the
    // device.  Only an empty cluster (rare) hands control back to the host for faiss' split protocol.
    int it = 0;
    while (it < niter) {
        PB_CUDA_OK(cudaMemcpyAsync(d_cen, cen.data(), cen.size() * sizeof(float), cudaMemcpyHostToDevice, st));
        PB_CUDA_OK(cudaMemsetAsync(d_stop, 0, sizeof(int), st));
        for (int i = it; i < niter; i++) {
            { PbProfScope _prof("k_cen4", st, false);
            k_cen4<<<(K + 255) / 256, 256, 0, st>>>(d_cen, K, d_cen4, d_stop);
            }
            { PbProfScope _prof("k_assign", st);
            if (cen_in_smem) k_assign<true><<<grid, 256, smem, st>>>(d_x0, d_x1, d_x2, nx, d_cen4, K, nx < 20, d_assign, d_stop);
            else k_assign<false><<<grid, 256, 0, st>>>(d_x0, d_x1, d_x2, nx, d_cen4, K, nx < 20, d_assign, d_stop);
            }
            pb_launch_class_rank(PB_CLS_BUCKET, K, d_seg, 1, (uint32_t)nx, tiles, d_assign, nullptr, nullptr, d_tile_hist,
                                 d_cstart, st);
            pb_launch_scatter_ord(PB_CLS_BUCKET, K, d_seg, 1, (uint32_t)nx, d_assign, nullptr, nullptr, d_tile_hist,
                                  d_cstart, d_ord, st);
            const int cg = (K + KM_WARPS - 1) / KM_WARPS;
            { PbProfScope _prof("k_centroid_chains", st);
            if (d_wf) k_centroid_chains<true><<<cg, KM_WARPS * 32, 0, st>>>(d_aos, d_ord, d_cstart, K, d_sums, d_stop);
            else k_centroid_chains<false><<<cg, KM_WARPS * 32, 0, st>>>(d_aos, d_ord, d_cstart, K, d_sums, d_stop);
            }
            { PbProfScope _prof("k_finalize_centroids", st, false);
            k_finalize_centroids<<<1, 256, 0, st>>>(d_sums, K, i, d_cen, d_stop);
            }
        }
        PB_CUDA_OK(cudaGetLastError());
        int stop = 0;
        PB_CUDA_OK(cudaMemcpyAsync(&stop, d_stop, sizeof(int), cudaMemcpyDeviceToHost, st));
        PB_CUDA_OK(cudaMemcpyAsync(cen.data(), d_cen, cen.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
        PB_CUDA_OK(cudaMemcpyAsync(sums.data(), d_sums, sums.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
        PB_CUDA_OK(cudaStreamSynchronize(st));
        if (!stop) break; // all remaining iterations completed on the device
        // iteration stop - 1 produced an empty cluster: finish it on the host exactly as faiss does
        std::vector<float> hassign(K);
        for (int c = 0; c < K; c++) { // compute_centroids epilogue (Clustering.cpp:194-203)
            hassign[c] = sums[4 * c];
            for (int j = 0; j < 3; j++) cen[3 * c + j] = sums[4 * c + 1 + j];
            if (hassign[c] == 0) continue;
            const float norm = 1 / hassign[c];
            for (int j = 0; j < 3; j++) cen[3 * c + j] *= norm;
        }
        Mt19937 rng(1234u); // split_clusters (Clustering.cpp:216-263)
        for (int ci = 0; ci < K; ci++) {
            if (hassign[ci] != 0) continue;
            int cj;
            for (cj = 0; true; cj = (cj + 1) % K) {
                float p = (hassign[cj] - 1.0) / (float)(nx - (size_t)K);
                float r = (float)(uint64_t)rng.next() / 4294967295.0f; // mt() / float(mt.max())
                if (r < p) break;
            }
            memcpy(&cen[3 * ci], &cen[3 * cj], 3 * sizeof(float));
            for (int j = 0; j < 3; j++) {
                if (j % 2 == 0) { cen[3 * ci + j] *= 1 + (1 / 1024.); cen[3 * cj + j] *= 1 - (1 / 1024.); }
                else { cen[3 * ci + j] *= 1 - (1 / 1024.); cen[3 * cj + j] *= 1 + (1 / 1024.); }
            }
            hassign[ci] = hassign[cj] / 2;
            hassign[cj] -= hassign[ci];
        }
        it = stop; // resume with the next iteration
    }
}

void pb_kmeans_refine(const double *const planes[3], const double *d_w, size_t n, std::vector<double> &pal_rm,
                      int niter, int max_points_per_centroid, int sm_count, cudaStream_t st, long *launches) {
    const size_t K = pal_rm.size() / 3;
    std::vector<float> cen(3 * K);
    for (size_t j = 0; j < 3 * K; j++) cen[j] = (float)pal_rm[j]; // refine.c:108-120
    auto finish = [&]() { // refine.c:202-212 runs whatever faiss did (error code ignored, bug B8)
        for (size_t j = 0; j < 3 * K; j++) pal_rm[j] = (double)cen[j];
    };
    // Clustering.cpp:273-279 throws for n < K.  K > PB_KMEANS_MAX_K with n >= K never gets here: patolette()
    // rejects it with exit code -6 (16-bit assignments, per-warp class counters of the stable sort).
    if (n < K || K > PB_KMEANS_MAX_K) { finish(); return; }
    DevMem mem;
    int *d_flag = mem.alloc<int>(1);
    PB_CUDA_OK(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
    size_t want = (n + 255) / 256, cap = (size_t)sm_count * 8;
    const int grid = (int)(want < cap ? want : cap);
    size_t nx = n;
    uint32_t *d_pick = nullptr;
    const bool subsample = n > K * (size_t)max_points_per_centroid; // Clustering.cpp:311
    if (subsample) {
        { PbProfScope _prof("k_scan_finite", st);
        k_scan_finite<<<grid, 256, 0, st>>>(planes[0], planes[1], planes[2], n, d_flag);
        }
        nx = K * (size_t)max_points_per_centroid;
        std::vector<uint32_t> pick = rand_perm_prefix(n, nx, 1234u);
        d_pick = mem.alloc<uint32_t>(nx);
        PB_CUDA_OK(cudaMemcpyAsync(d_pick, pick.data(), nx * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        PB_CUDA_OK(cudaStreamSynchronize(st));
    }
    float *d_x0 = mem.alloc<float>(nx), *d_x1 = mem.alloc<float>(nx), *d_x2 = mem.alloc<float>(nx);
    float *d_wf = d_w ? mem.alloc<float>(nx) : nullptr;
    { PbProfScope _prof("k_to_f32", st);
    k_to_f32<<<grid, 256, 0, st>>>(planes[0], planes[1], planes[2], d_w, d_pick, nx, d_x0, d_x1, d_x2, d_wf, d_flag);
    }
    PB_CUDA_OK(cudaGetLastError());
    int flag = 0;
    PB_CUDA_OK(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    PB_CUDA_OK(cudaStreamSynchronize(st));
    if (flag) { finish(); return; } // "input contains NaN's or Inf's" -> exception -> unrefined centres
    if (nx == K) { // Clustering.cpp:330-352: centroids = the first k input vectors
        std::vector<float> h(3 * K);
        for (int j = 0; j < 3; j++) {
            std::vector<double> tmp(K);
            PB_CUDA_OK(cudaMemcpy(tmp.data(), planes[j], K * sizeof(double), cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < K; i++) h[3 * i + j] = (float)tmp[i];
        }
        cen = h;
        finish();
        return;
    }
    pb_kmeans_device(d_x0, d_x1, d_x2, d_wf, nx, cen, (int)K, niter, sm_count, st, launches);
    finish();
}
