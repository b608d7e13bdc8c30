// pb_dither.cu - Riemersma dither along a Hilbert curve (reference: lib/src/dither/riemersma.c).
//
// The reference walks a 2^level x 2^level Hilbert curve from (0,0) (riemersma.c:176-257),
// dithering the pixels it meets inside the image (:146-156); each pixel depends on the 16
// previous errors along the curve (:275-341), so the recurrence itself is strictly
// sequential - there is no wavefront to exploit.  What CAN be parallel is
// everything around it:
//   1. k_hilbert_rank: every pixel computes, in closed form, its position in the walk
//      (its Hilbert index minus the out-of-image cells before it: a quadtree descent that
//      adds the clipped areas of the sibling quadrants the curve visits first);
//   2. k_permute: pixels are laid out in walk order (coalesced reads for step 3);
//   3. k_riemersma_chain: ONE warp runs the recurrence: lanes 0..2 own the R/G/B error
//      queues, all 32 lanes split the K palette candidates of the exact nearest-neighbour
//      search (K/32 each) and butterfly-reduce the argmin;
//   4. k_unpermute: indices return to raster order.
// Arithmetic order is the reference's: 16-tap error sum oldest first with separately
// rounded products (:292-297), no clamping (:299-312), query scaled by the double
// sqrt-luma weights, palette by the float-rounded ones (:315-317 vs :419-425, bug B5),
// squared L2 summed R,G,B, lowest index on ties.
#include <math.h>

#include <vector>

#include "pb_common.cuh"
#include "pb_kernels.h"
#include "pb_prof.h"
#include "pb_pipeline.h"
#include "pb_pool.h"

namespace {

enum { D_UP = 0, D_LEFT = 1, D_RIGHT = 2, D_DOWN = 3 };
// per direction: the four quadrants in visiting order as (qx | qy << 1), and their sub-directions
__constant__ uint8_t c_quad[4][4] = {
    {0, 2, 3, 1}, // UP:    (0,0) (0,1) (1,1) (1,0)
    {0, 1, 3, 2}, // LEFT:  (0,0) (1,0) (1,1) (0,1)
    {3, 2, 0, 1}, // RIGHT: (1,1) (0,1) (0,0) (1,0)
    {3, 1, 0, 2}, // DOWN:  (1,1) (1,0) (0,0) (0,1)
};
__constant__ uint8_t c_sub[4][4] = {
    {D_LEFT, D_UP, D_UP, D_RIGHT},
    {D_UP, D_LEFT, D_LEFT, D_DOWN},
    {D_DOWN, D_RIGHT, D_RIGHT, D_UP},
    {D_RIGHT, D_DOWN, D_DOWN, D_LEFT},
};

__device__ __forceinline__ unsigned long long clipped_area(uint32_t ox, uint32_t oy, uint32_t s, uint32_t W,
                                                          uint32_t H) {
    if (ox >= W || oy >= H) return 0;
    const uint32_t w = min(ox + s, W) - ox, h = min(oy + s, H) - oy;
    return (unsigned long long)w * h;
}

__global__ void k_hilbert_rank(uint32_t W, uint32_t H, int level, uint32_t *__restrict__ rank) {
    const size_t n = (size_t)W * H;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
        const uint32_t x = (uint32_t)(p % W), y = (uint32_t)(p / W);
        uint32_t ox = 0, oy = 0;
        int dir = D_UP;
        unsigned long long r = 0;
        for (int l = level - 1; l >= 0; l--) {
            const uint32_t s = 1u << l;
            const uint32_t me = ((x >> l) & 1u) | (((y >> l) & 1u) << 1);
            int k = 0;
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const uint32_t q = c_quad[dir][t];
                if (q == me) { k = t; break; }
                r += clipped_area(ox + (q & 1u) * s, oy + (q >> 1) * s, s, W, H);
            }
            ox += (me & 1u) * s;
            oy += (me >> 1) * s;
            dir = c_sub[dir][k];
        }
        rank[p] = (uint32_t)r;
    }
}

__global__ void k_permute(const double *__restrict__ c0, const double *__restrict__ c1,
                          const double *__restrict__ c2, const uint32_t *__restrict__ rank, size_t n,
                          double *__restrict__ h0, double *__restrict__ h1, double *__restrict__ h2) {
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
        const uint32_t r = rank[p];
        h0[r] = c0[p]; h1[r] = c1[p]; h2[r] = c2[p];
    }
}

__global__ void k_unpermute(const uint32_t *__restrict__ rank, const uint32_t *__restrict__ hidx, size_t n,
                            unsigned long long *__restrict__ map) {
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x)
        map[p] = hidx[rank[p]];
}

constexpr int DT_TILE = 128; // pixels staged per tile

// One warp.  pal: K x 3 row-major (linear Rec2020); palw: same scaled by the float weights.
__global__ void __launch_bounds__(32) k_riemersma_chain(const double *__restrict__ h0, const double *__restrict__ h1,
                                                        const double *__restrict__ h2, size_t n,
                                                        const double *__restrict__ pal,
                                                        const double *__restrict__ palw, int K,
                                                        const double *__restrict__ qweights,
                                                        uint32_t *__restrict__ hidx) {
    extern __shared__ double s_mem[];
    double *s_pal = s_mem;                 // K * 3
    double *s_palw = s_mem + (size_t)K * 3; // K * 3
    double *s_px = s_palw + (size_t)K * 3;  // 2 * 3 * DT_TILE
    const int lane = threadIdx.x;
    for (int i = lane; i < K * 3; i += 32) { s_pal[i] = pal[i]; s_palw[i] = palw[i]; }
    double w[16];
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = qweights[i];
    // riemersma.c:30-34 (double weights on the query side)
    const double cw = lane == 0 ? 0.51254268114958 : (lane == 1 ? 0.8234075540095561 : 0.2435159132377184);
    double q[16]; // error queue of this lane's channel (lanes 0..2), oldest first
#pragma unroll
    for (int i = 0; i < 16; i++) q[i] = 0.0;

    const size_t ntiles = (n + DT_TILE - 1) / DT_TILE;
    auto stage = [&](size_t t, int buf) {
        const size_t base = t * DT_TILE;
        for (int e = lane; e < DT_TILE; e += 32) {
            const size_t i = base + e;
            double a = 0, b = 0, c = 0;
            if (i < n) { a = h0[i]; b = h1[i]; c = h2[i]; }
            s_px[(buf * 3 + 0) * DT_TILE + e] = a;
            s_px[(buf * 3 + 1) * DT_TILE + e] = b;
            s_px[(buf * 3 + 2) * DT_TILE + e] = c;
        }
    };
    if (ntiles) stage(0, 0);
    __syncwarp();
    for (size_t t = 0; t < ntiles; t++) {
        const int cur = (int)(t & 1);
        if (t + 1 < ntiles) stage(t + 1, cur ^ 1); // loads overlap the chain below
        const int cnt = (int)min((size_t)DT_TILE, n - t * DT_TILE);
        const int ch = lane < 3 ? lane : 0;
        for (int e = 0; e < cnt; e++) {
            // riemersma.c:292-297: error = sum_i queue[i] * weight[i], i ascending
            double err = 0.0;
#pragma unroll
            for (int i = 0; i < 16; i++) err = __dadd_rn(err, __dmul_rn(q[i], w[i]));
            const double P = s_px[(cur * 3 + ch) * DT_TILE + e];
            const double C = __dadd_rn(P, err);      // :310-312, no clamping
            const double Cw = __dmul_rn(cw, C);      // :315-317
            const double x = __shfl_sync(0xffffffffu, Cw, 0), y = __shfl_sync(0xffffffffu, Cw, 1),
                         z = __shfl_sync(0xffffffffu, Cw, 2);
            double bd = 0.0;
            int best = 0x7fffffff;
            for (int j = lane; j < K; j += 32) {
                const double dx = __dsub_rn(x, s_palw[3 * j]), dy = __dsub_rn(y, s_palw[3 * j + 1]),
                             dz = __dsub_rn(z, s_palw[3 * j + 2]);
                const double dd = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                if (best == 0x7fffffff || dd < bd) { bd = dd; best = j; }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const double od = __shfl_xor_sync(0xffffffffu, bd, o);
                const int oj = __shfl_xor_sync(0xffffffffu, best, o);
                // exact argmin, lowest index on ties; lanes with no candidate carry best = INT_MAX
                const bool take = oj != 0x7fffffff && (best == 0x7fffffff || od < bd || (od == bd && oj < best));
                if (take) { bd = od; best = oj; }
            }
            if (lane == 0) hidx[t * DT_TILE + e] = (uint32_t)best;
            // :332-340: shift the queue, append P - palette[best]
#pragma unroll
            for (int i = 0; i < 15; i++) q[i] = q[i + 1];
            q[15] = __dsub_rn(P, s_pal[3 * best + ch]);
        }
        __syncwarp();
    }
}

} // namespace

void pb_dither_riemersma(const double *const planes[3], size_t width, size_t height,
                         const std::vector<double> &pal_rm, unsigned long long *d_map, int sm_count,
                         cudaStream_t st, long *launches) {
    const size_t n = width * height;
    const int K = (int)(pal_rm.size() / 3);
    // riemersma.c:124-144
    int level = 0;
    size_t mx = width > height ? width : height, value = mx;
    while (value > 1) { value >>= 1; level++; }
    if (((size_t)1 << level) < mx) level++;
    if (level == 0 || K == 0) return; // :452-456: a 1x1 image is left untouched (bug B6)

    // host-side constants: queue weights (:360-373) and the float-rounded palette scaling (:419-425)
    double qw[16];
    {
        double m = exp(log((double)16) / ((double)16 - 1)), v = 1;
        for (int i = 0; i < 16; i++) { qw[i] = v / (double)16; v *= m; }
    }
    const double fx = (double)(float)0.51254268114958, fy = (double)(float)0.8234075540095561,
                 fz = (double)(float)0.2435159132377184;
    std::vector<double> palw(pal_rm.size());
    for (int j = 0; j < K; j++) { // nearest.c:32-61
        palw[3 * j] = pal_rm[3 * j] * fx;
        palw[3 * j + 1] = pal_rm[3 * j + 1] * fy;
        palw[3 * j + 2] = pal_rm[3 * j + 2] * fz;
    }
    double *d_h[3] = {nullptr, nullptr, nullptr}, *d_pal = nullptr, *d_palw = nullptr, *d_qw = nullptr;
    uint32_t *d_rank = nullptr, *d_hidx = nullptr;
    auto cleanup = [&]() {
        for (int j = 0; j < 3; j++) pb_pool_free(d_h[j]);
        pb_pool_free(d_pal); pb_pool_free(d_palw); pb_pool_free(d_qw); pb_pool_free(d_rank); pb_pool_free(d_hidx);
    };
    try {
        for (int j = 0; j < 3; j++) d_h[j] = (double *)pb_pool_alloc(n * sizeof(double));
        d_pal = (double *)pb_pool_alloc(pal_rm.size() * sizeof(double));
        d_palw = (double *)pb_pool_alloc(pal_rm.size() * sizeof(double));
        d_qw = (double *)pb_pool_alloc(sizeof qw);
        d_rank = (uint32_t *)pb_pool_alloc(n * sizeof(uint32_t));
        d_hidx = (uint32_t *)pb_pool_alloc((n + DT_TILE) * sizeof(uint32_t));
        PB_CUDA_OK(cudaMemcpyAsync(d_pal, pal_rm.data(), pal_rm.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        PB_CUDA_OK(cudaMemcpyAsync(d_palw, palw.data(), palw.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        PB_CUDA_OK(cudaMemcpyAsync(d_qw, qw, sizeof qw, cudaMemcpyHostToDevice, st));
        size_t want = (n + 255) / 256, cap = (size_t)sm_count * 8;
        const int grid = (int)(want < cap ? want : cap);
        { PbProfScope _prof("k_hilbert_rank", st);
        k_hilbert_rank<<<grid, 256, 0, st>>>((uint32_t)width, (uint32_t)height, level, d_rank);
        }
        { PbProfScope _prof("k_permute", st);
        k_permute<<<grid, 256, 0, st>>>(planes[0], planes[1], planes[2], d_rank, n, d_h[0], d_h[1], d_h[2]);
        }
        const size_t smem = ((size_t)K * 6 + 2 * 3 * DT_TILE) * sizeof(double);
        if (smem > 48 * 1024)
            PB_CUDA_OK(cudaFuncSetAttribute(k_riemersma_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        { PbProfScope _prof("k_riemersma_chain", st);
        k_riemersma_chain<<<1, 32, smem, st>>>(d_h[0], d_h[1], d_h[2], n, d_pal, d_palw, K, d_qw, d_hidx);
        }
        { PbProfScope _prof("k_unpermute", st);
        k_unpermute<<<grid, 256, 0, st>>>(d_rank, d_hidx, n, d_map);
        }
        PB_CUDA_OK(cudaGetLastError());
        PB_CUDA_OK(cudaStreamSynchronize(st));
    } catch (...) {
        cleanup();
        throw;
    }
    cleanup();
}
