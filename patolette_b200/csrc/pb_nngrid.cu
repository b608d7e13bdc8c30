// pb_nngrid.cu - exact nearest-palette assignment with candidate lists (palette/nearest.c:150-209).
//
// The reference asks for the exact squared-L2 1-NN of every pixel among the K palette colours, in f64,
// lowest index on ties (oracle: brute force).  Brute force is 9*K FP64 operations per pixel - FP64-pipe
// bound (2.8 ms at 4096^2, K=256).  Here the bounding box of the pixels is cut into 16^3 cells and every
// cell gets the list of palette entries that can be nearest to SOME point of the cell:
//     keep j  iff  mind(cell, j) <= min_k maxd(cell, k)         (with a relative margin, see below)
// where mind / maxd are the smallest / largest squared distance from entry j to the cell's box.  For a
// point x of the cell and k* = argmin maxd: d(x, nearest) <= d(x, k*) <= maxd(k*), and d(x, j) >= mind(j),
// so the true nearest entry - and every entry that ties with it - is on the list.  A pixel then evaluates
// the reference's own expression  (dx*dx + dy*dy) + dz*dz  only for the ~10-20 entries of its cell, in
// ascending index order with a strict <, which is what the brute force does restricted to a superset of
// the possible winners: the result is identical.
//
// Floating point: every box edge, and every difference between a palette coordinate and a box edge, carries
// an ABSOLUTE error of a few ulps of the coordinates themselves (which matters when the pixel range is tiny
// next to the coordinates, e.g. colours 0.5 +- 1e-9).  The boxes are therefore inflated by
// 1e-9 * range + 1e-15 * max|coordinate| (a pixel whose cell index was rounded across a border is still
// inside the box it was assigned to), near distances are shrunk and far distances grown by
// 1e-15 * (max|coordinate| + |p|) before squaring, and the comparison keeps j when
// mind * (1 - 1e-9) <= best_maxd * (1 + 1e-9).  Pixels that are not inside the grid at all (NaN, or outside
// the box by rounding) take the brute-force loop.
#include "pb_common.cuh"
#include "pb_kernels.h"
#include "pb_nngrid.cuh"
#include "pb_prof.h"

namespace {

using GridHdr = PbGridHdr;
using GridGeom = PbGridGeom;
__device__ __forceinline__ GridGeom grid_geom(const GridHdr *h) { return pb_grid_geom(h); }

__global__ void k_nn_bbox_init(GridHdr *h, double s0, double s1, double s2, double expand, int ng) {
    if (threadIdx.x < 3) { h->mn[threadIdx.x] = ~0ULL; h->mx[threadIdx.x] = 0ULL; }
    if (threadIdx.x == 0) { h->scale[0] = s0; h->scale[1] = s1; h->scale[2] = s2; h->expand = expand; h->ng = ng; }
}

__global__ void __launch_bounds__(256) k_nn_bbox(const double *__restrict__ c0, const double *__restrict__ c1,
                                                 const double *__restrict__ c2, size_t n, GridHdr *h) {
    unsigned long long mn[3] = {~0ULL, ~0ULL, ~0ULL}, mx[3] = {0, 0, 0};
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double v[3] = {c0[i], c1[i], c2[i]};
#pragma unroll
        for (int d = 0; d < 3; d++) {
            if (v[d] != v[d]) continue; // NaN: not part of the box
            const unsigned long long e = pb_ord_encode(v[d]);
            mn[d] = e < mn[d] ? e : mn[d];
            mx[d] = e > mx[d] ? e : mx[d];
        }
    }
#pragma unroll
    for (int d = 0; d < 3; d++) {
        for (int o = 16; o; o >>= 1) {
            const unsigned long long a = __shfl_xor_sync(0xffffffffu, mn[d], o), b = __shfl_xor_sync(0xffffffffu, mx[d], o);
            mn[d] = a < mn[d] ? a : mn[d];
            mx[d] = b > mx[d] ? b : mx[d];
        }
        if ((threadIdx.x & 31) == 0) { atomicMin(&h->mn[d], mn[d]); atomicMax(&h->mx[d], mx[d]); }
    }
}

// one CTA per cell: candidate list in ascending palette index
__global__ void __launch_bounds__(128) k_nn_cells(const GridHdr *__restrict__ hdr, const double *__restrict__ pal, int K,
                                                  unsigned short *__restrict__ cnt, unsigned short *__restrict__ list) {
    extern __shared__ unsigned char s_raw[];
    double *s_mind = reinterpret_cast<double *>(s_raw);      // [K]
    __shared__ double s_red[4];
    __shared__ double s_best;
    const GridGeom g = grid_geom(hdr);
    const int NG = g.ng;
    const int cell = blockIdx.x, cz = cell % NG, cy = (cell / NG) % NG, cx = cell / (NG * NG);
    const int cc[3] = {cx, cy, cz};
    double lo[3], hi[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const double pad = 1e-9 * g.w[d] * NG + 1e-15 * g.cmax[d] + 1e-300;
        lo[d] = g.lo[d] + cc[d] * g.w[d] - pad;
        hi[d] = g.lo[d] + (cc[d] + 1) * g.w[d] + pad;
    }
    double best = 1e308;
    for (int j = threadIdx.x; j < K; j += blockDim.x) {
        double mind = 0.0, maxd = 0.0;
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const double p = pal[3 * j + d];
            const double err = 1e-15 * (g.cmax[d] + fabs(p));
            const double below = lo[d] - p - err, above = p - hi[d] - err;
            const double out = below > 0 ? below : (above > 0 ? above : 0.0);
            const double f1 = fabs(p - lo[d]), f2 = fabs(p - hi[d]);
            const double far = (f1 > f2 ? f1 : f2) + err;
            mind += out * out;
            maxd += far * far;
        }
        if (!(mind == mind)) mind = 0.0;       // NaN palette entries stay on every list
        if (!(maxd == maxd)) maxd = 1e308;
        s_mind[j] = mind;
        best = maxd < best ? maxd : best;
    }
    for (int o = 16; o; o >>= 1) { const double v = __shfl_xor_sync(0xffffffffu, best, o); best = v < best ? v : best; }
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        double b = s_red[0];
        for (int w = 1; w < 4; w++) b = s_red[w] < b ? s_red[w] : b;
        s_best = b * (1.0 + 1e-9) + 1e-300;
    }
    __syncthreads();
    if (threadIdx.x < 32) { // ordered compaction by one warp
        const int lane = threadIdx.x;
        unsigned int n = 0;
        unsigned short *out = list + (size_t)cell * K;
        for (int j0 = 0; j0 < K; j0 += 32) {
            const int j = j0 + lane;
            const bool keep = j < K && s_mind[j] * (1.0 - 1e-9) <= s_best;
            const unsigned int m = __ballot_sync(0xffffffffu, keep);
            if (keep) out[n + __popc(m & ((1u << lane) - 1u))] = (unsigned short)j;
            n += __popc(m);
        }
        if (lane == 0) cnt[cell] = (unsigned short)n;
    }
}

__global__ void __launch_bounds__(256) k_nearest_grid(const double *__restrict__ c0, const double *__restrict__ c1,
                                                      const double *__restrict__ c2, size_t n,
                                                      const double *__restrict__ pal, int K,
                                                      const GridHdr *__restrict__ hdr, const unsigned short *__restrict__ cnt,
                                                      const unsigned short *__restrict__ list,
                                                      unsigned long long *__restrict__ map) {
    extern __shared__ double s_pal[];
    for (int i = threadIdx.x; i < K * 3; i += blockDim.x) s_pal[i] = pal[i];
    __syncthreads();
    const GridGeom g = grid_geom(hdr);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double x = c0[i], y = c1[i], z = c2[i];
        double bd = 0.0;
        int best = 0;
        const int cell = pb_grid_cell(g.ok, g.lo, g.inv, g.ng, x, y, z);
        if (cell >= 0) {
            const int m = cnt[cell];
            const unsigned short *L = list + (size_t)cell * K;
            for (int t = 0; t < m; t++) {
                const int j = L[t];
                const double dx = __dsub_rn(x, s_pal[3 * j]), dy = __dsub_rn(y, s_pal[3 * j + 1]),
                             dz = __dsub_rn(z, s_pal[3 * j + 2]);
                const double dd = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                if (t == 0 || dd < bd) { bd = dd; best = j; }
            }
        } else {
#pragma unroll 4
            for (int j = 0; j < K; j++) {
                const double dx = __dsub_rn(x, s_pal[3 * j]), dy = __dsub_rn(y, s_pal[3 * j + 1]),
                             dz = __dsub_rn(z, s_pal[3 * j + 2]);
                const double dd = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                if (j == 0 || dd < bd) { bd = dd; best = j; }
            }
        }
        map[i] = (unsigned long long)best;
    }
}

} // namespace

size_t pb_nngrid_scratch_bytes(int K) {
    const size_t ncell = (size_t)pb_grid_ng(K) * pb_grid_ng(K) * pb_grid_ng(K);
    return 256 + ncell * 2 + ncell * (size_t)K * 2;
}

// pixels -> bounding box -> per-cell candidate lists (d_scratch: pb_nngrid_scratch_bytes(K))
void pb_launch_nngrid_build(const double *const planes[3], size_t n, const double *d_palette_rm, int K, void *d_scratch,
                            int sm_count, cudaStream_t st, const double *scale, double expand) {
    GridHdr *hdr = (GridHdr *)d_scratch;
    const int ng = pb_grid_ng(K), NCELL = ng * ng * ng;
    unsigned short *cnt = (unsigned short *)((char *)d_scratch + 256);
    unsigned short *list = cnt + NCELL;
    { PbProfScope _prof("k_nn_bbox", st, false);
      k_nn_bbox_init<<<1, 32, 0, st>>>(hdr, scale ? scale[0] : 1.0, scale ? scale[1] : 1.0, scale ? scale[2] : 1.0, expand, ng);
      size_t want = (n + 256 * 8 - 1) / (256 * 8), cap = (size_t)sm_count * 8;
      k_nn_bbox<<<(int)(want < cap ? (want ? want : 1) : cap), 256, 0, st>>>(planes[0], planes[1], planes[2], n, hdr); }
    { PbProfScope _prof("k_nn_cells", st, false);
      const size_t smem = (size_t)K * sizeof(double);
      if (smem > 32 * 1024) PB_CUDA_OK(cudaFuncSetAttribute(k_nn_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_nn_cells<<<NCELL, 128, smem, st>>>(hdr, d_palette_rm, K, cnt, list); }
    PB_CUDA_OK(cudaGetLastError());
}

void pb_launch_nearest_grid(const double *const planes[3], size_t n, const double *d_palette_rm, int K, const void *d_scratch,
                            unsigned long long *d_map, int sm_count, cudaStream_t st) {
    if (n == 0) return;
    const GridHdr *hdr = (const GridHdr *)d_scratch;
    const int NCELL = pb_grid_ng(K) * pb_grid_ng(K) * pb_grid_ng(K);
    const unsigned short *cnt = (const unsigned short *)((const char *)d_scratch + 256);
    const unsigned short *list = cnt + NCELL;
    size_t want = (n + 255) / 256, cap = (size_t)sm_count * 8;
    const int grid = (int)(want < cap ? want : cap);
    const size_t smem = (size_t)K * 3 * sizeof(double);
    if (smem > 32 * 1024)
        PB_CUDA_OK(cudaFuncSetAttribute(k_nearest_grid, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    { PbProfScope _prof("k_nearest", st);
      k_nearest_grid<<<grid, 256, smem, st>>>(planes[0], planes[1], planes[2], n, d_palette_rm, K, hdr, cnt, list, d_map); }
    PB_CUDA_OK(cudaGetLastError());
}
