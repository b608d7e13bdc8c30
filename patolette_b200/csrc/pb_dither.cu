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
#include "pb_nccl.h"
#include "pb_nngrid.cuh"
#include "pb_prof.h"
#include "pb_pipeline.h"
#include "pb_pool.h"
#include "pb_tma.cuh"

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

__global__ void k_unpermute(const uint32_t *__restrict__ rank, const uint32_t *__restrict__ hidx, size_t first, size_t n,
                            unsigned long long *__restrict__ map) {
    for (size_t p = first + (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < first + n; p += (size_t)gridDim.x * blockDim.x)
        map[p] = hidx[rank[p]];
}

// Tile versions of the two permutations.  An aligned 32 x 32 block of the image is ONE sub-square of the recursion,
// so its in-image pixels occupy a contiguous range of the walk (k_hilbert_rank: the levels above 5 add the same
// areas for all of them): a CTA reads the block row by row (coalesced), finds the range's start (the smallest rank),
// reorders in shared memory and writes the range as one contiguous run - instead of 8-byte stores scattered over
// 32-byte sectors (measured 1.5 TB/s for k_permute at 16384^2).
__device__ __forceinline__ uint32_t block_min_u32(uint32_t v, uint32_t *s_red /* [8] */) {
    v = __reduce_min_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    uint32_t r = s_red[0];
#pragma unroll
    for (int w = 1; w < 8; w++) r = min(r, s_red[w]);
    return r;
}

__global__ void __launch_bounds__(256) k_permute_tile(const double *__restrict__ c0, const double *__restrict__ c1,
                                                      const double *__restrict__ c2, const uint32_t *__restrict__ rank,
                                                      uint32_t W, uint32_t H, double *__restrict__ h0,
                                                      double *__restrict__ h1, double *__restrict__ h2) {
    __shared__ __align__(128) double sm[3][1024];
    __shared__ uint32_t s_red[8];
    const uint32_t x = blockIdx.x * 32 + (threadIdx.x & 31), y0 = blockIdx.y * 32 + (threadIdx.x >> 5);
    uint32_t r[4], rmin = 0xffffffffu;
    double v0[4], v1[4], v2[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t y = y0 + 8 * k;
        const bool in = x < W && y < H;
        const size_t p = (size_t)y * W + x;
        r[k] = in ? rank[p] : 0xffffffffu;
        v0[k] = in ? c0[p] : 0.0; v1[k] = in ? c1[p] : 0.0; v2[k] = in ? c2[p] : 0.0;
        rmin = min(rmin, r[k]);
    }
    const uint32_t r0 = block_min_u32(rmin, s_red);
    const uint32_t count = (min(blockIdx.x * 32 + 32, W) - blockIdx.x * 32) * (min(blockIdx.y * 32 + 32, H) - blockIdx.y * 32);
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (r[k] != 0xffffffffu && r[k] - r0 < 1024u) { sm[0][r[k] - r0] = v0[k]; sm[1][r[k] - r0] = v1[k]; sm[2][r[k] - r0] = v2[k]; }
    if (((r0 | count) & 1u) == 0u) {
        // the block's run leaves as three bulk copies (TMA, shared -> global: 8 KB each for a whole block): 16-byte
        // aligned on both sides because the run starts and ends on an even walk position
        pb_fence_proxy_async();
        __syncthreads();
        if (threadIdx.x == 0) {
            pb_bulk_store(h0 + r0, sm[0], count * 8u);
            pb_bulk_store(h1 + r0, sm[1], count * 8u);
            pb_bulk_store(h2 + r0, sm[2], count * 8u);
            pb_bulk_commit();
            pb_bulk_wait_read();
        }
        return;
    }
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < count; t += 256) {
        h0[(size_t)r0 + t] = sm[0][t]; h1[(size_t)r0 + t] = sm[1][t]; h2[(size_t)r0 + t] = sm[2][t];
    }
}

__global__ void __launch_bounds__(256) k_unpermute_tile(const uint32_t *__restrict__ rank, const uint32_t *__restrict__ hidx,
                                                        uint32_t W, uint32_t H, size_t first, size_t n,
                                                        unsigned long long *__restrict__ map) {
    __shared__ __align__(128) uint32_t sm[1024];
    __shared__ uint32_t s_red[8];
    __shared__ __align__(8) unsigned long long s_bar;
    const uint32_t x = blockIdx.x * 32 + (threadIdx.x & 31), y0 = blockIdx.y * 32 + (threadIdx.x >> 5);
    uint32_t r[4], rmin = 0xffffffffu;
    if (threadIdx.x == 0) pb_mbar_init(&s_bar, 1);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t y = y0 + 8 * k;
        const bool in = x < W && y < H;
        r[k] = in ? rank[(size_t)y * W + x] : 0xffffffffu;
        rmin = min(rmin, r[k]);
    }
    const uint32_t r0 = block_min_u32(rmin, s_red); // (its barrier also publishes the mbarrier's initialisation)
    const uint32_t count = (min(blockIdx.x * 32 + 32, W) - blockIdx.x * 32) * (min(blockIdx.y * 32 + 32, H) - blockIdx.y * 32);
    if (((r0 | count) & 3u) == 0u) {
        // the block's choices arrive as one bulk copy (TMA, global -> shared, 4 KB for a whole block) that signals
        // the mbarrier every thread then waits on
        if (threadIdx.x == 0) {
            pb_mbar_expect_tx(&s_bar, count * 4u);
            pb_bulk_load(sm, hidx + r0, count * 4u, &s_bar);
        }
        pb_mbar_wait(&s_bar, 0);
    } else {
        for (uint32_t t = threadIdx.x; t < count; t += 256) sm[t] = hidx[(size_t)r0 + t];
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const size_t p = (size_t)(y0 + 8 * k) * W + x;
        if (r[k] != 0xffffffffu && r[k] - r0 < 1024u && p >= first && p < first + n) map[p] = sm[r[k] - r0];
    }
}

// ---- the recurrence ------------------------------------------------------------------------------
// State of a chain = the 16-entry error queue; entry s is P_s - palette[idx_s], so the state after
// pixel t is a function of the last 16 CHOICES only.  Two runs over the same pixels that make the
// same 16 consecutive choices are in bit-identical states from then on.  That finite memory is what
// lets the walk be cut into segments:
//   k_riemersma_spec   : segment g is run by its own warp from an EMPTY queue, starting DT_WARM
//                        pixels early; by the time it reaches its segment the warm-up has (almost
//                        always) locked onto the true trajectory.  It also records its choices for
//                        the 16 pixels just before the segment.
//   k_riemersma_repair : one warp walks the segment boundaries in order.  If segment g's 16 recorded
//                        warm-up choices equal what segment g-1 produced there, g started in the
//                        true state - nothing to do.  Otherwise the true chain is continued from
//                        the boundary, overwriting choices, until 16 consecutive choices agree with
//                        the speculative ones again (measured: ~150 pixels on average).
// Correct for ANY data - the speculation only buys parallelism; with no agreement at all the repair
// kernel degenerates into the plain sequential walk.
constexpr int DT_WARPS = 8;

struct DitherLane {
    double q[16];  // error queue of this lane's channel (lanes 0..2), oldest first
    double w[16];  // queue weights (riemersma.c:360-373)
    double cw;     // sqrt-luma weight of this lane's channel, double precision (riemersma.c:30-34)
    int ch;
};

// Candidate lists of the exact nearest-neighbour search (pb_nngrid.cu), built over the weighted query space:
// a query inside the grid looks at the ~10-20 entries of its cell instead of all K.  geom = {lo[3], inv[3]}
// in shared memory; cnt == nullptr: brute force.
struct DitherGrid {
    const double *geom;
    const unsigned short *cnt, *list;
    int ng;
};

// One pixel: returns the chosen palette index (uniform across the warp) and updates the queue.
__device__ __forceinline__ int dither_step(DitherLane &L, double P, const double *__restrict__ s_pal,
                                           const double *__restrict__ s_palw, int K, int lane, const DitherGrid &G) {
    // riemersma.c:292-297: error = sum_i queue[i] * weight[i], i ascending
    double err = 0.0;
#pragma unroll
    for (int i = 0; i < 16; i++) err = __dadd_rn(err, __dmul_rn(L.q[i], L.w[i]));
    const double C = __dadd_rn(P, err);   // :310-312, no clamping
    const double Cw = __dmul_rn(L.cw, C); // :315-317
    const double x = __shfl_sync(0xffffffffu, Cw, 0), y = __shfl_sync(0xffffffffu, Cw, 1),
                 z = __shfl_sync(0xffffffffu, Cw, 2);
    double bd = 0.0;
    int best = 0x7fffffff;
    const int cell = G.cnt ? pb_grid_cell(true, G.geom, G.geom + 3, G.ng, x, y, z) : -1; // warp-uniform
    if (cell >= 0) {
        const unsigned short *Lst = G.list + (size_t)cell * K;
        const int j0 = Lst[lane]; // K >= 64: in bounds; issued together with the count
        const int m = G.cnt[cell];
        for (int t = lane; t < m; t += 32) {
            const int j = t == lane ? j0 : (int)Lst[t];
            const double dx = __dsub_rn(x, s_palw[3 * j]), dy = __dsub_rn(y, s_palw[3 * j + 1]),
                         dz = __dsub_rn(z, s_palw[3 * j + 2]);
            const double dd = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            if (best == 0x7fffffff || dd < bd) { bd = dd; best = j; }
        }
    } else {
        for (int j = lane; j < K; j += 32) {
            const double dx = __dsub_rn(x, s_palw[3 * j]), dy = __dsub_rn(y, s_palw[3 * j + 1]),
                         dz = __dsub_rn(z, s_palw[3 * j + 2]);
            const double dd = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            if (best == 0x7fffffff || dd < bd) { bd = dd; best = j; }
        }
    }
    // exact warp argmin, lowest index on ties, via integer reductions: squared distances are
    // non-negative doubles, whose bit patterns order like unsigned integers
    {
        const unsigned long long key = best == 0x7fffffff ? ~0ULL : (unsigned long long)__double_as_longlong(bd);
        const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
        const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
        const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
        const bool win = hi == mhi && lo == mlo;
        best = (int)__reduce_min_sync(0xffffffffu, win ? (unsigned)best : 0x7fffffffu);
    }
    // :332-340: shift the queue, append P - palette[best]
#pragma unroll
    for (int i = 0; i < 15; i++) L.q[i] = L.q[i + 1];
    L.q[15] = __dsub_rn(P, s_pal[3 * best + L.ch]);
    return best;
}

__device__ __forceinline__ void dither_lane_init(DitherLane &L, const double *__restrict__ qweights, int lane) {
#pragma unroll
    for (int i = 0; i < 16; i++) { L.w[i] = qweights[i]; L.q[i] = 0.0; }
    L.cw = lane == 0 ? 0.51254268114958 : (lane == 1 ? 0.8234075540095561 : 0.2435159132377184);
    L.ch = lane < 3 ? lane : 0;
}

// Runs pixels [from, to) of the walk: 32 at a time, each lane fetches one pixel (coalesced) and the
// values reach the channel lanes by shuffle; choices are collected one per lane and stored coalesced.
// emit(pos, idx) decides where a choice goes.
template <typename Emit>
__device__ __forceinline__ void dither_run(DitherLane &L, const double *__restrict__ h0, const double *__restrict__ h1,
                                           const double *__restrict__ h2, size_t from, size_t to,
                                           const double *__restrict__ s_pal, const double *__restrict__ s_palw, int K,
                                           int lane, const DitherGrid &G, Emit emit) {
    for (size_t base = from; base < to; base += 32) {
        const size_t i = base + lane;
        double a = 0, b = 0, c = 0;
        if (i < to) { a = h0[i]; b = h1[i]; c = h2[i]; }
        const int cnt = (int)min((size_t)32, to - base);
        int mine = 0;
        for (int e = 0; e < cnt; e++) {
            const double pa = __shfl_sync(0xffffffffu, a, e), pb = __shfl_sync(0xffffffffu, b, e),
                         pc = __shfl_sync(0xffffffffu, c, e);
            const double P = L.ch == 0 ? pa : (L.ch == 1 ? pb : pc);
            const int best = dither_step(L, P, s_pal, s_palw, K, lane, G);
            if (lane == e) mine = best;
        }
        if (lane < cnt) emit(i, mine);
    }
}

__global__ void __launch_bounds__(DT_WARPS * 32) k_riemersma_spec(const double *__restrict__ h0, const double *__restrict__ h1,
                                                                 const double *__restrict__ h2, size_t n, size_t seg,
                                                                 size_t warm, const double *__restrict__ pal,
                                                                 const double *__restrict__ palw, int K,
                                                                 const double *__restrict__ qweights,
                                                                 uint32_t *__restrict__ hidx, uint32_t *__restrict__ overlap,
                                                                 const void *__restrict__ nngrid, bool pal_in_smem) {
    extern __shared__ double s_mem[];
    // palettes too large for shared memory (48 B x K) are read from global memory
    const double *s_pal = pal_in_smem ? s_mem : pal, *s_palw = pal_in_smem ? s_mem + (size_t)K * 3 : palw;
    __shared__ double s_geom[6];
    __shared__ int s_grid_ok, s_grid_ng;
    if (pal_in_smem)
        for (int i = threadIdx.x; i < K * 3; i += blockDim.x) { s_mem[i] = pal[i]; s_mem[(size_t)K * 3 + i] = palw[i]; }
    if (threadIdx.x == 0) {
        s_grid_ok = 0;
        if (nngrid) {
            const PbGridGeom g = pb_grid_geom((const PbGridHdr *)nngrid);
            for (int d = 0; d < 3; d++) { s_geom[d] = g.lo[d]; s_geom[3 + d] = g.inv[d]; }
            s_grid_ok = g.ok;
            s_grid_ng = g.ng;
        }
    }
    __syncthreads();
    DitherGrid G{s_geom, nullptr, nullptr, 0};
    if (s_grid_ok) {
        G.ng = s_grid_ng;
        G.cnt = (const unsigned short *)((const char *)nngrid + 256);
        G.list = G.cnt + G.ng * G.ng * G.ng;
    }
    const int lane = threadIdx.x & 31;
    const size_t g = (size_t)blockIdx.x * DT_WARPS + (threadIdx.x >> 5);
    const size_t a = g * seg;
    if (a >= n) return;
    const size_t end = min(n, a + seg), start = a > warm ? a - warm : 0;
    DitherLane L;
    dither_lane_init(L, qweights, lane);
    uint32_t *ov = overlap + g * 16;
    dither_run(L, h0, h1, h2, start, end, s_pal, s_palw, K, lane, G, [&](size_t pos, int idx) {
        if (pos >= a) hidx[pos] = (uint32_t)idx;
        else if (pos + 16 >= a) ov[pos + 16 - a] = (uint32_t)idx;
    });
}

// ---- sub-warp speculation: 4 lanes per chain, 8 chains per warp --------------------------------------------
// In k_riemersma_spec a whole warp executes every instruction of ONE chain although only three lanes own an
// error queue and a candidate list has ~15 entries: ~180 warp instructions per pixel, and the kernel is bound
// by instruction issue, not by FP64 or memory.  Here a GROUP of four lanes runs a chain - lanes 0..2 own the
// R/G/B queues (the 16-tap sum is the same dependent chain of separately rounded products, riemersma.c:292-297),
// all four split the candidates of the exact search and reduce the argmin by two shuffles - so one warp
// instruction advances eight chains.  The queue lives in registers and is never shifted: the walk is unrolled
// sixteen pixels at a time and pixel k overwrites slot k (the oldest).  Pixels are staged through shared memory
// in batches of 16 (each lane fetches 4 consecutive pixels of every channel: 128-byte coalesced per group) with
// the next batch in registers; choices leave as one 64-byte store per group and batch.
constexpr int DS_GROUP = 4;
constexpr int DS_CHAINS = 32 / DS_GROUP;
constexpr int DS_WARPS = 4;
__constant__ double c_qw[16]; // queue weights (riemersma.c:360-373), oldest first

// squared distance to palette entry j.  The weighted palette is stored as K (x, y) pairs followed by K z values: one
// 16-byte and one 8-byte load per candidate, consecutive entries in consecutive bank quads / pairs (the first layout,
// x y z pad per entry, put every 16-byte load of a warp on four of the eight quads: ncu showed the kernel at 72 %
// of the shared-memory data pipe's peak)
__device__ __forceinline__ double dither_dist4(const double *__restrict__ pxy, const double *__restrict__ pz, int j, double x, double y,
                                               double z) {
    const double2 a = reinterpret_cast<const double2 *>(pxy)[j];
    const double dx = __dsub_rn(x, a.x), dy = __dsub_rn(y, a.y), dz = __dsub_rn(z, pz[j]);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

__device__ __forceinline__ int dither_nn4(double x, double y, double z, const double *__restrict__ s_palw, int K, int gl,
                                          const DitherGrid &G) {
    double bd = 0.0;
    int best = 0x7fffffff;
    const double *pz = s_palw + 2 * (size_t)K;
    const int cell = G.cnt ? pb_grid_cell(true, G.geom, G.geom + 3, G.ng, x, y, z) : -1; // uniform within the group
    if (cell >= 0) {
        const unsigned short *Lst = G.list + (size_t)cell * K;
        // the first four entries of this lane are requested together with the count (K >= 64: in bounds), so the
        // step pays one memory round trip, not two
        const int pre[4] = {Lst[gl], Lst[gl + DS_GROUP], Lst[gl + 2 * DS_GROUP], Lst[gl + 3 * DS_GROUP]};
        const int m = G.cnt[cell];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int t = gl + u * DS_GROUP;
            if (t < m) {
                const int j = pre[u];
                const double dd = dither_dist4(s_palw, pz, j, x, y, z);
                if (best == 0x7fffffff || dd < bd) { bd = dd; best = j; }
            }
        }
        for (int t = gl + 4 * DS_GROUP; t < m; t += DS_GROUP) {
            const int j = Lst[t];
            const double dd = dither_dist4(s_palw, pz, j, x, y, z);
            if (best == 0x7fffffff || dd < bd) { bd = dd; best = j; }
        }
    } else {
        for (int j = gl; j < K; j += DS_GROUP) {
            const double dd = dither_dist4(s_palw, pz, j, x, y, z);
            if (best == 0x7fffffff || dd < bd) { bd = dd; best = j; }
        }
    }
    // exact argmin over the group, lowest index on ties: squared distances are non-negative doubles, whose bit
    // patterns order like unsigned integers
    unsigned long long key = best == 0x7fffffff ? ~0ULL : (unsigned long long)__double_as_longlong(bd);
#pragma unroll
    for (int o = 1; o < DS_GROUP; o <<= 1) {
        const unsigned long long ok = __shfl_xor_sync(0xffffffffu, key, o);
        const int oj = __shfl_xor_sync(0xffffffffu, best, o);
        if (ok < key || (ok == key && oj < best)) { key = ok; best = oj; }
    }
    return best;
}

#ifndef PB_DS_MINB
#define PB_DS_MINB 4
#endif
template <bool PAL_SMEM> // palette in shared memory (LDS) or - too large - in global memory
__global__ void __launch_bounds__(DS_WARPS * 32, PB_DS_MINB) k_riemersma_spec4(const double *__restrict__ h0, const double *__restrict__ h1,
                                                                  const double *__restrict__ h2, size_t n, size_t seg,
                                                                  size_t warm, const double *__restrict__ pal,
                                                                  const double *__restrict__ palw, int K,
                                                                  uint32_t *__restrict__ hidx, uint32_t *__restrict__ overlap,
                                                                  const void *__restrict__ nngrid,
                                                                  size_t g_first, size_t g_end) {
    extern __shared__ __align__(16) double s_mem[]; // weighted palette: K (x, y) pairs, K z values; then [K][3] palette
    const double *s_palw = PAL_SMEM ? s_mem : palw, *s_pal = PAL_SMEM ? s_mem + (size_t)K * 3 : pal;
    __shared__ double s_geom[6];
    __shared__ int s_grid_ok, s_grid_ng;
    // (rows padded to 17: with 16 the eight chains' rows of one channel - and the three channels of a chain - start in
    // the same bank, and the per-step loads px[ch][k] / stores choice[k] of a warp serialise 24 / 4 ways)
    __shared__ double s_px[DS_WARPS][DS_CHAINS][3][17];
    __shared__ int s_choice[DS_WARPS][DS_CHAINS][17];
    if (PAL_SMEM) {
        for (int i = threadIdx.x; i < K * 3; i += blockDim.x) s_mem[i] = palw[i];
        for (int i = threadIdx.x; i < K * 3; i += blockDim.x) s_mem[(size_t)K * 3 + i] = pal[i];
    }
    if (threadIdx.x == 0) {
        s_grid_ok = 0;
        if (nngrid) {
            const PbGridGeom g = pb_grid_geom((const PbGridHdr *)nngrid);
            for (int d = 0; d < 3; d++) { s_geom[d] = g.lo[d]; s_geom[3 + d] = g.inv[d]; }
            s_grid_ok = g.ok;
            s_grid_ng = g.ng;
        }
    }
    __syncthreads();
    DitherGrid G{s_geom, nullptr, nullptr, 0};
    if (s_grid_ok) {
        G.ng = s_grid_ng;
        G.cnt = (const unsigned short *)((const char *)nngrid + 256);
        G.list = G.cnt + G.ng * G.ng * G.ng;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gl = lane & (DS_GROUP - 1), grp = lane / DS_GROUP;
    const int ch = gl < 3 ? gl : 0;
    const double cw = ch == 0 ? 0.51254268114958 : (ch == 1 ? 0.8234075540095561 : 0.2435159132377184);
    const size_t g = g_first + ((size_t)blockIdx.x * DS_WARPS + warp) * DS_CHAINS + grp; // (an image-sharded run: this rank's chains)
    const size_t a = g * seg;                       // first pixel this chain owns
    const bool live = g < g_end && a < n;
    const size_t end = live ? min(n, a + seg) : 0;
    size_t pos = live ? (a > warm ? a - warm : 0) : 0; // a - pos is a multiple of 16 (seg and warm are multiples of 128)
    double(*px)[17] = s_px[warp][grp];
    int *choice = s_choice[warp][grp];
    uint32_t *ov = overlap + g * 16;
    double q[16];
#pragma unroll
    for (int i = 0; i < 16; i++) q[i] = 0.0;
    double nx[3][4]; // the next batch: pixels pos + 4 * gl + {0..3} of every channel
    auto prefetch = [&](size_t at) {
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const size_t i = at + 4 * gl + e;
            const bool in = i < end;
            nx[0][e] = in ? h0[i] : 0.0;
            nx[1][e] = in ? h1[i] : 0.0;
            nx[2][e] = in ? h2[i] : 0.0;
        }
    };
    prefetch(pos);
    while (__any_sync(0xffffffffu, pos < end)) { // warp-uniform: finished groups idle through the rest
        __syncwarp();
#pragma unroll
        for (int e = 0; e < 4; e++) { px[0][4 * gl + e] = nx[0][e]; px[1][4 * gl + e] = nx[1][e]; px[2][4 * gl + e] = nx[2][e]; }
        __syncwarp();
        prefetch(pos + 16);
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const double P = px[ch][k];
            // riemersma.c:292-297: error = sum_i queue[i] * weight[i], oldest first; slot k holds the oldest entry
            // (summing the 15 older taps of the NEXT pixel while this pixel's candidate list is on its way from L2 was
            // tried: 40 more registers - spills or a CTA less per SM - cost more than the hidden DADD chain saves)
            double err = 0.0;
#pragma unroll
            for (int t = 0; t < 16; t++) err = __dadd_rn(err, __dmul_rn(q[(k + t) & 15], c_qw[t]));
            const double Cw = __dmul_rn(cw, __dadd_rn(P, err)); // :310-317, no clamping
            const double x = __shfl_sync(0xffffffffu, Cw, 0, DS_GROUP), y = __shfl_sync(0xffffffffu, Cw, 1, DS_GROUP),
                         z = __shfl_sync(0xffffffffu, Cw, 2, DS_GROUP);
            const int best = dither_nn4(x, y, z, s_palw, K, gl, G);
            q[k] = __dsub_rn(P, s_pal[3 * best + ch]); // :334-340: the newest entry takes the oldest one's slot
            if (gl == 0) choice[k] = best;
        }
        __syncwarp();
        if (pos < end) { // this lane's four choices of the batch
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const size_t i = pos + 4 * gl + e;
                if (i < end) {
                    if (i >= a) hidx[i] = (uint32_t)choice[4 * gl + e];
                    else if (i + 16 >= a) ov[i + 16 - a] = (uint32_t)choice[4 * gl + e];
                }
            }
        }
        pos += 16;
    }
}

// boundary g needs the sequential repair iff the 16 choices chain g made just before its segment differ from
// what chain g - 1 produced there
__global__ void k_riemersma_check(const uint32_t *__restrict__ hidx, const uint32_t *__restrict__ overlap, size_t nseg, size_t seg,
                                  unsigned char *__restrict__ flags) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nseg) return;
    bool differ = false;
    if (g >= 1) {
        const size_t a = g * seg;
#pragma unroll
        for (int i = 0; i < 16; i++) differ |= overlap[g * 16 + i] != hidx[a - 16 + i];
    }
    flags[g] = differ ? 1 : 0;
}

__global__ void __launch_bounds__(32) k_riemersma_repair(const double *__restrict__ h0, const double *__restrict__ h1,
                                                         const double *__restrict__ h2, size_t n, size_t seg,
                                                         const double *__restrict__ pal, const double *__restrict__ palw,
                                                         int K, const double *__restrict__ qweights,
                                                         uint32_t *__restrict__ hidx, const uint32_t *__restrict__ overlap,
                                                         unsigned long long *__restrict__ stats,
                                                         const void *__restrict__ nngrid, bool pal_in_smem,
                                                         const unsigned char *__restrict__ flags) {
    extern __shared__ double s_mem[];
    const double *s_pal = pal_in_smem ? s_mem : pal, *s_palw = pal_in_smem ? s_mem + (size_t)K * 3 : palw;
    __shared__ double s_geom[6];
    __shared__ int s_grid_ok, s_grid_ng;
    const int lane = threadIdx.x;
    if (pal_in_smem)
        for (int i = lane; i < K * 3; i += 32) { s_mem[i] = pal[i]; s_mem[(size_t)K * 3 + i] = palw[i]; }
    if (lane == 0) {
        s_grid_ok = 0;
        if (nngrid) {
            const PbGridGeom g = pb_grid_geom((const PbGridHdr *)nngrid);
            for (int d = 0; d < 3; d++) { s_geom[d] = g.lo[d]; s_geom[3 + d] = g.inv[d]; }
            s_grid_ok = g.ok;
            s_grid_ng = g.ng;
        }
    }
    __syncwarp();
    DitherGrid G{s_geom, nullptr, nullptr, 0};
    if (s_grid_ok) {
        G.ng = s_grid_ng;
        G.cnt = (const unsigned short *)((const char *)nngrid + 256);
        G.list = G.cnt + G.ng * G.ng * G.ng;
    }
    const size_t nseg = (n + seg - 1) / seg;
    DitherLane L;
    dither_lane_init(L, qweights, lane);
    const double *hp = L.ch == 0 ? h0 : (L.ch == 1 ? h1 : h2);
    unsigned long long repaired_segments = 0, repaired_pixels = 0;
    size_t g = 1;
    while (g < nseg) {
        if (flags) {
            // k_riemersma_check compared every boundary in parallel against the speculative choices.  A boundary it
            // passed stays passed: its 16-pixel window only changes if an earlier repair runs across it, and then
            // the walk below jumps past it anyway.  32 flags per step.
            const unsigned m = __ballot_sync(0xffffffffu, g + lane < nseg && flags[g + lane]);
            if (!m) { g += 32; continue; }
            g += __ffs(m) - 1;
        }
        const size_t a = g * seg;
        const bool differ = lane < 16 && overlap[g * 16 + lane] != hidx[a - 16 + lane];
        if (!__any_sync(0xffffffffu, differ)) { g++; continue; }
        repaired_segments++;
        // true state at a: queue entry i is P - palette[choice] of pixel a - 16 + i (riemersma.c:334-340)
#pragma unroll
        for (int i = 0; i < 16; i++) L.q[i] = __dsub_rn(hp[a - 16 + i], s_pal[3 * hidx[a - 16 + i] + L.ch]);
        // continue the true chain until it makes the same 16 consecutive choices as the speculative
        // run that owns those pixels (agreements must not straddle a segment boundary: the stored
        // choices on either side come from different speculative runs)
        size_t pos = a;
        int agree = 0;
        while (pos < n && agree < 16) {
            if (pos % seg == 0) agree = 0;
            const double P = hp[pos];
            const int best = dither_step(L, P, s_pal, s_palw, K, lane, G);
            const uint32_t old = hidx[pos];
            if ((uint32_t)best == old) agree++;
            else { agree = 0; if (lane == 0) hidx[pos] = (uint32_t)best; }
            __syncwarp();
            pos++;
            repaired_pixels++;
        }
        g = (pos - 1) / seg + 1; // the segment holding the last repaired pixel is true through its end
    }
    if (lane == 0 && stats) { stats[0] = repaired_segments; stats[1] = repaired_pixels; }
}

} // namespace

// The position of every pixel in the walk depends on (width, height) only: the table of the last image size is
// kept between calls (4 B per pixel; patolette_b200_release_cache() drops it).
namespace {
struct RankCache {
    int device = -1;
    size_t width = 0, height = 0;
    uint32_t *rank = nullptr;
} g_rank_cache;
} // namespace
void pb_dither_release_cache() {
    if (g_rank_cache.rank) pb_pool_free(g_rank_cache.rank);
    g_rank_cache = RankCache{};
}

static bool g_dither_grid = true; // patolette_b200_set_option "dither_grid"
void pb_dither_set_grid(bool on) { g_dither_grid = on; }
static bool g_dither_subwarp = true; // "dither_subwarp": 4 lanes per chain (k_riemersma_spec4) or a warp per chain
void pb_dither_set_subwarp(bool on) { g_dither_subwarp = on; }
static bool g_dither_tiles = true;    // "dither_tiles": tile-wise Hilbert permutation kernels (default) or the per-pixel scatter / gather
static bool g_dither_one_wave = true; // "dither_one_wave": segment length from the chip's resident chain capacity (default) or n / 2048
void pb_dither_set_tiles(bool on) { g_dither_tiles = on; }
void pb_dither_set_one_wave(bool on) { g_dither_one_wave = on; }

void pb_dither_riemersma(const double *const planes[3], size_t width, size_t height,
                         const std::vector<double> &pal_rm, unsigned long long *d_map, int sm_count,
                         cudaStream_t st, long *launches, const PbDitherShard *shard) {
    const size_t n = width * height;
    // image-sharded runs: the speculative chains are dealt out over the ranks (contiguous ranges of segments), the
    // choices and warm-up records all-gathered over NVLink; the boundary check / repair is replicated (it needs
    // every rank's choices and is sequential anyway) and every rank un-permutes the pixels it has to return
    const int world = shard && shard->world > 1 && g_dither_subwarp ? shard->world : 1, srank = world > 1 ? shard->rank : 0;
    const size_t out_first = shard ? shard->out_first : 0, out_count = shard ? shard->out_count : n;
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
    double *d_h[3] = {nullptr, nullptr, nullptr}, *d_pal = nullptr, *d_palw = nullptr, *d_qw = nullptr, *d_palw4 = nullptr;
    void *d_grid = nullptr;
    uint32_t *d_rank = nullptr, *d_hidx = nullptr, *d_overlap = nullptr;
    unsigned long long *d_stats = nullptr;
    unsigned char *d_flags = nullptr;
    auto cleanup = [&]() {
        pb_pool_free(d_flags);
        pb_pool_free(d_palw4);
        for (int j = 0; j < 3; j++) pb_pool_free(d_h[j]);
        pb_pool_free(d_pal); pb_pool_free(d_palw); pb_pool_free(d_qw); pb_pool_free(d_hidx); pb_pool_free(d_overlap); pb_pool_free(d_stats); pb_pool_free(d_grid);
    };
    try {
        for (int j = 0; j < 3; j++) d_h[j] = (double *)pb_pool_alloc(n * sizeof(double));
        d_pal = (double *)pb_pool_alloc(pal_rm.size() * sizeof(double));
        d_palw = (double *)pb_pool_alloc(pal_rm.size() * sizeof(double));
        d_qw = (double *)pb_pool_alloc(sizeof qw);
        int dev = 0;
        PB_CUDA_OK(cudaGetDevice(&dev));
        const bool rank_cached = g_rank_cache.rank && g_rank_cache.device == dev && g_rank_cache.width == width && g_rank_cache.height == height;
        if (!rank_cached) {
            pb_dither_release_cache();
            g_rank_cache.rank = (uint32_t *)pb_pool_alloc(n * sizeof(uint32_t));
            g_rank_cache.device = dev; g_rank_cache.width = width; g_rank_cache.height = height;
        }
        d_rank = g_rank_cache.rank;
        size_t seg = ((n / 2048 + 127) / 128) * 128;
        seg = seg < 1024 ? 1024 : (seg > 8192 ? 8192 : seg);
        if (g_dither_subwarp && g_dither_one_wave) {
            // one wave: as many chains as the chip (all ranks' chips) holds resident at once - a second, partly filled
            // wave costs as much as the first, and fewer, longer segments mean less warm-up work
            const size_t smem4 = (size_t)K * 6 * sizeof(double);
            const bool smem4_ok = smem4 <= PB_SMEM_PALETTE_LIMIT;
            if (smem4_ok && smem4 > 32 * 1024)
                PB_CUDA_OK(cudaFuncSetAttribute(k_riemersma_spec4<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4));
            int per_sm = 0;
            if (smem4_ok) PB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_riemersma_spec4<true>, DS_WARPS * 32, smem4));
            else PB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_riemersma_spec4<false>, DS_WARPS * 32, 0));
            if (per_sm < 1) per_sm = 1;
            const size_t capacity = (size_t)sm_count * per_sm * DS_WARPS * DS_CHAINS * (size_t)world;
            seg = (((n + capacity - 1) / capacity + 127) / 128) * 128;
            if (seg < 1024) seg = 1024;
        }
        const size_t warm = seg < 2048 ? seg : 2048;
        const size_t nseg = (n + seg - 1) / seg;
        const size_t per_rank = (nseg + (size_t)world - 1) / (size_t)world; // chains per rank
        d_hidx = (uint32_t *)pb_pool_alloc((per_rank * (size_t)world * seg + 64) * sizeof(uint32_t));
        PB_CUDA_OK(cudaMemcpyAsync(d_pal, pal_rm.data(), pal_rm.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        PB_CUDA_OK(cudaMemcpyAsync(d_palw, palw.data(), palw.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        PB_CUDA_OK(cudaMemcpyAsync(d_qw, qw, sizeof qw, cudaMemcpyHostToDevice, st));
        size_t want = (n + 255) / 256, cap = (size_t)sm_count * 8;
        const int grid = (int)(want < cap ? want : cap);
        pb_prof_next_bytes(4.0 * (double)n);
        if (!rank_cached)
        { PbProfScope _prof("k_hilbert_rank", st);
        k_hilbert_rank<<<grid, 256, 0, st>>>((uint32_t)width, (uint32_t)height, level, d_rank);
        }
        pb_prof_next_bytes(52.0 * (double)n); // 24 B read + 4 B rank + 24 B written
        { PbProfScope _prof("k_permute", st);
        if (g_dither_tiles)
            k_permute_tile<<<dim3((unsigned)((width + 31) / 32), (unsigned)((height + 31) / 32)), 256, 0, st>>>(
                planes[0], planes[1], planes[2], d_rank, (uint32_t)width, (uint32_t)height, d_h[0], d_h[1], d_h[2]);
        else
            k_permute<<<grid, 256, 0, st>>>(planes[0], planes[1], planes[2], d_rank, n, d_h[0], d_h[1], d_h[2]);
        }
        // segment / warm-up lengths: enough segments to occupy the chip, warm-up long enough that
        // almost every segment locks on before it starts (cold starts converge in ~150 pixels on
        // average, ~1600 worst observed)
        d_overlap = (uint32_t *)pb_pool_alloc((per_rank * (size_t)world + 1) * 16 * sizeof(uint32_t));
        d_stats = (unsigned long long *)pb_pool_alloc(2 * sizeof(unsigned long long));
        const bool pal_in_smem = (size_t)K * 6 * sizeof(double) <= PB_SMEM_PALETTE_LIMIT;
        const size_t smem = pal_in_smem ? (size_t)K * 6 * sizeof(double) : 0;
        if (smem > 32 * 1024) {
            PB_CUDA_OK(cudaFuncSetAttribute(k_riemersma_spec, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            PB_CUDA_OK(cudaFuncSetAttribute(k_riemersma_repair, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
        // candidate lists over the weighted query space: the pixels' box scaled by the sqrt-luma weights and
        // widened by a quarter of its range on every side (queries = pixel + diffused error; a query that
        // still falls outside searches all K entries)
        if (g_dither_grid && K >= 64 && K <= 4096 && n >= 16384) {
            d_grid = pb_pool_alloc(pb_nngrid_scratch_bytes(K));
            const double cw[3] = {0.51254268114958, 0.8234075540095561, 0.2435159132377184};
            const double *hp[3] = {d_h[0], d_h[1], d_h[2]};
            pb_launch_nngrid_build(hp, n, d_palw, K, d_grid, sm_count, st, cw, 0.25);
        }
        pb_prof_next_bytes(28.0 * (double)n); // 24 B of colours read + 4 B index written per pixel of the walk
        if (g_dither_subwarp) {
            PB_CUDA_OK(cudaMemcpyToSymbolAsync(c_qw, qw, sizeof qw, 0, cudaMemcpyHostToDevice, st)); // (per device; 128 B)
            // the weighted palette as K (x, y) pairs + K z values (one 16-byte and one 8-byte load per candidate)
            std::vector<double> palw4((size_t)K * 3, 0.0);
            for (int j = 0; j < K; j++) { palw4[2 * j] = palw[3 * j]; palw4[2 * j + 1] = palw[3 * j + 1]; palw4[2 * (size_t)K + j] = palw[3 * j + 2]; }
            d_palw4 = (double *)pb_pool_alloc(palw4.size() * sizeof(double));
            PB_CUDA_OK(cudaMemcpyAsync(d_palw4, palw4.data(), palw4.size() * sizeof(double), cudaMemcpyHostToDevice, st));
            PB_CUDA_OK(cudaStreamSynchronize(st)); // (palw4 is a local)
            const size_t smem4 = (size_t)K * 6 * sizeof(double);
            const bool smem4_ok = smem4 <= PB_SMEM_PALETTE_LIMIT;
            if (smem4_ok && smem4 > 32 * 1024)
                PB_CUDA_OK(cudaFuncSetAttribute(k_riemersma_spec4<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4));
            const size_t per_cta = (size_t)DS_WARPS * DS_CHAINS;
            const size_t g_first = (size_t)srank * per_rank, g_end = g_first + per_rank < nseg ? g_first + per_rank : nseg;
            const size_t mine = g_end > g_first ? g_end - g_first : 0;
            if (mine) {
                PbProfScope _prof("k_riemersma_spec", st);
                const unsigned gridx = (unsigned)((mine + per_cta - 1) / per_cta);
                if (smem4_ok)
                    k_riemersma_spec4<true><<<gridx, DS_WARPS * 32, smem4, st>>>(d_h[0], d_h[1], d_h[2], n, seg, warm, d_pal, d_palw4, K, d_hidx,
                                                                                 d_overlap, d_grid, g_first, g_end);
                else
                    k_riemersma_spec4<false><<<gridx, DS_WARPS * 32, 0, st>>>(d_h[0], d_h[1], d_h[2], n, seg, warm, d_pal, d_palw4, K, d_hidx,
                                                                              d_overlap, d_grid, g_first, g_end);
            }
            if (world > 1) { // every rank's choices and warm-up records, in place
                pb_nccl_allgather(d_hidx + g_first * seg, d_hidx, per_rank * seg * sizeof(uint32_t), st);
                pb_nccl_allgather(d_overlap + g_first * 16, d_overlap, per_rank * 16 * sizeof(uint32_t), st);
            }
        } else
        { PbProfScope _prof("k_riemersma_spec", st);
        k_riemersma_spec<<<(unsigned)((nseg + DT_WARPS - 1) / DT_WARPS), DT_WARPS * 32, smem, st>>>(
            d_h[0], d_h[1], d_h[2], n, seg, warm, d_pal, d_palw, K, d_qw, d_hidx, d_overlap, d_grid, pal_in_smem);
        }
        d_flags = (unsigned char *)pb_pool_alloc(nseg + 64);
        { PbProfScope _prof("k_riemersma_check", st, false);
        k_riemersma_check<<<(unsigned)((nseg + 255) / 256), 256, 0, st>>>(d_hidx, d_overlap, nseg, seg, d_flags);
        }
        { PbProfScope _prof("k_riemersma_repair", st);
        k_riemersma_repair<<<1, 32, smem, st>>>(d_h[0], d_h[1], d_h[2], n, seg, d_pal, d_palw, K, d_qw, d_hidx,
                                                d_overlap, d_stats, d_grid, pal_in_smem, d_flags);
        }
        pb_prof_next_bytes(16.0 * (double)out_count);
        if (out_count)
        { PbProfScope _prof("k_unpermute", st);
        if (g_dither_tiles)
            k_unpermute_tile<<<dim3((unsigned)((width + 31) / 32), (unsigned)((height + 31) / 32)), 256, 0, st>>>(
                d_rank, d_hidx, (uint32_t)width, (uint32_t)height, out_first, out_count, d_map);
        else
            k_unpermute<<<grid, 256, 0, st>>>(d_rank, d_hidx, out_first, out_count, d_map);
        }
        PB_CUDA_OK(cudaGetLastError());
        PB_CUDA_OK(cudaStreamSynchronize(st));
    } catch (...) {
        cleanup();
        pb_dither_release_cache(); // the table may not have been filled
        throw;
    }
    cleanup();
}
