// pb_parallel.cu - the order-independent (exactly parallelisable) kernels:
// projections + min/max, bucket ids, split objective, stable multi-class scatter,
// nearest-palette assignment.  All are HBM-streaming kernels over planar f64.
#include "pb_common.cuh"
#include "pb_kernels.h"
#include "pb_pipeline.h"
#include "pb_prof.h"

namespace {

constexpr int PAR_THREADS = 256;
constexpr int PAR_CHUNK = 4096; // pixels per CTA visit

// ---------------------------------------------------------------------------------
// Projections on the principal axis + exact min / max (sort.c:43-59).
// min/max are order-independent, so a block reduction + one ordered-uint atomic per
// CTA is bit-exact.  24 B read per pixel.
// ---------------------------------------------------------------------------------
__global__ void k_split_init(PbSplit *sp, int nseg) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nseg) {
        sp[i].mn_enc = ~0ULL;
        sp[i].mx_enc = 0ULL;
        sp[i].split = 0;
        sp[i].nleft = 0;
        sp[i].pad = PB_ROUTE_EXACT;
    }
}

__global__ void __launch_bounds__(PAR_THREADS) k_dots_minmax(PbPlanes b0, PbPlanes b1,
                                                             const PbSeg *__restrict__ segs,
                                                             const double *__restrict__ axes,
                                                             PbSplit *__restrict__ sp) {
    const int seg = blockIdx.y;
    const PbSeg sg = segs[seg];
    const PbPlanes &P = sg.buf ? b1 : b0;
    const double x0 = axes[seg * 3], x1 = axes[seg * 3 + 1], x2 = axes[seg * 3 + 2];
    const double *c0 = P.c[0] + sg.lo, *c1 = P.c[1] + sg.lo, *c2 = P.c[2] + sg.lo;
    unsigned long long mn = ~0ULL, mx = 0ULL;
    const uint32_t tail0 = sg.n - (sg.n & 3u); // first row of dgemv's scalar tail
    for (uint32_t base = blockIdx.x * PAR_CHUNK; base < sg.n; base += gridDim.x * PAR_CHUNK) {
        const uint32_t end = min(base + (uint32_t)PAR_CHUNK, sg.n);
        for (uint32_t i = base + threadIdx.x; i < end; i += PAR_THREADS) {
            const unsigned long long e = pb_ord_encode(pb_dgemv_row3(c0[i], c1[i], c2[i], x0, x1, x2, i >= tail0));
            mn = min(mn, e);
            mx = max(mx, e);
        }
    }
    for (int o = 16; o; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    __shared__ unsigned long long smn[PAR_THREADS / 32], smx[PAR_THREADS / 32];
    if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < PAR_THREADS / 32; w++) { mn = min(mn, smn[w]); mx = max(mx, smx[w]); }
        if (mn != ~0ULL || mx != 0ULL) {
            atomicMin(&sp[seg].mn_enc, mn);
            atomicMax(&sp[seg].mx_enc, mx);
        }
    }
}

// ---------------------------------------------------------------------------------
// Bucket ids (sort.c:61-87).  24 B read + 2 B written per pixel - plus a 32-byte interleaved copy
// (c0, c1, c2, w) of every pixel for the per-bucket sums that follow: they GATHER the members of a bucket,
// and one 32-byte sector per member instead of one per plane cuts their sector traffic by three.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(PAR_THREADS) k_buckets(PbPlanes b0, PbPlanes b1,
                                                         const PbSeg *__restrict__ segs,
                                                         const double *__restrict__ axes,
                                                         PbSplit *__restrict__ sp,
                                                         uint16_t *__restrict__ bucket, double *__restrict__ aos) {
    const int seg = blockIdx.y;
    const PbSeg sg = segs[seg];
    const PbPlanes &P = sg.buf ? b1 : b0;
    const double x0 = axes[seg * 3], x1 = axes[seg * 3 + 1], x2 = axes[seg * 3 + 2];
    const double *c0 = P.c[0] + sg.lo, *c1 = P.c[1] + sg.lo, *c2 = P.c[2] + sg.lo;
    uint16_t *bk = bucket + sg.lo;
    const double *cw = P.w ? P.w + sg.lo : nullptr;
    double2 *ao = reinterpret_cast<double2 *>(aos) + 2 * (size_t)sg.lo;
    const double mn = pb_ord_decode(sp[seg].mn_enc), mx = pb_ord_decode(sp[seg].mx_enc);
    const bool degenerate = __dsub_rn(mx, mn) < PB_DELTA;
    const double s = 1.0 / __dsub_rn(mx, mn);
    const uint32_t tail0 = sg.n - (sg.n & 3u);
    if (blockIdx.x == 0 && threadIdx.x == 0) sp[seg].degenerate = degenerate;
    for (uint32_t base = blockIdx.x * PAR_CHUNK; base < sg.n; base += gridDim.x * PAR_CHUNK) {
        const uint32_t end = min(base + (uint32_t)PAR_CHUNK, sg.n);
        for (uint32_t i = base + threadIdx.x; i < end; i += PAR_THREADS) {
            uint32_t b;
            const double v0 = c0[i], v1 = c1[i], v2 = c2[i];
            if (aos) { // (the gathering per-bucket sums; the default route sorts the payload itself instead)
                ao[2 * (size_t)i] = make_double2(v0, v1);
                ao[2 * (size_t)i + 1] = make_double2(v2, cw ? cw[i] : 1.0);
            }
            if (degenerate) {
                b = i % PB_BUCKETS; // sort.c:66-75 round-robin
            } else {
                const double dot = pb_dgemv_row3(v0, v1, v2, x0, x1, x2, i >= tail0);
                const double ratio = __dmul_rn(__dsub_rn(dot, mn), s);
                const unsigned long long q = (unsigned long long)__dmul_rn((double)PB_BUCKETS, ratio);
                b = q < PB_BUCKETS - 1 ? (uint32_t)q : PB_BUCKETS - 1;
            }
            bk[i] = (uint16_t)b;
        }
    }
}

// ---------------------------------------------------------------------------------
// Split objective over the 512 cumulative bucket sums (local.c:137-171).
// One CTA of 512 threads per cluster.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(PB_BUCKETS) k_split_select(const double *__restrict__ bsums,
                                                             const uint32_t *__restrict__ class_start,
                                                             PbSplit *__restrict__ sp) {
    __shared__ double cs[3][PB_BUCKETS];
    __shared__ unsigned long long sz[PB_BUCKETS];
    __shared__ double obj[PB_BUCKETS];
    __shared__ int loc[PB_BUCKETS];
    const int seg = blockIdx.x, i = threadIdx.x;
    const double *in = bsums + (size_t)seg * PB_BUCKETS * 4;
    sz[i] = (unsigned long long)__double_as_longlong(in[i * 4]);
    cs[0][i] = in[i * 4 + 1];
    cs[1][i] = in[i * 4 + 2];
    cs[2][i] = in[i * 4 + 3];
    __syncthreads();
    if (i < 3) { // local.c:137-141: sums[i] += sums[i-1], strictly left to right
        double run = cs[i][0];
        for (int b = 1; b < PB_BUCKETS; b++) { run = __dadd_rn(cs[i][b], run); cs[i][b] = run; }
    } else if (i == 3) { // local.c:144-146
        unsigned long long run = sz[0];
        for (int b = 1; b < PB_BUCKETS; b++) { run += sz[b]; sz[b] = run; }
    }
    __syncthreads();
    double o = 0.0;
    const double sl = (double)sz[i], sr = (double)(sz[PB_BUCKETS - 1] - sz[i]);
#pragma unroll
    for (int j = 0; j < 3; j++) { // local.c:149-168
        const double csl = cs[j][i];
        const double csr = __dsub_rn(cs[j][PB_BUCKETS - 1], csl);
        double v = 0.0;
        if (sl != 0) v = __dadd_rn(v, __ddiv_rn(__dmul_rn(csl, csl), sl));
        if (sr != 0) v = __dadd_rn(v, __ddiv_rn(__dmul_rn(csr, csr), sr));
        o = __dadd_rn(o, v);
    }
    obj[i] = o;
    loc[i] = i;
    __syncthreads();
    // vector.c:26-46 maxloc: strict '>' scanning upward == lowest index among the maxima
    for (int half = PB_BUCKETS / 2; half; half >>= 1) {
        if (i < half) {
            const double a = obj[i], b = obj[i + half];
            if (b > a || (b == a && loc[i + half] < loc[i])) { obj[i] = b; loc[i] = loc[i + half]; }
        }
        __syncthreads();
    }
    if (i == 0) {
        const uint32_t *cst = class_start + (size_t)seg * (PB_BUCKETS + 1);
        sp[seg].split = (uint32_t)loc[0];
        sp[seg].nleft = cst[loc[0] + 1] - cst[0];
    }
}

// ---------------------------------------------------------------------------------
// Stable multi-class ranking and scatter.
// A warp owns a tile of SC_TILE consecutive positions and walks it row by row (32
// positions per row) so that ranks follow position order.  Phase 1 counts classes per
// tile, phase 2 turns the [tile][class] table into exclusive offsets (class-major
// order, i.e. all of class 0 first), phase 3 re-walks each tile and uses
// __match_any_sync to rank same-class lanes inside a row against a running per-class
// counter in shared memory.
// ---------------------------------------------------------------------------------
constexpr int SC_ROWS = 64;
constexpr int SC_TILE = SC_ROWS * 32;

struct ClsCtx {
    int mode;
    const uint16_t *bucket;
    const PbSplit *sp;
    const uint8_t *lut;
};
__device__ __forceinline__ uint32_t cls_of(const ClsCtx &c, int seg, uint32_t pos) {
    const uint32_t b = c.bucket[pos];
    if (c.mode == PB_CLS_BUCKET) return b;
    if (c.mode == PB_CLS_SPLIT) return b <= c.sp[seg].split ? 0u : 1u;
    return c.lut[b];
}

__global__ void k_tile_hist(ClsCtx cc, int nclass, const PbSeg *__restrict__ segs, uint32_t tiles_cap,
                            uint32_t *__restrict__ tile_hist) {
    extern __shared__ uint32_t s_cnt[];
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int seg = blockIdx.y;
    const PbSeg sg = segs[seg];
    const uint32_t tile = blockIdx.x * warps + warp;
    uint32_t *cnt = s_cnt + warp * nclass;
    if ((size_t)tile * SC_TILE >= sg.n) return;
    for (int c = lane; c < nclass; c += 32) cnt[c] = 0;
    __syncwarp();
    const uint32_t beg = tile * SC_TILE, end = min(beg + (uint32_t)SC_TILE, sg.n);
    for (uint32_t i = beg + lane; i < end; i += 32) atomicAdd(&cnt[cls_of(cc, seg, sg.lo + i)], 1u);
    __syncwarp();
    uint32_t *out = tile_hist + ((size_t)sg.tbase + tile) * nclass;
    for (int c = lane; c < nclass; c += 32) out[c] = cnt[c];
}

// Exclusive scan of the [tile][class] table down its columns (class-major order of the output), with
// ROW-wise (coalesced) accesses: (a) column sums of every chunk of SC_CHUNK tiles, (b) scan of the
// chunk totals (one thread per class), (c) exclusive offsets inside each chunk.
constexpr int SC_CHUNK = 32;
__device__ __forceinline__ uint32_t chunk_row0(const PbSeg &sg, int seg) { return sg.tbase / SC_CHUNK + (uint32_t)seg; }

__global__ void __launch_bounds__(256) k_tile_scan_a(int nclass, const PbSeg *__restrict__ segs,
                                                     const uint32_t *__restrict__ tile_hist,
                                                     uint32_t *__restrict__ chunk_tot) {
    const int seg = blockIdx.y;
    const PbSeg sg = segs[seg];
    const uint32_t ntiles = (sg.n + SC_TILE - 1) / SC_TILE;
    const uint32_t t0 = blockIdx.x * SC_CHUNK;
    if (t0 >= ntiles) return;
    const uint32_t cnt = min((uint32_t)SC_CHUNK, ntiles - t0);
    const uint32_t *h = tile_hist + ((size_t)sg.tbase + t0) * nclass;
    uint32_t *out = chunk_tot + (size_t)(chunk_row0(sg, seg) + blockIdx.x) * nclass;
    for (int c = threadIdx.x; c < nclass; c += blockDim.x) {
        uint32_t v[SC_CHUNK]; // every load of the column in flight at once (one memory round trip, not 32)
#pragma unroll
        for (int t = 0; t < SC_CHUNK; t++) v[t] = t < (int)cnt ? h[(size_t)t * nclass + c] : 0u;
        uint32_t sum = 0;
#pragma unroll
        for (int t = 0; t < SC_CHUNK; t++) sum += v[t];
        out[c] = sum;
    }
}

// exclusive scan of the chunk totals down every class column: one WARP per column, every lane a contiguous run
// of chunks (independent loads), a warp scan of the 32 partial sums.  (The first version walked the chunks with
// one thread per column: thousands of dependent load-add-store steps for a 128 M pixel segment.)
__global__ void __launch_bounds__(256) k_tile_scan_b(int nclass, const PbSeg *__restrict__ segs,
                                                     uint32_t *__restrict__ chunk_tot,
                                                     uint32_t *__restrict__ class_tot) {
    const int seg = blockIdx.y;
    const PbSeg sg = segs[seg];
    const uint32_t ntiles = (sg.n + SC_TILE - 1) / SC_TILE, nchunks = (ntiles + SC_CHUNK - 1) / SC_CHUNK;
    uint32_t *ct = chunk_tot + (size_t)chunk_row0(sg, seg) * nclass;
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= nclass) return;
    const uint32_t per = (nchunks + 31) / 32;
    const uint32_t k0 = min((uint32_t)lane * per, nchunks), k1 = min(k0 + per, nchunks);
    uint32_t s = 0;
#pragma unroll 8
    for (uint32_t k = k0; k < k1; k++) s += ct[(size_t)k * nclass + c];
    uint32_t incl = s;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    uint32_t run = incl - s;
#pragma unroll 8
    for (uint32_t k = k0; k < k1; k++) {
        const uint32_t v = ct[(size_t)k * nclass + c];
        ct[(size_t)k * nclass + c] = run;
        run += v;
    }
    if (lane == 31) class_tot[(size_t)seg * (nclass + 1) + c] = incl;
}

__global__ void __launch_bounds__(256) k_tile_scan_c(int nclass, const PbSeg *__restrict__ segs,
                                                     uint32_t *__restrict__ tile_hist,
                                                     const uint32_t *__restrict__ chunk_tot) {
    const int seg = blockIdx.y;
    const PbSeg sg = segs[seg];
    const uint32_t ntiles = (sg.n + SC_TILE - 1) / SC_TILE;
    const uint32_t t0 = blockIdx.x * SC_CHUNK;
    if (t0 >= ntiles) return;
    const uint32_t cnt = min((uint32_t)SC_CHUNK, ntiles - t0);
    uint32_t *h = tile_hist + ((size_t)sg.tbase + t0) * nclass;
    const uint32_t *in = chunk_tot + (size_t)(chunk_row0(sg, seg) + blockIdx.x) * nclass;
    for (int c = threadIdx.x; c < nclass; c += blockDim.x) {
        uint32_t v[SC_CHUNK];
#pragma unroll
        for (int t = 0; t < SC_CHUNK; t++) v[t] = t < (int)cnt ? h[(size_t)t * nclass + c] : 0u;
        uint32_t run = in[c];
#pragma unroll
        for (int t = 0; t < SC_CHUNK; t++) {
            if (t < (int)cnt) h[(size_t)t * nclass + c] = run;
            run += v[t];
        }
    }
}

// class_start[seg][c] = seg.lo + sum_{c' < c} tot[c']   (in place over class_tot)
__global__ void k_class_start(int nclass, const PbSeg *__restrict__ segs, uint32_t *__restrict__ class_start) {
    const int seg = blockIdx.x;
    if (threadIdx.x != 0) return;
    uint32_t *cs = class_start + (size_t)seg * (nclass + 1);
    uint32_t run = segs[seg].lo;
    for (int c = 0; c < nclass; c++) { const uint32_t v = cs[c]; cs[c] = run; run += v; }
    cs[nclass] = run;
}

// MODE 0: ord[dst] = source position; 1: move the planar payload (planes + weight + index); 2: write the source
// pixel as an interleaved (c0, c1, c2, w) record at dst - the bucket-sorted copy the per-bucket sums stream through
// (one full 32-byte sector per pixel: no more traffic than the 4-byte index it replaces, and the sums that follow
// read sequentially instead of gathering)
enum { SCAT_ORD = 0, SCAT_PLANES = 1, SCAT_SORTED = 2 };
template <int MODE>
__global__ void k_scatter(ClsCtx cc, int nclass, const PbSeg *__restrict__ segs, uint32_t tiles_cap,
                          const uint32_t *__restrict__ tile_hist, const uint32_t *__restrict__ class_start,
                          uint32_t *__restrict__ ord, PbPlanes src0, PbPlanes src1, PbPlanes dst0, PbPlanes dst1,
                          bool src_is_identity, double *__restrict__ sorted) {
    extern __shared__ uint32_t s_cnt[];
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int seg = blockIdx.y;
    const PbSeg sg = segs[seg];
    const uint32_t tile = blockIdx.x * warps + warp;
    if ((size_t)tile * SC_TILE >= sg.n) return;
    uint32_t *cnt = s_cnt + warp * (nclass + 1);
    const uint32_t *hist = tile_hist + ((size_t)sg.tbase + tile) * nclass;
    const uint32_t *cst = class_start + (size_t)seg * (nclass + 1);
    for (int c = lane; c < nclass; c += 32) cnt[c] = hist[c] + cst[c];
    if (lane == 0) cnt[nclass] = 0;
    __syncwarp();
    const PbPlanes &S = sg.buf ? src1 : src0;
    const PbPlanes &D = sg.buf ? dst1 : dst0;
    const uint32_t beg = tile * SC_TILE, end = min(beg + (uint32_t)SC_TILE, sg.n);
    constexpr int RB = 8; // rows whose class ids are fetched together (one memory round trip per 8 rows)
    for (uint32_t row0 = beg; row0 < end; row0 += 32 * RB) {
        uint32_t cls[RB];
#pragma unroll
        for (int r = 0; r < RB; r++) {
            const uint32_t i = row0 + r * 32 + lane;
            cls[r] = i < end ? cls_of(cc, seg, sg.lo + i) : (uint32_t)nclass;
        }
#pragma unroll
        for (int r = 0; r < RB; r++) {
            const uint32_t row = row0 + r * 32;
            if (row >= end) break; // warp-uniform
            const uint32_t i = row + lane;
            const bool valid = i < end;
            const uint32_t pos = sg.lo + i;
            const uint32_t c = cls[r];
            const uint32_t peers = __match_any_sync(0xffffffffu, c);
            const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
            const uint32_t base = cnt[c];
            __syncwarp();
            if (rank == 0) cnt[c] = base + __popc(peers);
            __syncwarp();
            if (valid) {
                const uint32_t dst = base + rank;
                if (MODE == SCAT_PLANES) {
                    D.c[0][dst] = S.c[0][pos];
                    D.c[1][dst] = S.c[1][pos];
                    D.c[2][dst] = S.c[2][pos];
                    if (S.w) D.w[dst] = S.w[pos];
                    if (D.idx) D.idx[dst] = src_is_identity ? pos : S.idx[pos];
                } else if (MODE == SCAT_SORTED) {
                    double2 *o = reinterpret_cast<double2 *>(sorted) + 2 * (size_t)dst;
                    o[0] = make_double2(S.c[0][pos], S.c[1][pos]);
                    o[1] = make_double2(S.c[2][pos], S.w ? S.w[pos] : 1.0);
                } else {
                    ord[dst] = pos;
                }
            }
        }
    }
}


// Bucket-sorted interleaved copy, CTA-level: the warp-level walk above writes every pixel's 32-byte record on its
// own, and with 512 destinations per tile the partially written lines fall out of L2 before their neighbours
// arrive (measured 1.4 TB/s).  Here a CTA sorts a tile (one scatter tile, 2048 pixels) in shared memory first -
// payload staged planar, a stable rank per pixel (per-warp class histograms, a prefix over the warps, the
// __match_any walk) turned into a local permutation - and then writes the tile's pixels in SORTED order: the
// members of one bucket leave as one contiguous run (4 pixels = one 128-byte line on average).
#ifndef PB_SS_TILES
#define PB_SS_TILES 1 // scatter tiles per CTA: 1 -> 92 KB of shared memory, two CTAs per SM (one CTA's phases overlap the other's)
#endif
constexpr int SS_WARPS = 8, SS_THREADS = 32 * SS_WARPS, SS_TILE = PB_SS_TILES * SC_TILE, SS_ROWS = SS_TILE / 32 / SS_WARPS;
struct SsSmem {
    double pay[4][SS_TILE];            // planar payload of the tile (c0, c1, c2, w)
    uint32_t cnt[SS_WARPS][PB_BUCKETS]; // per-warp class counts, then running local slots
    uint32_t lstart[PB_BUCKETS + 1];   // first local slot of every class
    uint32_t gbase[PB_BUCKETS];        // first global position of the tile's members of every class
    uint16_t perm[SS_TILE];            // local source index of every local slot
    uint16_t cls[SS_TILE];             // class of every local slot
};
__global__ void __launch_bounds__(SS_THREADS) k_scatter_sorted_cta(const uint16_t *__restrict__ bucket, const PbSeg *__restrict__ segs,
                                                                  const uint32_t *__restrict__ tile_hist,
                                                                  const uint32_t *__restrict__ class_start, PbPlanes src0,
                                                                  PbPlanes src1, double *__restrict__ sorted) {
    extern __shared__ __align__(16) unsigned char ss_raw[];
    SsSmem &sm = *reinterpret_cast<SsSmem *>(ss_raw);
    const int seg = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const PbSeg sg = segs[seg];
    const uint32_t beg = blockIdx.x * SS_TILE;
    if (beg >= sg.n) return;
    const uint32_t cnt_t = min((uint32_t)SS_TILE, sg.n - beg);
    const PbPlanes &S = sg.buf ? src1 : src0;
    const size_t g0 = (size_t)sg.lo + beg;
    // global bases of the tile's first scatter tile (the second one continues them) + zeroed counters
    {
        const uint32_t *hist = tile_hist + ((size_t)sg.tbase + PB_SS_TILES * (size_t)blockIdx.x) * PB_BUCKETS;
        const uint32_t *cst = class_start + (size_t)seg * (PB_BUCKETS + 1);
        for (int c = tid; c < PB_BUCKETS; c += SS_THREADS) sm.gbase[c] = hist[c] + cst[c];
        for (int i = tid; i < SS_WARPS * PB_BUCKETS; i += SS_THREADS) (&sm.cnt[0][0])[i] = 0u;
    }
    // payload -> shared memory (coalesced), class ids of this warp's 16 rows -> registers
#pragma unroll 4
    for (uint32_t i = tid; i < cnt_t; i += SS_THREADS) {
        sm.pay[0][i] = S.c[0][g0 + i];
        sm.pay[1][i] = S.c[1][g0 + i];
        sm.pay[2][i] = S.c[2][g0 + i];
        sm.pay[3][i] = S.w ? S.w[g0 + i] : 1.0;
    }
    uint32_t cls[SS_ROWS];
#pragma unroll
    for (int r = 0; r < SS_ROWS; r++) {
        const uint32_t i = (warp * SS_ROWS + r) * 32 + lane;
        cls[r] = i < cnt_t ? (uint32_t)bucket[g0 + i] : 0xffffu;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SS_ROWS; r++)
        if (cls[r] != 0xffffu) atomicAdd(&sm.cnt[warp][cls[r]], 1u);
    __syncthreads();
    // per class: exclusive prefix over the warps (in place) and the class total
    for (int c = tid; c < PB_BUCKETS; c += SS_THREADS) {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < SS_WARPS; w++) { const uint32_t v = sm.cnt[w][c]; sm.cnt[w][c] = run; run += v; }
        sm.lstart[c + 1] = run; // totals, scanned below
    }
    if (tid == 0) sm.lstart[0] = 0;
    __syncthreads();
    if (warp == 0) { // inclusive scan of the 512 totals: 16 per lane + a warp scan
        uint32_t v[PB_BUCKETS / 32], s = 0;
#pragma unroll
        for (int k = 0; k < PB_BUCKETS / 32; k++) { s += sm.lstart[1 + lane * (PB_BUCKETS / 32) + k]; v[k] = s; }
        uint32_t incl = s;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        const uint32_t off = incl - s;
#pragma unroll
        for (int k = 0; k < PB_BUCKETS / 32; k++) sm.lstart[1 + lane * (PB_BUCKETS / 32) + k] = v[k] + off;
    }
    __syncthreads();
    for (int i = tid; i < SS_WARPS * PB_BUCKETS; i += SS_THREADS) (&sm.cnt[0][0])[i] += sm.lstart[i % PB_BUCKETS]; // running local slots
    __syncthreads();
    // the stable ranking walk: every warp over its own rows, in order
#pragma unroll
    for (int r = 0; r < SS_ROWS; r++) {
        const uint32_t i = (warp * SS_ROWS + r) * 32 + lane;
        if ((uint32_t)(warp * SS_ROWS + r) * 32 >= cnt_t) break; // warp-uniform
        const uint32_t c = cls[r];
        const uint32_t peers = __match_any_sync(0xffffffffu, c);
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        uint32_t base = 0;
        if (c != 0xffffu) base = sm.cnt[warp][c];
        __syncwarp();
        if (c != 0xffffu && rank == 0) sm.cnt[warp][c] = base + __popc(peers);
        __syncwarp();
        if (c != 0xffffu) {
            sm.perm[base + rank] = (uint16_t)i;
            sm.cls[base + rank] = (uint16_t)c;
        }
    }
    __syncthreads();
    // write out in sorted order: consecutive slots of one class -> consecutive global positions
    double2 *out = reinterpret_cast<double2 *>(sorted);
    for (uint32_t t = tid; t < cnt_t; t += SS_THREADS) {
        const uint32_t c = sm.cls[t], i = sm.perm[t];
        const size_t dst = (size_t)sm.gbase[c] + (t - sm.lstart[c]);
        out[2 * dst] = make_double2(sm.pay[0][i], sm.pay[1][i]);
        out[2 * dst + 1] = make_double2(sm.pay[2][i], sm.pay[3][i]);
    }
}

// Two-class payload scatter (the split partition, local.c:219-245): ranks come from two ballots per row of
// 32 positions and two warp-uniform running offsets in registers - no shared-memory counters, no
// __match_any, no __syncwarp - and four rows are loaded before any is stored (memory-level parallelism).
#ifndef PB_SC2_UNROLL
#define PB_SC2_UNROLL 4
#endif
constexpr int SC2_UNROLL = PB_SC2_UNROLL;
__global__ void __launch_bounds__(256) k_scatter2(const uint16_t *__restrict__ bucket, const PbSplit *__restrict__ sp,
                                                  const PbSeg *__restrict__ segs, const uint32_t *__restrict__ tile_hist,
                                                  const uint32_t *__restrict__ class_start, PbPlanes src0, PbPlanes src1,
                                                  PbPlanes dst0, PbPlanes dst1, bool src_is_identity) {
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int seg = blockIdx.y;
    const PbSeg sg = segs[seg];
    const uint32_t tile = blockIdx.x * warps + warp;
    if ((size_t)tile * SC_TILE >= sg.n) return;
    const uint32_t *hist = tile_hist + ((size_t)sg.tbase + tile) * 2;
    const uint32_t *cst = class_start + (size_t)seg * 3;
    uint32_t base0 = hist[0] + cst[0], base1 = hist[1] + cst[1];
    const uint32_t split = sp[seg].split;
    const PbPlanes &S = sg.buf ? src1 : src0;
    const PbPlanes &D = sg.buf ? dst1 : dst0;
    const bool has_w = S.w != nullptr, has_idx = D.idx != nullptr;
    const uint32_t beg = tile * SC_TILE, end = min(beg + (uint32_t)SC_TILE, sg.n);
    const uint32_t below = (1u << lane) - 1u;
    for (uint32_t row = beg; row < end; row += 32 * SC2_UNROLL) {
        bool valid[SC2_UNROLL], right[SC2_UNROLL];
        double v0[SC2_UNROLL], v1[SC2_UNROLL], v2[SC2_UNROLL], vw[SC2_UNROLL];
        uint32_t vi[SC2_UNROLL];
#pragma unroll
        for (int u = 0; u < SC2_UNROLL; u++) {
            const uint32_t i = row + u * 32 + lane;
            valid[u] = i < end;
            const uint32_t pos = sg.lo + i;
            right[u] = false; v0[u] = v1[u] = v2[u] = vw[u] = 0.0; vi[u] = 0;
            if (valid[u]) {
                right[u] = bucket[pos] > split;
                v0[u] = S.c[0][pos];
                v1[u] = S.c[1][pos];
                v2[u] = S.c[2][pos];
                if (has_w) vw[u] = S.w[pos];
                if (has_idx) vi[u] = src_is_identity ? pos : S.idx[pos];
            }
        }
#pragma unroll
        for (int u = 0; u < SC2_UNROLL; u++) {
            const uint32_t m1 = __ballot_sync(0xffffffffu, valid[u] && right[u]);
            const uint32_t m0 = __ballot_sync(0xffffffffu, valid[u] && !right[u]);
            if (valid[u]) {
                const uint32_t dst = right[u] ? base1 + __popc(m1 & below) : base0 + __popc(m0 & below);
                D.c[0][dst] = v0[u];
                D.c[1][dst] = v1[u];
                D.c[2][dst] = v2[u];
                if (has_w) D.w[dst] = vw[u];
                if (has_idx) D.idx[dst] = vi[u];
            }
            base0 += __popc(m0);
            base1 += __popc(m1);
        }
    }
}

__global__ void k_make_children(const PbSeg *__restrict__ segs, int nseg, const PbSplit *__restrict__ sp,
                                PbSeg *__restrict__ children) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nseg) return;
    const PbSeg sg = segs[i];
    const uint32_t nl = sp[i].nleft;
    // each child gets an ordered-sum block region as large as its parent's (its own size is only
    // known on the device): regions of a batch stay disjoint and total 2 * sum(parent blocks + 1)
    const uint32_t pb = (sg.n + 511u) / 512u + 1u;
    children[2 * i] = PbSeg{sg.lo, nl, sg.buf ^ 1u, 0u, 2u * sg.bbase, sp[i].pad};
    children[2 * i + 1] = PbSeg{sg.lo + nl, sg.n - nl, sg.buf ^ 1u, 0u, 2u * sg.bbase + pb, sp[i].pad};
}

// ---------------------------------------------------------------------------------
// Nearest palette entry (nearest.c:150-209 + the FLANN exact-1-NN contract): squared
// L2 in f64, dimensions summed in order 0,1,2, strict '<' over ascending palette index.
// Palette in shared memory (broadcast reads); 24 B read + 8 B written per pixel but
// 9*K FP64 operations, so FP64-pipe bound for K >= ~16.
// ---------------------------------------------------------------------------------
template <bool IN_SMEM> // palettes too large for shared memory are read from global memory (broadcast loads)
__global__ void __launch_bounds__(256) k_nearest(const double *__restrict__ c0, const double *__restrict__ c1,
                                                 const double *__restrict__ c2, size_t n,
                                                 const double *__restrict__ pal, int K,
                                                 unsigned long long *__restrict__ map) {
    extern __shared__ double s_pal_buf[];
    if (IN_SMEM) {
        for (int i = threadIdx.x; i < K * 3; i += blockDim.x) s_pal_buf[i] = pal[i];
        __syncthreads();
    }
    const double *s_pal = IN_SMEM ? s_pal_buf : pal;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double x = c0[i], y = c1[i], z = c2[i];
        double bd = 0.0;
        int best = 0;
#pragma unroll 4
        for (int j = 0; j < K; j++) {
            const double dx = __dsub_rn(x, s_pal[3 * j]), dy = __dsub_rn(y, s_pal[3 * j + 1]),
                         dz = __dsub_rn(z, s_pal[3 * j + 2]);
            const double dd = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            if (j == 0 || dd < bd) { bd = dd; best = j; }
        }
        map[i] = (unsigned long long)best;
    }
}

__global__ void k_labels(PbPlanes b0, PbPlanes b1, const PbSeg *__restrict__ segs, uint32_t *__restrict__ labels) {
    const int seg = blockIdx.y;
    const PbSeg sg = segs[seg];
    const PbPlanes &P = sg.buf ? b1 : b0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < sg.n; i += gridDim.x * blockDim.x)
        labels[P.idx[sg.lo + i]] = sg.pad; // the caller stores the cluster slot in PbSeg::pad
}

inline int blocks_for(uint32_t max_n, int sm_count) {
    uint32_t want = (max_n + PAR_CHUNK - 1) / PAR_CHUNK;
    uint32_t cap = (uint32_t)sm_count * 8;
    return (int)(want < cap ? (want ? want : 1) : cap);
}

} // namespace

void pb_launch_dots_minmax(const PbPlanes bufs[2], const PbSeg *d_segs, int nseg, uint32_t max_n,
                           const double *d_axes, PbSplit *d_split, int sm_count, cudaStream_t st) {
    if (nseg <= 0) return;
    { PbProfScope _prof("k_split_init", st, false);
    k_split_init<<<(nseg + 63) / 64, 64, 0, st>>>(d_split, nseg);
    }
    dim3 grid(blocks_for(max_n, sm_count), nseg);
    { PbProfScope _prof("k_dots_minmax", st);
    k_dots_minmax<<<grid, PAR_THREADS, 0, st>>>(bufs[0], bufs[1], d_segs, d_axes, d_split);
    }
    PB_CUDA_OK(cudaGetLastError());
}

void pb_launch_buckets(const PbPlanes bufs[2], const PbSeg *d_segs, int nseg, uint32_t max_n,
                       const double *d_axes, PbSplit *d_split, uint16_t *d_bucket, double *d_aos, int sm_count,
                       cudaStream_t st) {
    if (nseg <= 0) return;
    dim3 grid(blocks_for(max_n, sm_count), nseg);
    { PbProfScope _prof("k_buckets", st);
    k_buckets<<<grid, PAR_THREADS, 0, st>>>(bufs[0], bufs[1], d_segs, d_axes, d_split, d_bucket, d_aos);
    }
    PB_CUDA_OK(cudaGetLastError());
}

void pb_launch_split_select(const double *d_bucket_sums, const uint32_t *d_class_start, int nseg,
                            PbSplit *d_split, cudaStream_t st) {
    if (nseg <= 0) return;
    { PbProfScope _prof("k_split_select", st);
    k_split_select<<<nseg, PB_BUCKETS, 0, st>>>(d_bucket_sums, d_class_start, d_split);
    }
    PB_CUDA_OK(cudaGetLastError());
}

size_t pb_scatter_tiles(uint32_t n) { return ((size_t)n + SC_TILE - 1) / SC_TILE; }

static int scatter_warps(int nclass) { return nclass <= 16 ? 8 : (nclass <= 1024 ? 4 : (nclass <= 8192 ? 2 : 1)); }

size_t pb_scatter_chunk_offset(size_t total_tiles, int nclass) { return total_tiles * (size_t)nclass; }

size_t pb_scatter_table_words(size_t total_tiles, int nseg, int nclass) {
    return (total_tiles + total_tiles / SC_CHUNK + (size_t)nseg + 2) * (size_t)nclass + 64;
}

void pb_launch_class_rank(int cls_mode, int nclass, const PbSeg *d_segs, int nseg, uint32_t max_n,
                          size_t total_tiles, const uint16_t *d_bucket, const PbSplit *d_split,
                          const uint8_t *d_lut, uint32_t *d_tile_hist, uint32_t *d_class_start, cudaStream_t st) {
    if (nseg <= 0) return;
    const uint32_t tiles_cap = (uint32_t)pb_scatter_tiles(max_n ? max_n : 1);
    const int warps = scatter_warps(nclass);
    ClsCtx cc{cls_mode, d_bucket, d_split, d_lut};
    dim3 g1((tiles_cap + warps - 1) / warps, nseg);
    if ((size_t)warps * nclass * 4 > 32 * 1024)
        PB_CUDA_OK(cudaFuncSetAttribute(k_tile_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, warps * nclass * 4));
    { PbProfScope _prof("k_tile_hist", st, false);
    k_tile_hist<<<g1, warps * 32, (size_t)warps * nclass * 4, st>>>(cc, nclass, d_segs, tiles_cap, d_tile_hist);
    }
    // the chunk totals live behind the tile table (pb_scatter_table_words reserves the room)
    uint32_t *d_chunk_tot = d_tile_hist + pb_scatter_chunk_offset(total_tiles, nclass);
    dim3 g2((tiles_cap + SC_CHUNK - 1) / SC_CHUNK, nseg);
    { PbProfScope _prof("k_tile_scan", st, false);
    k_tile_scan_a<<<g2, 256, 0, st>>>(nclass, d_segs, d_tile_hist, d_chunk_tot);
    }
    { PbProfScope _prof("k_tile_scan", st, false);
    k_tile_scan_b<<<dim3((nclass + 7) / 8, nseg), 256, 0, st>>>(nclass, d_segs, d_chunk_tot, d_class_start);
    }
    { PbProfScope _prof("k_tile_scan", st, false);
    k_tile_scan_c<<<g2, 256, 0, st>>>(nclass, d_segs, d_tile_hist, d_chunk_tot);
    }
    { PbProfScope _prof("k_class_start", st, false);
    k_class_start<<<nseg, 32, 0, st>>>(nclass, d_segs, d_class_start);
    }
    PB_CUDA_OK(cudaGetLastError());
}

void pb_launch_scatter_ord(int cls_mode, int nclass, const PbSeg *d_segs, int nseg, uint32_t max_n,
                           const uint16_t *d_bucket, const PbSplit *d_split, const uint8_t *d_lut,
                           const uint32_t *d_tile_hist, const uint32_t *d_class_start, uint32_t *d_ord,
                           cudaStream_t st) {
    if (nseg <= 0) return;
    const uint32_t tiles_cap = (uint32_t)pb_scatter_tiles(max_n ? max_n : 1);
    const int warps = scatter_warps(nclass);
    ClsCtx cc{cls_mode, d_bucket, d_split, d_lut};
    dim3 g((tiles_cap + warps - 1) / warps, nseg);
    PbPlanes none{};
    if ((size_t)warps * (nclass + 1) * 4 > 32 * 1024)
        PB_CUDA_OK(cudaFuncSetAttribute(k_scatter<SCAT_ORD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        warps * (nclass + 1) * 4));
    { PbProfScope _prof("k_scatter_ord", st);
    k_scatter<SCAT_ORD><<<g, warps * 32, (size_t)warps * (nclass + 1) * 4, st>>>(
        cc, nclass, d_segs, tiles_cap, d_tile_hist, d_class_start, d_ord, none, none, none, none, false, nullptr);
    }
    PB_CUDA_OK(cudaGetLastError());
}

static bool g_scatter_cta = true; // "scatter_cta": CTA-level sort in shared memory for the bucket-sorted copy (default) or the warp-level walk
void pb_scatter_set_cta(bool on) { g_scatter_cta = on; }

void pb_launch_scatter_sorted(int cls_mode, int nclass, const PbPlanes src[2], const PbSeg *d_segs, int nseg, uint32_t max_n,
                              const uint16_t *d_bucket, const PbSplit *d_split, const uint8_t *d_lut,
                              const uint32_t *d_tile_hist, const uint32_t *d_class_start, double *d_sorted, cudaStream_t st) {
    if (nseg <= 0) return;
    const uint32_t tiles_cap = (uint32_t)pb_scatter_tiles(max_n ? max_n : 1);
    if (cls_mode == PB_CLS_BUCKET && nclass == PB_BUCKETS && g_scatter_cta) {
        PB_CUDA_OK(cudaFuncSetAttribute(k_scatter_sorted_cta, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SsSmem)));
        PbProfScope _prof("k_scatter_sorted", st);
        k_scatter_sorted_cta<<<dim3((tiles_cap + PB_SS_TILES - 1) / PB_SS_TILES, nseg), SS_THREADS, sizeof(SsSmem), st>>>(d_bucket, d_segs, d_tile_hist, d_class_start,
                                                                                          src[0], src[1], d_sorted);
        PB_CUDA_OK(cudaGetLastError());
        return;
    }
    const int warps = scatter_warps(nclass);
    ClsCtx cc{cls_mode, d_bucket, d_split, d_lut};
    dim3 g((tiles_cap + warps - 1) / warps, nseg);
    PbPlanes none{};
    if ((size_t)warps * (nclass + 1) * 4 > 32 * 1024)
        PB_CUDA_OK(cudaFuncSetAttribute(k_scatter<SCAT_SORTED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        warps * (nclass + 1) * 4));
    { PbProfScope _prof("k_scatter_sorted", st);
    k_scatter<SCAT_SORTED><<<g, warps * 32, (size_t)warps * (nclass + 1) * 4, st>>>(
        cc, nclass, d_segs, tiles_cap, d_tile_hist, d_class_start, nullptr, src[0], src[1], none, none, false, d_sorted);
    }
    PB_CUDA_OK(cudaGetLastError());
}

void pb_launch_scatter_payload(int cls_mode, int nclass, const PbPlanes src[2], const PbPlanes dst[2],
                               bool src_is_identity, const PbSeg *d_segs, int nseg, uint32_t max_n,
                               const uint16_t *d_bucket, const PbSplit *d_split, const uint8_t *d_lut,
                               const uint32_t *d_tile_hist, const uint32_t *d_class_start, cudaStream_t st) {
    if (nseg <= 0) return;
    const uint32_t tiles_cap = (uint32_t)pb_scatter_tiles(max_n ? max_n : 1);
    const int warps = scatter_warps(nclass);
    ClsCtx cc{cls_mode, d_bucket, d_split, d_lut};
    dim3 g((tiles_cap + warps - 1) / warps, nseg);
    if (cls_mode == PB_CLS_SPLIT && nclass == 2) {
        PbProfScope _prof("k_scatter2", st);
        dim3 g2((tiles_cap + 7) / 8, nseg);
        k_scatter2<<<g2, 256, 0, st>>>(d_bucket, d_split, d_segs, d_tile_hist, d_class_start, src[0], src[1], dst[0], dst[1],
                                      src_is_identity);
    } else {
        PbProfScope _prof("k_scatter_payload", st);
        k_scatter<SCAT_PLANES><<<g, warps * 32, (size_t)warps * (nclass + 1) * 4, st>>>(
            cc, nclass, d_segs, tiles_cap, d_tile_hist, d_class_start, nullptr, src[0], src[1], dst[0], dst[1],
            src_is_identity, nullptr);
    }
    PB_CUDA_OK(cudaGetLastError());
}

void pb_launch_make_children(const PbSeg *d_segs, int nseg, const PbSplit *d_split, PbSeg *d_children,
                             cudaStream_t st) {
    if (nseg <= 0) return;
    { PbProfScope _prof("k_make_children", st, false);
    k_make_children<<<(nseg + 63) / 64, 64, 0, st>>>(d_segs, nseg, d_split, d_children);
    }
    PB_CUDA_OK(cudaGetLastError());
}

void pb_launch_nearest(const double *const planes[3], size_t n, const double *d_palette_rm, int K,
                       unsigned long long *d_map, int sm_count, cudaStream_t st) {
    if (n == 0) return;
    size_t want = (n + 255) / 256, cap = (size_t)sm_count * 8;
    const int grid = (int)(want < cap ? want : cap);
    const size_t smem = (size_t)K * 3 * sizeof(double);
    const bool in_smem = smem <= PB_SMEM_PALETTE_LIMIT;
    if (in_smem && smem > 32 * 1024)
        PB_CUDA_OK(cudaFuncSetAttribute(k_nearest<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    { PbProfScope _prof("k_nearest", st);
    if (in_smem) k_nearest<true><<<grid, 256, smem, st>>>(planes[0], planes[1], planes[2], n, d_palette_rm, K, d_map);
    else k_nearest<false><<<grid, 256, 0, st>>>(planes[0], planes[1], planes[2], n, d_palette_rm, K, d_map);
    }
    PB_CUDA_OK(cudaGetLastError());
}

void pb_launch_labels(const PbPlanes bufs[2], const PbSeg *d_segs, int nseg, uint32_t max_n,
                      uint32_t *d_labels, int sm_count, cudaStream_t st) {
    if (nseg <= 0) return;
    dim3 grid(blocks_for(max_n, sm_count), nseg);
    { PbProfScope _prof("k_labels", st);
    k_labels<<<grid, 256, 0, st>>>(bufs[0], bufs[1], d_segs, d_labels);
    }
    PB_CUDA_OK(cudaGetLastError());
}
