// pb_chain.cu - ordered ("chain") per-bucket sums.  (The per-cluster mean / covariance /
// distortion passes live in pb_ordered.cu, which adds the binade-speculative fast path.)
//
// Every statistic the reference derives from a cluster - weighted mean
// (array/matrix2D.c:200-233), centred covariance (math/pca.c:84-97), distortion
// (quantize/cluster.c:135-148), per-bucket moments (quantize/local.c:124-134,
// quantize/cells.c:78-116) - is a plain left-to-right f64 accumulation in ascending
// pixel order.  Floating-point addition is not associative, so a tree / warp-shuffle /
// atomic reduction gives different low-order bits (it is MORE accurate, but the
// parity bar is bit-exact, and a flipped low bit can flip a bucket decision three
// stages later).  These kernels therefore keep each accumulation a true sequential
// chain and find their parallelism ACROSS chains:
//   * one warp per cluster; lanes 0..C-1 each own one chain (C = 4 for the mean pass,
//     7 for the centred pass, 4/10 for the per-bucket passes);
//   * all 32 lanes cooperate on the memory side: coalesced (or gathered) loads of a
//     tile into shared memory, double-buffered through registers so the next tile's
//     HBM latency hides behind the current tile's dependent-add chain;
//   * chain lanes then walk the tile reading operands as shared-memory broadcasts -
//     the critical path per element is exactly one dependent DADD.
// Cost model: ~8-10 SM cycles per element per cluster (DADD latency), independent of
// the chain count, so the big early clusters are latency-bound (see DESIGN.md for the
// planned binade-speculative block summaries that lift this to HBM speed while
// staying bit-identical).
#include "pb_common.cuh"
#include "pb_kernels.h"
#include "pb_prof.h"

namespace {

// ------------------------------------------------------------------------------------
// Per-bucket ordered sums.  The members of bucket b of a cluster, in ascending pixel
// order, are ord[class_start[b] .. class_start[b+1]) (stable bucket sort, pb_scatter.cu).
// One warp per (cluster, bucket): lanes gather a tile of members, chain lanes add.
// ------------------------------------------------------------------------------------
constexpr int BK_TILE = 128;           // members staged per step (4 per lane)
constexpr int BK_STRIDE = BK_TILE + 2;
constexpr int BK_WARPS = 4;
constexpr int BK_PER = BK_TILE / 32;

// The chain lanes run ONE uniform loop `acc = acc + term[lane][e]`: every per-element term is formed by the
// gathering lanes (in parallel) before it is staged, so the critical path per element is a shared-memory
// broadcast read + one dependent DADD.  The gather of tile t+1 is issued into registers before the chain
// over tile t starts, so its latency hides behind the chain.
//   LQ (local.c:124-134): term 0 = w (bucket "size"), terms 1..3 = c_j * w.
//     The reference accumulates the size as size_t += double (local.c:133), i.e.
//     size = trunc((double)size + w) at every step.  Unweighted that is the member count (exact), and
//     c_j * 1.0 == c_j: three plain chains.  Weighted, every chain lane runs the trunc-select stream.
template <bool WEIGHTED>
__global__ void __launch_bounds__(BK_WARPS * 32) k_bucket_chains_lq(PbPlanes b0, PbPlanes b1,
                                                                   const PbSeg *__restrict__ segs, int nseg,
                                                                   const uint32_t *__restrict__ ord,
                                                                   const uint32_t *__restrict__ class_start,
                                                                   double *__restrict__ out) {
    constexpr int NT = WEIGHTED ? 4 : 3;
    __shared__ double sm_all[BK_WARPS][NT][BK_STRIDE];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gw = blockIdx.x * BK_WARPS + warp;
    if (gw >= nseg * PB_BUCKETS) return;
    const int seg = gw / PB_BUCKETS, b = gw % PB_BUCKETS;
    double(*sm)[BK_STRIDE] = sm_all[warp];
    const PbSeg sg = segs[seg];
    const PbPlanes &P = sg.buf ? b1 : b0;
    const uint32_t *cs = class_start + (size_t)seg * (PB_BUCKETS + 1);
    const uint32_t beg = cs[b], end = cs[b + 1];
    double acc = 0.0;
    double g[BK_PER][NT];
    auto gather = [&](uint32_t i0) {
#pragma unroll
        for (int q = 0; q < BK_PER; q++) {
            const uint32_t i = i0 + q * 32 + lane;
#pragma unroll
            for (int t = 0; t < NT; t++) g[q][t] = 0.0;
            if (i < end) {
                const uint32_t p = ord[i];
                if (WEIGHTED) {
                    const double w = P.w[p];
                    g[q][0] = w;
                    g[q][1] = __dmul_rn(P.c[0][p], w);
                    g[q][2] = __dmul_rn(P.c[1][p], w);
                    g[q][3] = __dmul_rn(P.c[2][p], w);
                } else {
                    g[q][0] = P.c[0][p];
                    g[q][1] = P.c[1][p];
                    g[q][2] = P.c[2][p];
                }
            }
        }
    };
    if (beg < end) gather(beg);
    for (uint32_t i0 = beg; i0 < end; i0 += BK_TILE) {
        const uint32_t cnt = min((uint32_t)BK_TILE, end - i0);
#pragma unroll
        for (int q = 0; q < BK_PER; q++)
#pragma unroll
            for (int t = 0; t < NT; t++) sm[t][q * 32 + lane] = g[q][t];
        __syncwarp();
        if (i0 + BK_TILE < end) gather(i0 + BK_TILE); // in flight during the chain below
        if (lane < NT) {
            const double *vp = sm[lane];
            if (WEIGHTED) {
#pragma unroll 8
                for (uint32_t e = 0; e < cnt; e++) {
                    const double t = __dadd_rn(acc, vp[e]);
                    acc = lane == 0 ? trunc(t) : t;
                }
            } else {
#pragma unroll 16
                for (uint32_t e = 0; e < cnt; e++) acc = __dadd_rn(acc, vp[e]);
            }
        }
        __syncwarp();
    }
    double *o = out + ((size_t)seg * PB_BUCKETS + b) * 4;
    if (WEIGHTED) {
        if (lane == 0) o[0] = __longlong_as_double((long long)(unsigned long long)acc);
        else if (lane < 4) o[lane] = acc;
    } else {
        if (lane == 0) o[0] = __longlong_as_double((long long)(unsigned long long)(end - beg));
        if (lane < 3) o[1 + lane] = acc;
    }
}

// GQ cell moments (cells.c:78-116), unweighted, over the whole image:
// terms 0..2 = c_j ; 3 = (cx^2 + cy^2) + cz^2 ; 4..9 = c_r * c_s for (r,s) = (0,0)(0,1)(1,1)(0,2)(1,2)(2,2).
__global__ void __launch_bounds__(BK_WARPS * 32) k_bucket_chains_gq(PbPlanes src, const uint32_t *__restrict__ ord,
                                                                   const uint32_t *__restrict__ class_start,
                                                                   double *__restrict__ out) {
    constexpr int NT = 10;
    constexpr int GT = 64, GPER = GT / 32, GSTRIDE = GT + 2; // smaller tile: 10 term rows per warp
    __shared__ double sm_all[BK_WARPS][NT][GSTRIDE];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * BK_WARPS + warp;
    if (b >= PB_BUCKETS) return;
    double(*sm)[GSTRIDE] = sm_all[warp];
    const uint32_t beg = class_start[b], end = class_start[b + 1];
    double acc = 0.0;
    double g[GPER][3];
    auto gather = [&](uint32_t i0) {
#pragma unroll
        for (int q = 0; q < GPER; q++) {
            const uint32_t i = i0 + q * 32 + lane;
            g[q][0] = g[q][1] = g[q][2] = 0.0;
            if (i < end) {
                const uint32_t p = ord[i];
                g[q][0] = src.c[0][p];
                g[q][1] = src.c[1][p];
                g[q][2] = src.c[2][p];
            }
        }
    };
    if (beg < end) gather(beg);
    for (uint32_t i0 = beg; i0 < end; i0 += GT) {
        const uint32_t cnt = min((uint32_t)GT, end - i0);
#pragma unroll
        for (int q = 0; q < GPER; q++) {
            const int e = q * 32 + lane;
            const double x = g[q][0], y = g[q][1], z = g[q][2];
            sm[0][e] = x;
            sm[1][e] = y;
            sm[2][e] = z;
            sm[3][e] = __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));
            sm[4][e] = __dmul_rn(x, x);
            sm[5][e] = __dmul_rn(x, y);
            sm[6][e] = __dmul_rn(y, y);
            sm[7][e] = __dmul_rn(x, z);
            sm[8][e] = __dmul_rn(y, z);
            sm[9][e] = __dmul_rn(z, z);
        }
        __syncwarp();
        if (i0 + GT < end) gather(i0 + GT);
        if (lane < NT) {
            const double *vp = sm[lane];
#pragma unroll 16
            for (uint32_t e = 0; e < cnt; e++) acc = __dadd_rn(acc, vp[e]);
        }
        __syncwarp();
    }
    if (lane < 10) out[(size_t)b * 10 + lane] = acc;
}

} // namespace

void pb_launch_bucket_chains_lq(const PbPlanes bufs[2], const PbSeg *d_segs, int nseg, bool weighted,
                                const uint32_t *d_ord, const uint32_t *d_class_start, double *d_out,
                                cudaStream_t st) {
    if (nseg <= 0) return;
    const int grid = (nseg * PB_BUCKETS + BK_WARPS - 1) / BK_WARPS;
    if (weighted)
        { PbProfScope _prof("k_bucket_chains_lq", st);
        k_bucket_chains_lq<true><<<grid, BK_WARPS * 32, 0, st>>>(bufs[0], bufs[1], d_segs, nseg, d_ord, d_class_start, d_out);
        }
    else
        { PbProfScope _prof("k_bucket_chains_lq", st);
        k_bucket_chains_lq<false><<<grid, BK_WARPS * 32, 0, st>>>(bufs[0], bufs[1], d_segs, nseg, d_ord, d_class_start, d_out);
        }
    PB_CUDA_OK(cudaGetLastError());
}

void pb_launch_bucket_chains_gq(const PbPlanes &src, const uint32_t *d_ord, const uint32_t *d_class_start,
                                double *d_out, cudaStream_t st) {
    { PbProfScope _prof("k_bucket_chains_gq", st);
    k_bucket_chains_gq<<<PB_BUCKETS / BK_WARPS, BK_WARPS * 32, 0, st>>>(src, d_ord, d_class_start, d_out);
    }
    PB_CUDA_OK(cudaGetLastError());
}
