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
constexpr int BK_TILE = 64;
constexpr int BK_STRIDE = BK_TILE + 2;
constexpr int BK_WARPS = 4;

template <bool WEIGHTED>
__global__ void __launch_bounds__(BK_WARPS * 32) k_bucket_chains_lq(PbPlanes b0, PbPlanes b1,
                                                                   const PbSeg *__restrict__ segs, int nseg,
                                                                   const uint32_t *__restrict__ ord,
                                                                   const uint32_t *__restrict__ class_start,
                                                                   double *__restrict__ out) {
    __shared__ double sm_all[BK_WARPS][4][BK_STRIDE];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gw = blockIdx.x * BK_WARPS + warp;
    if (gw >= nseg * PB_BUCKETS) return;
    const int seg = gw / PB_BUCKETS, b = gw % PB_BUCKETS;
    double(*sm)[BK_STRIDE] = sm_all[warp];
    const PbSeg sg = segs[seg];
    const PbPlanes &P = sg.buf ? b1 : b0;
    const uint32_t *cs = class_start + (size_t)seg * (PB_BUCKETS + 1);
    const uint32_t beg = cs[b], end = cs[b + 1];
    // lane 0: bucket "size" - the reference accumulates a size_t with += double
    // (local.c:133), i.e. size = trunc((double)size + w) at every step; kept here as an
    // integer-valued double so all four chain lanes run the same instruction stream.
    // lanes 1..3: sum c_j * w (local.c:130-132).
    double acc = 0.0;
    for (uint32_t i0 = beg; i0 < end; i0 += BK_TILE) {
        const uint32_t cnt = min((uint32_t)BK_TILE, end - i0);
#pragma unroll
        for (int q = 0; q < BK_TILE / 32; q++) {
            const uint32_t e = q * 32 + lane;
            if (e < cnt) {
                const uint32_t p = ord[i0 + e];
                sm[0][e] = WEIGHTED ? P.w[p] : 1.0;
                sm[1][e] = P.c[0][p];
                sm[2][e] = P.c[1][p];
                sm[3][e] = P.c[2][p];
            }
        }
        __syncwarp();
        if (lane < 4) {
            const double *vp = sm[lane];
#pragma unroll 8
            for (uint32_t e = 0; e < cnt; e++) {
                const double w = sm[0][e];
                const double t = __dadd_rn(acc, lane == 0 ? w : __dmul_rn(vp[e], w));
                acc = lane == 0 ? trunc(t) : t;
            }
        }
        __syncwarp();
    }
    double *o = out + ((size_t)seg * PB_BUCKETS + b) * 4;
    if (lane == 0) o[0] = __longlong_as_double((long long)(unsigned long long)acc);
    else if (lane < 4) o[lane] = acc;
}

// GQ cell moments (cells.c:78-116), unweighted, over the whole image.
__global__ void __launch_bounds__(BK_WARPS * 32) k_bucket_chains_gq(PbPlanes src, const uint32_t *__restrict__ ord,
                                                                   const uint32_t *__restrict__ class_start,
                                                                   double *__restrict__ out) {
    __shared__ double sm_all[BK_WARPS][3][BK_STRIDE];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * BK_WARPS + warp;
    if (b >= PB_BUCKETS) return;
    double(*sm)[BK_STRIDE] = sm_all[warp];
    const uint32_t beg = class_start[b], end = class_start[b + 1];
    // lanes 0..2: sum c_j ; lane 3: sum (cx^2 + cy^2) + cz^2 ; lanes 4..9: sum c_r * c_s
    const int r = (lane == 4 || lane == 5 || lane == 7) ? 0 : ((lane == 6 || lane == 8) ? 1 : 2);
    const int s = lane == 4 ? 0 : ((lane == 5 || lane == 6) ? 1 : 2);
    double acc = 0.0;
    for (uint32_t i0 = beg; i0 < end; i0 += BK_TILE) {
        const uint32_t cnt = min((uint32_t)BK_TILE, end - i0);
#pragma unroll
        for (int q = 0; q < BK_TILE / 32; q++) {
            const uint32_t e = q * 32 + lane;
            if (e < cnt) {
                const uint32_t p = ord[i0 + e];
                sm[0][e] = src.c[0][p];
                sm[1][e] = src.c[1][p];
                sm[2][e] = src.c[2][p];
            }
        }
        __syncwarp();
        if (lane < 3) {
            const double *vp = sm[lane];
#pragma unroll 8
            for (uint32_t e = 0; e < cnt; e++) acc = __dadd_rn(acc, vp[e]);
        } else if (lane == 3) {
#pragma unroll 4
            for (uint32_t e = 0; e < cnt; e++) {
                const double x = sm[0][e], y = sm[1][e], z = sm[2][e];
                acc = __dadd_rn(acc, __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
            }
        } else if (lane < 10) {
            const double *rp = sm[r], *sp = sm[s];
#pragma unroll 8
            for (uint32_t e = 0; e < cnt; e++) acc = __dadd_rn(acc, __dmul_rn(rp[e], sp[e]));
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
