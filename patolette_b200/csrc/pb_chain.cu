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
constexpr int BK_TILE = 64;            // members per stage (2 per lane)
constexpr int BK_STRIDE = BK_TILE + 2;
constexpr int BK_WARPS = 4;
constexpr int BK_PER = BK_TILE / 32;
constexpr int BK_DEPTH = 4;            // stages in flight per warp

// The members of a bucket are gathered (random 8-byte reads of the planes, measured 135-157 cycles per
// element when only one tile was in flight) through a ring of BK_DEPTH stages filled with cp.async: the
// gathers of tile t + BK_DEPTH - 1 are issued before the chain over tile t starts, and the positions (`ord`)
// they need were loaded one iteration earlier, so the only latency left on the warp's critical path is the
// chain itself: a shared-memory broadcast read + one dependent DADD per element, ONE uniform loop
// `acc += term[lane][e]` for all chain lanes (every term is formed by the gathering lanes, in parallel).
__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// Generic driver: NP planes gathered per member, `emit(e, raw)` turns the staged values of element e into
// the NT term rows, chain lanes accumulate with `step(acc, term, lane)`.
template <int NP, int NT, typename Planes, typename Emit, typename Step>
__device__ __forceinline__ double bucket_chain(const uint32_t *__restrict__ ord, uint32_t beg, uint32_t end, Planes plane,
                                               double (*raw)[NP][BK_TILE], double (*term)[BK_STRIDE], Emit emit, Step step,
                                               int lane) {
    double acc = 0.0;
    if (beg >= end) return acc;
    const uint32_t ntile = (end - beg + BK_TILE - 1) / BK_TILE;
    uint32_t ordv[BK_PER]; // positions of the tile whose gathers are issued next
    auto load_ord = [&](uint32_t t) {
#pragma unroll
        for (int q = 0; q < BK_PER; q++) {
            const uint32_t i = beg + t * BK_TILE + q * 32 + lane;
            ordv[q] = (t < ntile && i < end) ? ord[i] : 0xffffffffu;
        }
    };
    auto issue = [&](uint32_t t) { // gathers of tile t into stage t % BK_DEPTH (always one commit per call)
#pragma unroll
        for (int q = 0; q < BK_PER; q++) {
            if (ordv[q] != 0xffffffffu) {
#pragma unroll
                for (int j = 0; j < NP; j++) cp_async8(&raw[t % BK_DEPTH][j][q * 32 + lane], plane(j) + ordv[q]);
            }
        }
        cp_async_commit();
    };
    for (uint32_t t = 0; t + 1 < BK_DEPTH; t++) { load_ord(t); issue(t); }
    load_ord(BK_DEPTH - 1);
    for (uint32_t t = 0; t < ntile; t++) {
        const uint32_t cnt = min((uint32_t)BK_TILE, end - (beg + t * BK_TILE));
        cp_async_wait<BK_DEPTH - 2>(); // tile t has landed (this lane's copies)
        __syncwarp();                  // ... and everybody else's
#pragma unroll
        for (int q = 0; q < BK_PER; q++) {
            const int e = q * 32 + lane;
            double v[NP];
#pragma unroll
            for (int j = 0; j < NP; j++) v[j] = raw[t % BK_DEPTH][j][e];
            emit(e, v, term);
        }
        __syncwarp();
        issue(t + BK_DEPTH - 1);      // refills the stage consumed at iteration t - 1
        load_ord(t + BK_DEPTH);
        if (lane < NT) {
            const double *vp = term[lane];
#pragma unroll 8
            for (uint32_t e = 0; e < cnt; e++) acc = step(acc, vp[e]);
        }
        __syncwarp();
    }
    cp_async_wait<0>();
    return acc;
}

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
    __shared__ double raw_all[BK_WARPS][BK_DEPTH][NT][BK_TILE];
    __shared__ double term_all[BK_WARPS][NT][BK_STRIDE];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gw = blockIdx.x * BK_WARPS + warp;
    if (gw >= nseg * PB_BUCKETS) return;
    const int seg = gw / PB_BUCKETS, b = gw % PB_BUCKETS;
    const PbSeg sg = segs[seg];
    const PbPlanes &P = sg.buf ? b1 : b0;
    const uint32_t *cs = class_start + (size_t)seg * (PB_BUCKETS + 1);
    const uint32_t beg = cs[b], end = cs[b + 1];
    auto plane = [&](int j) -> const double * { return j < 3 ? P.c[j] : P.w; };
    auto emit = [&](int e, const double *v, double (*term)[BK_STRIDE]) {
        if (WEIGHTED) {
            term[0][e] = v[3];
            term[1][e] = __dmul_rn(v[0], v[3]);
            term[2][e] = __dmul_rn(v[1], v[3]);
            term[3][e] = __dmul_rn(v[2], v[3]);
        } else {
            term[0][e] = v[0];
            term[1][e] = v[1];
            term[2][e] = v[2];
        }
    };
    auto step = [&](double acc, double t) {
        const double r = __dadd_rn(acc, t);
        return (WEIGHTED && lane == 0) ? trunc(r) : r;
    };
    const double acc = bucket_chain<NT, NT>(ord, beg, end, plane, raw_all[warp], term_all[warp], emit, step, lane);
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
    __shared__ double raw_all[BK_WARPS][BK_DEPTH][3][BK_TILE];
    __shared__ double term_all[BK_WARPS][NT][BK_STRIDE];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * BK_WARPS + warp;
    if (b >= PB_BUCKETS) return;
    const uint32_t beg = class_start[b], end = class_start[b + 1];
    auto plane = [&](int j) -> const double * { return src.c[j]; };
    auto emit = [&](int e, const double *v, double (*term)[BK_STRIDE]) {
        const double x = v[0], y = v[1], z = v[2];
        term[0][e] = x;
        term[1][e] = y;
        term[2][e] = z;
        term[3][e] = __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));
        term[4][e] = __dmul_rn(x, x);
        term[5][e] = __dmul_rn(x, y);
        term[6][e] = __dmul_rn(y, y);
        term[7][e] = __dmul_rn(x, z);
        term[8][e] = __dmul_rn(y, z);
        term[9][e] = __dmul_rn(z, z);
    };
    auto step = [&](double acc, double t) { return __dadd_rn(acc, t); };
    const double acc = bucket_chain<3, NT>(ord, beg, end, plane, raw_all[warp], term_all[warp], emit, step, lane);
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
