// pb_ordered.cu - bit-exact LEFT-TO-RIGHT f64 sums at parallel speed.
//
// The reference's cluster statistics - weighted mean (array/matrix2D.c:200-233), centred
// covariance (math/pca.c:84-97) and distortion (quantize/cluster.c:135-148) - are naive
// sequential accumulations over up to N pixels in ascending pixel order.  The parity bar is
// bit-exact, and fl(fl(a+b)+c) != fl(a+fl(b+c)), so a tree / shuffle / atomic reduction is
// out.  A literal sequential chain costs one dependent DADD (~8 cycles) per pixel per pass:
// seconds per image.  This file gets the SAME BITS in parallel.  The arithmetic (and the proof
// obligations) live in pb_span.h, which also compiles on the host: tests/native/test_span.cpp
// checks it against the literal loop on adversarial data.
//
//   Speculate, summarise, verify.
//     S1  k_ord_blocksum : plain (unordered) f64 sum of every block of OB elements, per chain.
//     S2  k_ord_prefix   : approximate running total at each block start.
//     S3  k_ord_summary  : one warp per block, 16 consecutive elements per lane: approximate running sum at
//                          every ELEMENT -> predicted binade of every partial sum; the block's unit is the ulp
//                          of the lowest one.  Each lane turns its elements into a span (integer translation +
//                          interval of start states for which every prediction is right), spans are
//                          concatenated in element order (an associative monoid): one record per block and
//                          chain.  Chains that cannot be summarised are flagged and their 512 terms written
//                          to a dump slot for the replay.
//     S3b k_ord_summary2 : (block, chain) pairs in which a step depends on the parity of the state (a tie, or
//                          a step up from the lowest binade) are redone for both parities (work list).
//     S3c k_ord_group    : 32 records -> one group record; for every record the composition of the run of
//                          usable records that starts there (run record, both parities).
//     S4  k_ord_resolve  : one warp per chain walks the blocks in order with the exact state (integer * unit):
//                          32 group records at a time (scan + ballot), then inside a group run by run (one
//                          interval check per run); a record that cannot be applied means its block is
//                          REPLAYED: the literal sequential loop over its (dumped) terms.
//   The predictions only decide SPEED: an accepted block is proven step by step to be what the
//   sequential loop computes (every rounding used the right grid), everything else is the loop itself.
//
// Small clusters skip S1-S3 and run S4 in replay-only mode (one launch).
#include <stdlib.h>

#include "pb_common.cuh"
#include "pb_kernels.h"
#include "pb_prof.h"
#include "pb_span.h"

namespace {

constexpr int OB = 512;         // elements per summary block
constexpr int OB_THREADS = 128; // S1: 4 elements per thread

// 0 accepted, 1 replayed, 2 unusable record, 3 state not expressible in the block's unit, 4 interval,
// 5 replay rounds, 6 element-wise sub-chunks, 7 blocks accepted through the two-parity record
// 8 cycles in scan-walk, 9 cycles in two-parity records, 10 cycles in replays, 11 record-group loads,
// 12 cycles of the slowest resolving warp seen
__device__ unsigned long long g_ord_counts[16];
// per chain of the centred pass (debug): cycles in {scan walk, record walk, replays}, replays, general-path records
__device__ unsigned long long g_ord_chain[7][5];

// One block of one chain.  flag: see F_*.
struct OrdRec {
    long long sum, lo, hi;
    int eref; // binade whose ulp is the unit
    int flag;
};
enum { F_OK = 0, F_REPLAY = 1, F_PENDING = 2, F_SENSITIVE = 3 };

enum { KIND_MEAN = 0, KIND_CENTERED = 1 };
template <int KIND> struct NChains { static constexpr int C = KIND == KIND_MEAN ? 4 : 7; };
// chain 0 of the mean pass is the weight sum; unweighted it is the row count (matrix2D.c:230) - not summed
template <int KIND, bool W> __host__ __device__ constexpr bool chain_live(int c) { return !(KIND == KIND_MEAN && !W && c == 0); }

// record of (chain, block) of a segment: chain-major inside the segment's region of the packed table, so
// that the 32 lanes of a resolving warp read 32 consecutive records
__device__ __forceinline__ size_t rec_row(const PbSeg &sg, int C, int chain, uint32_t nblk, uint32_t blk) {
    return (size_t)sg.bbase * C + (size_t)chain * nblk + blk;
}

// All chain terms of one element (S1 / S3).
//   MEAN:     t0 = w, t1..3 = c_j * w                                  (matrix2D.c:222-228, vector.c:97-109)
//   CENTERED: t0..5 = (w * c^_j) * c^_k for (j,k) = (0,0)(1,0)(1,1)(2,0)(2,1)(2,2)   (pca.c:88-93)
//             t6    = ((c^_0^2 + c^_1^2) + c^_2^2) * w                 (cluster.c:141-147)
template <int KIND, bool W>
__device__ __forceinline__ void terms_all(double w, double c0, double c1, double c2, double m0, double m1,
                                          double m2, double *t) {
    if (KIND == KIND_MEAN) {
        t[0] = W ? w : 1.0;
        t[1] = W ? __dmul_rn(c0, w) : c0;
        t[2] = W ? __dmul_rn(c1, w) : c1;
        t[3] = W ? __dmul_rn(c2, w) : c2;
    } else {
        const double d0 = __dsub_rn(c0, m0), d1 = __dsub_rn(c1, m1), d2 = __dsub_rn(c2, m2);
        const double w0 = W ? __dmul_rn(w, d0) : d0, w1 = W ? __dmul_rn(w, d1) : d1, w2 = W ? __dmul_rn(w, d2) : d2;
        t[0] = __dmul_rn(w0, d0);
        t[1] = __dmul_rn(w1, d0);
        t[2] = __dmul_rn(w1, d1);
        t[3] = __dmul_rn(w2, d0);
        t[4] = __dmul_rn(w2, d1);
        t[5] = __dmul_rn(w2, d2);
        const double ss = __dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2));
        t[6] = W ? __dmul_rn(ss, w) : ss;
    }
}

// The term of ONE chain (lane) for one element, branch-free in `lane` (S4 replay).
template <int KIND, bool W>
__device__ __forceinline__ double term_one(int lane, double w, double c0, double c1, double c2, double m0,
                                           double m1, double m2) {
    if (KIND == KIND_MEAN) {
        const double v = lane == 1 ? c0 : (lane == 2 ? c1 : c2);
        const double p = W ? __dmul_rn(v, w) : v;
        return lane == 0 ? (W ? w : 1.0) : p;
    } else {
        const double d0 = __dsub_rn(c0, m0), d1 = __dsub_rn(c1, m1), d2 = __dsub_rn(c2, m2);
        const int j = lane == 0 ? 0 : (lane <= 2 ? 1 : 2);
        const int k = (lane == 0 || lane == 1 || lane == 3) ? 0 : ((lane == 2 || lane == 4) ? 1 : 2);
        const double dj = j == 0 ? d0 : (j == 1 ? d1 : d2);
        const double dk = k == 0 ? d0 : (k == 1 ? d1 : d2);
        const double tc = W ? __dmul_rn(__dmul_rn(w, dj), dk) : __dmul_rn(dj, dk);
        const double ss = __dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2));
        const double td = W ? __dmul_rn(ss, w) : ss;
        return lane == 6 ? td : tc;
    }
}

__device__ __forceinline__ double block_reduce_sum(double v, double *sm /* [OB_THREADS/32] */) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = sm[0];
    for (int w = 1; w < OB_THREADS / 32; w++) r += sm[w];
    __syncthreads();
    return r;
}

// ---- S1: unordered block sums -----------------------------------------------------------------
template <int KIND, bool W>
__global__ void __launch_bounds__(OB_THREADS) k_ord_blocksum(PbPlanes b0, PbPlanes b1, const PbSeg *__restrict__ segs,
                                                             const PbStats *__restrict__ stats,
                                                             double *__restrict__ psum, unsigned cmask) {
    constexpr int C = NChains<KIND>::C;
    __shared__ double red[OB_THREADS / 32];
    const int seg = blockIdx.y;
    const PbSeg sg = segs[seg];
    const uint32_t base = blockIdx.x * OB;
    if (base >= sg.n) return;
    const PbPlanes &P = sg.buf ? b1 : b0;
    double m0 = 0, m1 = 0, m2 = 0;
    if (KIND == KIND_CENTERED) { m0 = stats[seg].mean[0]; m1 = stats[seg].mean[1]; m2 = stats[seg].mean[2]; }
    double acc[C];
#pragma unroll
    for (int c = 0; c < C; c++) acc[c] = 0.0;
#pragma unroll
    for (int k = 0; k < OB / OB_THREADS; k++) {
        const uint32_t i = base + k * OB_THREADS + threadIdx.x;
        if (i < sg.n) {
            const size_t p = (size_t)sg.lo + i;
            double t[C];
            terms_all<KIND, W>(W ? P.w[p] : 1.0, P.c[0][p], P.c[1][p], P.c[2][p], m0, m1, m2, t);
#pragma unroll
            for (int c = 0; c < C; c++) acc[c] += t[c];
        }
    }
    double *out = psum + ((size_t)sg.bbase + blockIdx.x) * C;
#pragma unroll
    for (int c = 0; c < C; c++) {
        if (!chain_live<KIND, W>(c) || !(cmask >> c & 1u)) continue; // (cmask: chains this rank owns, CTA-uniform)
        const double r = block_reduce_sum(acc[c], red);
        if (threadIdx.x == 0) out[c] = r;
    }
}

// ---- S1': one warp per block, every load of the block in flight at once; the MEAN pass also leaves the block's
// raw second moments behind, from which the centred pass derives ITS block sums without reading a pixel:
//     sum w (c_j - m_j)(c_k - m_k) = Q_jk - m_j S_k - m_k S_j + m_j m_k S_w        (Q_jk = sum w c_j c_k)
// The cancellation costs digits (|c|^2 / sigma^2 of them), which is fine: these sums only steer predictions.
constexpr int OBW_WARPS = 4;
constexpr int RAW_N = 10; // S_w, S_0, S_1, S_2, Q_00, Q_10, Q_11, Q_20, Q_21, Q_22
template <bool W>
__global__ void __launch_bounds__(32 * OBW_WARPS) k_ord_blocksum_raw(PbPlanes b0, PbPlanes b1, const PbSeg *__restrict__ segs,
                                                                     double *__restrict__ psum /* [block][4] */,
                                                                     double *__restrict__ raw /* [block][RAW_N] */) {
    const int seg = blockIdx.y, lane = threadIdx.x & 31;
    const PbSeg sg = segs[seg];
    const uint32_t blk = blockIdx.x * OBW_WARPS + (threadIdx.x >> 5);
    if ((size_t)blk * OB >= sg.n) return; // warp-uniform
    const PbPlanes &P = sg.buf ? b1 : b0;
    const uint32_t bcnt = min((uint32_t)OB, sg.n - blk * OB);
    const size_t g0 = (size_t)sg.lo + (size_t)blk * OB;
    double a[RAW_N];
#pragma unroll
    for (int i = 0; i < RAW_N; i++) a[i] = 0.0;
    auto add = [&](double w, double c0, double c1, double c2) {
        const double w0 = W ? c0 * w : c0, w1 = W ? c1 * w : c1, w2 = W ? c2 * w : c2;
        a[0] += W ? w : 1.0; a[1] += w0; a[2] += w1; a[3] += w2;
        a[4] = __fma_rn(w0, c0, a[4]); a[5] = __fma_rn(w1, c0, a[5]); a[6] = __fma_rn(w1, c1, a[6]);
        a[7] = __fma_rn(w2, c0, a[7]); a[8] = __fma_rn(w2, c1, a[8]); a[9] = __fma_rn(w2, c2, a[9]);
    };
    if (bcnt == OB) {
#pragma unroll
        for (int h = 0; h < 2; h++) { // two rounds of 8 rows: 24 - 32 loads in flight per lane
            double v0[8], v1[8], v2[8], vw[W ? 8 : 1];
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const size_t g = g0 + (h * 8 + q) * 32 + lane;
                v0[q] = P.c[0][g]; v1[q] = P.c[1][g]; v2[q] = P.c[2][g];
                if (W) vw[q] = P.w[g];
            }
#pragma unroll
            for (int q = 0; q < 8; q++) add(W ? vw[q] : 1.0, v0[q], v1[q], v2[q]);
        }
    } else {
        for (uint32_t i = lane; i < bcnt; i += 32) add(W ? P.w[g0 + i] : 1.0, P.c[0][g0 + i], P.c[1][g0 + i], P.c[2][g0 + i]);
    }
#pragma unroll
    for (int i = 0; i < RAW_N; i++)
#pragma unroll
        for (int o = 16; o; o >>= 1) a[i] += __shfl_xor_sync(0xffffffffu, a[i], o);
    const size_t row = (size_t)sg.bbase + blk;
    double mine = 0.0;
#pragma unroll
    for (int i = 0; i < RAW_N; i++) if (lane == i) mine = a[i];
    if (lane < RAW_N) raw[row * RAW_N + lane] = mine;
    if (lane < 4) psum[row * 4 + lane] = mine; // chains of the mean pass: w, c0 w, c1 w, c2 w
}

// block sums of the centred pass from the raw moments and the (now exact) mean
__global__ void __launch_bounds__(256) k_ord_derive_centered(const PbSeg *__restrict__ segs, const PbStats *__restrict__ stats,
                                                             const double *__restrict__ raw, double *__restrict__ psum /* [block][7] */) {
    const int seg = blockIdx.y;
    const PbSeg sg = segs[seg];
    const uint32_t nblk = (sg.n + OB - 1) / OB;
    const double m0 = stats[seg].mean[0], m1 = stats[seg].mean[1], m2 = stats[seg].mean[2];
    for (uint32_t blk = blockIdx.x * blockDim.x + threadIdx.x; blk < nblk; blk += gridDim.x * blockDim.x) {
        const double *r = raw + ((size_t)sg.bbase + blk) * RAW_N;
        const double Sw = r[0], S0 = r[1], S1 = r[2], S2 = r[3];
        const double t00 = r[4] - 2.0 * m0 * S0 + m0 * m0 * Sw;
        const double t10 = r[5] - m1 * S0 - m0 * S1 + m1 * m0 * Sw;
        const double t11 = r[6] - 2.0 * m1 * S1 + m1 * m1 * Sw;
        const double t20 = r[7] - m2 * S0 - m0 * S2 + m2 * m0 * Sw;
        const double t21 = r[8] - m2 * S1 - m1 * S2 + m2 * m1 * Sw;
        const double t22 = r[9] - 2.0 * m2 * S2 + m2 * m2 * Sw;
        double *o = psum + ((size_t)sg.bbase + blk) * 7;
        o[0] = t00; o[1] = t10; o[2] = t11; o[3] = t20; o[4] = t21; o[5] = t22; o[6] = (t00 + t11) + t22;
    }
}

// ---- S2: approximate exclusive prefix per chain (in place over the block sums) ------------------
// One CTA per (chain, segment): every thread sums a contiguous run of block sums (independent loads, several
// in flight), the CTA scans the partials, the threads write the exclusive prefixes of their run.  (The first
// version was one warp per chain: a 128 M pixel cluster has 262 144 blocks, i.e. 8192 dependent strided loads
// per lane and pass - 8 ms of pure latency at 16384^2.)  Any summation order will do: these are predictions.
constexpr int OP_THREADS = 1024;
template <int C>
__global__ void __launch_bounds__(OP_THREADS) k_ord_prefix(const PbSeg *__restrict__ segs, double *__restrict__ psum, int first_chain,
                                                           unsigned cmask) {
    __shared__ double s_part[OP_THREADS / 32];
    const int seg = blockIdx.y, c = blockIdx.x + first_chain, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (!(cmask >> c & 1u)) return;
    const uint32_t nblk = (segs[seg].n + OB - 1) / OB;
    double *io = psum + (size_t)segs[seg].bbase * C + c;
    const uint32_t per = (nblk + OP_THREADS - 1) / OP_THREADS;
    const uint32_t b0 = min((uint32_t)tid * per, nblk), b1 = min(b0 + per, nblk);
    double s = 0.0;
#pragma unroll 8
    for (uint32_t b = b0; b < b1; b++) s += io[(size_t)b * C];
    double incl = s;
    for (int o = 1; o < 32; o <<= 1) {
        const double v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_part[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        double w = s_part[lane];
        for (int o = 1; o < 32; o <<= 1) {
            const double v = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += v;
        }
        s_part[lane] = w; // inclusive over the warps
    }
    __syncthreads();
    double run = (incl - s) + (warp ? s_part[warp - 1] : 0.0);
    for (uint32_t b = b0; b < b1; b++) {
        const double v = io[(size_t)b * C];
        io[(size_t)b * C] = run;
        run += v;
    }
}

// Long segments: the scan of one chain is cut into S slabs so that S CTAs share it (one CTA per chain left the
// top of the tree - a 268 M pixel cluster has 524 288 blocks - scanning on 7 of 148 SMs for 1 ms per pass):
// slab sums first, then every slab scans itself on top of the sums of the slabs before it.
constexpr int OPS_THREADS = 256;
constexpr int OPS_MAX_SLABS = 32;
template <int C>
__global__ void __launch_bounds__(OPS_THREADS) k_ord_prefix_part(const PbSeg *__restrict__ segs, const double *__restrict__ psum,
                                                                 int first_chain, unsigned cmask, int S, double *__restrict__ part) {
    __shared__ double s_part[OPS_THREADS / 32];
    const int seg = blockIdx.y, c = blockIdx.x + first_chain, slab = blockIdx.z, tid = threadIdx.x;
    if (!(cmask >> c & 1u)) return;
    const uint32_t nblk = (segs[seg].n + OB - 1) / OB, L = (nblk + S - 1) / S;
    const uint32_t b0 = min((uint32_t)slab * L, nblk), b1 = min(b0 + L, nblk);
    const double *io = psum + (size_t)segs[seg].bbase * C + c;
    double acc = 0.0;
#pragma unroll 4
    for (uint32_t b = b0 + tid; b < b1; b += OPS_THREADS) acc += io[(size_t)b * C];
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((tid & 31) == 0) s_part[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < OPS_THREADS / 32; w++) t += s_part[w];
        part[((size_t)seg * S + slab) * C + c] = t;
    }
}
template <int C>
__global__ void __launch_bounds__(OPS_THREADS) k_ord_prefix_slab(const PbSeg *__restrict__ segs, double *__restrict__ psum, int first_chain,
                                                                 unsigned cmask, int S, const double *__restrict__ part) {
    __shared__ double s_part[OPS_THREADS / 32];
    const int seg = blockIdx.y, c = blockIdx.x + first_chain, slab = blockIdx.z, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (!(cmask >> c & 1u)) return;
    const uint32_t nblk = (segs[seg].n + OB - 1) / OB, L = (nblk + S - 1) / S;
    const uint32_t s0 = min((uint32_t)slab * L, nblk), s1 = min(s0 + L, nblk);
    double *io = psum + (size_t)segs[seg].bbase * C + c;
    double offset = 0.0;
    for (int q = 0; q < slab; q++) offset += part[((size_t)seg * S + q) * C + c];
    const uint32_t cnt = s1 - s0, per = (cnt + OPS_THREADS - 1) / OPS_THREADS;
    const uint32_t b0 = s0 + min((uint32_t)tid * per, cnt), b1 = min(b0 + per, s1);
    double s = 0.0;
#pragma unroll 8
    for (uint32_t b = b0; b < b1; b++) s += io[(size_t)b * C];
    double incl = s;
    for (int o = 1; o < 32; o <<= 1) {
        const double v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_part[warp] = incl;
    __syncthreads();
    double before = 0.0;
    for (int w = 0; w < warp; w++) before += s_part[w];
    double run = offset + before + (incl - s);
    for (uint32_t b = b0; b < b1; b++) {
        const double v = io[(size_t)b * C];
        io[(size_t)b * C] = run;
        run += v;
    }
}

// ---- S3: block summaries with per-element binade prediction ------------------------------------
// One WARP per block: 16 consecutive elements per lane.  (ncu on the first version - 64 threads x 8 elements,
// everything unrolled - showed two thirds of the instructions in the per-thread fixed part (span set-up and
// the in-order composition) and 30 % of the stall samples on instruction fetch: longer per-lane runs halve
// the fixed share, rolled element loops keep the body inside the instruction cache, and a warp needs no
// shared memory or CTA barrier to compose its block.)
#ifndef PB_OS_WARPS
#define PB_OS_WARPS 2
#endif
constexpr int OS_WARPS = PB_OS_WARPS;       // blocks per CTA
constexpr int OS_THREADS = 32 * OS_WARPS;
constexpr int OS_PER = OB / 32;             // consecutive elements per lane
constexpr int OS_STRIDE = OS_PER + 1;       // padded lane stride of the staged planes (doubles)
constexpr int OS_PLANE = 32 * OS_STRIDE;
template <int KIND, int NC> __host__ __device__ constexpr bool summary_staged() { return KIND == 0 /* KIND_MEAN */ && NC != 1; }
#ifndef PB_OS_UNROLL_MEAN
#define PB_OS_UNROLL_MEAN 4 // small loop body: unrolled for load-level parallelism
#endif
#ifndef PB_OS_UNROLL_CEN
#define PB_OS_UNROLL_CEN 1  // large loop body: rolled, it has to stay inside the instruction cache
#endif
#ifndef PB_OS_MINB
#define PB_OS_MINB 8 // minimum resident CTAs per SM asked of ptxas for the summary kernel (register cap)
#endif

__device__ __forceinline__ PbSpan shfl_down_span(const PbSpan &v, int o) {
    PbSpan r;
    r.sum = __shfl_down_sync(0xffffffffu, v.sum, o);
    r.lo = __shfl_down_sync(0xffffffffu, v.lo, o);
    r.hi = __shfl_down_sync(0xffffffffu, v.hi, o);
    return r;
}
__device__ __forceinline__ PbSpan shfl_up_span(const PbSpan &v, int o) {
    PbSpan r;
    r.sum = __shfl_up_sync(0xffffffffu, v.sum, o);
    r.lo = __shfl_up_sync(0xffffffffu, v.lo, o);
    r.hi = __shfl_up_sync(0xffffffffu, v.hi, o);
    return r;
}

// Blocks whose record is F_REPLAY are known before the resolve runs: their terms are written out here, in
// parallel, so that the sequential replay is a coalesced 4 KB read + the dependent adds (no loads of the
// pixel planes, no term arithmetic on the resolving warp's critical path).
struct Dump {
    double *terms;          // [cap][OB]
    unsigned int *count;    // slots handed out (may run past cap: those blocks are replayed from the planes)
    unsigned int cap;
};

// The general element step, out of line: it is the rare path (lanes whose predictions change level or
// sign) and keeping it out of the main body keeps the common path's register footprint small.
template <int NV>
__device__ __noinline__ void run_push_slow(PbRun *r, double term, double approx, int eref) {
    pb_run_push<NV>(*r, term, approx, eref);
}

// Summarises block `blk` of segment `sg`; called by a whole warp.  NC = C: every live chain, one record each
// (NV = 1; chains with a parity-dependent step are left F_PENDING and returned as a bit mask).  NC = 1: the
// single chain `ch0`, for both start parities (NV = 2, the work-list pass).
template <int KIND, bool W, int NV, int NC, bool MASKED = false>
__device__ __forceinline__ unsigned summarise_block(const PbPlanes &P, const PbSeg &sg, uint32_t blk, int ch0, double m0,
                                                    double m1, double m2, const double *__restrict__ pstart,
                                                    OrdRec *__restrict__ rec0, OrdRec *__restrict__ rec1, const Dump &dump,
                                                    double *stage /* this warp's [3 or 4][OS_PLANE] */, unsigned cmask) {
    constexpr int C = NChains<KIND>::C;
    constexpr int NOLEVEL = -(1 << 20);
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int OS_UNROLL = NC == 1 ? 4 : (KIND == KIND_MEAN ? PB_OS_UNROLL_MEAN : PB_OS_UNROLL_CEN);
    static_assert(NC == 1 || NC == C, "all chains or one");
    auto chain_of = [&](int slot) { return NC == 1 ? ch0 : slot; };
    // MASKED (chain-sharded runs only): the chains this rank owns; otherwise a compile-time predicate
    auto live = [&](int slot) { return NC == 1 ? true : (chain_live<KIND, W>(slot) && (!MASKED || (cmask >> slot & 1u))); };
    const uint32_t nblk = (sg.n + OB - 1) / OB;
    const int lane = threadIdx.x & 31;
    const uint32_t i0 = blk * OB + lane * OS_PER; // this lane's consecutive elements
    const bool have = i0 < sg.n;
    const int mycnt = have ? min(OS_PER, (int)(sg.n - i0)) : 0;
    // Mean pass (few operations per element): the block's planes go through shared memory - coalesced
    // 256-byte reads (two L1 wavefronts per instruction), then every lane reads ITS 16 consecutive elements
    // at stride 17 (conflict-free).  Straight from global memory every load instruction costs 32 wavefronts
    // (one 128-byte line per lane) and the load/store unit becomes the bottleneck.  The centred pass hides
    // that behind its arithmetic and measured slower with the extra staging step, so it reads directly.
    constexpr bool STAGED = summary_staged<KIND, NC>();
    const size_t p0 = (size_t)sg.lo + i0;
    if (STAGED) {
        const size_t b0 = (size_t)sg.lo + (size_t)blk * OB;
        const uint32_t bcnt = min((uint32_t)OB, sg.n - blk * OB);
        __syncwarp();
#pragma unroll 4
        for (int q = 0; q < OB / 32; q++) {
            const uint32_t idx = q * 32 + lane;
            if (idx < bcnt) {
                const int at = (int)(idx >> 4) * OS_STRIDE + (int)(idx & 15);
                stage[0 * OS_PLANE + at] = P.c[0][b0 + idx];
                stage[1 * OS_PLANE + at] = P.c[1][b0 + idx];
                stage[2 * OS_PLANE + at] = P.c[2][b0 + idx];
                if (W) stage[3 * OS_PLANE + at] = P.w[b0 + idx];
            }
        }
        __syncwarp();
    }
    const double *mine = stage + lane * OS_STRIDE;
    struct Px { double w, c0, c1, c2; };
    auto fetch = [&](int k) {
        Px x;
        if (STAGED) { x.w = W ? mine[3 * OS_PLANE + k] : 1.0; x.c0 = mine[k]; x.c1 = mine[OS_PLANE + k]; x.c2 = mine[2 * OS_PLANE + k]; }
        else { const size_t p = p0 + k; x.w = W ? P.w[p] : 1.0; x.c0 = P.c[0][p]; x.c1 = P.c[1][p]; x.c2 = P.c[2][p]; }
        return x;
    };
    auto terms_of = [&](const Px &x, double *t) {
        if (NC == 1) t[0] = term_one<KIND, W>(ch0, x.w, x.c0, x.c1, x.c2, m0, m1, m2);
        else terms_all<KIND, W>(x.w, x.c0, x.c1, x.c2, m0, m1, m2, t);
    };
    auto terms = [&](int k, double *t) { terms_of(fetch(k), t); };
    // ---- phase 1: approximate running sum at the start of this lane's elements ---------------------
    double tstart[NC];
    {
        double tl[NC];
#pragma unroll
        for (int c = 0; c < NC; c++) tl[c] = 0.0;
#pragma unroll OS_UNROLL
        for (int k = 0; k < mycnt; k++) { // (a register prefetch of element k + 1 measured slower: 128-register cap)
            double t[C];
            terms(k, t);
#pragma unroll
            for (int c = 0; c < NC; c++)
                if (live(c)) tl[c] += t[c];
        }
#pragma unroll
        for (int c = 0; c < NC; c++) {
            tstart[c] = 0.0;
            if (!live(c)) continue;
            double incl = tl[c];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double v = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += v;
            }
            tstart[c] = pstart[chain_of(c)] + (incl - tl[c]);
        }
    }
    // ---- phase 2: predicted binade of every partial sum (start states included) -> range over the block;
    //      at the same time the cheap summary, valid if this lane's predictions all sit on one level:
    //      it works in the ulp of the lane's own level, the block's unit is not needed yet ------------------
    int tlevel[NC]; // the lane's (absolute) level if uniform in binade and sign, else NOLEVEL
    int eref[NC];   // lowest predicted binade of the block
    bool usable[NC];
    PbUni uni[NC];
    {
        int emin[NC], emax[NC];
        double run[NC], uscale[NC];
        bool flip[NC];
#pragma unroll
        for (int c = 0; c < NC; c++) {
            run[c] = tstart[c];
            const int e = pb_exponent_of(run[c]);
            emin[c] = have ? e : (1 << 20);
            emax[c] = have ? e : NOLEVEL;
            flip[c] = false;
            pb_uni_begin(uni[c]);
            uscale[c] = pb_eref_ok(e) ? pb_pow2(52 - e) : 0.0;
        }
#pragma unroll OS_UNROLL
        for (int k = 0; k < mycnt; k++) {
            double t[C];
            terms(k, t);
#pragma unroll
            for (int c = 0; c < NC; c++) {
                if (!live(c)) continue;
                run[c] += t[c];
                const int e = pb_exponent_of(run[c]);
                emin[c] = min(emin[c], e);
                emax[c] = max(emax[c], e);
                flip[c] |= (run[c] < 0) != (tstart[c] < 0);
                pb_uni_push(uni[c], t[c], uscale[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < NC; c++) {
            tlevel[c] = (have && emin[c] == emax[c] && !flip[c]) ? emin[c] : NOLEVEL;
            eref[c] = 0;
            usable[c] = false;
            if (!live(c)) continue;
            const int lo = __reduce_min_sync(FULL, emin[c]), hi = __reduce_max_sync(FULL, emax[c]);
            // zero / subnormal / non-finite predictions, or too wide a range: replay
            usable[c] = pb_eref_ok(lo) && pb_eref_ok(hi) && hi - lo <= PB_SPAN_MAX_LEVEL;
            if (usable[c]) eref[c] = lo;
        }
    }
    // ---- phase 3: spans in units of the block's lowest binade -----------------------------------------
    bool slow[NC];
    PbRun run_st[NC];
    bool any_slow = false;
#pragma unroll
    for (int c = 0; c < NC; c++) {
        // a lane whose predictions sit on one level is a plain translation for BOTH start parities unless
        // it holds a tie on the lowest level; only then (NV = 2) it is redone by the general path
        bool fast = usable[c] && tlevel[c] != NOLEVEL;
        if (fast) {
            pb_uni_end(uni[c], run_st[c], tlevel[c] - eref[c], tstart[c] < 0);
            if (NV == 2 && run_st[c].sensitive) fast = false;
        }
        if (!fast) pb_run_begin(run_st[c], tstart[c], eref[c]);
        slow[c] = usable[c] && have && !fast;
        any_slow |= slow[c];
    }
    if (any_slow) { // the general path, element by element, for the chains of this lane that need it
        double run[NC];
#pragma unroll
        for (int c = 0; c < NC; c++) run[c] = tstart[c];
#pragma unroll 1
        for (int k = 0; k < mycnt; k++) {
            double t[C];
            terms(k, t);
#pragma unroll
            for (int c = 0; c < NC; c++) {
                if (!live(c)) continue;
                run[c] += t[c]; // same operations as phase 2: same predictions
                if (slow[c] && !run_st[c].bad) run_push_slow<NV>(&run_st[c], t[c], run[c], eref[c]);
            }
        }
    }
    __syncwarp();
    // in-order composition over the warp (lane i absorbs lane i + o); lane 0 ends up with the block
    unsigned pending = 0;
    int slot[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) {
        slot[c] = -1;
        if (!live(c)) continue;
        PbSpan2 v = pb_span2_identity();
        unsigned f = 0;
        if (usable[c] && have) {
            v = pb_run_span<NV>(run_st[c]);
            f = (run_st[c].bad ? 1u : 0u) | (run_st[c].sensitive ? 2u : 0u);
        }
        f = __reduce_or_sync(FULL, f);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            PbSpan2 r;
            r.p[0] = shfl_down_span(v.p[0], o);
            if (NV == 2) r.p[1] = shfl_down_span(v.p[1], o);
            else r.p[1] = r.p[0];
            if ((lane & (2 * o - 1)) == 0) {
                if (NV == 2) v = pb_span2_cat(v, r);
                else { v.p[0] = pb_span_cat(v.p[0], r.p[0]); v.p[1] = v.p[0]; }
            }
        }
        int flag;
        if (!usable[c] || (f & 1u)) flag = F_REPLAY;
        else if (NV == 1) flag = (f & 2u) ? F_PENDING : F_OK;
        // a parity-dependent block whose start state sits above its lowest binade (enforced by its own
        // start constraint) only ever sees parity 0: variant 0 of the two-parity composition is a plain record
        else flag = pb_exponent_of(pstart[chain_of(c)]) - eref[c] < 1 ? F_SENSITIVE : F_OK;
        if (lane == 0) {
            // contradictory predictions (empty interval): never applicable.  A two-parity record stays
            // usable if one parity is valid; the resolve checks the interval of the parity it needs.
            if (flag == F_OK && !pb_span_valid(v.p[0])) flag = F_REPLAY;
            if (flag == F_SENSITIVE && !pb_span_valid(v.p[0]) && !pb_span_valid(v.p[1])) flag = F_REPLAY;
            const size_t row = rec_row(sg, C, chain_of(c), nblk, blk);
            OrdRec o0;
            o0.sum = v.p[0].sum; o0.lo = v.p[0].lo; o0.hi = v.p[0].hi; o0.eref = eref[c]; o0.flag = flag;
            if (flag == F_REPLAY) { // the record carries the dump slot instead of a translation
                const unsigned int sl = atomicAdd(dump.count, 1u);
                o0.sum = sl < dump.cap ? (long long)sl : -1LL;
                slot[c] = (int)o0.sum;
            }
            rec0[row] = o0;
            if (NV == 2) {
                OrdRec o1;
                o1.sum = v.p[1].sum; o1.lo = v.p[1].lo; o1.hi = v.p[1].hi; o1.eref = eref[c]; o1.flag = flag;
                rec1[row] = o1;
            }
        }
        flag = __shfl_sync(FULL, flag, 0);
        slot[c] = __shfl_sync(FULL, slot[c], 0);
        if (flag == F_PENDING) pending |= 1u << c;
    }
    {
        bool any = false;
#pragma unroll
        for (int c = 0; c < NC; c++) any |= slot[c] >= 0;
        if (any) { // warp-uniform: write this lane's terms of the dumped chains
#pragma unroll 1
            for (int k = 0; k < mycnt; k++) {
                double t[C];
                terms(k, t);
#pragma unroll
                for (int c = 0; c < NC; c++)
                    if (slot[c] >= 0) dump.terms[(size_t)slot[c] * OB + lane * OS_PER + k] = t[c];
            }
        }
    }
    return pending; // warp-uniform: chains left F_PENDING
}

template <int KIND, bool W, bool MASKED>
__global__ void __launch_bounds__(OS_THREADS, PB_OS_MINB) k_ord_summary(PbPlanes b0, PbPlanes b1, const PbSeg *__restrict__ segs,
                                                            const PbStats *__restrict__ stats,
                                                            const double *__restrict__ psum, OrdRec *__restrict__ rec0,
                                                            unsigned int *__restrict__ list_count,
                                                            uint2 *__restrict__ list, Dump dump, unsigned cmask) {
    constexpr int C = NChains<KIND>::C;
    __shared__ double stage[OS_WARPS][summary_staged<KIND, C>() ? (W ? 4 : 3) * OS_PLANE : 1];
    const int seg = blockIdx.y;
    const PbSeg sg = segs[seg];
    const uint32_t blk = blockIdx.x * OS_WARPS + (threadIdx.x >> 5);
    if ((size_t)blk * OB >= sg.n) return; // warp-uniform
    const PbPlanes &P = sg.buf ? b1 : b0;
    double m0 = 0, m1 = 0, m2 = 0;
    if (KIND == KIND_CENTERED) { m0 = stats[seg].mean[0]; m1 = stats[seg].mean[1]; m2 = stats[seg].mean[2]; }
    const unsigned pending = summarise_block<KIND, W, 1, C, MASKED>(P, sg, blk, 0, m0, m1, m2, psum + ((size_t)sg.bbase + blk) * C,
                                                            rec0, nullptr, dump, stage[threadIdx.x >> 5], cmask);
    if ((threadIdx.x & 31) == 0 && pending) { // one work item per (block, chain): block index < 2^28
        const unsigned int at = atomicAdd(list_count, (unsigned)__popc(pending));
        unsigned int k = 0;
        for (int c = 0; c < C; c++)
            if (pending >> c & 1u) list[at + k++] = make_uint2((unsigned)seg, blk | ((unsigned)c << 28));
    }
}


// ---- S3-fast: block-uniform summaries ------------------------------------------------------------------
// The common case by far: EVERY partial sum of a block of one chain lies in the binade of the block's start
// state (a running sum of millions of terms moves by a fraction of itself inside 512 elements).  Then every
// step is the translation S <- S + RN_u(a_i) on ONE grid u = ulp(start) (pb_span.h, "translation"), and
//     RN_u(a) = (a + M) - M        with M = 1.5 * 2^e:  the FPU itself quantises onto the grid,
// two dependent additions per element instead of the level bookkeeping of the general summary.  The sum of
// quantised terms is exact in floating point (multiples of u below 2^53 u).  A tie (a_i exactly halfway
// between grid points: its winner depends on the parity of the state) shows up as |a_i - RN_u(a_i)| = u / 2
// and sends the (block, chain) pair to the general path; so does everything else that is not this case.
// The claim "all partial sums stay inside binade e" is again an INTERVAL for the exact start state, built
// from the exact extremes of the in-order prefix sums; the resolve checks it.  Soundness is pb_span.h's:
// this is pb_run_uniform with k = 0 on a whole block (tests/native/test_span.cpp: `fast block` cases).
//
// One warp per block, 16 consecutive elements per lane (in-order prefixes inside the lane, span monoid
// across lanes).  The block's planes are staged through shared memory: coalesced 256-byte global reads,
// then every lane reads ITS run at stride 17 (conflict-free).
#ifndef PB_OF_WARPS
#define PB_OF_WARPS 2
#endif
constexpr int OF_WARPS = PB_OF_WARPS;
constexpr int OF_THREADS = 32 * OF_WARPS;
#ifndef PB_OF_MINB
#define PB_OF_MINB 8 // resident CTAs per SM asked of ptxas: 16 warps at 128 registers; shared memory allows 8 (6 weighted)
#endif
constexpr long long OF_MARGIN = 1LL << 20; // units between the predicted start state and the interval's ends

// terms of chain c never negative (given non-negative weights): prefix extremes are 0 and the total
template <int KIND> __host__ __device__ constexpr bool chain_monotone(int c) {
    return KIND == KIND_MEAN ? c == 0 : (c == 0 || c == 2 || c == 5 || c == 6);
}

template <int KIND, bool W, bool MASKED>
__global__ void __launch_bounds__(OF_THREADS, PB_OF_MINB) k_ord_fast(PbPlanes b0, PbPlanes b1, const PbSeg *__restrict__ segs,
                                                                     const PbStats *__restrict__ stats,
                                                                     const double *__restrict__ psum, OrdRec *__restrict__ rec0,
                                                                     unsigned int *__restrict__ list_count,
                                                                     uint2 *__restrict__ list, unsigned cmask) {
    constexpr int C = NChains<KIND>::C;
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ double stage_all[OF_WARPS][(W ? 4 : 3) * OS_PLANE];
    const int seg = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const PbSeg sg = segs[seg];
    const uint32_t blk = blockIdx.x * OF_WARPS + warp;
    if ((size_t)blk * OB >= sg.n) return; // warp-uniform
    const PbPlanes &P = sg.buf ? b1 : b0;
    double *stage = stage_all[warp];
    const uint32_t nblk = (sg.n + OB - 1) / OB;
    const uint32_t bcnt = min((uint32_t)OB, sg.n - blk * OB);
    const size_t g0 = (size_t)sg.lo + (size_t)blk * OB;
    // ---- stage the block ------------------------------------------------------------------------------
    // a full block: every load of the block is issued before the first store (48 - 64 independent 256-byte
    // requests per warp in flight: one memory round trip per block instead of four; nothing else is live yet)
    if (bcnt == OB) {
        double v0[OB / 32], v1[OB / 32], v2[OB / 32], vw[W ? OB / 32 : 1];
#pragma unroll
        for (int q = 0; q < OB / 32; q++) {
            const size_t g = g0 + q * 32 + lane;
            v0[q] = P.c[0][g]; v1[q] = P.c[1][g]; v2[q] = P.c[2][g];
            if (W) vw[q] = P.w[g];
        }
#pragma unroll
        for (int q = 0; q < OB / 32; q++) {
            const int idx = q * 32 + lane;
            const int at = (idx >> 4) * OS_STRIDE + (idx & 15);
            stage[0 * OS_PLANE + at] = v0[q]; stage[1 * OS_PLANE + at] = v1[q]; stage[2 * OS_PLANE + at] = v2[q];
            if (W) stage[3 * OS_PLANE + at] = vw[q];
        }
    } else {
#pragma unroll 4
        for (int q = 0; q < OB / 32; q++) {
            const uint32_t idx = q * 32 + lane;
            if (idx < bcnt) {
                const int at = (int)(idx >> 4) * OS_STRIDE + (int)(idx & 15);
                stage[0 * OS_PLANE + at] = P.c[0][g0 + idx];
                stage[1 * OS_PLANE + at] = P.c[1][g0 + idx];
                stage[2 * OS_PLANE + at] = P.c[2][g0 + idx];
                if (W) stage[3 * OS_PLANE + at] = P.w[g0 + idx];
            }
        }
    }
    double m0 = 0, m1 = 0, m2 = 0;
    if (KIND == KIND_CENTERED) { m0 = stats[seg].mean[0]; m1 = stats[seg].mean[1]; m2 = stats[seg].mean[2]; }
    // ---- per chain: the grid of the predicted start state -------------------------------------------------
    const double *pstart = psum + ((size_t)sg.bbase + blk) * C;
    PbFastGrid grid[C];
#pragma unroll
    for (int c = 0; c < C; c++) grid[c] = pb_fast_grid(chain_live<KIND, W>(c) ? pstart[c] : 0.0);
    __syncwarp();
    // ---- this lane's 16 consecutive elements: quantised prefix sums and their extremes ---------------------
    const int mycnt = min(OS_PER, max(0, (int)bcnt - lane * OS_PER));
    const double *mine = stage + lane * OS_STRIDE;
    double ps[C], mn[C], mx[C];
    bool tie_any = false; // a tie in ANY chain sends all chains of the block to the general pass (ties are rare)
    bool wneg = false;
#pragma unroll
    for (int c = 0; c < C; c++) ps[c] = mn[c] = mx[c] = 0.0;
    auto element = [&](int k) {
        const double w = W ? mine[3 * OS_PLANE + k] : 1.0;
        double t[C];
        terms_all<KIND, W>(w, mine[k], mine[OS_PLANE + k], mine[2 * OS_PLANE + k], m0, m1, m2, t);
        if (W) wneg |= w < 0.0;
#pragma unroll
        for (int c = 0; c < C; c++) {
            if (!chain_live<KIND, W>(c)) continue;
            bool tie_c;
            ps[c] += pb_fast_quant(grid[c], t[c], &tie_c); // exact while below 2^53 u (checked at the end)
            tie_any |= tie_c;
            if (!chain_monotone<KIND>(c)) {
                mn[c] = ps[c] < mn[c] ? ps[c] : mn[c];
                mx[c] = ps[c] > mx[c] ? ps[c] : mx[c];
            }
        }
    };
    if (mycnt == OS_PER) { // (every lane of a whole block) straight-line code: constant shared-memory offsets
#pragma unroll
        for (int k = 0; k < OS_PER; k++) element(k);
    } else {
        for (int k = 0; k < mycnt; k++) element(k);
    }
    // which chains hold the tie(s): a second pass over the block, only when there was one (a few percent of the blocks)
    unsigned tie = 0;
    if (__any_sync(FULL, tie_any)) {
        for (int k = 0; k < mycnt; k++) {
            double t[C];
            terms_all<KIND, W>(W ? mine[3 * OS_PLANE + k] : 1.0, mine[k], mine[OS_PLANE + k], mine[2 * OS_PLANE + k], m0, m1, m2, t);
#pragma unroll
            for (int c = 0; c < C; c++) {
                if (!chain_live<KIND, W>(c)) continue;
                bool tie_c;
                pb_fast_quant(grid[c], t[c], &tie_c);
                if (tie_c) tie |= 1u << c;
            }
        }
        tie = __reduce_or_sync(FULL, tie);
    }
    const bool any_wneg = W && __any_sync(FULL, wneg);
    // ---- compose the lanes in element order; lane c finishes chain c ----------------------------------------
    double r_sum = 0.0, r_mn = 0.0, r_mx = 0.0; // of chain `lane`
    PbFastGrid r_grid = grid[0];
#pragma unroll
    for (int c = 0; c < C; c++) {
        if (!chain_live<KIND, W>(c)) continue;
        double tot, lo_ext, hi_ext;
        if (chain_monotone<KIND>(c)) { // terms >= 0 (negative weights are checked): extremes are 0 and the total
            tot = ps[c];
#pragma unroll
            for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync(FULL, tot, o);
            lo_ext = 0.0;
            hi_ext = tot;
        } else {
            double incl = ps[c];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double v = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += v;
            }
            const double excl = incl - ps[c];
            lo_ext = excl + mn[c];
            hi_ext = excl + mx[c];
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const double a = __shfl_xor_sync(FULL, lo_ext, o), b = __shfl_xor_sync(FULL, hi_ext, o);
                lo_ext = a < lo_ext ? a : lo_ext;
                hi_ext = b > hi_ext ? b : hi_ext;
            }
            tot = __shfl_sync(FULL, incl, 31);
        }
        if (lane == c) { r_sum = tot; r_mn = lo_ext; r_mx = hi_ext; r_grid = grid[c]; }
    }
    // ---- the record (lane c: chain c) -----------------------------------------------------------------------
    bool general = false;
    if (lane < C && chain_live<KIND, W>(lane) && (!MASKED || (cmask >> lane & 1u))) {
        PbSpan sp;
        const bool accept = !any_wneg && pb_fast_finish(r_grid, pstart[lane], r_sum, r_mn, r_mx, (tie >> lane & 1u) != 0, OF_MARGIN, sp);
        if (accept) {
            OrdRec o;
            o.sum = sp.sum; o.lo = sp.lo; o.hi = sp.hi; o.eref = r_grid.e; o.flag = F_OK;
            rec0[rec_row(sg, C, lane, nblk, blk)] = o;
        }
        general = !accept;
    }
    const unsigned gm = __ballot_sync(FULL, general);
    if (gm) { // one work item per (block, chain) for the general two-parity pass: block index < 2^28
        unsigned int at = 0;
        if (lane == 0) at = atomicAdd(list_count, (unsigned)__popc(gm));
        at = __shfl_sync(FULL, at, 0);
        if (general) list[at + __popc(gm & ((1u << lane) - 1u))] = make_uint2((unsigned)seg, blk | ((unsigned)lane << 28));
    }
    if (lane == 0) {
        int live = 0;
#pragma unroll
        for (int c = 0; c < C; c++) live += (chain_live<KIND, W>(c) && (!MASKED || (cmask >> c & 1u))) ? 1 : 0;
        atomicAdd(&g_ord_counts[13], (unsigned long long)(live - __popc(gm)));
        atomicAdd(&g_ord_counts[14], (unsigned long long)__popc(gm));
    }
}


// ---- S1+S2+S3 fused: single-pass summaries with a decoupled look-back ----------------------------------------
// k_ord_fast needs the approximate running sum at every block start, which cost a whole extra sweep over the
// planes (k_ord_blocksum) plus a scan (k_ord_prefix).  Here ONE kernel reads every pixel once:
//   * a warp takes a TILE of two consecutive blocks of a segment (a ticket per segment hands tiles out in launch
//     order, so every predecessor of a tile is already running - the look-back below cannot deadlock);
//   * both blocks are fetched into shared memory with asynchronous copies (cp.async, 8 bytes per lane and request,
//     256 coalesced bytes per warp request; the second block lands while the first is being summed);
//   * sweep A: unordered f64 sums of the tile's terms per chain -> the tile's AGGREGATE is published;
//   * look-back (Merrill & Garland's single-pass scan): the warp reads the status of the 32 tiles before it, adds
//     aggregates back to the nearest tile that already knows its inclusive prefix, publishes its own;
//   * sweep B: the block-uniform summary of k_ord_fast on the data still sitting in shared memory.
// The running sums only steer predictions (which grid a block is summarised on, whether it is offered at all):
// their summation order - which depends on timing here - cannot change a bit of the result.
constexpr int OT_BLOCKS = 2;              // blocks per tile
constexpr int OT_WARPS = 2;               // tiles per CTA
constexpr int OT_THREADS = 32 * OT_WARPS;
struct TileStat {
    double agg[8];  // chain sums of the tile (7 used)
    double incl[8]; // running sums up to and including the tile
};
enum { TS_NONE = 0, TS_AGG = 1, TS_INCL = 2 };

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <bool W>
__device__ __forceinline__ void stage_block_async(const PbPlanes &P, size_t g0, uint32_t bcnt, double *stage, int lane) {
#pragma unroll
    for (int q = 0; q < OB / 32; q++) {
        const uint32_t idx = q * 32 + lane;
        if (idx < bcnt) {
            const int at = (int)(idx >> 4) * OS_STRIDE + (int)(idx & 15);
            cp_async8(&stage[0 * OS_PLANE + at], &P.c[0][g0 + idx]);
            cp_async8(&stage[1 * OS_PLANE + at], &P.c[1][g0 + idx]);
            cp_async8(&stage[2 * OS_PLANE + at], &P.c[2][g0 + idx]);
            if (W) cp_async8(&stage[3 * OS_PLANE + at], &P.w[g0 + idx]);
        }
    }
    cp_async_commit();
}

// sweep A over one staged block: per chain the warp's unordered sum of the terms (every lane ends up with it)
template <int KIND, bool W>
__device__ __forceinline__ void block_sums(const double *stage, uint32_t bcnt, int lane, double m0, double m1, double m2,
                                           double *out /* [C] */) {
    constexpr int C = NChains<KIND>::C;
    const int mycnt = min(OS_PER, max(0, (int)bcnt - lane * OS_PER));
    const double *mine = stage + lane * OS_STRIDE;
    double acc[C];
#pragma unroll
    for (int c = 0; c < C; c++) acc[c] = 0.0;
#pragma unroll 4
    for (int k = 0; k < mycnt; k++) {
        double t[C];
        terms_all<KIND, W>(W ? mine[3 * OS_PLANE + k] : 1.0, mine[k], mine[OS_PLANE + k], mine[2 * OS_PLANE + k], m0, m1, m2, t);
#pragma unroll
        for (int c = 0; c < C; c++)
            if (chain_live<KIND, W>(c)) acc[c] += t[c];
    }
#pragma unroll
    for (int c = 0; c < C; c++) {
        out[c] = 0.0;
        if (!chain_live<KIND, W>(c)) continue;
        double v = acc[c];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        out[c] = v;
    }
}

// sweep B over one staged block: the block-uniform summary (see k_ord_fast) given the predicted start states
template <int KIND, bool W, bool MASKED>
__device__ __forceinline__ void fast_block(const double *stage, const PbSeg &sg, int seg, uint32_t blk, uint32_t bcnt, uint32_t nblk,
                                           int lane, double m0, double m1, double m2, const double *pstart /* [C], registers */,
                                           OrdRec *__restrict__ rec0, unsigned int *__restrict__ list_count,
                                           uint2 *__restrict__ list, unsigned cmask) {
    constexpr int C = NChains<KIND>::C;
    constexpr unsigned FULL = 0xffffffffu;
    PbFastGrid grid[C];
#pragma unroll
    for (int c = 0; c < C; c++) grid[c] = pb_fast_grid(chain_live<KIND, W>(c) ? pstart[c] : 0.0);
    const int mycnt = min(OS_PER, max(0, (int)bcnt - lane * OS_PER));
    const double *mine = stage + lane * OS_STRIDE;
    double ps[C], mn[C], mx[C];
    unsigned tie = 0;
    bool wneg = false;
#pragma unroll
    for (int c = 0; c < C; c++) ps[c] = mn[c] = mx[c] = 0.0;
#pragma unroll 2
    for (int k = 0; k < mycnt; k++) {
        const double w = W ? mine[3 * OS_PLANE + k] : 1.0;
        double t[C];
        terms_all<KIND, W>(w, mine[k], mine[OS_PLANE + k], mine[2 * OS_PLANE + k], m0, m1, m2, t);
        if (W) wneg |= w < 0.0;
#pragma unroll
        for (int c = 0; c < C; c++) {
            if (!chain_live<KIND, W>(c)) continue;
            bool tie_c;
            ps[c] += pb_fast_quant(grid[c], t[c], &tie_c);
            if (tie_c) tie |= 1u << c;
            if (!chain_monotone<KIND>(c)) {
                mn[c] = ps[c] < mn[c] ? ps[c] : mn[c];
                mx[c] = ps[c] > mx[c] ? ps[c] : mx[c];
            }
        }
    }
    tie = __reduce_or_sync(FULL, tie);
    const bool any_wneg = W && __any_sync(FULL, wneg);
    double r_sum = 0.0, r_mn = 0.0, r_mx = 0.0, r_start = 0.0;
    PbFastGrid r_grid = grid[0];
#pragma unroll
    for (int c = 0; c < C; c++) {
        if (!chain_live<KIND, W>(c)) continue;
        double tot, lo_ext, hi_ext;
        if (chain_monotone<KIND>(c)) {
            tot = ps[c];
#pragma unroll
            for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync(FULL, tot, o);
            lo_ext = 0.0;
            hi_ext = tot;
        } else {
            double incl = ps[c];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double v = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += v;
            }
            const double excl = incl - ps[c];
            lo_ext = excl + mn[c];
            hi_ext = excl + mx[c];
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const double a = __shfl_xor_sync(FULL, lo_ext, o), b = __shfl_xor_sync(FULL, hi_ext, o);
                lo_ext = a < lo_ext ? a : lo_ext;
                hi_ext = b > hi_ext ? b : hi_ext;
            }
            tot = __shfl_sync(FULL, incl, 31);
        }
        if (lane == c) { r_sum = tot; r_mn = lo_ext; r_mx = hi_ext; r_grid = grid[c]; r_start = pstart[c]; }
    }
    bool general = false;
    if (lane < C && chain_live<KIND, W>(lane) && (!MASKED || (cmask >> lane & 1u))) {
        PbSpan sp;
        const bool accept = !any_wneg && pb_fast_finish(r_grid, r_start, r_sum, r_mn, r_mx, (tie >> lane & 1u) != 0, OF_MARGIN, sp);
        if (accept) {
            OrdRec o;
            o.sum = sp.sum; o.lo = sp.lo; o.hi = sp.hi; o.eref = r_grid.e; o.flag = F_OK;
            rec0[rec_row(sg, C, lane, nblk, blk)] = o;
        }
        general = !accept;
    }
    const unsigned gm = __ballot_sync(FULL, general);
    if (gm) {
        unsigned int at = 0;
        if (lane == 0) at = atomicAdd(list_count, (unsigned)__popc(gm));
        at = __shfl_sync(FULL, at, 0);
        if (general) list[at + __popc(gm & ((1u << lane) - 1u))] = make_uint2((unsigned)seg, blk | ((unsigned)lane << 28));
    }
    if (lane == 0) {
        int live = 0;
#pragma unroll
        for (int c = 0; c < C; c++) live += (chain_live<KIND, W>(c) && (!MASKED || (cmask >> c & 1u))) ? 1 : 0;
        atomicAdd(&g_ord_counts[13], (unsigned long long)(live - __popc(gm)));
        atomicAdd(&g_ord_counts[14], (unsigned long long)__popc(gm));
    }
}

template <int KIND, bool W, bool MASKED>
__global__ void __launch_bounds__(OT_THREADS) k_ord_fused(PbPlanes b0, PbPlanes b1, const PbSeg *__restrict__ segs,
                                                          const PbStats *__restrict__ stats, double *__restrict__ psum,
                                                          OrdRec *__restrict__ rec0, unsigned int *__restrict__ list_count,
                                                          uint2 *__restrict__ list, unsigned int *__restrict__ tickets,
                                                          int *__restrict__ tflag, TileStat *__restrict__ tstat, unsigned cmask) {
    constexpr int C = NChains<KIND>::C;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int PLANES = W ? 4 : 3;
    extern __shared__ __align__(16) double ot_smem[]; // [OT_WARPS][OT_BLOCKS][PLANES * OS_PLANE]
    const int seg = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const PbSeg sg = segs[seg];
    const uint32_t nblk = (sg.n + OB - 1) / OB, ntile = (nblk + OT_BLOCKS - 1) / OT_BLOCKS;
    if ((uint32_t)blockIdx.x * OT_WARPS >= ntile) return; // no tile left for this CTA's first warp: none for any
    // the ticket: tiles of a segment are handed out in the order warps start running
    unsigned int tile = 0;
    if (lane == 0) tile = atomicAdd(&tickets[seg], 1u);
    tile = __shfl_sync(FULL, tile, 0);
    if (tile >= ntile) return;
    const PbPlanes &P = sg.buf ? b1 : b0;
    double *stage0 = ot_smem + (size_t)(warp * OT_BLOCKS) * PLANES * OS_PLANE, *stage1 = stage0 + PLANES * OS_PLANE;
    const uint32_t blk0 = tile * OT_BLOCKS, blk1 = blk0 + 1;
    const uint32_t cnt0 = min((uint32_t)OB, sg.n - blk0 * OB);
    const uint32_t cnt1 = blk1 < nblk ? min((uint32_t)OB, sg.n - blk1 * OB) : 0u;
    stage_block_async<W>(P, (size_t)sg.lo + (size_t)blk0 * OB, cnt0, stage0, lane);
    stage_block_async<W>(P, (size_t)sg.lo + (size_t)blk1 * OB, cnt1, stage1, lane); // (an empty group when there is no second block)
    double m0 = 0, m1 = 0, m2 = 0;
    if (KIND == KIND_CENTERED) { m0 = stats[seg].mean[0]; m1 = stats[seg].mean[1]; m2 = stats[seg].mean[2]; }
    // ---- sweep A -----------------------------------------------------------------------------------------
    double s0[C], s1[C];
    cp_async_wait<1>();
    __syncwarp();
    block_sums<KIND, W>(stage0, cnt0, lane, m0, m1, m2, s0);
    cp_async_wait<0>();
    __syncwarp();
    block_sums<KIND, W>(stage1, cnt1, lane, m0, m1, m2, s1);
    // ---- publish the aggregate, look back, publish the inclusive prefix ---------------------------------------
    const size_t trow = (size_t)sg.bbase + tile; // one row per tile; tiles <= blocks, so the block table's packing serves
    double mine_agg = 0.0; // lane c: chain c
#pragma unroll
    for (int c = 0; c < C; c++) if (lane == c) mine_agg = s0[c] + s1[c];
    if (lane < C) tstat[trow].agg[lane] = mine_agg;
    __threadfence();
    __syncwarp();
    if (lane == 0 && tile + 1 < ntile) atomicExch(&tflag[trow], TS_AGG); // (the last tile has no reader)
    double ex[C]; // running sums before this tile (every lane)
#pragma unroll
    for (int c = 0; c < C; c++) ex[c] = 0.0;
    {
        long long look = (long long)tile - 1; // nearest tile not yet accounted for
        while (look >= 0) {
            const long long t = look - lane; // lane 0: the nearest predecessor
            int f = TS_INCL;                 // tiles before the segment: an inclusive prefix of zero
            if (t >= 0) {
                const volatile int *fp = &tflag[(size_t)sg.bbase + (size_t)t];
                f = *fp;
            }
            const unsigned has_incl = __ballot_sync(FULL, f == TS_INCL);
            const int p = has_incl ? __ffs(has_incl) - 1 : 32;                  // nearest tile with an inclusive prefix
            const unsigned need = p >= 32 ? FULL : ((2u << p) - 1u);           // lanes 0..p
            const unsigned missing = __ballot_sync(FULL, f == TS_NONE) & need;
            if (missing) { __nanosleep(40); continue; }                         // a predecessor has not published yet
            __threadfence();
            // lanes < p add their tile's aggregate, lane p its inclusive prefix
            double v[C];
#pragma unroll
            for (int c = 0; c < C; c++) v[c] = 0.0;
            if (t >= 0 && lane <= p) {
                const volatile double *src = lane == p ? tstat[(size_t)sg.bbase + (size_t)t].incl : tstat[(size_t)sg.bbase + (size_t)t].agg;
#pragma unroll
                for (int c = 0; c < C; c++)
                    if (chain_live<KIND, W>(c)) v[c] = src[c];
            }
#pragma unroll
            for (int c = 0; c < C; c++) {
                if (!chain_live<KIND, W>(c)) continue;
                double x = v[c];
#pragma unroll
                for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
                ex[c] += x;
            }
            if (p < 32) break;
            look -= 32;
        }
    }
    {
        double mine_incl = 0.0;
#pragma unroll
        for (int c = 0; c < C; c++) if (lane == c) mine_incl = ex[c] + (s0[c] + s1[c]);
        if (lane < C) tstat[trow].incl[lane] = mine_incl;
        __threadfence();
        __syncwarp();
        if (lane == 0 && tile + 1 < ntile) atomicExch(&tflag[trow], TS_INCL);
    }
    // the predicted start states of both blocks (the general pass reads them from the table)
    double ps1[C];
#pragma unroll
    for (int c = 0; c < C; c++) ps1[c] = ex[c] + s0[c];
    {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int c = 0; c < C; c++) if (lane == c) { a = ex[c]; b = ps1[c]; }
        if (lane < C) {
            psum[((size_t)sg.bbase + blk0) * C + lane] = a;
            if (cnt1) psum[((size_t)sg.bbase + blk1) * C + lane] = b;
        }
    }
    // ---- sweep B --------------------------------------------------------------------------------------------
    fast_block<KIND, W, MASKED>(stage0, sg, seg, blk0, cnt0, nblk, lane, m0, m1, m2, ex, rec0, list_count, list, cmask);
    if (cnt1) fast_block<KIND, W, MASKED>(stage1, sg, seg, blk1, cnt1, nblk, lane, m0, m1, m2, ps1, rec0, list_count, list, cmask);
}

// ---- S3b: (block, chain) pairs with a parity-dependent step: both start parities ---------------------
template <int KIND, bool W>
__global__ void __launch_bounds__(OS_THREADS) k_ord_summary2(PbPlanes b0, PbPlanes b1, const PbSeg *__restrict__ segs,
                                                             const PbStats *__restrict__ stats,
                                                             const double *__restrict__ psum, OrdRec *__restrict__ rec0,
                                                             OrdRec *__restrict__ rec1,
                                                             const unsigned int *__restrict__ list_count,
                                                             const uint2 *__restrict__ list, Dump dump) {
    constexpr int C = NChains<KIND>::C;
    __shared__ double stage[OS_WARPS][summary_staged<KIND, 1>() ? (W ? 4 : 3) * OS_PLANE : 1];
    const unsigned int nitem = *list_count, stride = gridDim.x * OS_WARPS;
    for (unsigned int item = blockIdx.x * OS_WARPS + (threadIdx.x >> 5); item < nitem; item += stride) { // persistent warps
        const int seg = (int)list[item].x;
        const uint32_t blk = list[item].y & 0x0fffffffu;
        const int chain = (int)(list[item].y >> 28);
        const PbSeg sg = segs[seg];
        const PbPlanes &P = sg.buf ? b1 : b0;
        double m0 = 0, m1 = 0, m2 = 0;
        if (KIND == KIND_CENTERED) { m0 = stats[seg].mean[0]; m1 = stats[seg].mean[1]; m2 = stats[seg].mean[2]; }
        summarise_block<KIND, W, 2, 1>(P, sg, blk, chain, m0, m1, m2, psum + ((size_t)sg.bbase + blk) * C, rec0, rec1, dump,
                                       stage[threadIdx.x >> 5], ~0u);
    }
}

// ---- S3c: group records ------------------------------------------------------------------------------
// 32 consecutive block records of a chain composed into one (same monoid), in parallel over all groups, so
// that the sequential walk below moves 1024 blocks per step where nothing special happens.  A group that
// holds anything but plain records of one unit is marked F_REPLAY ("walk its records").
__device__ __forceinline__ size_t group_row(const PbSeg &sg, int C, int seg, int chain, uint32_t nblk) {
    // disjoint ranges of ceil(nblk / 32) rows per (segment, chain) in a table of total_blocks * C / 32 + nseg * (C + 1) rows
    return ((size_t)sg.bbase * C + (size_t)chain * nblk) / 32 + (size_t)chain + (size_t)seg * (C + 1);
}

// Besides the group record, every record i gets a RUN record: the composition of the maximal run of usable
// records (plain or two-parity) of one unit that starts at i (inside its group) and its length - a backward
// segmented scan of the two-state transducer, done here in parallel for all groups.  The sequential walk
// then crosses a whole run with one interval check on the variant its state's parity selects.
__device__ __forceinline__ PbSpan2 shfl_down_span2(const PbSpan2 &v, int o) {
    PbSpan2 r;
    r.p[0] = shfl_down_span(v.p[0], o);
    r.p[1] = shfl_down_span(v.p[1], o);
    return r;
}

template <int KIND, bool W>
__global__ void __launch_bounds__(32 * NChains<KIND>::C) k_ord_group(const PbSeg *__restrict__ segs,
                                                                     const OrdRec *__restrict__ rec0,
                                                                     const OrdRec *__restrict__ rec1,
                                                                     OrdRec *__restrict__ grec, OrdRec *__restrict__ rrec0,
                                                                     OrdRec *__restrict__ rrec1, unsigned cmask) {
    constexpr int C = NChains<KIND>::C;
    const int seg = blockIdx.y, chain = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const PbSeg sg = segs[seg];
    const uint32_t nblk = (sg.n + OB - 1) / OB, g0 = blockIdx.x * 32;
    if (g0 >= nblk || !chain_live<KIND, W>(chain) || !(cmask >> chain & 1u)) return;
    const uint32_t gcnt = min(32u, nblk - g0);
    PbSpan2 v = pb_span2_identity();
    int eref = 0, len = 0, flag = F_OK;
    const size_t row = rec_row(sg, C, chain, nblk, g0 + lane);
    if (lane < (int)gcnt) {
        const OrdRec r = rec0[row];
        v.p[0].sum = r.sum; v.p[0].lo = r.lo; v.p[0].hi = r.hi;
        v.p[1] = v.p[0];
        eref = r.eref;
        flag = r.flag;
        if (flag == F_SENSITIVE) {
            const OrdRec q = rec1[row];
            v.p[1].sum = q.sum; v.p[1].lo = q.lo; v.p[1].hi = q.hi;
        }
        len = (flag == F_OK || flag == F_SENSITIVE) ? 1 : 0;
    }
    const bool all_plain = __all_sync(0xffffffffu, lane >= (int)gcnt || flag == F_OK);
    bool open = len > 0; // the run may still extend: then it covers exactly [lane, lane + o) at round o
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const PbSpan2 pv = shfl_down_span2(v, o);
        const int plen = __shfl_down_sync(0xffffffffu, len, o), peref = __shfl_down_sync(0xffffffffu, eref, o);
        const bool popen = __shfl_down_sync(0xffffffffu, (int)open, o) != 0;
        if (open) {
            if (lane + o < 32 && plen > 0 && peref == eref) {
                v = pb_span2_cat(v, pv);
                len += plen;
                open = popen;
            } else open = false;
        }
    }
    if (lane < (int)gcnt) {
        OrdRec o;
        o.sum = v.p[0].sum; o.lo = v.p[0].lo; o.hi = v.p[0].hi; o.eref = eref;
        o.flag = len; // 0: record `lane` cannot be applied from its summary
        rrec0[row] = o;
        o.sum = v.p[1].sum; o.lo = v.p[1].lo; o.hi = v.p[1].hi;
        rrec1[row] = o;
    }
    if (lane == 0) { // the group record serves the plain level-2 scan: only groups without two-parity records
        OrdRec o;
        o.sum = v.p[0].sum; o.lo = v.p[0].lo; o.hi = v.p[0].hi; o.eref = eref;
        o.flag = (all_plain && len == (int)gcnt && pb_span_valid(v.p[0])) ? F_OK : F_REPLAY;
        grec[group_row(sg, C, seg, chain, nblk) + blockIdx.x] = o;
    }
}

// ---- S4: ordered resolve ---------------------------------------------------------------------------
// One CTA per cluster, one WARP per chain; the warp carries the exact running sum as an integer in the
// unit of the records it is walking (pb_span.h: PbState; every lane holds the same state).
//   level 2: 32 group records at a time - in-order warp scan of the span monoid + a ballot for the first
//            group whose interval does not hold (or that is not plain); everything before it is applied
//            in one step;
//   level 1: the records of that group, the same way: scan over the plain run that starts at the current
//            record, then the record the scan stopped at on its own (two-parity record, unit change,
//            interval failure, unusable record), then scan again;
//   a record that cannot be applied means its block is REPLAYED: the block's terms (written out by the
//   summary kernels for the blocks they flagged, recomputed from the planes otherwise) are staged in
//   shared memory and added one by one - the reference loop.
struct ResolveShared {
    double terms[7][OB];
    double res[7];
};

// Applies the longest acceptable prefix of the plain records held by lanes [first, cnt) (one record per
// lane, in order) to the state; returns how many were applied.
__device__ __forceinline__ uint32_t scan_apply(const OrdRec &r, uint32_t cnt, uint32_t first, PbState &st, int lane) {
    const int e0 = __shfl_sync(0xffffffffu, r.eref, (int)first);
    if (!pb_state_rebase(st, e0)) return 0; // warp-uniform
    PbSpan v = pb_span_identity(); // lanes before `first`
    if (lane >= (int)first) {
        if (lane < (int)cnt && r.flag == F_OK && r.eref == e0) { v.sum = r.sum; v.lo = r.lo; v.hi = r.hi; }
        else v = pb_span_invalid(); // absorbing: nothing from here on is accepted
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { // in-order inclusive scan of the monoid
        const PbSpan up = shfl_up_span(v, o);
        if (lane >= o) v = pb_span_cat(up, v);
    }
    const bool valid = pb_span_valid(v) && st.S >= v.lo && st.S <= v.hi;
    const unsigned fails = __ballot_sync(0xffffffffu, lane >= (int)first && !valid);
    const uint32_t f = fails ? (uint32_t)(__ffs(fails) - 1) : 32u; // lanes >= cnt always fail: f <= cnt when cnt < 32
    if (f > first) st.S += __shfl_sync(0xffffffffu, v.sum, (int)f - 1);
    return f - first;
}

// the sequential loop over terms staged in shared memory: every lane runs the same chain (broadcast reads)
__device__ __forceinline__ double chain_terms(const double *sm, uint32_t cnt, double s) {
    const double2 *sm2 = reinterpret_cast<const double2 *>(sm);
    uint32_t i = 0;
#pragma unroll 8
    for (; i + 2 <= cnt; i += 2) {
        const double2 v = sm2[i >> 1];
        s = __dadd_rn(s, v.x);
        s = __dadd_rn(s, v.y);
    }
    if (i < cnt) s = __dadd_rn(s, sm[i]);
    return s;
}

template <int KIND, bool W>
__device__ __forceinline__ double replay_block(const PbPlanes &P, size_t first, uint32_t cnt, int chain, double m0,
                                               double m1, double m2, double s, int lane, double *sm) {
    double t[OB / 32];
#pragma unroll
    for (int q = 0; q < OB / 32; q++) { // all loads in flight at once: nothing is stored until every term is in a register
        const uint32_t k = q * 32 + lane;
        t[q] = 0.0;
        if (k < cnt) {
            const size_t p = first + k;
            t[q] = term_one<KIND, W>(chain, W ? P.w[p] : 1.0, P.c[0][p], P.c[1][p], P.c[2][p], m0, m1, m2);
        }
    }
#pragma unroll
    for (int q = 0; q < OB / 32; q++) sm[q * 32 + lane] = t[q];
    __syncwarp();
    s = chain_terms(sm, cnt, s);
    __syncwarp();
    return s;
}

// replay of a block whose terms were written out by the summary kernels
__device__ __forceinline__ double replay_dump(const double *__restrict__ terms, uint32_t cnt, double s, int lane, double *sm) {
    const double2 *src = reinterpret_cast<const double2 *>(terms);
    double2 *dst = reinterpret_cast<double2 *>(sm);
    double2 t[OB / 64];
#pragma unroll
    for (int q = 0; q < OB / 64; q++) {
        const uint32_t k = q * 32 + lane;
        t[q] = make_double2(0.0, 0.0);
        if (2 * k < cnt) t[q] = src[k];
    }
#pragma unroll
    for (int q = 0; q < OB / 64; q++) dst[q * 32 + lane] = t[q];
    __syncwarp();
    s = chain_terms(sm, cnt, s);
    __syncwarp();
    return s;
}

template <int KIND, bool W>
__global__ void __launch_bounds__(32 * NChains<KIND>::C) k_ord_resolve(PbPlanes b0, PbPlanes b1,
                                                                       const PbSeg *__restrict__ segs,
                                                                       PbStats *__restrict__ stats,
                                                                       const OrdRec *__restrict__ rec0,
                                                                       const OrdRec *__restrict__ rec1,
                                                                       const OrdRec *__restrict__ grec,
                                                                       const OrdRec *__restrict__ rrec,
                                                                       const OrdRec *__restrict__ rrec1,
                                                                       const double *__restrict__ dump_terms,
                                                                       bool use_summaries, unsigned cmask, bool raw_mean) {
    constexpr int C = NChains<KIND>::C;
    __shared__ __align__(16) ResolveShared sh;
    const int seg = blockIdx.x, chain = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const PbSeg sg = segs[seg];
    const PbPlanes &P = sg.buf ? b1 : b0;
    const uint32_t n = sg.n, nblk_all = (n + OB - 1) / OB,
                   nblk = (chain_live<KIND, W>(chain) && (cmask >> chain & 1u)) ? nblk_all : 0u;
    double m0 = 0, m1 = 0, m2 = 0;
    if (KIND == KIND_CENTERED) { m0 = stats[seg].mean[0]; m1 = stats[seg].mean[1]; m2 = stats[seg].mean[2]; }
    double sd = 0.0;                          // the exact running sum, authoritative while !st.ok
    PbState st = pb_state_from_double(sd);    // ... and as integer * unit while st.ok
    unsigned int n_acc = 0, n_rep = 0, n_acc2 = 0, n_gen = 0, n_why[3] = {0, 0, 0};
    long long cyc[3] = {0, 0, 0}, t_begin = clock64();
    const size_t row0 = rec_row(sg, C, chain, nblk_all, 0);
    const size_t grow0 = group_row(sg, C, seg, chain, nblk_all);
    const uint32_t ngrp = (nblk + 31) / 32;
    OrdRec dummy;
    dummy.sum = -1; dummy.lo = 1; dummy.hi = 0; dummy.eref = 0; dummy.flag = F_REPLAY;
    uint32_t g0 = 0; // next block; a multiple of 32 at the top of the loop
    while (g0 < nblk) {
        // ---- level 2: up to 32 groups in one step ------------------------------------------------------
        if (use_summaries && nblk - g0 > 64) {
            const long long t0 = clock64();
            const uint32_t G = g0 >> 5, gc = min(32u, ngrp - G);
            OrdRec q = dummy;
            if (lane < (int)gc) q = grec[grow0 + G + lane];
            const uint32_t a = scan_apply(q, gc, 0, st, lane);
            const uint32_t blocks = min(32u * a, nblk - g0);
            n_acc += blocks;
            g0 += blocks;
            cyc[0] += clock64() - t0;
            if (a == gc || g0 >= nblk) continue;
        }
        // ---- level 1: the records of one group --------------------------------------------------------
        const uint32_t gcnt = min(32u, nblk - g0);
        OrdRec r = dummy, r1 = dummy, rr = dummy, rq = dummy; // r1 is only meaningful where r.flag == F_SENSITIVE
        rr.flag = 0;
        if (use_summaries && lane < (int)gcnt) {
            r = rec0[row0 + g0 + lane];
            rr = rrec[row0 + g0 + lane];
            rq = rrec1[row0 + g0 + lane];
            if (r.flag == F_SENSITIVE) r1 = rec1[row0 + g0 + lane];
        }
        uint32_t next = 0;
        while (next < gcnt) {
            long long t0 = clock64();
            const int run = __shfl_sync(0xffffffffu, rr.flag, (int)next); // length of the usable run that starts here
            if (run > 0) {
                // the whole run with one interval check on the variant the state's parity selects (its
                // composition was prepared by k_ord_group) ...
                const int er = __shfl_sync(0xffffffffu, rr.eref, (int)next);
                uint32_t a = 0;
                bool crossed = false;
                if (pb_state_rebase(st, er)) {
                    const OrdRec &pick = (st.S & 1LL) ? rq : rr; // warp-uniform choice
                    const long long rsum = __shfl_sync(0xffffffffu, pick.sum, (int)next),
                                    rlo = __shfl_sync(0xffffffffu, pick.lo, (int)next),
                                    rhi = __shfl_sync(0xffffffffu, pick.hi, (int)next);
                    if (st.S >= rlo && st.S <= rhi) { // (an empty interval has lo > hi)
                        st.S += rsum;
                        a = (uint32_t)run;
                        crossed = true;
                    }
                }
                if (!crossed && __shfl_sync(0xffffffffu, r.flag, (int)next) == F_OK)
                    a = scan_apply(r, gcnt, next, st, lane); // ... or as far as the plain records go, one by one
                n_acc += a;
                next += a;
                cyc[0] += clock64() - t0;
                if (next >= gcnt) break;
                if (crossed) continue; // the next record starts another run or is unusable: look again
                t0 = clock64();
            }
            // record `next` on its own (warp-uniform)
            const int src = (int)next;
            const int fl1 = __shfl_sync(0xffffffffu, r.flag, src), er = __shfl_sync(0xffffffffu, r.eref, src);
            const long long s0 = __shfl_sync(0xffffffffu, r.sum, src);
            bool ok = false;
            int why = 0;
            n_gen++;
            if (fl1 == F_OK || fl1 == F_SENSITIVE) {
                if (pb_state_rebase(st, er)) {
                    const bool odd = (st.S & 1LL) && fl1 == F_SENSITIVE;
                    const OrdRec &pick = odd ? r1 : r;
                    const long long vsum = odd ? __shfl_sync(0xffffffffu, pick.sum, src) : s0;
                    const long long vlo = __shfl_sync(0xffffffffu, pick.lo, src), vhi = __shfl_sync(0xffffffffu, pick.hi, src);
                    if (st.S >= vlo && st.S <= vhi) {
                        st.S += vsum;
                        ok = true;
                        n_acc++;
                        n_acc2 += fl1 == F_SENSITIVE;
                    } else why = 2;
                } else why = 1;
            }
            if (!ok) {
                const long long t1 = clock64();
                const double s = st.ok ? pb_state_to_double(st) : sd;
                const uint32_t base = (g0 + next) * OB;
                const uint32_t cnt = min((uint32_t)OB, n - base);
                {   // the next block of this group that will be replayed from the dump: pull its 4 KB towards L1 now,
                    // one 128-byte line per lane, so that its load does not start cold after this chain
                    const unsigned nx = __ballot_sync(0xffffffffu, lane > src && r.flag == F_REPLAY && r.sum >= 0);
                    if (nx) {
                        const long long sl = __shfl_sync(0xffffffffu, r.sum, __ffs(nx) - 1);
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(dump_terms + (size_t)sl * OB + lane * 16));
                    }
                }
                if (fl1 == F_REPLAY && s0 >= 0) // the record names a dump slot
                    sd = replay_dump(dump_terms + (size_t)s0 * OB, cnt, s, lane, sh.terms[chain]);
                else
                    sd = replay_block<KIND, W>(P, (size_t)sg.lo + base, cnt, chain, m0, m1, m2, s, lane, sh.terms[chain]);
                st = pb_state_from_double(sd);
                n_rep++;
                n_why[why]++;
                const long long dt = clock64() - t1;
                cyc[2] += dt;
                cyc[1] -= dt;
            }
            next++;
            cyc[1] += clock64() - t0;
        }
        g0 += 32;
    }
    if (lane == 0) {
        sh.res[chain] = st.ok ? pb_state_to_double(st) : sd;
        if (use_summaries) {
            atomicAdd(&g_ord_counts[0], (unsigned long long)n_acc);
            atomicAdd(&g_ord_counts[1], (unsigned long long)n_rep);
            for (int q = 0; q < 3; q++) atomicAdd(&g_ord_counts[2 + q], (unsigned long long)n_why[q]);
            atomicAdd(&g_ord_counts[7], (unsigned long long)n_acc2);
            for (int q = 0; q < 3; q++) atomicAdd(&g_ord_counts[8 + q], (unsigned long long)cyc[q]);
            atomicAdd(&g_ord_counts[11], (unsigned long long)n_gen);
            atomicMax(&g_ord_counts[12], (unsigned long long)(clock64() - t_begin));
            if (KIND == KIND_CENTERED) {
                for (int q = 0; q < 3; q++) atomicAdd(&g_ord_chain[chain][q], (unsigned long long)cyc[q]);
                atomicAdd(&g_ord_chain[chain][3], (unsigned long long)n_rep);
                atomicAdd(&g_ord_chain[chain][4], (unsigned long long)n_gen);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (KIND == KIND_MEAN) {
            // matrix2D.c:230-231: scale = 1 / wsum (1 / rows when unweighted); mean *= scale
            const double wsum = W ? sh.res[0] : (double)n;
            const double inv = 1.0 / wsum;
            stats[seg].wsum = wsum;
            // raw_mean (chain-sharded runs): the sums leave unscaled, the owner of the weight sum may be another rank
            for (int j = 0; j < 3; j++) stats[seg].mean[j] = raw_mean ? sh.res[1 + j] : __dmul_rn(sh.res[1 + j], inv);
        } else {
            for (int j = 0; j < 6; j++) stats[seg].cov[j] = sh.res[j];
            stats[seg].dist = sh.res[6];
        }
    }
}

struct Scratch {
    double *psum;
    OrdRec *rec0, *rec1;
    uint2 *list;
    unsigned int *list_count; // [0] work list of summary2, [1] dump slots
    OrdRec *grec, *rrec, *rrec1;
    Dump dump;
    double *raw;           // raw second moments of every block (written by the mean pass, read by the centred pass)
    unsigned int *tickets; // fused pass: next tile of every segment
    int *tflag;            // ... status of every tile
    TileStat *tstat;       // ... its aggregate / inclusive prefix
};
constexpr size_t OT_MAX_SEGS = 256;
size_t group_rows(size_t total_blocks) { return total_blocks * 7 / 32 + 4096; } // + nseg * (C + 1), nseg <= 2 * 64
long long g_dump_cap_override = -1; // debug/test knob (patolette_b200_set_option "dump_cap")
// "fused_pass": single-pass summaries with a decoupled look-back (k_ord_fused) instead of blocksum + prefix + k_ord_fast.
// Measured SLOWER at 16384^2 (107 vs 85 ms for the two sweeps): the summaries are bound by instruction issue, not by HBM,
// and the fused kernel's 52 KB of staging per CTA halves the resident warps.  Kept as a tested route, off by default.
bool g_fused_pass = false;
int g_prefix_slabs = 1; // "prefix_slabs": long segments' block-sum scans shared by several CTAs (1, default: slabs of >= 8192 blocks), one CTA per chain (0), or a forced slab count (tests)
bool g_raw_moments = true;          // "raw_moments": the centred pass derives its block sums from the mean pass's raw moments
struct RawTag {
    const void *segs = nullptr;
    int nseg = 0;
    uint32_t max_n = 0, total_blocks = 0;
    bool weighted = false;
    bool operator==(const RawTag &o) const {
        return segs && segs == o.segs && nseg == o.nseg && max_n == o.max_n && total_blocks == o.total_blocks && weighted == o.weighted;
    }
};
RawTag &raw_tag(const void *scratch) { // which mean pass last filled the raw table of this scratch (two scratches per image)
    static const void *key[4] = {nullptr, nullptr, nullptr, nullptr};
    static RawTag tags[4];
    for (int i = 0; i < 4; i++)
        if (key[i] == scratch) return tags[i];
    static int next = 0;
    const int i = next++ & 3;
    key[i] = scratch;
    tags[i] = RawTag{};
    return tags[i];
}
bool g_fast_summary = true;         // "fast_summary": block-uniform summaries (k_ord_fast) + general work list; 0 = per-element summaries for every block
size_t dump_slots(size_t total_blocks) { return total_blocks / 4 + 1024; }
unsigned int dump_cap(size_t total_blocks) {
    const size_t n = dump_slots(total_blocks);
    return (unsigned int)(g_dump_cap_override >= 0 && (size_t)g_dump_cap_override < n ? (size_t)g_dump_cap_override : n);
}
Scratch carve(void *d_scratch, size_t total_blocks) {
    Scratch s;
    char *p = (char *)d_scratch;
    s.dump.terms = (double *)p; p += dump_slots(total_blocks) * OB * sizeof(double); // 4 KB slots: keeps 16 B alignment
    s.psum = (double *)p; p += total_blocks * 7 * sizeof(double);
    s.rec0 = (OrdRec *)p; p += total_blocks * 7 * sizeof(OrdRec);
    s.rec1 = (OrdRec *)p; p += total_blocks * 7 * sizeof(OrdRec);
    s.grec = (OrdRec *)p; p += group_rows(total_blocks) * sizeof(OrdRec);
    s.rrec = (OrdRec *)p; p += total_blocks * 7 * sizeof(OrdRec);
    s.rrec1 = (OrdRec *)p; p += total_blocks * 7 * sizeof(OrdRec);
    s.list = (uint2 *)p; p += total_blocks * 7 * sizeof(uint2);
    s.list_count = (unsigned int *)p; p += 64;
    s.dump.count = s.list_count + 1;
    s.dump.cap = dump_cap(total_blocks);
    s.raw = (double *)p; p += total_blocks * RAW_N * sizeof(double);
    s.tstat = (TileStat *)p; p += total_blocks * sizeof(TileStat);
    s.tflag = (int *)p; p += total_blocks * sizeof(int);
    s.tickets = (unsigned int *)p; // [OT_MAX_SEGS], directly behind the flags: one memset clears both
    return s;
}

// Unused dynamic shared memory that caps the resident CTAs of the summary kernel: every lane streams its
// own 128-byte line per plane, so the L1 working set is 12-16 KB per warp and more resident warps than L1
// can hold turn the second sweep over the block into L2 traffic.
size_t summary_pad_smem(int kind) {
    static int pad[2] = {-1, -1};
    if (pad[kind] < 0) {
        const char *e = getenv(kind == KIND_MEAN ? "PB200_SUMMARY_PAD_MEAN" : "PB200_SUMMARY_PAD_CENTERED");
        pad[kind] = e ? atoi(e) : 0;
    }
    return (size_t)pad[kind];
}

template <int KIND, bool W>
void launch_pass(const PbPlanes bufs[2], const PbSeg *d_segs, int nseg, uint32_t max_n, uint32_t total_blocks,
                 PbStats *d_stats, void *d_scratch, size_t scratch_bytes, cudaStream_t st, unsigned cmask, bool raw_mean) {
    constexpr int C = NChains<KIND>::C;
    const uint32_t blk_cap = (max_n + OB - 1) / OB; // grid width; the tables are packed by PbSeg::bbase
    const size_t need = pb_ordered_scratch_bytes(total_blocks);
    const bool speculative = max_n >= 8 * OB && d_scratch && need <= scratch_bytes;
    Scratch sc{};
    if (speculative) {
        sc = carve(d_scratch, total_blocks);
        dim3 grid(blk_cap, nseg), sgrid((blk_cap + OS_WARPS - 1) / OS_WARPS, nseg);
        const bool fused = g_fast_summary && g_fused_pass && (size_t)nseg <= OT_MAX_SEGS;
        if (fused) {
            PB_CUDA_OK(cudaMemsetAsync(sc.list_count, 0, 2 * sizeof(unsigned int), st));
            PB_CUDA_OK(cudaMemsetAsync(sc.tflag, 0, (size_t)total_blocks * sizeof(int) + OT_MAX_SEGS * sizeof(unsigned int), st));
            const uint32_t tile_cap = (blk_cap + OT_BLOCKS - 1) / OT_BLOCKS;
            dim3 tgrid((tile_cap + OT_WARPS - 1) / OT_WARPS, nseg);
            const size_t smem = (size_t)OT_WARPS * OT_BLOCKS * (W ? 4 : 3) * OS_PLANE * sizeof(double);
            if (cmask == ~0u) PB_CUDA_OK(cudaFuncSetAttribute(k_ord_fused<KIND, W, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            else PB_CUDA_OK(cudaFuncSetAttribute(k_ord_fused<KIND, W, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            PbProfScope p(KIND == KIND_MEAN ? "k_ord_fused_mean" : "k_ord_fused_centered", st);
            if (cmask == ~0u)
                k_ord_fused<KIND, W, false><<<tgrid, OT_THREADS, smem, st>>>(bufs[0], bufs[1], d_segs, d_stats, sc.psum, sc.rec0, sc.list_count, sc.list, sc.tickets, sc.tflag, sc.tstat, cmask);
            else
                k_ord_fused<KIND, W, true><<<tgrid, OT_THREADS, smem, st>>>(bufs[0], bufs[1], d_segs, d_stats, sc.psum, sc.rec0, sc.list_count, sc.list, sc.tickets, sc.tflag, sc.tstat, cmask);
        } else {
        // the mean pass leaves the raw second moments of every block in the scratch; the centred pass over the SAME
        // segments and scratch (the pipeline always runs them back to back) derives its block sums from them
        RawTag &tag = raw_tag(d_scratch);
        const RawTag now{d_segs, nseg, max_n, total_blocks, W};
        if (KIND == KIND_MEAN && g_raw_moments) {
            PbProfScope p("k_ord_blocksum_mean", st, false);
            k_ord_blocksum_raw<W><<<dim3((blk_cap + OBW_WARPS - 1) / OBW_WARPS, nseg), 32 * OBW_WARPS, 0, st>>>(bufs[0], bufs[1], d_segs, sc.psum, sc.raw);
            tag = now;
        } else if (KIND == KIND_CENTERED && g_raw_moments && tag == now) {
            PbProfScope p("k_ord_derive_centered", st, false);
            const uint32_t gx = (blk_cap + 255) / 256;
            k_ord_derive_centered<<<dim3(gx < 64 ? gx : 64, nseg), 256, 0, st>>>(d_segs, d_stats, sc.raw, sc.psum);
            tag = RawTag{};
        } else {
            PbProfScope p(KIND == KIND_MEAN ? "k_ord_blocksum_mean" : "k_ord_blocksum_centered", st, false);
            k_ord_blocksum<KIND, W><<<grid, OB_THREADS, 0, st>>>(bufs[0], bufs[1], d_segs, d_stats, sc.psum, cmask);
            tag = RawTag{};
        }
        {
            const int nch = chain_live<KIND, W>(0) ? C : C - 1, first = chain_live<KIND, W>(0) ? 0 : 1;
            int S = g_prefix_slabs >= 2 ? g_prefix_slabs : (int)(blk_cap / 8192);
            S = S > OPS_MAX_SLABS ? OPS_MAX_SLABS : S;
            if (g_prefix_slabs && S >= 2 && (size_t)nseg * S * C * sizeof(double) <= (size_t)total_blocks * sizeof(TileStat)) {
                double *part = reinterpret_cast<double *>(sc.tstat); // (the fused pass's table: idle on this route)
                { PbProfScope p("k_ord_prefix", st, false);
                  k_ord_prefix_part<C><<<dim3(nch, nseg, S), OPS_THREADS, 0, st>>>(d_segs, sc.psum, first, cmask, S, part); }
                { PbProfScope p("k_ord_prefix", st, false);
                  k_ord_prefix_slab<C><<<dim3(nch, nseg, S), OPS_THREADS, 0, st>>>(d_segs, sc.psum, first, cmask, S, part); }
            } else {
                PbProfScope p("k_ord_prefix", st, false);
                k_ord_prefix<C><<<dim3(nch, nseg), OP_THREADS, 0, st>>>(d_segs, sc.psum, first, cmask);
            }
        }
        PB_CUDA_OK(cudaMemsetAsync(sc.list_count, 0, 2 * sizeof(unsigned int), st));
        if (g_fast_summary) {
            dim3 fgrid((blk_cap + OF_WARPS - 1) / OF_WARPS, nseg);
            PbProfScope p(KIND == KIND_MEAN ? "k_ord_fast_mean" : "k_ord_fast_centered", st);
            if (cmask == ~0u)
                k_ord_fast<KIND, W, false><<<fgrid, OF_THREADS, 0, st>>>(bufs[0], bufs[1], d_segs, d_stats, sc.psum, sc.rec0, sc.list_count, sc.list, cmask);
            else
                k_ord_fast<KIND, W, true><<<fgrid, OF_THREADS, 0, st>>>(bufs[0], bufs[1], d_segs, d_stats, sc.psum, sc.rec0, sc.list_count, sc.list, cmask);
        } else
        { PbProfScope p(KIND == KIND_MEAN ? "k_ord_summary_mean" : "k_ord_summary_centered", st);
          if (cmask == ~0u) // every chain: the single-GPU instantiation, chain liveness known at compile time
              k_ord_summary<KIND, W, false><<<sgrid, OS_THREADS, summary_pad_smem(KIND), st>>>(bufs[0], bufs[1], d_segs, d_stats, sc.psum, sc.rec0, sc.list_count, sc.list, sc.dump, cmask);
          else
              k_ord_summary<KIND, W, true><<<sgrid, OS_THREADS, summary_pad_smem(KIND), st>>>(bufs[0], bufs[1], d_segs, d_stats, sc.psum, sc.rec0, sc.list_count, sc.list, sc.dump, cmask); }
        }
        { PbProfScope p("k_ord_summary2", st, false);
          k_ord_summary2<KIND, W><<<148 * 16, OS_THREADS, 0, st>>>(bufs[0], bufs[1], d_segs, d_stats, sc.psum, sc.rec0, sc.rec1, sc.list_count, sc.list, sc.dump); }
        { PbProfScope p("k_ord_group", st, false);
          k_ord_group<KIND, W><<<dim3((blk_cap + 31) / 32, nseg), 32 * C, 0, st>>>(d_segs, sc.rec0, sc.rec1, sc.grec, sc.rrec, sc.rrec1, cmask); }
    }
    { PbProfScope p(KIND == KIND_MEAN ? "k_ord_resolve_mean" : "k_ord_resolve_centered", st, !speculative);
      k_ord_resolve<KIND, W><<<nseg, 32 * C, 0, st>>>(bufs[0], bufs[1], d_segs, d_stats, sc.rec0, sc.rec1, sc.grec, sc.rrec, sc.rrec1, sc.dump.terms, speculative, cmask, raw_mean); }
    PB_CUDA_OK(cudaGetLastError());
}

} // namespace

void pb_ordered_counts(unsigned long long out[16], bool reset) {
    PB_CUDA_OK(cudaMemcpyFromSymbol(out, g_ord_counts, sizeof(unsigned long long) * 16));
    if (reset) {
        unsigned long long z[16] = {0};
        PB_CUDA_OK(cudaMemcpyToSymbol(g_ord_counts, z, sizeof z));
    }
}

void pb_ordered_chain_debug(unsigned long long out[35], bool reset) {
    PB_CUDA_OK(cudaMemcpyFromSymbol(out, g_ord_chain, sizeof(unsigned long long) * 35));
    if (reset) {
        unsigned long long z[35] = {0};
        PB_CUDA_OK(cudaMemcpyToSymbol(g_ord_chain, z, sizeof z));
    }
}

void pb_ordered_set_dump_cap(long long slots) { g_dump_cap_override = slots; }
void pb_ordered_set_fast(bool on) { g_fast_summary = on; }
void pb_ordered_set_fused(bool on) { g_fused_pass = on; }
void pb_ordered_set_prefix_slabs(int mode) { g_prefix_slabs = mode < 0 ? 0 : (mode > OPS_MAX_SLABS ? OPS_MAX_SLABS : mode); }
void pb_ordered_set_raw_moments(bool on) { g_raw_moments = on; }

uint32_t pb_ordered_blocks(uint32_t n) { return (n + OB - 1) / OB; }

size_t pb_ordered_scratch_bytes(size_t total_blocks) {
    return dump_slots(total_blocks) * OB * sizeof(double) + group_rows(total_blocks) * sizeof(OrdRec) +
           total_blocks * 7 * (sizeof(double) + 4 * sizeof(OrdRec) + sizeof(uint2)) + 256 +
           total_blocks * (sizeof(TileStat) + sizeof(int) + RAW_N * sizeof(double)) + OT_MAX_SEGS * sizeof(unsigned int);
}

void pb_launch_pass_mean(const PbPlanes bufs[2], const PbSeg *d_segs, int nseg, uint32_t max_n,
                         uint32_t total_blocks, bool weighted, PbStats *d_stats, void *d_scratch,
                         size_t scratch_bytes, cudaStream_t st, unsigned cmask, bool raw_mean) {
    if (nseg <= 0) return;
    if (weighted) launch_pass<KIND_MEAN, true>(bufs, d_segs, nseg, max_n, total_blocks, d_stats, d_scratch, scratch_bytes, st, cmask, raw_mean);
    else launch_pass<KIND_MEAN, false>(bufs, d_segs, nseg, max_n, total_blocks, d_stats, d_scratch, scratch_bytes, st, cmask, raw_mean);
}

void pb_launch_pass_centered(const PbPlanes bufs[2], const PbSeg *d_segs, int nseg, uint32_t max_n,
                             uint32_t total_blocks, bool weighted, PbStats *d_stats, void *d_scratch,
                             size_t scratch_bytes, cudaStream_t st, unsigned cmask) {
    if (nseg <= 0) return;
    if (weighted) launch_pass<KIND_CENTERED, true>(bufs, d_segs, nseg, max_n, total_blocks, d_stats, d_scratch, scratch_bytes, st, cmask, false);
    else launch_pass<KIND_CENTERED, false>(bufs, d_segs, nseg, max_n, total_blocks, d_stats, d_scratch, scratch_bytes, st, cmask, false);
}
