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
//     S3  k_ord_summary  : approximate running sum at every ELEMENT -> predicted binade of every
//                          partial sum; the block's unit is the ulp of the lowest one.  Each thread
//                          turns its 8 consecutive elements into a span (integer translation +
//                          interval of start states for which every prediction is right), spans are
//                          concatenated in element order (an associative monoid): one record per block.
//     S3b k_ord_summary2 : blocks in which a step depends on the parity of the state (a tie, or a step
//                          up from the lowest binade) are redone for both parities (work list).
//     S4  k_ord_resolve  : one warp per chain walks the blocks in order with the exact state: 32 block
//                          records at a time, in-order warp scan of the spans, ballot -> the first block
//                          whose interval does not hold; everything before it is applied in one step,
//                          that block is REPLAYED: the same idea with the exact binade at 16-element
//                          granularity, down to the literal sequential loop for the sub-chunk where a
//                          prediction breaks.
//   The predictions only decide SPEED: an accepted block is proven step by step to be what the
//   sequential loop computes (every rounding used the right grid), everything else is the loop itself.
//
// Small clusters skip S1-S3 and run S4 in replay-only mode (one launch).
#include "pb_common.cuh"
#include "pb_kernels.h"
#include "pb_prof.h"
#include "pb_span.h"

namespace {

constexpr int OB = 512;         // elements per summary block
constexpr int OB_THREADS = 128; // S1: 4 elements per thread
constexpr int E_NOGUESS = 0x7fffffff;
constexpr double MAGIC = 6755399441055744.0; // 1.5 * 2^52: (t + MAGIC) - MAGIC == rint(t) for |t| < 2^51
constexpr double TWO51 = 2251799813685248.0;
constexpr long long TWO52 = 1LL << 52, TWO53 = 1LL << 53;

// 0 accepted, 1 replayed, 2 unusable record, 3 state not expressible in the block's unit, 4 interval,
// 5 replay rounds, 6 element-wise sub-chunks, 7 blocks accepted through the two-parity record
__device__ unsigned long long g_ord_counts[8];

// One block of one chain.  flag: see F_*.
struct OrdRec {
    long long sum, lo, hi;
    int eref; // binade whose ulp is the unit
    int flag;
};
enum { F_OK = 0, F_REPLAY = 1, F_PENDING = 2, F_SENSITIVE = 3 };

enum { KIND_MEAN = 0, KIND_CENTERED = 1 };
template <int KIND> struct NChains { static constexpr int C = KIND == KIND_MEAN ? 4 : 7; };

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
                                                             double *__restrict__ psum) {
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
        const double r = block_reduce_sum(acc[c], red);
        if (threadIdx.x == 0) out[c] = r;
    }
}

// ---- S2: approximate exclusive prefix per chain (in place over the block sums) ------------------
template <int C>
__global__ void __launch_bounds__(32) k_ord_prefix(const PbSeg *__restrict__ segs, double *__restrict__ psum) {
    const int seg = blockIdx.y, c = blockIdx.x, lane = threadIdx.x;
    const uint32_t nblk = (segs[seg].n + OB - 1) / OB;
    double *io = psum + (size_t)segs[seg].bbase * C + c;
    const uint32_t per = (nblk + 31) / 32;
    const uint32_t b0 = min(lane * per, nblk), b1 = min(b0 + per, nblk);
    double s = 0.0;
    for (uint32_t b = b0; b < b1; b++) s += io[(size_t)b * C];
    double incl = s;
    for (int o = 1; o < 32; o <<= 1) {
        const double v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    double run = incl - s;
    for (uint32_t b = b0; b < b1; b++) {
        const double v = io[(size_t)b * C];
        io[(size_t)b * C] = run;
        run += v;
    }
}

// ---- S3: block summaries with per-element binade prediction ------------------------------------
constexpr int OS_THREADS = 64; // 8 consecutive elements per thread, two warps per block
constexpr int OS_PER = OB / OS_THREADS;

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

struct SumShared {
    double wsum[7];          // warp 0's total (approximate prefix hand-over)
    int emin[2][7], emax[2][7];
    PbSpan2 span[7];         // warp 0's span
    int flag[7];
};

// Summarises block `blk` of segment `sg` for every chain with need[c] (NV = 1: one record, parity-dependent
// blocks are left F_PENDING; NV = 2: both parities).  All threads of the CTA take part.
template <int KIND, bool W, int NV>
__device__ __forceinline__ bool summarise_block(const PbPlanes &P, const PbSeg &sg, uint32_t blk, double m0, double m1,
                                                double m2, const double *__restrict__ pstart, OrdRec *__restrict__ rec0,
                                                OrdRec *__restrict__ rec1, const bool *need, SumShared &sh) {
    constexpr int C = NChains<KIND>::C;
    const uint32_t nblk = (sg.n + OB - 1) / OB;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t i0 = blk * OB + threadIdx.x * OS_PER; // this thread's consecutive elements
    const bool have = i0 < sg.n;
    if (threadIdx.x < C) sh.flag[threadIdx.x] = 0;

    // ---- phase 1: approximate running sum at the start of this thread's elements ------------------
    double tstart[C];
    {
        double tl[C];
#pragma unroll
        for (int c = 0; c < C; c++) tl[c] = 0.0;
#pragma unroll
        for (int k = 0; k < OS_PER; k++) {
            if (i0 + k < sg.n) {
                const size_t p = (size_t)sg.lo + i0 + k;
                double t[C];
                terms_all<KIND, W>(W ? P.w[p] : 1.0, P.c[0][p], P.c[1][p], P.c[2][p], m0, m1, m2, t);
#pragma unroll
                for (int c = 0; c < C; c++) tl[c] += t[c];
            }
        }
#pragma unroll
        for (int c = 0; c < C; c++) {
            double incl = tl[c];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (warp == 0 && lane == 31) sh.wsum[c] = incl;
            tstart[c] = incl - tl[c];
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < C; c++) tstart[c] += pstart[c] + (warp ? sh.wsum[c] : 0.0);
    }
    // ---- phase 2: binade range of the predicted partial sums (start states included) --------------
    {
        int emin[C], emax[C];
        double run[C];
#pragma unroll
        for (int c = 0; c < C; c++) {
            run[c] = tstart[c];
            const int e = pb_exponent_of(run[c]);
            emin[c] = have ? e : (1 << 20);
            emax[c] = have ? e : -(1 << 20);
        }
#pragma unroll
        for (int k = 0; k < OS_PER; k++) {
            if (i0 + k < sg.n) {
                const size_t p = (size_t)sg.lo + i0 + k;
                double t[C];
                terms_all<KIND, W>(W ? P.w[p] : 1.0, P.c[0][p], P.c[1][p], P.c[2][p], m0, m1, m2, t);
#pragma unroll
                for (int c = 0; c < C; c++) {
                    run[c] += t[c];
                    const int e = pb_exponent_of(run[c]);
                    emin[c] = min(emin[c], e);
                    emax[c] = max(emax[c], e);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < C; c++) {
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                emin[c] = min(emin[c], __shfl_xor_sync(0xffffffffu, emin[c], o));
                emax[c] = max(emax[c], __shfl_xor_sync(0xffffffffu, emax[c], o));
            }
            if (lane == 0) { sh.emin[warp][c] = emin[c]; sh.emax[warp][c] = emax[c]; }
        }
        __syncthreads();
    }
    // ---- phase 3: spans on the predicted grids, in units of the lowest binade ----------------------
    int eref[C];
    bool usable[C];
    PbRun run_st[C];
#pragma unroll
    for (int c = 0; c < C; c++) {
        const int lo = min(sh.emin[0][c], sh.emin[1][c]), hi = max(sh.emax[0][c], sh.emax[1][c]);
        eref[c] = lo;
        // zero / subnormal / non-finite predictions, or too wide a range: replay
        usable[c] = need[c] && pb_eref_ok(lo) && pb_eref_ok(hi) && hi - lo <= PB_SPAN_MAX_LEVEL;
        if (!usable[c]) eref[c] = 0;
        pb_run_begin(run_st[c], tstart[c], eref[c]);
    }
    if (have) {
        double run[C];
#pragma unroll
        for (int c = 0; c < C; c++) run[c] = tstart[c];
#pragma unroll
        for (int k = 0; k < OS_PER; k++) {
            if (i0 + k < sg.n) {
                const size_t p = (size_t)sg.lo + i0 + k;
                double t[C];
                terms_all<KIND, W>(W ? P.w[p] : 1.0, P.c[0][p], P.c[1][p], P.c[2][p], m0, m1, m2, t);
#pragma unroll
                for (int c = 0; c < C; c++) {
                    run[c] += t[c]; // same operations as phase 2: same predictions
                    if (usable[c] && !run_st[c].bad) pb_run_push<NV>(run_st[c], t[c], run[c], eref[c]);
                }
            }
        }
    }
    bool pending = false;
#pragma unroll
    for (int c = 0; c < C; c++) {
        if (!need[c]) continue; // CTA-uniform
        PbSpan2 v = pb_span2_identity();
        if (usable[c] && have) {
            v = pb_run_span<NV>(run_st[c]);
            const int f = (run_st[c].bad ? 1 : 0) | (run_st[c].sensitive ? 2 : 0);
            if (f) atomicOr(&sh.flag[c], f);
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { // in-order tree: lane i absorbs lane i + o
            PbSpan2 r;
            r.p[0] = shfl_down_span(v.p[0], o);
            if (NV == 2) r.p[1] = shfl_down_span(v.p[1], o);
            else r.p[1] = r.p[0];
            if ((lane & (2 * o - 1)) == 0) {
                if (NV == 2) v = pb_span2_cat(v, r);
                else { v.p[0] = pb_span_cat(v.p[0], r.p[0]); v.p[1] = v.p[0]; }
            }
        }
        if (warp == 0 && lane == 0) sh.span[c] = v;
        __syncthreads();
        if (warp == 1 && lane == 0) {
            PbSpan2 w;
            if (NV == 2) w = pb_span2_cat(sh.span[c], v);
            else { w.p[0] = pb_span_cat(sh.span[c].p[0], v.p[0]); w.p[1] = w.p[0]; }
            const int f = sh.flag[c];
            int flag;
            if (!usable[c] || (f & 1)) flag = F_REPLAY;
            else if (NV == 1) flag = (f & 2) ? F_PENDING : F_OK;
            else flag = F_SENSITIVE;
            // contradictory predictions (empty interval): never applicable.  A two-parity record stays
            // usable if one parity is valid; the resolve checks the interval of the parity it needs.
            if (flag == F_OK && !pb_span_valid(w.p[0])) flag = F_REPLAY;
            if (flag == F_SENSITIVE && !pb_span_valid(w.p[0]) && !pb_span_valid(w.p[1])) flag = F_REPLAY;
            const size_t row = rec_row(sg, C, c, nblk, blk);
            OrdRec o0;
            o0.sum = w.p[0].sum; o0.lo = w.p[0].lo; o0.hi = w.p[0].hi; o0.eref = eref[c]; o0.flag = flag;
            rec0[row] = o0;
            if (NV == 2) {
                OrdRec o1;
                o1.sum = w.p[1].sum; o1.lo = w.p[1].lo; o1.hi = w.p[1].hi; o1.eref = eref[c]; o1.flag = flag;
                rec1[row] = o1;
            }
            pending |= flag == F_PENDING;
        }
    }
    return pending; // meaningful on (warp 1, lane 0)
}

template <int KIND, bool W>
__global__ void __launch_bounds__(OS_THREADS) k_ord_summary(PbPlanes b0, PbPlanes b1, const PbSeg *__restrict__ segs,
                                                            const PbStats *__restrict__ stats,
                                                            const double *__restrict__ psum, OrdRec *__restrict__ rec0,
                                                            unsigned int *__restrict__ list_count,
                                                            uint2 *__restrict__ list) {
    constexpr int C = NChains<KIND>::C;
    __shared__ SumShared sh;
    const int seg = blockIdx.y;
    const PbSeg sg = segs[seg];
    if (blockIdx.x * OB >= sg.n) return;
    const PbPlanes &P = sg.buf ? b1 : b0;
    double m0 = 0, m1 = 0, m2 = 0;
    if (KIND == KIND_CENTERED) { m0 = stats[seg].mean[0]; m1 = stats[seg].mean[1]; m2 = stats[seg].mean[2]; }
    bool need[C];
#pragma unroll
    for (int c = 0; c < C; c++) need[c] = true;
    const bool pending = summarise_block<KIND, W, 1>(P, sg, blockIdx.x, m0, m1, m2,
                                                     psum + ((size_t)sg.bbase + blockIdx.x) * C, rec0, nullptr, need, sh);
    if (threadIdx.x == 32 && pending) list[atomicAdd(list_count, 1u)] = make_uint2((unsigned)seg, blockIdx.x);
}

// ---- S3b: blocks with a parity-dependent step: both start parities ---------------------------------
template <int KIND, bool W>
__global__ void __launch_bounds__(OS_THREADS) k_ord_summary2(PbPlanes b0, PbPlanes b1, const PbSeg *__restrict__ segs,
                                                             const PbStats *__restrict__ stats,
                                                             const double *__restrict__ psum, OrdRec *__restrict__ rec0,
                                                             OrdRec *__restrict__ rec1,
                                                             const unsigned int *__restrict__ list_count,
                                                             const uint2 *__restrict__ list) {
    constexpr int C = NChains<KIND>::C;
    __shared__ SumShared sh;
    for (unsigned int item = blockIdx.x; item < *list_count; item += gridDim.x) { // persistent CTAs over the work list
        const int seg = (int)list[item].x;
        const uint32_t blk = list[item].y;
        const PbSeg sg = segs[seg];
        const PbPlanes &P = sg.buf ? b1 : b0;
        const uint32_t nblk = (sg.n + OB - 1) / OB;
        double m0 = 0, m1 = 0, m2 = 0;
        if (KIND == KIND_CENTERED) { m0 = stats[seg].mean[0]; m1 = stats[seg].mean[1]; m2 = stats[seg].mean[2]; }
        __syncthreads(); // sh is reused across items
        bool need[C]; // only the chains that asked for it are redone (uniform across the CTA)
#pragma unroll
        for (int c = 0; c < C; c++) need[c] = rec0[rec_row(sg, C, c, nblk, blk)].flag == F_PENDING;
        __syncthreads(); // every thread has read the flags before (warp 1, lane 0) rewrites the records
        summarise_block<KIND, W, 2>(P, sg, blk, m0, m1, m2, psum + ((size_t)sg.bbase + blk) * C, rec0, rec1, need, sh);
    }
}

// ---- S4: ordered resolve ---------------------------------------------------------------------------
constexpr int SUB = OB / 32; // elements per lane in a replay

__device__ __forceinline__ long long warp_incl_scan(long long v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
    }
    return v;
}

// One lane's 16 elements quantised against binade e: total, prefix extremes, exclusive prefix of the
// totals over the lanes before it, and whether the sub-chunk is unusable (unquantisable term, or a tie,
// whose rounding depends on the parity of the state - such a sub-chunk is simply added element-wise).
// The state provably stays inside binade e (prefix extremes checked against (2^52, 2^53)), so every
// step is the translation rint(a / q).
struct SubVer {
    int e;
    int bad;
    long long sum, mn, mx, pre;
};

__device__ __forceinline__ SubVer quantise_sub(const double *t, int my, int e, int lane) {
    SubVer v;
    v.e = e;
    const double scale = scalbn(1.0, 52 - e);
    double sum = 0.0, mn = 1e300, mx = -1e300;
    int bad = my == 0;
#pragma unroll
    for (int k = 0; k < SUB; k++) {
        if (k < my) {
            const double u = __dmul_rn(t[k], scale);
            const double d = __dsub_rn(__dadd_rn(u, MAGIC), MAGIC);
            bad |= !(fabs(u) < TWO51) | (fabs(__dsub_rn(u, d)) == 0.5);
            sum += d;
            mn = fmin(mn, sum);
            mx = fmax(mx, sum);
        }
    }
    v.bad = bad;
    v.sum = bad ? 0 : (long long)sum;
    v.mn = bad ? 0 : (long long)mn;
    v.mx = bad ? 0 : (long long)mx;
    v.pre = warp_incl_scan(v.sum, lane) - v.sum;
    return v;
}

// Replays one block exactly.  The exact state s is known, so each lane quantises its 16 consecutive
// elements against the TRUE binade; prefix totals are scanned once per binade, after which finding the
// first sub-chunk that cannot be applied is one ballot: lane l checks its own prefix extremes against
// the state it would start from if every lane before it is applied.  Accepted sub-chunks are applied in
// one step, the failing one (where the binade changes, or a tie sits) is added element by element - the
// literal reference loop - and the walk resumes behind it with the quantisation of the new binade.
// Sums that wander around a power of two bounce between adjacent binades, so the last three
// quantisations are kept.
template <int KIND, bool W>
__device__ __forceinline__ double replay_block(const PbPlanes &P, size_t first, uint32_t cnt, int chain, double m0,
                                               double m1, double m2, double s, int lane) {
    double t[SUB];
    const int my = max(0, min(SUB, (int)cnt - lane * SUB));
#pragma unroll
    for (int k = 0; k < SUB; k++) {
        t[k] = 0.0;
        if (k < my) {
            const size_t p = first + (size_t)lane * SUB + k;
            t[k] = term_one<KIND, W>(chain, W ? P.w[p] : 1.0, P.c[0][p], P.c[1][p], P.c[2][p], m0, m1, m2);
        }
    }
    const uint32_t nl = (cnt + SUB - 1) / SUB; // lanes that hold elements
    uint32_t next = 0;
    unsigned int rounds = 0, elementwise = 0;
    SubVer v0, v1, v2;
    v0.e = v1.e = v2.e = E_NOGUESS;
    int victim = 0;
    while (next < nl) {
        rounds++;
        const long long bits = __double_as_longlong(s);
        const int ef = (int)((bits >> 52) & 0x7ff);
        uint32_t f = next;
        if (ef > 24 && ef < 2000) { // a normal, finite state
            const int es = ef - 1023;
            if (v0.e != es && v1.e != es && v2.e != es) {
                const SubVer nv = quantise_sub(t, my, es, lane);
                if (victim == 0) v0 = nv; else if (victim == 1) v1 = nv; else v2 = nv;
                victim = victim == 2 ? 0 : victim + 1;
            }
            const SubVer &v = v0.e == es ? v0 : (v1.e == es ? v1 : v2);
            const long long M = (bits & 0x000fffffffffffffLL) | TWO52; // |s| / q
            const bool negs = bits < 0;
            const long long p = v.pre - __shfl_sync(0xffffffffu, v.pre, (int)next); // lanes [next, lane)
            const long long cur = negs ? M - p : M + p;
            const long long vmin = negs ? cur - v.mx : cur + v.mn, vmax = negs ? cur - v.mn : cur + v.mx;
            const bool mine = lane >= (int)next && lane < (int)nl;
            const bool valid = !v.bad && vmin > TWO52 && vmax < TWO53;
            const unsigned fails = __ballot_sync(0xffffffffu, mine && !valid);
            f = fails ? (uint32_t)(__ffs(fails) - 1) : nl;
            if (f > next) {
                const long long acc = __shfl_sync(0xffffffffu, p + v.sum, (int)f - 1); // total of lanes [next, f)
                const long long M2 = negs ? M - acc : M + acc;
                s = __longlong_as_double((bits & 0xfff0000000000000LL) | (M2 & 0x000fffffffffffffLL));
            }
        }
        if (f < nl) { // sub-chunk f: element by element (binade change, tie, or a zero / subnormal state)
            double v = s;
            if (lane == (int)f) {
#pragma unroll
                for (int k = 0; k < SUB; k++)
                    if (k < my) v = __dadd_rn(v, t[k]);
            }
            s = __shfl_sync(0xffffffffu, v, (int)f);
            next = f + 1;
            elementwise++;
        } else {
            next = nl;
        }
    }
    if (lane == 0) {
        atomicAdd(&g_ord_counts[5], (unsigned long long)rounds);
        atomicAdd(&g_ord_counts[6], (unsigned long long)elementwise);
    }
    return s;
}

template <int KIND, bool W>
__global__ void __launch_bounds__(32 * NChains<KIND>::C) k_ord_resolve(PbPlanes b0, PbPlanes b1,
                                                                       const PbSeg *__restrict__ segs,
                                                                       PbStats *__restrict__ stats,
                                                                       const OrdRec *__restrict__ rec0,
                                                                       const OrdRec *__restrict__ rec1,
                                                                       bool use_summaries) {
    constexpr int C = NChains<KIND>::C;
    __shared__ double s_res[C];
    const int seg = blockIdx.x, chain = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const PbSeg sg = segs[seg];
    const PbPlanes &P = sg.buf ? b1 : b0;
    const uint32_t n = sg.n, nblk = (n + OB - 1) / OB;
    double m0 = 0, m1 = 0, m2 = 0;
    if (KIND == KIND_CENTERED) { m0 = stats[seg].mean[0]; m1 = stats[seg].mean[1]; m2 = stats[seg].mean[2]; }
    double s = 0.0; // exact running sum of this warp's chain
    unsigned int n_acc = 0, n_rep = 0, n_acc2 = 0, n_why[3] = {0, 0, 0};
    const size_t row0 = rec_row(sg, C, chain, nblk, 0);
    for (uint32_t g0 = 0; g0 < nblk; g0 += 32) {
        const uint32_t gcnt = min(32u, nblk - g0);
        OrdRec r, r1;
        r.sum = 0; r.lo = 1; r.hi = 0; r.eref = 0; r.flag = F_REPLAY;
        if (use_summaries && lane < (int)gcnt) r = rec0[row0 + g0 + lane];
        r1 = r;
        if (r.flag == F_SENSITIVE) r1 = rec1[row0 + g0 + lane];
        uint32_t next = 0;
        while (next < gcnt) {
            const int fl0 = __shfl_sync(0xffffffffu, r.flag, (int)next), e0 = __shfl_sync(0xffffffffu, r.eref, (int)next);
            bool handled = false;
            int why = 0;
            if (fl0 == F_OK) {
                long long S = 0;
                if (pb_eref_ok(e0) && pb_state_to_units(s, e0, S)) {
                    // maximal run of plain records with the same unit starting at `next`
                    const bool okl = lane >= (int)next && r.flag == F_OK && r.eref == e0; // lanes >= gcnt hold F_REPLAY
                    const unsigned notok = __ballot_sync(0xffffffffu, lane >= (int)next && !okl);
                    const uint32_t runend = notok ? (uint32_t)(__ffs(notok) - 1) : 32u;
                    PbSpan v = pb_span_identity();
                    if (lane >= (int)next && lane < (int)runend) { v.sum = r.sum; v.lo = r.lo; v.hi = r.hi; }
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { // in-order inclusive scan of the monoid
                        const PbSpan up = shfl_up_span(v, o);
                        if (lane >= o) v = pb_span_cat(up, v);
                    }
                    const bool mine = lane >= (int)next && lane < (int)runend;
                    const bool valid = pb_span_valid(v) && S >= v.lo && S <= v.hi;
                    const unsigned fails = __ballot_sync(0xffffffffu, mine && !valid);
                    const uint32_t f = fails ? (uint32_t)(__ffs(fails) - 1) : runend;
                    if (f > next) {
                        const long long acc = __shfl_sync(0xffffffffu, v.sum, (int)f - 1);
                        s = pb_units_to_state(S + acc, e0); // exact: at most 53 significant bits by the last constraint
                        n_acc += f - next;
                        next = f;
                        handled = true;
                    } else {
                        why = 2; // the very first record's interval does not hold
                    }
                } else {
                    why = 1;
                }
            } else if (fl0 == F_SENSITIVE) {
                double s2 = s;
                int ok = 0;
                if (lane == (int)next) {
                    PbSpan2 sp;
                    sp.p[0].sum = r.sum; sp.p[0].lo = r.lo; sp.p[0].hi = r.hi;
                    sp.p[1].sum = r1.sum; sp.p[1].lo = r1.lo; sp.p[1].hi = r1.hi;
                    ok = pb_span2_apply(sp, r.eref, s2) ? 1 : 0;
                }
                ok = __shfl_sync(0xffffffffu, ok, (int)next);
                if (ok) {
                    s = __shfl_sync(0xffffffffu, s2, (int)next);
                    n_acc++;
                    n_acc2++;
                    next++;
                    handled = true;
                } else {
                    why = 2;
                }
            }
            if (!handled) {
                const uint32_t base = (g0 + next) * OB;
                s = replay_block<KIND, W>(P, (size_t)sg.lo + base, min((uint32_t)OB, n - base), chain, m0, m1, m2, s, lane);
                n_rep++;
                n_why[why]++;
                next++;
            }
        }
    }
    if (lane == 0) {
        s_res[chain] = s;
        if (use_summaries) {
            atomicAdd(&g_ord_counts[0], (unsigned long long)n_acc);
            atomicAdd(&g_ord_counts[1], (unsigned long long)n_rep);
            for (int q = 0; q < 3; q++) atomicAdd(&g_ord_counts[2 + q], (unsigned long long)n_why[q]);
            atomicAdd(&g_ord_counts[7], (unsigned long long)n_acc2);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (KIND == KIND_MEAN) {
            // matrix2D.c:230-231: scale = 1 / wsum (1 / rows when unweighted); mean *= scale
            const double wsum = W ? s_res[0] : (double)n;
            const double inv = 1.0 / wsum;
            stats[seg].wsum = wsum;
            for (int j = 0; j < 3; j++) stats[seg].mean[j] = __dmul_rn(s_res[1 + j], inv);
        } else {
            for (int j = 0; j < 6; j++) stats[seg].cov[j] = s_res[j];
            stats[seg].dist = s_res[6];
        }
    }
}

struct Scratch {
    double *psum;
    OrdRec *rec0, *rec1;
    uint2 *list;
    unsigned int *list_count;
};
Scratch carve(void *d_scratch, size_t total_blocks) {
    Scratch s;
    char *p = (char *)d_scratch;
    s.psum = (double *)p; p += total_blocks * 7 * sizeof(double);
    s.rec0 = (OrdRec *)p; p += total_blocks * 7 * sizeof(OrdRec);
    s.rec1 = (OrdRec *)p; p += total_blocks * 7 * sizeof(OrdRec);
    s.list = (uint2 *)p; p += total_blocks * sizeof(uint2);
    s.list_count = (unsigned int *)p;
    return s;
}

template <int KIND, bool W>
void launch_pass(const PbPlanes bufs[2], const PbSeg *d_segs, int nseg, uint32_t max_n, uint32_t total_blocks,
                 PbStats *d_stats, void *d_scratch, size_t scratch_bytes, cudaStream_t st) {
    constexpr int C = NChains<KIND>::C;
    const uint32_t blk_cap = (max_n + OB - 1) / OB; // grid width; the tables are packed by PbSeg::bbase
    const size_t need = pb_ordered_scratch_bytes(total_blocks);
    const bool speculative = max_n >= 8 * OB && d_scratch && need <= scratch_bytes;
    Scratch sc{};
    if (speculative) {
        sc = carve(d_scratch, total_blocks);
        dim3 grid(blk_cap, nseg);
        { PbProfScope p(KIND == KIND_MEAN ? "k_ord_blocksum_mean" : "k_ord_blocksum_centered", st, false);
          k_ord_blocksum<KIND, W><<<grid, OB_THREADS, 0, st>>>(bufs[0], bufs[1], d_segs, d_stats, sc.psum); }
        { PbProfScope p("k_ord_prefix", st, false);
          k_ord_prefix<C><<<dim3(C, nseg), 32, 0, st>>>(d_segs, sc.psum); }
        PB_CUDA_OK(cudaMemsetAsync(sc.list_count, 0, sizeof(unsigned int), st));
        { PbProfScope p(KIND == KIND_MEAN ? "k_ord_summary_mean" : "k_ord_summary_centered", st);
          k_ord_summary<KIND, W><<<grid, OS_THREADS, 0, st>>>(bufs[0], bufs[1], d_segs, d_stats, sc.psum, sc.rec0, sc.list_count, sc.list); }
        { PbProfScope p("k_ord_summary2", st, false);
          k_ord_summary2<KIND, W><<<148 * 4, OS_THREADS, 0, st>>>(bufs[0], bufs[1], d_segs, d_stats, sc.psum, sc.rec0, sc.rec1, sc.list_count, sc.list); }
    }
    { PbProfScope p(KIND == KIND_MEAN ? "k_ord_resolve_mean" : "k_ord_resolve_centered", st, !speculative);
      k_ord_resolve<KIND, W><<<nseg, 32 * C, 0, st>>>(bufs[0], bufs[1], d_segs, d_stats, sc.rec0, sc.rec1, speculative); }
    PB_CUDA_OK(cudaGetLastError());
}

} // namespace

void pb_ordered_counts(unsigned long long out[8], bool reset) {
    PB_CUDA_OK(cudaMemcpyFromSymbol(out, g_ord_counts, sizeof(unsigned long long) * 8));
    if (reset) {
        unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        PB_CUDA_OK(cudaMemcpyToSymbol(g_ord_counts, z, sizeof z));
    }
}

uint32_t pb_ordered_blocks(uint32_t n) { return (n + OB - 1) / OB; }

size_t pb_ordered_scratch_bytes(size_t total_blocks) {
    return total_blocks * (7 * (sizeof(double) + 2 * sizeof(OrdRec)) + sizeof(uint2)) + 256;
}

void pb_launch_pass_mean(const PbPlanes bufs[2], const PbSeg *d_segs, int nseg, uint32_t max_n,
                         uint32_t total_blocks, bool weighted, PbStats *d_stats, void *d_scratch,
                         size_t scratch_bytes, cudaStream_t st) {
    if (nseg <= 0) return;
    if (weighted) launch_pass<KIND_MEAN, true>(bufs, d_segs, nseg, max_n, total_blocks, d_stats, d_scratch, scratch_bytes, st);
    else launch_pass<KIND_MEAN, false>(bufs, d_segs, nseg, max_n, total_blocks, d_stats, d_scratch, scratch_bytes, st);
}

void pb_launch_pass_centered(const PbPlanes bufs[2], const PbSeg *d_segs, int nseg, uint32_t max_n,
                             uint32_t total_blocks, bool weighted, PbStats *d_stats, void *d_scratch,
                             size_t scratch_bytes, cudaStream_t st) {
    if (nseg <= 0) return;
    if (weighted) launch_pass<KIND_CENTERED, true>(bufs, d_segs, nseg, max_n, total_blocks, d_stats, d_scratch, scratch_bytes, st);
    else launch_pass<KIND_CENTERED, false>(bufs, d_segs, nseg, max_n, total_blocks, d_stats, d_scratch, scratch_bytes, st);
}
