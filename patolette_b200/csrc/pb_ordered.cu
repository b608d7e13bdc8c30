// pb_ordered.cu - bit-exact LEFT-TO-RIGHT f64 sums at parallel speed.
//
// The reference's cluster statistics - weighted mean (array/matrix2D.c:200-233), centred
// covariance (math/pca.c:84-97) and distortion (quantize/cluster.c:135-148) - are naive
// sequential accumulations over up to N pixels in ascending pixel order.  The parity bar is
// bit-exact, and fl(fl(a+b)+c) != fl(a+fl(b+c)), so a tree / shuffle / atomic reduction is
// out.  A literal sequential chain costs one dependent DADD (~6 cycles) per pixel per pass:
// seconds per image.  This file gets the SAME BITS in parallel.
//
//   Observation.  fl(s + a) rounds the exact sum x = s + a to the grid of the binade that x lies in:
//   with q = ulp(binade of x), fl(x) = q * rint(x / q).  s is itself a multiple of a quantum that
//   divides... not necessarily q, but if every partial sum is expressed in ONE fine unit u = 2^(E-52)
//   (E = the lowest binade met), s = S*u with S an integer, and the step becomes
//       S' = S + (rint(a / q_i) << k_i),   q_i = 2^(e_i - 52),  k_i = e_i - E          (no tie)
//   i.e. sequential floating-point accumulation is INTEGER accumulation of terms quantised to the grid
//   of the binade the running sum is in at that step - and integer addition is associative.  The
//   binades e_i are not known in advance, but an ordinary (unordered, approximate) prefix sum predicts
//   them: it is within ~1e-12 of the true running sum, so it lands in the same binade unless the sum
//   sits that close to a power of two.
//
//   Speculate, summarise, verify.
//     S1  k_ord_blocksum : plain f64 sum of every block of OB elements, per chain.
//     S2  k_ord_prefix   : approximate running total at each block start.
//     S3  k_ord_summary  : per block and chain: approximate running sum at every ELEMENT -> predicted
//                          binade e_i and sign; contribution c_i = rint(a_i / q_i) << k_i; and the
//                          condition for the prediction to be right - the true partial sum after element
//                          i must lie strictly inside binade e_i:  S_start + C_i in (2^(k_i+52), 2^(k_i+53))
//                          (mirrored for negative sums).  Every element thus bounds S_start by an
//                          interval; the block summary is  (sum of c_i, max of lower bounds, min of upper
//                          bounds)  - an in-order monoid reduction over exact integers.
//     S3b k_ord_summary_tie : a term that lands exactly half-way is rounded to the EVEN neighbour, which
//                          depends on the parity of the state: blocks that hold ties (and stay in one
//                          binade) get both parities (a two-state transducer, still associative).
//     S4  k_ord_resolve  : one warp per chain walks the blocks in order with the exact state; a block is
//                          accepted iff lo <= S <= hi, then S += sum.  Otherwise that block is REPLAYED:
//                          the same idea with the exact binade at 16-element granularity, down to the
//                          literal sequential loop for the sub-chunk where the prediction breaks.
//   The prediction only decides SPEED: an accepted block is proven step by step to be what the
//   sequential loop computes (every rounding used the right grid), everything else is the loop itself.
//   Sums that hover around zero (off-diagonal covariances of uncorrelated channels change binade every
//   few elements) validate like any other, as long as a block spans at most 8 binades.
//
// Small clusters skip S1-S3 and run S4 in replay-only mode (one launch).
#include "pb_common.cuh"
#include "pb_kernels.h"
#include "pb_prof.h"

namespace {

constexpr int OB = 512;         // elements per summary block
constexpr int OB_THREADS = 128; // S1: 4 elements per thread
constexpr int E_NOGUESS = 0x7fffffff;
constexpr double MAGIC = 6755399441055744.0; // 1.5 * 2^52: (t + MAGIC) - MAGIC == rint(t) for |t| < 2^51
constexpr double TWO51 = 2251799813685248.0;
constexpr long long TWO52 = 1LL << 52, TWO53 = 1LL << 53;
constexpr int MAX_SPREAD = 8; // binades a block may span (states stay below 2^61 in the block's unit)

// 0 accepted, 1 replayed, 2 flagged, 3 state outside the block's unit range, 4 interval, 5 replay rounds,
// 6 element-wise sub-chunks
__device__ unsigned long long g_ord_counts[8];

// Effect of a run of terms on the integer state S (in units of 2^(eref-52)): the run is what the
// sequential loop computes iff lo <= S_start <= hi; afterwards S = S_start + sum.
struct Span { long long sum, lo, hi; };
constexpr long long SPAN_INF = 1LL << 61;
__device__ __forceinline__ Span span_empty() { return Span{0, -SPAN_INF, SPAN_INF}; }
// in-order concatenation a ++ b: b's condition applies to S_start + a.sum
__device__ __forceinline__ Span span_cat(const Span &a, const Span &b) {
    return Span{a.sum + b.sum, max(a.lo, b.lo - a.sum), min(a.hi, b.hi - a.sum)};
}
struct Span2 { Span p[2]; }; // by parity of S_start (differs only when the run holds a tie)
__device__ __forceinline__ Span2 span2_cat(const Span2 &a, const Span2 &b) {
    Span2 r;
#pragma unroll
    for (int p = 0; p < 2; p++) r.p[p] = span_cat(a.p[p], b.p[(p + (int)(a.p[p].sum & 1LL)) & 1]);
    return r;
}

struct OrdSummary {
    Span2 t;
    int eref; // binade whose ulp is the unit of t
    int flag; // 0 usable; bit 0: replay; 2: tie inside a single-binade block -> k_ord_summary_tie
};

enum { KIND_MEAN = 0, KIND_CENTERED = 1 };
template <int KIND> struct NChains { static constexpr int C = KIND == KIND_MEAN ? 4 : 7; };

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
                                                             const PbStats *__restrict__ stats, uint32_t blk_cap,
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
__global__ void __launch_bounds__(32) k_ord_prefix(const PbSeg *__restrict__ segs, uint32_t blk_cap,
                                                   double *__restrict__ psum) {
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

__device__ __forceinline__ int exponent_of(double v) { return (int)((__double_as_longlong(v) >> 52) & 0x7ff) - 1023; }
__device__ __forceinline__ double pow2(int e) { return __longlong_as_double((long long)(e + 1023) << 52); }

// One element: predicted running sum `approx` (after the element) fixes the binade and sign; the term is
// quantised on that binade's grid and expressed in units of binade eref.  C = thread-local running total.
//   returns flags: bit 0 unusable, bit 1 tie
__device__ __forceinline__ int span_push(Span &sp, long long &C, double term, double approx, int eref) {
    const int e = exponent_of(approx), k = e - eref;
    int bad = (k < 0) | (k > MAX_SPREAD);
    const double u = __dmul_rn(term, pow2(52 - e)); // a / q_i, exact (power of two)
    bad |= !(fabs(u) < 9007199254740992.0);         // beyond 2^53 the state bound is violated anyway (or NaN)
    const long long d = __double2ll_rn(u);          // rint(u); exact remainder below
    const int tie = fabs(__dsub_rn(u, (double)d)) == 0.5;
    const int kk = bad ? 0 : k;
    C += d << kk;
    // true partial sum after this element strictly inside the predicted binade (mirrored if negative)
    const long long A = 1LL << (kk + 52), B = 1LL << (kk + 53);
    const bool neg = approx < 0;
    const long long lo = (neg ? -B : A) + 1 - C, hi = (neg ? -A : B) - 1 - C;
    sp.lo = max(sp.lo, lo);
    sp.hi = min(sp.hi, hi);
    sp.sum = C;
    return bad | (tie << 1);
}

template <int KIND, bool W>
__global__ void __launch_bounds__(OS_THREADS) k_ord_summary(PbPlanes b0, PbPlanes b1, const PbSeg *__restrict__ segs,
                                                            const PbStats *__restrict__ stats, uint32_t blk_cap,
                                                            const double *__restrict__ pstart,
                                                            OrdSummary *__restrict__ sum,
                                                            unsigned int *__restrict__ tie_count,
                                                            uint2 *__restrict__ tie_list) {
    constexpr int C = NChains<KIND>::C;
    __shared__ double s_wsum[C];            // warp 0's total (approximate prefix hand-over)
    __shared__ int s_emin[2][C], s_emax[2][C];
    __shared__ Span s_span[C];              // warp 0's span
    __shared__ int s_flag[C];
    const int seg = blockIdx.y;
    const PbSeg sg = segs[seg];
    const uint32_t base = blockIdx.x * OB;
    if (base >= sg.n) return;
    const PbPlanes &P = sg.buf ? b1 : b0;
    const size_t row = ((size_t)sg.bbase + blockIdx.x) * C;
    OrdSummary *out = sum + row;
    double m0 = 0, m1 = 0, m2 = 0;
    if (KIND == KIND_CENTERED) { m0 = stats[seg].mean[0]; m1 = stats[seg].mean[1]; m2 = stats[seg].mean[2]; }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t i0 = base + threadIdx.x * OS_PER; // this thread's consecutive elements
    if (threadIdx.x < C) s_flag[threadIdx.x] = 0;

    // ---- phase 1: approximate running sum at the start of this thread's elements -----------------
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
            if (warp == 0 && lane == 31) s_wsum[c] = incl;
            tstart[c] = incl - tl[c];
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < C; c++) tstart[c] += pstart[row + c] + (warp ? s_wsum[c] : 0.0);
    }
    // binade range of the predicted running sums over the block
    {
        int emin[C], emax[C];
        double run[C];
#pragma unroll
        for (int c = 0; c < C; c++) { emin[c] = 1 << 20; emax[c] = -(1 << 20); run[c] = tstart[c]; }
#pragma unroll
        for (int k = 0; k < OS_PER; k++) {
            if (i0 + k < sg.n) {
                const size_t p = (size_t)sg.lo + i0 + k;
                double t[C];
                terms_all<KIND, W>(W ? P.w[p] : 1.0, P.c[0][p], P.c[1][p], P.c[2][p], m0, m1, m2, t);
#pragma unroll
                for (int c = 0; c < C; c++) {
                    run[c] += t[c];
                    const int e = exponent_of(run[c]);
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
            if (lane == 0) { s_emin[warp][c] = emin[c]; s_emax[warp][c] = emax[c]; }
        }
        __syncthreads();
    }
    // ---- phase 2: quantise on the predicted grids, in units of the lowest binade ------------------
    int eref[C], flag[C];
    bool uniform[C]; // the whole block is predicted to stay in one binade
    Span sp[C];
#pragma unroll
    for (int c = 0; c < C; c++) {
        const int lo = min(s_emin[0][c], s_emin[1][c]), hi = max(s_emax[0][c], s_emax[1][c]);
        eref[c] = lo;
        uniform[c] = lo == hi;
        // zero / subnormal / non-finite predictions, or too wide a range: replay
        flag[c] = (lo < -1000) | (hi > 1000) | (hi - lo > MAX_SPREAD);
        if (flag[c]) eref[c] = 0;
        sp[c] = span_empty();
    }
    {
        double run[C];
        long long Cacc[C];
#pragma unroll
        for (int c = 0; c < C; c++) { run[c] = tstart[c]; Cacc[c] = 0; }
#pragma unroll
        for (int k = 0; k < OS_PER; k++) {
            if (i0 + k < sg.n) {
                const size_t p = (size_t)sg.lo + i0 + k;
                double t[C];
                terms_all<KIND, W>(W ? P.w[p] : 1.0, P.c[0][p], P.c[1][p], P.c[2][p], m0, m1, m2, t);
#pragma unroll
                for (int c = 0; c < C; c++) {
                    run[c] += t[c]; // same operations as phase 1: same predictions
                    const int f = span_push(sp[c], Cacc[c], t[c], run[c], eref[c]);
                    // a tie can be resolved by the parity transducer only when the block stays in one binade
                    flag[c] |= (f & 1) | ((f & 2) ? (uniform[c] ? 2 : 1) : 0);
                }
            }
        }
    }
#pragma unroll
    for (int c = 0; c < C; c++) {
        if (flag[c]) atomicOr(&s_flag[c], flag[c]);
        Span v = sp[c];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { // in-order tree: lane i absorbs lane i + o
            Span r;
            r.sum = __shfl_down_sync(0xffffffffu, v.sum, o);
            r.lo = __shfl_down_sync(0xffffffffu, v.lo, o);
            r.hi = __shfl_down_sync(0xffffffffu, v.hi, o);
            if ((lane & (2 * o - 1)) == 0) v = span_cat(v, r);
        }
        if (warp == 0 && lane == 0) s_span[c] = v;
        sp[c] = v;
    }
    __syncthreads();
    if (warp == 1 && lane == 0) {
        bool tie = false;
#pragma unroll
        for (int c = 0; c < C; c++) {
            const Span v = span_cat(s_span[c], sp[c]);
            out[c].t.p[0] = v;
            out[c].t.p[1] = v;
            out[c].eref = eref[c];
            out[c].flag = s_flag[c];
            tie |= s_flag[c] == 2;
        }
        if (tie) tie_list[atomicAdd(tie_count, 1u)] = make_uint2((unsigned)seg, blockIdx.x);
    }
    // blocks shorter than one warp's share have no warp-1 elements: warp 1 spans are empty, fine
}

// ---- S3b: single-binade blocks that hold a tie: both start parities ------------------------------
// (the parity of the state decides which neighbour a half-way term is rounded to)
__device__ __forceinline__ int span2_push(Span2 &sp, long long C[2], double term, double approx, int eref) {
    const int e = exponent_of(approx);
    int bad = e != eref;
    const double u = __dmul_rn(term, pow2(52 - eref));
    bad |= !(fabs(u) < 9007199254740992.0);
    const long long d = __double2ll_rn(u);
    const bool tie = fabs(__dsub_rn(u, (double)d)) == 0.5;
    const long long lo_int = tie ? (long long)floor(u) : d; // k of u = k + 0.5
    const bool neg = approx < 0;
#pragma unroll
    for (int p = 0; p < 2; p++) {
        // state before this element has parity p + C[p]; a tie goes to the neighbour that makes it even
        const long long dd = (tie && (((p + C[p] + lo_int) & 1LL) != 0)) ? lo_int + 1 : lo_int;
        C[p] += dd;
        const long long lo = (neg ? -TWO53 : TWO52) + 1 - C[p], hi = (neg ? -TWO52 : TWO53) - 1 - C[p];
        sp.p[p].lo = max(sp.p[p].lo, lo);
        sp.p[p].hi = min(sp.p[p].hi, hi);
        sp.p[p].sum = C[p];
    }
    return bad;
}

template <int KIND, bool W>
__global__ void __launch_bounds__(OS_THREADS) k_ord_summary_tie(PbPlanes b0, PbPlanes b1, const PbSeg *__restrict__ segs,
                                                                const PbStats *__restrict__ stats, uint32_t blk_cap,
                                                                const double *__restrict__ pstart,
                                                                OrdSummary *__restrict__ sum,
                                                                const unsigned int *__restrict__ tie_count,
                                                                const uint2 *__restrict__ tie_list) {
    constexpr int C = NChains<KIND>::C;
    __shared__ double s_wsum[C];
    __shared__ Span2 s_span[C];
    __shared__ int s_flag[C];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (unsigned int item = blockIdx.x; item < *tie_count; item += gridDim.x) { // persistent CTAs over the work list
        const int seg = (int)tie_list[item].x;
        const uint32_t blk = tie_list[item].y;
        const PbSeg sg = segs[seg];
        const uint32_t base = blk * OB;
        const PbPlanes &P = sg.buf ? b1 : b0;
        const size_t row = ((size_t)sg.bbase + blk) * C;
        OrdSummary *out = sum + row;
        double m0 = 0, m1 = 0, m2 = 0;
        if (KIND == KIND_CENTERED) { m0 = stats[seg].mean[0]; m1 = stats[seg].mean[1]; m2 = stats[seg].mean[2]; }
        const uint32_t i0 = base + threadIdx.x * OS_PER;
        __syncthreads();
        if (threadIdx.x < C) s_flag[threadIdx.x] = 0;
        bool need[C]; // only the chains that actually hold a tie are redone (uniform across the CTA)
        int eref[C];
#pragma unroll
        for (int c = 0; c < C; c++) { need[c] = out[c].flag == 2; eref[c] = out[c].eref; }
        // the same approximate running sum as k_ord_summary (same operations, same order)
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
                if (warp == 0 && lane == 31) s_wsum[c] = incl;
                tstart[c] = incl - tl[c];
            }
            __syncthreads();
#pragma unroll
            for (int c = 0; c < C; c++) tstart[c] += pstart[row + c] + (warp ? s_wsum[c] : 0.0);
        }
        Span2 sp[C];
        int flag[C];
        {
            double run[C];
            long long Cacc[C][2];
#pragma unroll
            for (int c = 0; c < C; c++) {
                run[c] = tstart[c]; Cacc[c][0] = Cacc[c][1] = 0; flag[c] = 0;
                sp[c].p[0] = sp[c].p[1] = span_empty();
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
                        if (need[c]) flag[c] |= span2_push(sp[c], Cacc[c], t[c], run[c], eref[c]);
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < C; c++) {
            if (!need[c]) continue;
            if (flag[c]) atomicOr(&s_flag[c], 1);
            Span2 v = sp[c];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                Span2 r;
#pragma unroll
                for (int p = 0; p < 2; p++) {
                    r.p[p].sum = __shfl_down_sync(0xffffffffu, v.p[p].sum, o);
                    r.p[p].lo = __shfl_down_sync(0xffffffffu, v.p[p].lo, o);
                    r.p[p].hi = __shfl_down_sync(0xffffffffu, v.p[p].hi, o);
                }
                if ((lane & (2 * o - 1)) == 0) v = span2_cat(v, r);
            }
            if (warp == 0 && lane == 0) s_span[c] = v;
            sp[c] = v;
        }
        __syncthreads();
        if (warp == 1 && lane == 0) {
#pragma unroll
            for (int c = 0; c < C; c++) {
                if (!need[c]) continue;
                out[c].t = span2_cat(s_span[c], sp[c]);
                out[c].flag = s_flag[c];
            }
        }
    }
}

// ---- S4: ordered resolve ---------------------------------------------------------------------------
// One CTA per cluster, one WARP per chain; the warp's state is the exact running sum s (uniform across
// lanes).  Summaries are fetched 32 blocks at a time (lane b holds block b) and applied in order.
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
                                               double m1, double m2, double s, int lane, int &hover) {
    (void)hover;
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

// exact state -> integer in units of 2^(eref-52); false if it is not representable there
__device__ __forceinline__ bool state_to_units(double s, int eref, long long &S) {
    const long long bits = __double_as_longlong(s);
    if ((bits << 1) == 0) { S = 0; return true; }
    const int es = (int)((bits >> 52) & 0x7ff) - 1023, k0 = es - eref;
    if (es < -1000 || es > 1000 || k0 < 0 || k0 > MAX_SPREAD + 1) return false;
    const long long M = ((bits & 0x000fffffffffffffLL) | TWO52) << k0;
    S = bits < 0 ? -M : M;
    return true;
}

template <int KIND, bool W>
__global__ void __launch_bounds__(32 * NChains<KIND>::C) k_ord_resolve(PbPlanes b0, PbPlanes b1,
                                                                       const PbSeg *__restrict__ segs,
                                                                       PbStats *__restrict__ stats, uint32_t blk_cap,
                                                                       const OrdSummary *__restrict__ sum,
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
    unsigned int n_acc = 0, n_rep = 0, n_why[3] = {0, 0, 0};
    int hover = 0;
    const OrdSummary *srow = sum + (size_t)sg.bbase * C + chain;
    for (uint32_t g0 = 0; g0 < nblk; g0 += 32) {
        const uint32_t gcnt = min(32u, nblk - g0);
        OrdSummary sm;
        sm.t.p[0] = sm.t.p[1] = span_empty();
        sm.eref = 0;
        sm.flag = 1;
        if (use_summaries && lane < (int)gcnt) sm = srow[(size_t)(g0 + lane) * C];
        for (uint32_t b = 0; b < gcnt; b++) {
            bool accept = false;
            if (use_summaries) {
                const int flag = __shfl_sync(0xffffffffu, sm.flag, (int)b), eref = __shfl_sync(0xffffffffu, sm.eref, (int)b);
                long long S = 0;
                int why = 0;
                if (flag == 0) {
                    if (state_to_units(s, eref, S)) {
                        const int p = (int)(S & 1LL);
                        const long long lo = __shfl_sync(0xffffffffu, p ? sm.t.p[1].lo : sm.t.p[0].lo, (int)b);
                        const long long hi = __shfl_sync(0xffffffffu, p ? sm.t.p[1].hi : sm.t.p[0].hi, (int)b);
                        if (S >= lo && S <= hi) {
                            const long long d = __shfl_sync(0xffffffffu, p ? sm.t.p[1].sum : sm.t.p[0].sum, (int)b);
                            s = __dmul_rn((double)(S + d), pow2(eref - 52)); // exact: a valid double by construction
                            accept = true;
                        } else why = 2;
                    } else why = 1;
                }
                if (!accept) n_why[why]++;
            }
            if (accept) {
                n_acc++;
            } else {
                const uint32_t base = (g0 + b) * OB;
                s = replay_block<KIND, W>(P, (size_t)sg.lo + base, min((uint32_t)OB, n - base), chain, m0, m1, m2, s, lane, hover);
                n_rep++;
            }
        }
    }
    if (lane == 0) {
        s_res[chain] = s;
        if (use_summaries) {
            atomicAdd(&g_ord_counts[0], (unsigned long long)n_acc);
            atomicAdd(&g_ord_counts[1], (unsigned long long)n_rep);
            for (int r = 0; r < 3; r++) atomicAdd(&g_ord_counts[2 + r], (unsigned long long)n_why[r]);
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

template <int KIND, bool W>
void launch_pass(const PbPlanes bufs[2], const PbSeg *d_segs, int nseg, uint32_t max_n, uint32_t total_blocks,
                 PbStats *d_stats, void *d_scratch, size_t scratch_bytes, cudaStream_t st) {
    constexpr int C = NChains<KIND>::C;
    const uint32_t blk_cap = (max_n + OB - 1) / OB; // grid width; the tables are packed by PbSeg::bbase
    const size_t need = pb_ordered_scratch_bytes(total_blocks);
    const bool speculative = max_n >= 8 * OB && d_scratch && need <= scratch_bytes;
    double *psum = (double *)d_scratch;
    OrdSummary *sum = (OrdSummary *)((char *)d_scratch + (size_t)total_blocks * 7 * sizeof(double));
    uint2 *tie_list = (uint2 *)((char *)d_scratch + (size_t)total_blocks * 7 * (sizeof(double) + sizeof(OrdSummary)));
    unsigned int *tie_count = (unsigned int *)(tie_list + total_blocks);
    if (speculative) {
        dim3 grid(blk_cap, nseg);
        { PbProfScope p(KIND == KIND_MEAN ? "k_ord_blocksum_mean" : "k_ord_blocksum_centered", st, false);
          k_ord_blocksum<KIND, W><<<grid, OB_THREADS, 0, st>>>(bufs[0], bufs[1], d_segs, d_stats, blk_cap, psum); }
        { PbProfScope p("k_ord_prefix", st, false);
          k_ord_prefix<C><<<dim3(C, nseg), 32, 0, st>>>(d_segs, blk_cap, psum); }
        PB_CUDA_OK(cudaMemsetAsync(tie_count, 0, sizeof(unsigned int), st));
        { PbProfScope p(KIND == KIND_MEAN ? "k_ord_summary_mean" : "k_ord_summary_centered", st);
          k_ord_summary<KIND, W><<<grid, OS_THREADS, 0, st>>>(bufs[0], bufs[1], d_segs, d_stats, blk_cap, psum, sum, tie_count, tie_list); }
        { PbProfScope p("k_ord_summary_tie", st, false);
          k_ord_summary_tie<KIND, W><<<148 * 4, OS_THREADS, 0, st>>>(bufs[0], bufs[1], d_segs, d_stats, blk_cap, psum, sum, tie_count, tie_list); }
    }
    { PbProfScope p(KIND == KIND_MEAN ? "k_ord_resolve_mean" : "k_ord_resolve_centered", st, !speculative);
      k_ord_resolve<KIND, W><<<nseg, 32 * C, 0, st>>>(bufs[0], bufs[1], d_segs, d_stats, blk_cap, sum, speculative); }
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
    return total_blocks * (7 * (sizeof(double) + sizeof(OrdSummary)) + sizeof(uint2)) + 256;
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
