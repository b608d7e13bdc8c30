// pb_ordered.cu - bit-exact LEFT-TO-RIGHT f64 sums at parallel speed.
//
// The reference's cluster statistics - weighted mean (array/matrix2D.c:200-233), centred
// covariance (math/pca.c:84-97) and distortion (quantize/cluster.c:135-148) - are naive
// sequential accumulations over up to N pixels in ascending pixel order.  The parity bar is
// bit-exact, and fl(fl(a+b)+c) != fl(a+fl(b+c)), so a tree / shuffle / atomic reduction is
// out.  A literal sequential chain costs one dependent DADD (~8 cycles) per pixel per pass:
// seconds per image.  This file gets the SAME BITS in parallel:
//
//   Observation.  While the running sum s stays inside one binade [2^e, 2^(e+1)), its ulp
//   q = 2^(e-52) is constant and s = M*q with M an integer in [2^52, 2^53).  Then
//       fl(s + a) = (M + rint(a/q)) * q            (exactly, unless a/q is a tie x.5)
//   i.e. sequential floating-point accumulation degenerates into INTEGER accumulation of the
//   terms quantised to q - and integer addition is associative.
//
//   Speculate, summarise, verify.
//     S1  k_ord_blocksum : plain (unordered) f64 sum of every block of OB elements, per chain.
//     S2  k_ord_prefix   : approximate running total at each block start -> guessed binade e.
//     S3  k_ord_summary  : per block and chain, with q = 2^(e-52): D = sum of rint(a/q) and the
//                          min / max over the block's in-order prefix sums (exact integers; an
//                          in-order monoid reduction), plus a flag if any term was a tie or too
//                          large to quantise.
//     S4  k_ord_resolve  : one lane per chain walks the BLOCKS in order holding the exact
//                          state (M, e).  A block is accepted iff the guess was right, no flag
//                          is set and 2^52 < M + min .. M + max < 2^53 (every intermediate value
//                          provably stayed in the binade); then M += D.  Otherwise the lane
//                          replays that one block element by element - the literal reference loop.
//   The guess only decides SPEED: every accepted block is proven equal to the sequential
//   result, every other block IS the sequential loop.  Binade crossings (~log2 n per chain),
//   ties (~2 ln n) and the first block take the slow path; everything else is parallel.
//
// Small clusters skip S1-S3 and run S4 in replay-only mode (one launch).
#include "pb_common.cuh"
#include "pb_kernels.h"
#include "pb_prof.h"

namespace {

constexpr int OB = 512;         // elements per summary block
constexpr int OB_THREADS = 128; // S1/S3: 4 elements per thread
constexpr int E_NOGUESS = 0x7fffffff;
constexpr double MAGIC = 6755399441055744.0;      // 1.5 * 2^52: (t + MAGIC) - MAGIC == rint(t) for |t| < 2^51
constexpr double TWO51 = 2251799813685248.0;
constexpr long long TWO52 = 1LL << 52, TWO53 = 1LL << 53;

// blocks accepted from their summary / blocks replayed sequentially (per chain), since last reset
__device__ unsigned long long g_ord_counts[8]; // 0 accepted, 1 replayed, 2 flag, 3 binade guess, 4 bounds, 5 replay rounds, 6 element-wise sub-chunks

// Quantised effect of a run of terms on the integer state M, for both parities of M at its start:
// total, and min / max over the in-order prefixes.  The parity only matters through ties: a term
// that lands exactly half-way (a/q = k + 0.5) is rounded to the EVEN neighbour, i.e. it adds k or
// k + 1 depending on whether M + (everything before it) + k is even - a two-state transducer whose
// composition is still associative.  After any element the parity of the state is
// (p + prefix sum) mod 2, so no extra field is needed.
struct Tri { double sum, mn, mx; };
struct Tri2 { Tri p[2]; };

struct OrdSummary {
    Tri2 t;   // quantised terms of the block (integers stored as doubles)
    int e;    // guessed binade of the running sum across this block
    int flag; // non-zero: replay the block
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

// ---- S2: approximate exclusive prefix per chain -> guessed binade ------------------------------
template <int C>
__global__ void __launch_bounds__(32) k_ord_prefix(const PbSeg *__restrict__ segs, uint32_t blk_cap,
                                                   const double *__restrict__ psum, OrdSummary *__restrict__ sum) {
    const int seg = blockIdx.y, c = blockIdx.x, lane = threadIdx.x;
    const uint32_t nblk = (segs[seg].n + OB - 1) / OB;
    const double *in = psum + (size_t)segs[seg].bbase * C + c;
    OrdSummary *out = sum + (size_t)segs[seg].bbase * C + c;
    const uint32_t per = (nblk + 31) / 32;
    const uint32_t b0 = min(lane * per, nblk), b1 = min(b0 + per, nblk);
    double s = 0.0;
    for (uint32_t b = b0; b < b1; b++) s += in[(size_t)b * C];
    double incl = s;
    for (int o = 1; o < 32; o <<= 1) {
        const double v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    double run = incl - s;
    for (uint32_t b = b0; b < b1; b++) {
        int e = E_NOGUESS;
        const double a = fabs(run);
        if (a > 1e-280 && a < 1e280) e = ilogb(a);
        out[(size_t)b * C].e = e;
        run += in[(size_t)b * C];
    }
}

// ---- S3: quantised block summaries --------------------------------------------------------------
// (sum, min prefix, max prefix) of a sequence of integers is a monoid under in-order concatenation:
//   (a ++ b).sum = a.sum + b.sum ; (a ++ b).mn = min(a.mn, a.sum + b.mn) ; likewise mx.
// Each thread owns 4 CONSECUTIVE elements, warps reduce in lane order, warp 0..3 in warp order, so
// mn / mx are the exact extremes of the running integer sum in element order.
__device__ __forceinline__ int dparity(double v) { return (int)((long long)v & 1LL); }

// in-order concatenation a ++ b
__device__ __forceinline__ Tri2 tri2_cat(const Tri2 &a, const Tri2 &b) {
    Tri2 r;
#pragma unroll
    for (int p = 0; p < 2; p++) {
        const Tri &x = a.p[p];
        const Tri &y = b.p[(p + dparity(x.sum)) & 1];
        r.p[p] = Tri{x.sum + y.sum, fmin(x.mn, x.sum + y.mn), fmax(x.mx, x.sum + y.mx)};
    }
    return r;
}

// append one term u = a / q to the run
__device__ __forceinline__ void tri2_push(Tri2 &t, double u, int &flag, bool first) {
    const double d = __dsub_rn(__dadd_rn(u, MAGIC), MAGIC); // rint(u), ties to even of u itself
    const double r = __dsub_rn(u, d);                       // exact remainder
    flag |= !(fabs(u) < TWO51);                             // unquantisable (or NaN)
    const bool tie = fabs(r) == 0.5;
    const double lo = tie ? floor(u) : d;                   // k  (u = k + 0.5)
#pragma unroll
    for (int p = 0; p < 2; p++) {
        Tri &x = t.p[p];
        // ties go to the neighbour that makes the state even: parity before = p + x.sum
        const double dd = (tie && (((p + dparity(x.sum) + dparity(lo)) & 1) != 0)) ? lo + 1.0 : lo;
        const double ps = x.sum + dd;
        x.mn = first ? ps : fmin(x.mn, ps);
        x.mx = first ? ps : fmax(x.mx, ps);
        x.sum = ps;
    }
}

// ---- S3a: the common case - no tie anywhere in the block: one parity-independent summary ----------
constexpr int OS_THREADS = 64; // 8 consecutive elements per thread, two warps per block
struct TriPlain { double sum, mn, mx; };
__device__ __forceinline__ TriPlain trip_cat(const TriPlain &a, const TriPlain &b) {
    return TriPlain{a.sum + b.sum, fmin(a.mn, a.sum + b.mn), fmax(a.mx, a.sum + b.mx)};
}

template <int KIND, bool W>
__global__ void __launch_bounds__(OS_THREADS) k_ord_summary(PbPlanes b0, PbPlanes b1, const PbSeg *__restrict__ segs,
                                                            const PbStats *__restrict__ stats, uint32_t blk_cap,
                                                            OrdSummary *__restrict__ sum,
                                                            unsigned int *__restrict__ tie_count,
                                                            uint2 *__restrict__ tie_list) {
    constexpr int C = NChains<KIND>::C;
    constexpr int PER = OB / OS_THREADS;
    __shared__ TriPlain s_tri[OS_THREADS / 32][C];
    __shared__ int s_flag[C];
    const int seg = blockIdx.y;
    const PbSeg sg = segs[seg];
    const uint32_t base = blockIdx.x * OB;
    if (base >= sg.n) return;
    const PbPlanes &P = sg.buf ? b1 : b0;
    OrdSummary *out = sum + ((size_t)sg.bbase + blockIdx.x) * C;
    double m0 = 0, m1 = 0, m2 = 0;
    if (KIND == KIND_CENTERED) { m0 = stats[seg].mean[0]; m1 = stats[seg].mean[1]; m2 = stats[seg].mean[2]; }
    double scale[C];
    TriPlain tri[C];
    int flag[C]; // bit 0: unusable (no guess / unquantisable term), bit 1: a tie -> needs the tie-aware pass
#pragma unroll
    for (int c = 0; c < C; c++) {
        const int e = out[c].e;
        flag[c] = e == E_NOGUESS;
        scale[c] = flag[c] ? 0.0 : scalbn(1.0, 52 - e); // 1 / q
        tri[c] = TriPlain{0.0, 1e300, -1e300};          // empty run
    }
    if (threadIdx.x < C) s_flag[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PER; k++) {
        const uint32_t i = base + threadIdx.x * PER + k; // consecutive elements per thread
        if (i < sg.n) {
            const size_t p = (size_t)sg.lo + i;
            double t[C];
            terms_all<KIND, W>(W ? P.w[p] : 1.0, P.c[0][p], P.c[1][p], P.c[2][p], m0, m1, m2, t);
#pragma unroll
            for (int c = 0; c < C; c++) {
                const double u = __dmul_rn(t[c], scale[c]);             // a / q, exact (power of two)
                const double d = __dsub_rn(__dadd_rn(u, MAGIC), MAGIC); // rint(u)
                flag[c] |= (!(fabs(u) < TWO51) ? 1 : 0) | (fabs(__dsub_rn(u, d)) == 0.5 ? 2 : 0);
                const double ps = tri[c].sum + d;
                tri[c] = TriPlain{ps, fmin(tri[c].mn, ps), fmax(tri[c].mx, ps)};
            }
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < C; c++) {
        if (flag[c]) atomicOr(&s_flag[c], flag[c]);
        TriPlain v = tri[c];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { // in-order tree: lane i absorbs lane i + o
            TriPlain r;
            r.sum = __shfl_down_sync(0xffffffffu, v.sum, o);
            r.mn = __shfl_down_sync(0xffffffffu, v.mn, o);
            r.mx = __shfl_down_sync(0xffffffffu, v.mx, o);
            if ((lane & (2 * o - 1)) == 0) v = trip_cat(v, r);
        }
        if (lane == 0) s_tri[warp][c] = v;
    }
    __syncthreads();
    if (threadIdx.x < C) {
        TriPlain v = s_tri[0][threadIdx.x];
        for (int w = 1; w < OS_THREADS / 32; w++) v = trip_cat(v, s_tri[w][threadIdx.x]);
        const Tri tt{v.sum, v.mn, v.mx};
        out[threadIdx.x].t.p[0] = tt;
        out[threadIdx.x].t.p[1] = tt;
        out[threadIdx.x].flag = s_flag[threadIdx.x]; // 0 ok, odd: replay, 2: redo with k_ord_summary_tie
    }
    if (threadIdx.x == 0) { // blocks holding a tie go on the work list of the tie-aware pass
        bool tie = false;
        for (int c = 0; c < C; c++) tie |= s_flag[c] == 2;
        if (tie) tie_list[atomicAdd(tie_count, 1u)] = make_uint2((unsigned)seg, blockIdx.x);
    }
}

// ---- S3b: blocks that contain a tie: both start parities (the two-state transducer) --------------
template <int KIND, bool W>
__global__ void __launch_bounds__(OB_THREADS) k_ord_summary_tie(PbPlanes b0, PbPlanes b1, const PbSeg *__restrict__ segs,
                                                                const PbStats *__restrict__ stats, uint32_t blk_cap,
                                                                OrdSummary *__restrict__ sum,
                                                                const unsigned int *__restrict__ tie_count,
                                                                const uint2 *__restrict__ tie_list) {
    constexpr int C = NChains<KIND>::C;
    constexpr int PER = OB / OB_THREADS;
    __shared__ Tri2 s_tri[OB_THREADS / 32][C];
    __shared__ int s_flag[C];
  for (unsigned int item = blockIdx.x; item < *tie_count; item += gridDim.x) { // persistent CTAs over the work list
    const int seg = (int)tie_list[item].x;
    const uint32_t blk = tie_list[item].y;
    const PbSeg sg = segs[seg];
    const uint32_t base = blk * OB;
    const PbPlanes &P = sg.buf ? b1 : b0;
    OrdSummary *out = sum + ((size_t)sg.bbase + blk) * C;
    __syncthreads();
    double m0 = 0, m1 = 0, m2 = 0;
    if (KIND == KIND_CENTERED) { m0 = stats[seg].mean[0]; m1 = stats[seg].mean[1]; m2 = stats[seg].mean[2]; }
    double scale[C];
    Tri2 tri[C];
    int flag[C];
    bool need[C]; // only the chains that actually hold a tie are redone (uniform across the CTA)
    bool any = false;
#pragma unroll
    for (int c = 0; c < C; c++) {
        const int e = out[c].e;
        need[c] = out[c].flag == 2;
        flag[c] = e == E_NOGUESS;
        scale[c] = flag[c] ? 0.0 : scalbn(1.0, 52 - e); // 1 / q
        tri[c].p[0] = tri[c].p[1] = Tri{0.0, 1e300, -1e300}; // empty run
    }
    if (threadIdx.x < C) s_flag[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PER; k++) {
        const uint32_t i = base + threadIdx.x * PER + k; // consecutive elements per thread
        if (i < sg.n) {
            const size_t p = (size_t)sg.lo + i;
            double t[C];
            terms_all<KIND, W>(W ? P.w[p] : 1.0, P.c[0][p], P.c[1][p], P.c[2][p], m0, m1, m2, t);
#pragma unroll
            for (int c = 0; c < C; c++)
                if (need[c]) tri2_push(tri[c], __dmul_rn(t[c], scale[c]) /* a / q, exact */, flag[c], !any);
            any = true;
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < C; c++) {
        if (!need[c]) continue;
        if (flag[c]) atomicOr(&s_flag[c], 1);
        Tri2 v = tri[c];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { // in-order tree: lane i absorbs lane i + o
            Tri2 r;
#pragma unroll
            for (int p = 0; p < 2; p++) {
                r.p[p].sum = __shfl_down_sync(0xffffffffu, v.p[p].sum, o);
                r.p[p].mn = __shfl_down_sync(0xffffffffu, v.p[p].mn, o);
                r.p[p].mx = __shfl_down_sync(0xffffffffu, v.p[p].mx, o);
            }
            if ((lane & (2 * o - 1)) == 0) v = tri2_cat(v, r);
        }
        if (lane == 0) s_tri[warp][c] = v;
    }
    __syncthreads();
    if (threadIdx.x < C && out[threadIdx.x].flag == 2) { // == need[threadIdx.x]
        Tri2 v = s_tri[0][threadIdx.x];
        for (int w = 1; w < OB_THREADS / 32; w++) v = tri2_cat(v, s_tri[w][threadIdx.x]);
        out[threadIdx.x].t = v;
        out[threadIdx.x].flag = s_flag[threadIdx.x];
    }
  }
}

// ---- S4: ordered resolve ---------------------------------------------------------------------------
// One CTA per cluster, one WARP per chain.  The warp's state is the exact running sum s (uniform
// across lanes).  Blocks are taken 32 at a time, lane b holding the summary of block b: an in-order
// warp scan of the block totals gives every lane the exact state its block would start from IF all
// earlier blocks are accepted, each lane validates its own block against that state, and a ballot
// finds the first block that cannot be accepted.  Everything before it is applied in one step; that
// block is replayed; the walk resumes behind it.
//
// Replaying a block uses the same idea one level down, now with the EXACT binade (s is known):
// each lane quantises its 16 consecutive elements, the warp scans / validates / ballots, accepted
// sub-chunks are applied at once and only the sub-chunk where the binade changes (or a tie sits) is
// added element by element - the literal reference loop, 16 elements long.
constexpr int SUB = OB / 32; // elements per lane in a replay

__device__ __forceinline__ long long warp_incl_scan(long long v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
    }
    return v;
}

// Applies the items [next, limit) held one per lane (usable iff ok) to the state s as far as they
// validate.  Returns the first index that does not (limit if all do) and updates s.  Item l checks
// its own prefix extremes against the exact state it would start from if everything before it is
// accepted; that state comes from an in-order warp scan of the monoid.  `validate` turns the scanned
// run of a lane into (usable, total).
__device__ __forceinline__ uint32_t finish_run(double &s, long long bits, long long M, bool negs, bool mine, bool ok,
                                              long long tot, double rmn, double rmx, uint32_t next, uint32_t limit) {
    const long long lo = rmn > 9e299 ? 0 : (long long)rmn, hi = rmx < -9e299 ? 0 : (long long)rmx;
    const long long vmin = negs ? M - hi : M + lo, vmax = negs ? M - lo : M + hi;
    // every prefix up to and including this item strictly inside (2^52, 2^53): the unrounded value
    // must itself stay inside the binade
    const bool valid = ok && vmin > TWO52 && vmax < TWO53;
    const unsigned fails = __ballot_sync(0xffffffffu, mine && !valid);
    const uint32_t f = fails ? (uint32_t)(__ffs(fails) - 1) : limit;
    if (f > next) {
        const long long acc = __shfl_sync(0xffffffffu, tot, (int)f - 1); // inclusive total of lane f - 1
        const long long M2 = negs ? M - acc : M + acc;
        s = __longlong_as_double((bits & 0xfff0000000000000LL) | (M2 & 0x000fffffffffffffLL));
    }
    return f;
}

// items whose effect does not depend on the start parity (no tie inside): a 3-double scan
__device__ __forceinline__ uint32_t apply_run_plain(double &s, int lane, uint32_t next, uint32_t limit, bool ok,
                                                   const TriPlain &item) {
    const long long bits = __double_as_longlong(s);
    const long long M = (bits & 0x000fffffffffffffLL) | TWO52; // |s| / q
    const bool mine = lane >= (int)next && lane < (int)limit;
    TriPlain run = (mine && ok) ? item : TriPlain{0.0, 1e300, -1e300};
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        TriPlain up;
        up.sum = __shfl_up_sync(0xffffffffu, run.sum, o);
        up.mn = __shfl_up_sync(0xffffffffu, run.mn, o);
        up.mx = __shfl_up_sync(0xffffffffu, run.mx, o);
        if (lane >= o) run = trip_cat(up, run);
    }
    return finish_run(s, bits, M, bits < 0, mine, ok, (long long)run.sum, run.mn, run.mx, next, limit);
}

// general items: the parity each one starts from depends on the items before it, so the warp scans
// the two-parity monoid
__device__ __forceinline__ uint32_t apply_run(double &s, int lane, uint32_t next, uint32_t limit, bool ok,
                                             const Tri2 &item) {
    const long long bits = __double_as_longlong(s);
    const long long M = (bits & 0x000fffffffffffffLL) | TWO52; // |s| / q
    const int p0 = (int)(M & 1LL);
    const bool mine = lane >= (int)next && lane < (int)limit;
    const bool parity_matters = mine && ok && (item.p[0].sum != item.p[1].sum || item.p[0].mn != item.p[1].mn ||
                                               item.p[0].mx != item.p[1].mx);
    if (!__any_sync(0xffffffffu, parity_matters))
        return apply_run_plain(s, lane, next, limit, ok, TriPlain{item.p[0].sum, item.p[0].mn, item.p[0].mx});
    Tri2 inc;
    if (mine && ok) inc = item;
    else inc.p[0] = inc.p[1] = Tri{0.0, 1e300, -1e300};
    Tri2 run = inc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        Tri2 up;
#pragma unroll
        for (int p = 0; p < 2; p++) {
            up.p[p].sum = __shfl_up_sync(0xffffffffu, run.p[p].sum, o);
            up.p[p].mn = __shfl_up_sync(0xffffffffu, run.p[p].mn, o);
            up.p[p].mx = __shfl_up_sync(0xffffffffu, run.p[p].mx, o);
        }
        if (lane >= o) run = tri2_cat(up, run);
    }
    // run.p[p0] = items next..lane applied to a state of parity p0: prefix extremes included
    return finish_run(s, bits, M, bits < 0, mine, ok, (long long)run.p[p0].sum, run.p[p0].mn, run.p[p0].mx, next, limit);
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
        sm.t.p[0] = sm.t.p[1] = Tri{0.0, 0.0, 0.0};
        sm.e = E_NOGUESS;
        sm.flag = 1;
        if (use_summaries && lane < (int)gcnt) sm = srow[(size_t)(g0 + lane) * C];
        uint32_t next = 0;
        while (next < gcnt) {
            const long long bits = __double_as_longlong(s);
            const int es = (int)((bits >> 52) & 0x7ff) - 1023;
            const bool ok = use_summaries && sm.flag == 0 && sm.e == es && es > -1000 && es < 1000;
            const uint32_t f = apply_run(s, lane, next, gcnt, ok, sm.t);
            n_acc += f - next;
            if (f < gcnt && use_summaries) {
                const int fl = __shfl_sync(0xffffffffu, sm.flag, (int)f), fe = __shfl_sync(0xffffffffu, sm.e, (int)f);
                if (lane == 0) {
                    if (fl) n_why[0]++;
                    else if (fe != es) n_why[1]++;
                    else n_why[2]++;
                }
            }
            if (f < gcnt) {
                const uint32_t base = (g0 + f) * OB;
                s = replay_block<KIND, W>(P, (size_t)sg.lo + base, min((uint32_t)OB, n - base), chain, m0, m1, m2, s, lane, hover);
                n_rep++;
                next = f + 1;
            } else {
                next = gcnt;
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
    const double bytes = 0; // set by the caller through pb_prof_next_bytes for the resolve kernel
    (void)bytes;
    if (speculative) {
        dim3 grid(blk_cap, nseg);
        { PbProfScope p(KIND == KIND_MEAN ? "k_ord_blocksum_mean" : "k_ord_blocksum_centered", st, false);
          k_ord_blocksum<KIND, W><<<grid, OB_THREADS, 0, st>>>(bufs[0], bufs[1], d_segs, d_stats, blk_cap, psum); }
        { PbProfScope p("k_ord_prefix", st, false);
          k_ord_prefix<C><<<dim3(C, nseg), 32, 0, st>>>(d_segs, blk_cap, psum, sum); }
        PB_CUDA_OK(cudaMemsetAsync(tie_count, 0, sizeof(unsigned int), st));
        { PbProfScope p(KIND == KIND_MEAN ? "k_ord_summary_mean" : "k_ord_summary_centered", st);
          k_ord_summary<KIND, W><<<grid, OS_THREADS, 0, st>>>(bufs[0], bufs[1], d_segs, d_stats, blk_cap, sum, tie_count, tie_list); }
        { PbProfScope p("k_ord_summary_tie", st, false);
          k_ord_summary_tie<KIND, W><<<148 * 4, OB_THREADS, 0, st>>>(bufs[0], bufs[1], d_segs, d_stats, blk_cap, sum, tie_count, tie_list); }
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
