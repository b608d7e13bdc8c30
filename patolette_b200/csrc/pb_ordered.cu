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
constexpr int OB_STRIDE = OB + 2;
constexpr int E_NOGUESS = 0x7fffffff;
constexpr double MAGIC = 6755399441055744.0;      // 1.5 * 2^52: (t + MAGIC) - MAGIC == rint(t) for |t| < 2^51
constexpr double TWO51 = 2251799813685248.0;
constexpr long long TWO52 = 1LL << 52, TWO53 = 1LL << 53;

// blocks accepted from their summary / blocks replayed sequentially (per chain), since last reset
__device__ unsigned long long g_ord_counts[2];

struct OrdSummary {
    double sum, mn, mx; // quantised terms of the block: total, min and max over its in-order prefixes
    int e;              // guessed binade of the running sum across this block
    int flag;           // non-zero: replay the block
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
    double *out = psum + ((size_t)seg * blk_cap + blockIdx.x) * C;
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
    const double *in = psum + (size_t)seg * blk_cap * C + c;
    OrdSummary *out = sum + (size_t)seg * blk_cap * C + c;
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
struct Tri { double sum, mn, mx; };
__device__ __forceinline__ Tri tri_cat(const Tri &a, const Tri &b) {
    return Tri{a.sum + b.sum, fmin(a.mn, a.sum + b.mn), fmax(a.mx, a.sum + b.mx)};
}

template <int KIND, bool W>
__global__ void __launch_bounds__(OB_THREADS) k_ord_summary(PbPlanes b0, PbPlanes b1, const PbSeg *__restrict__ segs,
                                                            const PbStats *__restrict__ stats, uint32_t blk_cap,
                                                            OrdSummary *__restrict__ sum) {
    constexpr int C = NChains<KIND>::C;
    constexpr int PER = OB / OB_THREADS;
    __shared__ Tri s_tri[OB_THREADS / 32][C];
    __shared__ int s_flag[C];
    const int seg = blockIdx.y;
    const PbSeg sg = segs[seg];
    const uint32_t base = blockIdx.x * OB;
    if (base >= sg.n) return;
    const PbPlanes &P = sg.buf ? b1 : b0;
    OrdSummary *out = sum + ((size_t)seg * blk_cap + blockIdx.x) * C;
    double m0 = 0, m1 = 0, m2 = 0;
    if (KIND == KIND_CENTERED) { m0 = stats[seg].mean[0]; m1 = stats[seg].mean[1]; m2 = stats[seg].mean[2]; }
    double scale[C];
    Tri tri[C];
    int flag[C];
#pragma unroll
    for (int c = 0; c < C; c++) {
        const int e = out[c].e;
        flag[c] = e == E_NOGUESS;
        scale[c] = flag[c] ? 0.0 : scalbn(1.0, 52 - e); // 1 / q
        tri[c] = Tri{0.0, 1e300, -1e300};                 // empty sequence
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
                const double r = __dsub_rn(u, d);                       // exact remainder
                flag[c] |= !(fabs(u) < TWO51) | (fabs(r) == 0.5);       // unquantisable / tie / NaN
                const double ps = tri[c].sum + d;
                tri[c] = Tri{ps, fmin(tri[c].mn, ps), fmax(tri[c].mx, ps)};
            }
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < C; c++) {
        if (flag[c]) atomicOr(&s_flag[c], 1);
        Tri v = tri[c];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { // in-order tree: lane i absorbs lane i + o
            Tri r;
            r.sum = __shfl_down_sync(0xffffffffu, v.sum, o);
            r.mn = __shfl_down_sync(0xffffffffu, v.mn, o);
            r.mx = __shfl_down_sync(0xffffffffu, v.mx, o);
            if ((lane & (2 * o - 1)) == 0) v = tri_cat(v, r);
        }
        if (lane == 0) s_tri[warp][c] = v;
    }
    __syncthreads();
    if (threadIdx.x < C) {
        Tri v = s_tri[0][threadIdx.x];
        for (int w = 1; w < OB_THREADS / 32; w++) v = tri_cat(v, s_tri[w][threadIdx.x]);
        out[threadIdx.x].sum = v.sum;
        out[threadIdx.x].mn = v.mn;
        out[threadIdx.x].mx = v.mx;
        out[threadIdx.x].flag = s_flag[threadIdx.x];
    }
}

// ---- S4: ordered resolve (and the plain sequential chain when use_summaries == false) ---------------
template <int KIND, bool W>
__global__ void __launch_bounds__(32) k_ord_resolve(PbPlanes b0, PbPlanes b1, const PbSeg *__restrict__ segs,
                                                    PbStats *__restrict__ stats, uint32_t blk_cap,
                                                    const OrdSummary *__restrict__ sum, bool use_summaries) {
    constexpr int C = NChains<KIND>::C;
    __shared__ double tile[4][OB_STRIDE];
    const int seg = blockIdx.x, lane = threadIdx.x;
    const PbSeg sg = segs[seg];
    const PbPlanes &P = sg.buf ? b1 : b0;
    const uint32_t n = sg.n, nblk = (n + OB - 1) / OB;
    double m0 = 0, m1 = 0, m2 = 0;
    if (KIND == KIND_CENTERED) { m0 = stats[seg].mean[0]; m1 = stats[seg].mean[1]; m2 = stats[seg].mean[2]; }
    const bool chain = lane < C;
    // exact state of this lane's chain: s = M * 2^(e-52) while in integer mode, else the double s
    double s = 0.0;
    unsigned int n_acc = 0, n_rep = 0;
    constexpr int GRP = 32; // summaries of GRP blocks are staged in shared memory at a time
    __shared__ OrdSummary s_sum[GRP * C];
    const OrdSummary *sbase = sum + (size_t)seg * blk_cap * C;
    for (uint32_t b = 0; b < nblk; b++) {
        bool accept = false;
        if (use_summaries && (b % GRP) == 0) {
            __syncwarp();
            const uint32_t cnt = min((uint32_t)GRP, nblk - b) * C * (sizeof(OrdSummary) / 8);
            const unsigned long long *src = (const unsigned long long *)(sbase + (size_t)b * C);
            unsigned long long *dst = (unsigned long long *)s_sum;
            for (uint32_t i = lane; i < cnt; i += 32) dst[i] = src[i];
            __syncwarp();
        }
        if (use_summaries && chain) {
            const OrdSummary sm = s_sum[(b % GRP) * C + lane];
            const long long bits = __double_as_longlong(s);
            const int es = (int)((bits >> 52) & 0x7ff) - 1023;
            if (sm.flag == 0 && es == sm.e && es > -1000) {
                const long long M = (bits & 0x000fffffffffffffLL) | TWO52; // |s| / q
                const bool negs = bits < 0;
                // |s| / q after k elements = M + prefix_k (s > 0) or M - prefix_k (s < 0)
                const long long d = (long long)sm.sum, lo = (long long)sm.mn, hi = (long long)sm.mx;
                const long long vmin = negs ? M - hi : M + lo, vmax = negs ? M - lo : M + hi;
                // strictly above 2^52: the unrounded M + t must itself stay inside the binade
                if (vmin > TWO52 && vmax < TWO53) {
                    const long long M2 = negs ? M - d : M + d;
                    s = __longlong_as_double((bits & 0xfff0000000000000LL) | (M2 & 0x000fffffffffffffLL));
                    accept = true;
                }
            }
        }
        const bool need = chain && !accept;
        n_acc += accept;
        n_rep += need;
        if (__any_sync(0xffffffffu, need)) {
            // replay this block sequentially for the lanes that need it (the literal reference loop)
            const uint32_t base = b * OB, cnt = min((uint32_t)OB, n - base);
            __syncwarp();
            for (uint32_t i = lane; i < cnt; i += 32) {
                const size_t p = (size_t)sg.lo + base + i;
                tile[0][i] = W ? P.w[p] : 1.0;
                tile[1][i] = P.c[0][p];
                tile[2][i] = P.c[1][p];
                tile[3][i] = P.c[2][p];
            }
            __syncwarp();
            if (need) {
                uint32_t i = 0;
                for (; i + 4 <= cnt; i += 4) {
                    const double t0 = term_one<KIND, W>(lane, tile[0][i], tile[1][i], tile[2][i], tile[3][i], m0, m1, m2);
                    const double t1 = term_one<KIND, W>(lane, tile[0][i + 1], tile[1][i + 1], tile[2][i + 1], tile[3][i + 1], m0, m1, m2);
                    const double t2 = term_one<KIND, W>(lane, tile[0][i + 2], tile[1][i + 2], tile[2][i + 2], tile[3][i + 2], m0, m1, m2);
                    const double t3 = term_one<KIND, W>(lane, tile[0][i + 3], tile[1][i + 3], tile[2][i + 3], tile[3][i + 3], m0, m1, m2);
                    s = __dadd_rn(s, t0);
                    s = __dadd_rn(s, t1);
                    s = __dadd_rn(s, t2);
                    s = __dadd_rn(s, t3);
                }
                for (; i < cnt; i++)
                    s = __dadd_rn(s, term_one<KIND, W>(lane, tile[0][i], tile[1][i], tile[2][i], tile[3][i], m0, m1, m2));
            }
        }
    }
    if (chain && use_summaries) {
        atomicAdd(&g_ord_counts[0], (unsigned long long)n_acc);
        atomicAdd(&g_ord_counts[1], (unsigned long long)n_rep);
    }
    if (KIND == KIND_MEAN) {
        // matrix2D.c:230-231: s = 1 / wsum (1 / rows when unweighted); mean *= s
        double wsum = __shfl_sync(0xffffffffu, s, 0);
        if (!W) wsum = (double)n;
        const double inv = 1.0 / wsum;
        if (lane == 0) stats[seg].wsum = wsum;
        if (lane >= 1 && lane <= 3) stats[seg].mean[lane - 1] = __dmul_rn(s, inv);
    } else {
        if (lane < 6) stats[seg].cov[lane] = s;
        if (lane == 6) stats[seg].dist = s;
    }
}

template <int KIND, bool W>
void launch_pass(const PbPlanes bufs[2], const PbSeg *d_segs, int nseg, uint32_t max_n, PbStats *d_stats,
                 void *d_scratch, size_t scratch_bytes, cudaStream_t st) {
    constexpr int C = NChains<KIND>::C;
    const uint32_t blk_cap = (max_n + OB - 1) / OB;
    const size_t need = (size_t)nseg * blk_cap * C * (sizeof(double) + sizeof(OrdSummary));
    const bool speculative = max_n >= 8 * OB && d_scratch && need <= scratch_bytes;
    double *psum = (double *)d_scratch;
    OrdSummary *sum = (OrdSummary *)((char *)d_scratch + (size_t)nseg * blk_cap * C * sizeof(double));
    const double bytes = 0; // set by the caller through pb_prof_next_bytes for the resolve kernel
    (void)bytes;
    if (speculative) {
        dim3 grid(blk_cap, nseg);
        { PbProfScope p(KIND == KIND_MEAN ? "k_ord_blocksum_mean" : "k_ord_blocksum_centered", st, false);
          k_ord_blocksum<KIND, W><<<grid, OB_THREADS, 0, st>>>(bufs[0], bufs[1], d_segs, d_stats, blk_cap, psum); }
        { PbProfScope p("k_ord_prefix", st, false);
          k_ord_prefix<C><<<dim3(C, nseg), 32, 0, st>>>(d_segs, blk_cap, psum, sum); }
        { PbProfScope p(KIND == KIND_MEAN ? "k_ord_summary_mean" : "k_ord_summary_centered", st);
          k_ord_summary<KIND, W><<<grid, OB_THREADS, 0, st>>>(bufs[0], bufs[1], d_segs, d_stats, blk_cap, sum); }
    }
    { PbProfScope p(KIND == KIND_MEAN ? "k_ord_resolve_mean" : "k_ord_resolve_centered", st, !speculative);
      k_ord_resolve<KIND, W><<<nseg, 32, 0, st>>>(bufs[0], bufs[1], d_segs, d_stats, blk_cap, sum, speculative); }
    PB_CUDA_OK(cudaGetLastError());
}

} // namespace

void pb_ordered_counts(unsigned long long out[2], bool reset) {
    PB_CUDA_OK(cudaMemcpyFromSymbol(out, g_ord_counts, sizeof(unsigned long long) * 2));
    if (reset) {
        unsigned long long z[2] = {0, 0};
        PB_CUDA_OK(cudaMemcpyToSymbol(g_ord_counts, z, sizeof z));
    }
}

size_t pb_ordered_scratch_bytes(int nseg, uint32_t max_n) {
    const size_t blk_cap = ((size_t)max_n + OB - 1) / OB;
    return (size_t)nseg * blk_cap * 7 * (sizeof(double) + sizeof(OrdSummary)) + 256;
}

void pb_launch_pass_mean(const PbPlanes bufs[2], const PbSeg *d_segs, int nseg, uint32_t max_n, bool weighted,
                         PbStats *d_stats, void *d_scratch, size_t scratch_bytes, cudaStream_t st) {
    if (nseg <= 0) return;
    if (weighted) launch_pass<KIND_MEAN, true>(bufs, d_segs, nseg, max_n, d_stats, d_scratch, scratch_bytes, st);
    else launch_pass<KIND_MEAN, false>(bufs, d_segs, nseg, max_n, d_stats, d_scratch, scratch_bytes, st);
}

void pb_launch_pass_centered(const PbPlanes bufs[2], const PbSeg *d_segs, int nseg, uint32_t max_n, bool weighted,
                             PbStats *d_stats, void *d_scratch, size_t scratch_bytes, cudaStream_t st) {
    if (nseg <= 0) return;
    if (weighted) launch_pass<KIND_CENTERED, true>(bufs, d_segs, nseg, max_n, d_stats, d_scratch, scratch_bytes, st);
    else launch_pass<KIND_CENTERED, false>(bufs, d_segs, nseg, max_n, d_stats, d_scratch, scratch_bytes, st);
}
