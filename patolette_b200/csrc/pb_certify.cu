// pb_certify.cu - the split of a cluster WITHOUT the bucket sort: a certified argmax.
//
// get_optimal_bucket_index (quantize/local.c:102-177) sums w*c per bucket left to right, turns the 512 sums into
// prefixes, evaluates objective_i = sum_j csl_j^2/sl + csr_j^2/sr and returns ONE integer: the first maximum.  The
// exact route reproduces every rounding of those sums (stable 512-class bucket sort of the payload + one sequential
// chain per bucket: pb_parallel.cu / pb_chain.cu) - a third of the split loop's time for an index that almost never
// depends on the last bits.  Here the sums are formed in ANY order (warp-private shared-memory tables, no sort), and
// a second kernel PROVES that the reference's argmax is the one found:
//
//   * the reference's prefix P_ij and ours P'_ij are both roundings of the same exact sum of the same terms
//     fl(c*w): |P - exact| <= gamma_d * sum|t| with d the number of additions a term takes part in (Higham, Accuracy
//     and Stability, section 4.2: any order).  d <= n_b + 511 for the reference, <= 2 n_b + 600 here, and
//     sum|t| <= max|c_j| * sum w, so P_ij lies in [P'_ij - E_ij, P'_ij + E_ij] with E_ij computed below;
//   * every later operation of the reference (T - csl, squares, divisions, the five additions) is a correctly
//     rounded, monotone function of its operands: interval arithmetic with outward rounding (__dadd_rd / _ru ...)
//     encloses the reference's objective_i in [lo_i, hi_i];
//   * an empty bucket's objective equals its predecessor's bit for bit (same prefix, same size), so it can never be
//     the FIRST maximum: candidates are the non-empty buckets;
//   * if lo_best > hi_i for every other candidate, the reference's first maximum is `best` - certified.  Otherwise
//     the cluster is re-evaluated through the exact route (pb_pipeline.cu: eval_split), so the result never depends
//     on the certificate, only the time does.
//   * weighted runs: bucket sizes are size_t += double (local.c:133), i.e. trunc(fl(s + w)) per step = s + floor(w)
//     unless the fraction of w is within 2^(E-52) of 1 (s + w < 2^E): such "risky" pixels are counted and widen the
//     size intervals; negative / non-finite weights refuse the certificate.
#include "pb_common.cuh"
#include "pb_kernels.h"
#include "pb_prof.h"

namespace {

constexpr unsigned FULL = 0xffffffffu;
#ifndef PB_BH_WARPS
#define PB_BH_WARPS 4
#endif
#ifndef PB_BH_ROWS
#define PB_BH_ROWS 4
#endif
constexpr int BH_WARPS = PB_BH_WARPS, BH_THREADS = 32 * BH_WARPS;
constexpr int BH_ROWS = PB_BH_ROWS;        // rows of 32 pixels whose loads are in flight together
constexpr int BH_WARP_PX = 2048;           // pixels of a CTA visit per warp
constexpr int BH_CHUNK = BH_WARPS * BH_WARP_PX;

// One table per WARP (no atomics: ATOMS costs 2 cycles per lane, and 64-bit shared-memory atomics are CAS loops).
// Bucket b owns 32 bytes, two 16-byte halves {s0, s1} and {s2, count}; the halves of every other group of four
// buckets are swapped, so that a warp's 16-byte accesses to 32 random buckets spread over all eight bank quads.
// tag: one byte per bucket and row of the batch - which lane updates the bucket this round (see below).
template <bool W>
struct WarpTab {
    double2 m[2 * PB_BUCKETS];
    unsigned long long sz[W ? PB_BUCKETS : 1];
    uint32_t risky[W ? PB_BUCKETS : 1];
    uint8_t tag[BH_ROWS][PB_BUCKETS];
};

__device__ unsigned long long g_certify_counts[4]; // certified, refused, (spare)

// ---- bucket ids (sort.c:61-87, as k_buckets) + per-bucket sums in any order ------------------------------------
// A warp takes rows of 32 consecutive pixels.  Lanes of a row that share a bucket must not race on the table.
// (__match_any_sync would name them, but MATCH.ANY costs ~4 cycles per lane on B200: ncu showed the first version
// of this kernel waiting on it more than on HBM.)  Instead every lane writes its lane number into the bucket's tag
// byte and reads it back: exactly one lane per bucket finds itself - it updates the table; the others try again
// among themselves (one more round covers pairs), and what is left after two rounds (three or more pixels of a
// row in one bucket: flat image regions) is reduced bucket by bucket with warp shuffles.  Per batch of BH_ROWS rows
// the buckets and tags of all rows are formed first (independent work), then the updates run back to back.
template <bool W>
__global__ void __launch_bounds__(BH_THREADS) k_buckets_hist(PbPlanes b0, PbPlanes b1, const PbSeg *__restrict__ segs,
                                                             const double *__restrict__ axes, PbSplit *__restrict__ sp,
                                                             uint16_t *__restrict__ bucket, PbHist *__restrict__ hist) {
    extern __shared__ __align__(16) unsigned char bh_raw[];
    WarpTab<W> *tabs = reinterpret_cast<WarpTab<W> *>(bh_raw);
    const int seg = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const PbSeg sg = segs[seg];
    if ((size_t)blockIdx.x * BH_CHUNK >= sg.n) return; // CTA-uniform
    {
        uint32_t *z = reinterpret_cast<uint32_t *>(bh_raw);
        for (int i = tid; i < (int)(sizeof(WarpTab<W>) * BH_WARPS / 4); i += BH_THREADS) z[i] = 0u;
    }
    __syncthreads();
    WarpTab<W> &T = tabs[warp];
    const PbPlanes &P = sg.buf ? b1 : b0;
    const double x0 = axes[seg * 3], x1 = axes[seg * 3 + 1], x2 = axes[seg * 3 + 2];
    const double *c0 = P.c[0] + sg.lo, *c1 = P.c[1] + sg.lo, *c2 = P.c[2] + sg.lo;
    const double *cw = W ? P.w + sg.lo : nullptr;
    uint16_t *bk = bucket + sg.lo;
    const double mn = pb_ord_decode(sp[seg].mn_enc), mx = pb_ord_decode(sp[seg].mx_enc);
    const bool degenerate = __dsub_rn(mx, mn) < PB_DELTA;
    const double s = 1.0 / __dsub_rn(mx, mn);
    const uint32_t tail0 = sg.n - (sg.n & 3u);
    if (blockIdx.x == 0 && tid == 0) sp[seg].degenerate = degenerate;
    // fraction of w from which fl(s + w) may reach the next integer (s + w < 2^E, E in PbSeg::pad)
    const double thr = W ? 1.0 - scalbn(1.0, (int)sg.pad - 52) : 0.0;
    double am0 = 0.0, am1 = 0.0, am2 = 0.0;
    bool bad = false;

    auto rmw = [&](uint32_t b, double t0, double t1, double t2, uint32_t c, unsigned long long z, uint32_t rk) {
        const uint32_t sw = (b >> 2) & 1u;
        double2 *p = T.m + 2 * b;
        double2 A = p[sw], B = p[sw ^ 1u];
        A.x += t0;
        A.y += t1;
        B.x += t2;
        B.y = __longlong_as_double(__double_as_longlong(B.y) + (long long)c);
        p[sw] = A;
        p[sw ^ 1u] = B;
        if (W) { T.sz[b] += z; T.risky[b] += rk; }
    };

    // The warp walks batches of BH_ROWS rows: sub-chunks of BH_WARP_PX pixels inside the CTA's chunk, chunk after
    // chunk.  The loads of the NEXT batch are issued before the current one is worked on (register double buffer):
    // with only 12 warps per SM (the tables fill shared memory) the kernel cannot hide HBM latency by occupancy.
    double v0[BH_ROWS], v1[BH_ROWS], v2[BH_ROWS], vw[BH_ROWS];
    double n0[BH_ROWS], n1[BH_ROWS], n2[BH_ROWS], nw[BH_ROWS];
    auto load = [&](uint32_t r0, uint32_t e, double (&a0)[BH_ROWS], double (&a1)[BH_ROWS], double (&a2)[BH_ROWS], double (&aw)[BH_ROWS]) {
#pragma unroll
        for (int r = 0; r < BH_ROWS; r++) {
            const uint32_t i = r0 + r * 32 + lane;
            const bool valid = i < e;
            a0[r] = valid ? c0[i] : 0.0;
            a1[r] = valid ? c1[i] : 0.0;
            a2[r] = valid ? c2[i] : 0.0;
            aw[r] = (W && valid) ? cw[i] : 1.0;
        }
    };
    uint32_t cbase = blockIdx.x * BH_CHUNK;
    uint32_t row0 = cbase + warp * BH_WARP_PX;
    uint32_t wend = min(row0 + (uint32_t)BH_WARP_PX, sg.n);
    bool have = row0 < sg.n;
    if (have) load(row0, wend, v0, v1, v2, vw);
    while (have) { // (warp-uniform)
        {
            uint32_t nrow = row0 + 32 * BH_ROWS, nend = wend, nbase = cbase;
            if (nrow >= wend) {
                nbase = cbase + gridDim.x * BH_CHUNK;
                nrow = nbase + warp * BH_WARP_PX;
                nend = min(nrow + (uint32_t)BH_WARP_PX, sg.n);
            }
            const bool nhave = nrow < sg.n;
            if (nhave) load(nrow, nend, n0, n1, n2, nw);
            // ---- phase A: buckets and tags of every row of the batch ----
            uint32_t bb[BH_ROWS]; // bucket | 0x80000000 while this lane's pixel still has to reach the table
            unsigned long long zz[W ? BH_ROWS : 1];
            uint32_t rr[W ? BH_ROWS : 1];
#pragma unroll
            for (int r = 0; r < BH_ROWS; r++) {
                const uint32_t i = row0 + r * 32 + lane;
                const bool valid = i < wend;
                uint32_t b;
                if (degenerate) {
                    b = i % PB_BUCKETS; // sort.c:66-75 round-robin
                } else {
                    const double dot = pb_dgemv_row3(v0[r], v1[r], v2[r], x0, x1, x2, i >= tail0);
                    const double ratio = __dmul_rn(__dsub_rn(dot, mn), s);
                    const unsigned long long q = (unsigned long long)__dmul_rn((double)PB_BUCKETS, ratio);
                    b = q < PB_BUCKETS - 1 ? (uint32_t)q : PB_BUCKETS - 1;
                }
                am0 = fmax(am0, fabs(v0[r])); am1 = fmax(am1, fabs(v1[r])); am2 = fmax(am2, fabs(v2[r]));
                if (W) {
                    const double w = vw[r];
                    v0[r] = __dmul_rn(v0[r], w); v1[r] = __dmul_rn(v1[r], w); v2[r] = __dmul_rn(v2[r], w); // local.c:130-132
                    const bool okw = w >= 0.0 && w < 9.0e15;
                    bad |= valid && !okw;
                    const double fl = okw ? floor(w) : 0.0;
                    zz[r] = (unsigned long long)fl;
                    rr[r] = (okw && (w - fl) >= thr) ? 1u : 0u;
                }
                if (valid) {
                    bk[i] = (uint16_t)b;
                    T.tag[r][b] = (uint8_t)lane;
                }
                bb[r] = b | (valid ? 0x80000000u : 0u);
            }
            __syncwarp();
            bool win[BH_ROWS];
#pragma unroll
            for (int r = 0; r < BH_ROWS; r++) win[r] = (bb[r] >> 31) && T.tag[r][bb[r] & 0xffffu] == (uint8_t)lane;
            // ---- phase B: round 1 - one lane per bucket and row updates the table ----
            bool pending = false;
#pragma unroll
            for (int r = 0; r < BH_ROWS; r++) {
                if (win[r]) {
                    rmw(bb[r] & 0xffffu, v0[r], v1[r], v2[r], 1u, W ? zz[W ? r : 0] : 0ull, W ? rr[W ? r : 0] : 0u);
                    bb[r] &= 0xffffu;
                }
                pending |= (bb[r] >> 31) != 0u;
                __syncwarp();
            }
            if (__any_sync(FULL, pending)) {
                // ---- round 2 among the lanes that lost round 1 (covers every bucket with exactly two pixels in a row)
#pragma unroll
                for (int r = 0; r < BH_ROWS; r++)
                    if (bb[r] >> 31) T.tag[r][bb[r] & 0xffffu] = (uint8_t)lane;
                __syncwarp();
#pragma unroll
                for (int r = 0; r < BH_ROWS; r++) win[r] = (bb[r] >> 31) && T.tag[r][bb[r] & 0xffffu] == (uint8_t)lane;
                pending = false;
#pragma unroll
                for (int r = 0; r < BH_ROWS; r++) {
                    if (win[r]) {
                        rmw(bb[r] & 0xffffu, v0[r], v1[r], v2[r], 1u, W ? zz[W ? r : 0] : 0ull, W ? rr[W ? r : 0] : 0u);
                        bb[r] &= 0xffffu;
                    }
                    pending |= (bb[r] >> 31) != 0u;
                    __syncwarp();
                }
                if (__any_sync(FULL, pending)) {
                    // ---- the rest, bucket by bucket: one warp reduction per bucket that still has pixels waiting
#pragma unroll
                    for (int r = 0; r < BH_ROWS; r++) {
                        unsigned rem = __ballot_sync(FULL, (bb[r] >> 31) != 0u);
                        while (rem) {
                            const int L = __ffs(rem) - 1;
                            const uint32_t bL = __shfl_sync(FULL, bb[r], L);
                            const bool in = bb[r] == bL; // (flag bit included: waiting lanes of that bucket)
                            const unsigned mask = __ballot_sync(FULL, in);
                            double a0 = in ? v0[r] : 0.0, a1 = in ? v1[r] : 0.0, a2 = in ? v2[r] : 0.0;
                            unsigned long long az = (W && in) ? zz[W ? r : 0] : 0ull;
#pragma unroll
                            for (int o = 16; o; o >>= 1) {
                                a0 += __shfl_xor_sync(FULL, a0, o);
                                a1 += __shfl_xor_sync(FULL, a1, o);
                                a2 += __shfl_xor_sync(FULL, a2, o);
                                if (W) az += __shfl_xor_sync(FULL, az, o);
                            }
                            const uint32_t ark = W ? __reduce_add_sync(FULL, in ? rr[W ? r : 0] : 0u) : 0u;
                            if (lane == L) rmw(bL & 0xffffu, a0, a1, a2, (uint32_t)__popc(mask), az, ark);
                            __syncwarp();
                            rem &= ~mask;
                        }
                    }
                }
            }
            // hand over to the prefetched batch
#pragma unroll
            for (int r = 0; r < BH_ROWS; r++) { v0[r] = n0[r]; v1[r] = n1[r]; v2[r] = n2[r]; vw[r] = nw[r]; }
            row0 = nrow; wend = nend; cbase = nbase; have = nhave;
        }
    }
    __syncthreads();
    PbHist &H = hist[seg];
    for (int b = tid; b < PB_BUCKETS; b += BH_THREADS) {
        uint32_t rk = 0;
        unsigned long long z = 0;
        long long c = 0;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        const uint32_t sw = (b >> 2) & 1u;
#pragma unroll
        for (int w = 0; w < BH_WARPS; w++) {
            const double2 A = tabs[w].m[2 * b + sw], B = tabs[w].m[2 * b + (sw ^ 1u)];
            a0 += A.x; a1 += A.y; a2 += B.x;
            c += __double_as_longlong(B.y);
            if (W) { z += tabs[w].sz[b]; rk += tabs[w].risky[b]; }
        }
        if (c) {
            atomicAdd(&H.s[0][b], a0); atomicAdd(&H.s[1][b], a1); atomicAdd(&H.s[2][b], a2);
            atomicAdd(&H.cnt[b], (uint32_t)c);
            if (W) { atomicAdd(&H.sz[b], z); if (rk) atomicAdd(&H.risky[b], rk); }
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        am0 = fmax(am0, __shfl_xor_sync(FULL, am0, o));
        am1 = fmax(am1, __shfl_xor_sync(FULL, am1, o));
        am2 = fmax(am2, __shfl_xor_sync(FULL, am2, o));
    }
    if (lane == 0) { // non-negative doubles order like their bit patterns
        atomicMax(&H.absmax[0], (unsigned long long)__double_as_longlong(am0));
        atomicMax(&H.absmax[1], (unsigned long long)__double_as_longlong(am1));
        atomicMax(&H.absmax[2], (unsigned long long)__double_as_longlong(am2));
    }
    if (W && __any_sync(FULL, bad) && lane == 0) atomicOr(&H.bad, 1u);
}

// ---- the certificate -------------------------------------------------------------------------------------------
// inclusive prefix of 512 values by one warp: 16 per lane + a warp scan (our own prefix may use any order)
template <typename T>
__device__ __forceinline__ void warp_prefix512(T *a, int lane) {
    T v[16], run = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) { run += a[lane * 16 + k]; v[k] = run; }
    T incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const T u = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += u;
    }
    const T off = incl - run;
#pragma unroll
    for (int k = 0; k < 16; k++) a[lane * 16 + k] = v[k] + off;
}

struct Itv { double lo, hi; };
__device__ __forceinline__ Itv itv_sq(Itv x) { // x * x
    if (x.lo >= 0.0) return Itv{__dmul_rd(x.lo, x.lo), __dmul_ru(x.hi, x.hi)};
    if (x.hi <= 0.0) return Itv{__dmul_rd(x.hi, x.hi), __dmul_ru(x.lo, x.lo)};
    return Itv{0.0, fabs(x.lo) > fabs(x.hi) ? __dmul_ru(x.lo, x.lo) : __dmul_ru(x.hi, x.hi)};
}
__device__ __forceinline__ Itv itv_div_pos(Itv num /* >= 0 */, Itv den /* > 0 */) {
    return Itv{__ddiv_rd(num.lo, den.hi), __ddiv_ru(num.hi, den.lo)};
}
__device__ __forceinline__ Itv itv_add(Itv a, Itv b) { return Itv{__dadd_rd(a.lo, b.lo), __dadd_ru(a.hi, b.hi)}; }

template <bool W>
__global__ void __launch_bounds__(PB_BUCKETS) k_split_certify(const PbHist *__restrict__ hist, PbSplit *__restrict__ sp,
                                                              int distrust) {
    __shared__ double P[3][PB_BUCKETS];
    __shared__ unsigned long long C[PB_BUCKETS], Z[PB_BUCKETS], R[PB_BUCKETS];
    __shared__ double s_lo[PB_BUCKETS];
    __shared__ int s_loc[PB_BUCKETS];
    __shared__ unsigned s_red[PB_BUCKETS / 32];
    __shared__ int s_ok;
    const int seg = blockIdx.x, i = threadIdx.x, warp = i >> 5, lane = i & 31;
    const PbHist &H = hist[seg];
    const uint32_t cnt_i = H.cnt[i];
    P[0][i] = H.s[0][i]; P[1][i] = H.s[1][i]; P[2][i] = H.s[2][i];
    C[i] = cnt_i;
    Z[i] = W ? H.sz[i] : 0ull;
    R[i] = W ? (unsigned long long)H.risky[i] : 0ull;
    // largest bucket
    unsigned nbmax = __reduce_max_sync(FULL, cnt_i);
    if (lane == 0) s_red[warp] = nbmax;
    if (i == 0) s_ok = 1;
    __syncthreads();
    nbmax = 0;
#pragma unroll
    for (int w = 0; w < PB_BUCKETS / 32; w++) nbmax = max(nbmax, s_red[w]);
    if (warp < 3) warp_prefix512(P[warp], lane);
    else if (warp == 3) warp_prefix512(C, lane);
    else if (W && warp == 4) warp_prefix512(Z, lane);
    else if (W && warp == 5) warp_prefix512(R, lane);
    __syncthreads();
    // g = (d_ref + d_ours) * u * 1.01, d_ref <= nbmax + 511, d_ours <= 2 nbmax + 600; u = 2^-53
    const double g = __dmul_ru(__dmul_ru(__dadd_ru(__dmul_ru(3.0, (double)nbmax), 1200.0), 1.01), 0x1p-53);
    const unsigned long long Ctot = C[PB_BUCKETS - 1], Ztot = Z[PB_BUCKETS - 1], Rtot = R[PB_BUCKETS - 1];
    bool refuse = distrust != 0 || (W && H.bad != 0);
    if (W && (Ztot + Rtot + Ctot) >= (1ull << 52)) refuse = true;
    // sizes (local.c:144-146, :153-154): exact counts, or [sum floor(w), + risky pixels]
    Itv sl, sr, wl, wr_all; // wl: upper bound of sum w over buckets <= i (only .hi is used)
    if (W) {
        sl = Itv{(double)Z[i], (double)(Z[i] + R[i])};
        sr = Itv{(double)(Ztot - Z[i]), (double)((Ztot - Z[i]) + (Rtot - R[i]))};
        wl = Itv{0.0, (double)(Z[i] + C[i])};
        wr_all = Itv{0.0, (double)(Ztot + Ctot)};
    } else {
        sl = Itv{(double)C[i], (double)C[i]};
        sr = Itv{(double)(Ctot - C[i]), (double)(Ctot - C[i])};
        wl = sl;
        wr_all = Itv{0.0, (double)Ctot};
    }
    // a size interval that contains zero and something else: the reference's branch (local.c:157,161) is unknown
    if ((sl.lo == 0.0) != (sl.hi == 0.0) || (sr.lo == 0.0) != (sr.hi == 0.0)) refuse = refuse || (cnt_i > 0 || i == 0);
    Itv o{0.0, 0.0};
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const double M = __dmul_ru(__longlong_as_double((long long)H.absmax[j]), 1.0000001);
        const double gm = __dmul_ru(g, M);
        const double e_i = __dmul_ru(gm, wl.hi), e_t = __dmul_ru(gm, wr_all.hi);
        const Itv csl{__dsub_rd(P[j][i], e_i), __dadd_ru(P[j][i], e_i)};
        const Itv tot{__dsub_rd(P[j][PB_BUCKETS - 1], e_t), __dadd_ru(P[j][PB_BUCKETS - 1], e_t)};
        const Itv csr{__dsub_rd(tot.lo, csl.hi), __dsub_ru(tot.hi, csl.lo)};
        Itv v{0.0, 0.0};
        if (sl.hi != 0.0 && sl.lo != 0.0) v = itv_div_pos(itv_sq(csl), sl);                 // 0 + x is exact
        if (sr.hi != 0.0 && sr.lo != 0.0) v = itv_add(v, itv_div_pos(itv_sq(csr), sr));
        o = itv_add(o, v);
    }
    const bool cand = cnt_i > 0 || i == 0;
    // argmax of the lower bounds over the candidates (lowest index among equals)
    double key = (cand && o.lo == o.lo) ? o.lo : -INFINITY;
    s_lo[i] = key;
    s_loc[i] = i;
    __syncthreads();
    for (int half = PB_BUCKETS / 2; half; half >>= 1) {
        if (i < half) {
            const double a = s_lo[i], b = s_lo[i + half];
            if (b > a || (b == a && s_loc[i + half] < s_loc[i])) { s_lo[i] = b; s_loc[i] = s_loc[i + half]; }
        }
        __syncthreads();
    }
    const int best = s_loc[0];
    const double best_lo = s_lo[0];
    bool fine = !refuse;
    if (cand && i != best && !(best_lo > o.hi)) fine = false; // (NaN compares false: refused)
    if (!fine) s_ok = 0; // benign race: every writer stores 0
    __syncthreads();
    if (i == 0) {
        const bool ok = s_ok != 0 && best_lo > -INFINITY;
        sp[seg].split = (uint32_t)best;
        sp[seg].nleft = (uint32_t)C[best];
        sp[seg].pad = ok ? PB_ROUTE_CERTIFIED : PB_ROUTE_REFUSED;
        atomicAdd(&g_certify_counts[ok ? 0 : 1], 1ull);
    }
}

} // namespace

void pb_launch_buckets_hist(const PbPlanes bufs[2], const PbSeg *d_segs, int nseg, uint32_t max_n, bool weighted,
                            const double *d_axes, PbSplit *d_split, uint16_t *d_bucket, PbHist *d_hist, int sm_count,
                            cudaStream_t st) {
    if (nseg <= 0) return;
    PB_CUDA_OK(cudaMemsetAsync(d_hist, 0, (size_t)nseg * sizeof(PbHist), st));
    uint32_t want = (max_n + BH_CHUNK - 1) / BH_CHUNK;
    const uint32_t cap = (uint32_t)sm_count * 8;
    dim3 grid(want < cap ? (want ? want : 1) : cap, nseg);
    PbProfScope _prof("k_buckets_hist", st);
    if (weighted) {
        constexpr int smem = (int)(sizeof(WarpTab<true>) * BH_WARPS);
        PB_CUDA_OK(cudaFuncSetAttribute(k_buckets_hist<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        PB_CUDA_OK(cudaFuncSetAttribute(k_buckets_hist<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        k_buckets_hist<true><<<grid, BH_THREADS, smem, st>>>(bufs[0], bufs[1], d_segs, d_axes, d_split, d_bucket, d_hist);
    } else {
        constexpr int smem = (int)(sizeof(WarpTab<false>) * BH_WARPS);
        PB_CUDA_OK(cudaFuncSetAttribute(k_buckets_hist<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        PB_CUDA_OK(cudaFuncSetAttribute(k_buckets_hist<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        k_buckets_hist<false><<<grid, BH_THREADS, smem, st>>>(bufs[0], bufs[1], d_segs, d_axes, d_split, d_bucket, d_hist);
    }
    PB_CUDA_OK(cudaGetLastError());
}

void pb_launch_split_certify(const PbHist *d_hist, int nseg, bool weighted, PbSplit *d_split, int distrust, cudaStream_t st) {
    if (nseg <= 0) return;
    PbProfScope _prof("k_split_certify", st, false);
    if (weighted) k_split_certify<true><<<nseg, PB_BUCKETS, 0, st>>>(d_hist, d_split, distrust);
    else k_split_certify<false><<<nseg, PB_BUCKETS, 0, st>>>(d_hist, d_split, distrust);
    PB_CUDA_OK(cudaGetLastError());
}

void pb_certify_counts(unsigned long long out[4], bool reset) {
    PB_CUDA_OK(cudaMemcpyFromSymbol(out, g_certify_counts, sizeof(unsigned long long) * 4));
    if (reset) {
        const unsigned long long z[4] = {0, 0, 0, 0};
        PB_CUDA_OK(cudaMemcpyToSymbol(g_certify_counts, z, sizeof z));
    }
}
