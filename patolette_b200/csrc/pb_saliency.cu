// pb_saliency.cu - row N3: the saliency weights the reference's Python wrapper computes before it calls patolette()
// (src/patolette/patolette.pyx:54-313: `mbd` = three raster scans of a minimum-barrier distance on the channel mean,
// `get_weights` = that map + four border-colour Mahalanobis maps in CIELab, a centre prior and a sigmoid).
//
// Everything here is per-pixel work, a handful of global maxima - and ONE genuinely sequential part: a raster scan
// updates pixel (x, y) from its upper and its left neighbour, both already updated.  The data flow is a wavefront, and
// k_mbd_pass runs it as a systolic array: a warp owns 32 consecutive rows, lane t is row t, and at step s lane t works
// on column s - t.  The left neighbour is the lane's own previous result (registers), the upper neighbour is lane
// t-1's result of the previous step (one shuffle); the row above a warp's first row belongs to the previous CTA,
// which hands its last row's (U, L) pairs over through an edge buffer with the flag in the data.  rows + cols steps per
// scan instead of rows * cols.  The scan only takes float32 max / min / subtract, so the distance map is the
// reference's bit for bit.
//
// The rest follows the wrapper in f64 with CUDA's libm (`pow`, `cbrt`, `exp`, `sqrt`): numpy's own vector pow / cbrt and
// scikit-image's rgb2lab are not reproducible to the last bit from here (and scikit-image is not even installed to
// compare with), so the weights carry a floating-point tolerance, not bit parity (tests/test_saliency.py states it).
// Reductions are deterministic: maxima, and sums through per-CTA partials combined in a fixed order.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "pb_common.cuh"
#include "pb_pipeline.h"
#include "pb_pool.h"
#include "pb_prof.h"

namespace {

constexpr unsigned FULL = 0xffffffffu;

// skimage.color.rgb2lab, illuminant D65 / observer 2 (colorconv.py: rgb2xyz + xyz2lab)
__device__ __forceinline__ void rgb_to_lab(double r, double g, double b, double &L, double &A, double &B) {
    auto lin = [](double c) { return c > 0.04045 ? pow((c + 0.055) / 1.055, 2.4) : c / 12.92; };
    r = lin(r); g = lin(g); b = lin(b);
    double x = (r * 0.412453 + g * 0.357580 + b * 0.180423) / 0.95047;
    double y = (r * 0.212671 + g * 0.715160 + b * 0.072169) / 1.0;
    double z = (r * 0.019334 + g * 0.119193 + b * 0.950227) / 1.08883;
    auto f = [](double t) { return t > 0.008856 ? cbrt(t) : 7.787 * t + 16.0 / 116.0; };
    x = f(x); y = f(y); z = f(z);
    L = 116.0 * y - 16.0;
    A = 500.0 * (x - y);
    B = 200.0 * (y - z);
}

// patolette.pyx:204 (np.mean over the channels, cast to float32), :157-169 (L = U = img, D = inf inside / 0 on the
// border), :213 (rgb2lab)
__global__ void __launch_bounds__(256) k_sal_prepare(const double *__restrict__ c0, const double *__restrict__ c1,
                                                     const double *__restrict__ c2, uint32_t rows, uint32_t cols,
                                                     float *__restrict__ img, float *__restrict__ Lm, float *__restrict__ Um,
                                                     float *__restrict__ Dm, double *__restrict__ lab0, double *__restrict__ lab1,
                                                     double *__restrict__ lab2) {
    const size_t n = (size_t)rows * cols;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
        const double r = c0[p], g = c1[p], b = c2[p];
        float m = (float)(((r + g) + b) / 3.0);
        m = m == m ? m : 0.0f; // a NaN pixel (undefined in the reference too) must not enter the scans' U / L: NaN marks "not yet produced"
        const uint32_t x = (uint32_t)(p / cols), y = (uint32_t)(p % cols);
        img[p] = m; Lm[p] = m; Um[p] = m;
        Dm[p] = (x == 0 || y == 0 || x == rows - 1 || y == cols - 1) ? 0.0f : INFINITY;
        double L, A, B;
        rgb_to_lab(r, g, b, L, A, B);
        lab0[p] = L; lab1[p] = A; lab2[p] = B;
    }
}

// (relaxed.gpu, not a weak .cg access: ptxas may merge repeated WEAK loads of one address - it did, and hoisted the
//  poll out of its loop)
__device__ __forceinline__ float2 ld_edge(const float2 *p) { // one 8-byte load served by L2
    float2 v;
    asm volatile("ld.relaxed.gpu.global.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_edge(float2 *p, float u, float l) { // one 8-byte store: the pair appears at once
    asm volatile("st.relaxed.gpu.global.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(u), "f"(l) : "memory");
}
constexpr int MBD_SPIN_LIMIT = 1 << 23; // seconds of polling (a legitimate wait is at most one scan, ~20 ms): a bug must not hang the device
__device__ __forceinline__ int ld_flag(const int *p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_flag(int *p, int v) { asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// One raster scan (patolette.pyx:54-100) or inverse scan (:102-151) as a systolic wavefront; see the file header.
//
// In scan coordinates (i = visiting order of the rows, c = visiting order of the columns) pixel (i, c) needs
// (i - 1, c) and (i, c - 1).  A CTA owns rows 32 g .. 32 g + 31 (g is taken from a ticket, so a CTA's predecessor is
// always resident or done) and walks SKEWED tiles: in tile k, at step j, lane t of the compute warp works on column
// 32 k + j - t - all 32 lanes are busy in every step, the left neighbour is the lane's own previous result and the
// upper neighbour is lane t - 1's previous result (one shuffle).  Lane 0's upper neighbours (row 32 g - 1, columns
// 32 k .. 32 k + 31) belong to the previous CTA, which finishes them in ITS tiles k and k + 1.  They travel through an
// EDGE buffer with the flag in the data: before a scan every (U, L) pair of it is NaN (bytes 0xff), the producer's
// lane 31 stores its pair with one 8-byte store the moment the step has produced it, and the consumer polls the
// pairs it needs until they are numbers (U and L are maxima / minima of image values, which k_sal_prepare keeps
// NaN-free) - one store-to-load latency per hand-off, no fence, no separate flag.  A scan costs about
// cols / 32 + 2 rows / 32 tile times.
//
// The tile time is the length of ONE warp's dependent instruction stream (a first version that loaded, stepped and
// stored in a single warp spent most of it on copy instructions - profiles/r02_saliency.md), so the CTA is a
// pipeline of specialised warps over a ring of MBD_NBUF tile buffers in shared memory:
//   loaders (MBD_NLD warps)  cp.async the tile's 32 x 32 img / D / U / L values (coalesced row segments), warp 0 of
//                            them also polls the predecessor's edge pairs;
//   compute (1 warp)         32 branch-free dependent steps out of shared memory, results in place, edge pairs out;
//   storers (MBD_NST warps)  write D / U / L back.
// Hand-offs inside the CTA are monotonic tile counters in shared memory.  (Measured per tile after the last change:
// compute 1.75 us, one of two loaders 2.25 us, one of two storers 2.75 us - hence three of each.)
constexpr int MBD_T = 32, MBD_NBUF = 4, MBD_NLD = 3, MBD_NST = 3, MBD_WARPS = 8;
struct MbdBuf {
    float I[MBD_T][MBD_T + 1], D[MBD_T][MBD_T + 1], U[MBD_T][MBD_T + 1], L[MBD_T][MBD_T + 1];
    float upU[MBD_T], upL[MBD_T];
};
struct MbdSmem {
    MbdBuf buf[MBD_NBUF];
    volatile int ld_cnt[MBD_NLD]; // tiles a loader warp has landed
    volatile int comp_cnt;        // tiles computed
    volatile int st_cnt[MBD_NST]; // tiles a storer warp has written back
    int g;
};
__device__ __forceinline__ void cp_async4(float *smem_dst, const float *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
// Warp roles by physical warp: 0 = compute, 1..3 = loaders, 5..7 = storers; warp 4 leaves at once, so that the compute
// warp has its scheduler (warp id mod 4) to itself within the CTA.
//
// EVERY poll below is executed by all 32 lanes of the waiting warp (a broadcast read, or a vote): a warp in which only
// lane 0 polls comes back from the loop DIVERGED, and a diverged warp executes each shuffle of the step loop through
// the collective slow path (WARPSYNC.COLLECTIVE) - measured 13 us instead of 1.75 us per tile (profiles/r02_saliency.md).
__global__ void __launch_bounds__(32 * MBD_WARPS) k_mbd_pass(const float *__restrict__ img, float *Lm, float *Um, float *Dm, int rows,
                                                            int cols, int inverse, int *ctl /* [0] ticket, [1] error */,
                                                            float2 *edge /* [groups][cols] (U, L) of each group's last row */,
                                                            unsigned long long *dbg /* nullptr, or [4][ntiles][8] time stamps */) {
    extern __shared__ __align__(16) unsigned char mbd_smem_raw[];
    MbdSmem &S = *reinterpret_cast<MbdSmem *>(mbd_smem_raw);
    const int lane = threadIdx.x & 31, pwarp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        S.g = atomicAdd(ctl, 1);
        S.comp_cnt = 0;
        for (int w = 0; w < MBD_NLD; w++) S.ld_cnt[w] = 0;
        for (int w = 0; w < MBD_NST; w++) S.st_cnt[w] = 0;
    }
    __syncthreads();
    if (pwarp == 4) return;
    const int g = S.g;
    // the raster scan visits x = 1 .. rows-2, y = 1 .. cols-2; the inverse scan x = rows-2 .. 2, y = cols-2 .. 2
    const int R = inverse ? rows - 3 : rows - 2, Cn = inverse ? cols - 3 : cols - 2;
    const int ntiles = (Cn + 31 + 31) / 32; // steps 0 .. Cn + 30
    auto row_of = [&](int i) { return inverse ? rows - 2 - i : 1 + i; };
    auto col_of = [&](int c) { return inverse ? cols - 2 - c : 1 + c; };
    auto nap = [&]() { __nanosleep(32); }; // waiting warps leave the issue slots to the working ones
    // debug time stamps (PB_MBD_DEBUG=path): groups 0..3, per tile {compute start, compute end, loader-0 start, copies
    // issued, edge pairs seen, tile landed, storer-0 start, storer-0 end}, nanoseconds
    auto stamp = [&](int k, int slot) {
        if (dbg && g < 4 && lane == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            dbg[((size_t)g * ntiles + k) * 8 + slot] = t;
        }
    };

    if (pwarp == 0) {
        // ------------------------------------------------------------------ compute
        const int i = g * 32 + lane;
        const bool row_ok = i < R;
        float uleft = 0.f, lleft = 0.f; // U, L of the column this row visited last (starts on the border column)
        if (row_ok) {
            const size_t b = (size_t)row_of(i) * cols + (inverse ? cols - 1 : 0);
            uleft = Um[b];
            lleft = Lm[b];
        }
        float myU = 0.f, myL = 0.f; // this lane's U, L at the column it visited last (the next lane's upper neighbour)
        float2 *const my_edge = edge + (size_t)g * cols;
        for (int k = 0; k < ntiles; k++) {
            MbdBuf &B = S.buf[k % MBD_NBUF];
            bool ready;
            do {
                ready = true;
                for (int w = 0; w < MBD_NLD; w++) ready = ready && S.ld_cnt[w] > k;
            } while (!ready);
            __threadfence_block();
            __syncwarp();
            stamp(k, 0);
#pragma unroll 8
            for (int j = 0; j < 32; j++) {
                float upU = __shfl_up_sync(FULL, myU, 1), upL = __shfl_up_sync(FULL, myL, 1);
                const int c = 32 * k + j - lane;
                const bool active = row_ok && c >= 0 && c < Cn;
                const float tU = B.upU[j], tL = B.upL[j];
                upU = lane == 0 ? tU : upU;
                upL = lane == 0 ? tL : upL;
                const float ix = B.I[lane][j], d = B.D[lane][j], cu = B.U[lane][j], cl = B.L[lane][j];
                const float u1 = fmaxf(upU, ix), l1 = fminf(upL, ix), u2 = fmaxf(uleft, ix), l2 = fminf(lleft, ix);
                const float b1 = u1 - l1, b2 = u2 - l2;
                const bool keep = d <= b1 && d <= b2, use1 = b1 < d && b1 <= b2;
                const float nD = keep ? d : (use1 ? b1 : b2), nU = keep ? cu : (use1 ? u1 : u2), nL = keep ? cl : (use1 ? l1 : l2);
                B.D[lane][j] = nD; B.U[lane][j] = nU; B.L[lane][j] = nL; // (cells outside the scan are never written back)
                if (lane == 31 && active) st_edge(my_edge + c, nU, nL);
                uleft = active ? nU : uleft; lleft = active ? nL : lleft;
                myU = active ? nU : myU; myL = active ? nL : myL;
            }
            __syncwarp();
            stamp(k, 1);
            __threadfence_block();
            S.comp_cnt = k + 1; // (every lane stores the same value)
        }
    } else if (pwarp <= MBD_NLD) {
        // ------------------------------------------------------------------ loaders
        const int w = pwarp - 1;
        const size_t uprow = (size_t)(inverse ? row_of(g * 32) + 1 : row_of(g * 32) - 1) * cols; // the row above the CTA's first
        for (int k = 0; k < ntiles; k++) {
            MbdBuf &B = S.buf[k % MBD_NBUF];
            if (k >= MBD_NBUF) { // the buffer's previous tile must have been written back
                for (;;) {
                    bool free_ = true;
                    for (int q = 0; q < MBD_NST; q++) free_ = free_ && S.st_cnt[q] > k - MBD_NBUF;
                    if (free_) break;
                    nap();
                }
                __threadfence_block();
            }
            if (w == 0) stamp(k, 2);
            // tile in: row r of the tile holds columns 32 k - r .. 32 k - r + 31 of scan row 32 g + r
#pragma unroll 4
            for (int r = w; r < 32; r += MBD_NLD) {
                const int ir = g * 32 + r, c = 32 * k + lane - r;
                if (ir < R && c >= 0 && c < Cn) {
                    const size_t p = (size_t)row_of(ir) * cols + col_of(c);
                    cp_async4(&B.I[r][lane], img + p); cp_async4(&B.D[r][lane], Dm + p);
                    cp_async4(&B.U[r][lane], Um + p); cp_async4(&B.L[r][lane], Lm + p);
                }
            }
            if (w == 0) { // lane 0's upper neighbours
                stamp(k, 3);
                const int c = 32 * k + lane;
                const bool need = c < Cn;
                if (g > 0) { // the previous CTA's last row: poll the pairs until all of them have been produced
                    const float2 *src = edge + (size_t)(g - 1) * cols + (need ? c : 0);
                    float2 v = make_float2(__int_as_float(0x7fffffff), 0.f);
                    int spins = 0;
                    for (;;) {
                        if (need && v.x != v.x) v = ld_edge(src);
                        if (__all_sync(FULL, !need || v.x == v.x)) break;
                        // (gives up after seconds, or as soon as anyone else has: never hang the device)
                        const int f = (++spins & 1023) == 0 ? ld_flag(ctl + 1) : 0;
                        if (spins >= MBD_SPIN_LIMIT || __any_sync(FULL, f != 0)) {
                            if (lane == 0) st_flag(ctl + 1, 1);
                            break;
                        }
                        nap();
                    }
                    if (need) { B.upU[lane] = v.x; B.upL[lane] = v.y; }
                } else if (need) { // the border row
                    B.upU[lane] = Um[uprow + col_of(c)];
                    B.upL[lane] = Lm[uprow + col_of(c)];
                }
                __syncwarp();
                stamp(k, 4);
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            __threadfence_block();
            __syncwarp();
            if (w == 0) stamp(k, 5);
            S.ld_cnt[w] = k + 1; // (every lane stores the same value)
        }
    } else {
        // ------------------------------------------------------------------ storers
        const int w = pwarp - 5;
        for (int k = 0; k < ntiles; k++) {
            MbdBuf &B = S.buf[k % MBD_NBUF];
            while (S.comp_cnt <= k) nap();
            __threadfence_block();
            if (w == 0) stamp(k, 6);
#pragma unroll 4
            for (int r = w; r < 32; r += MBD_NST) {
                const int ir = g * 32 + r, c = 32 * k + lane - r;
                if (ir < R && c >= 0 && c < Cn) {
                    const size_t p = (size_t)row_of(ir) * cols + col_of(c);
                    Dm[p] = B.D[r][lane]; Um[p] = B.U[r][lane]; Lm[p] = B.L[r][lane];
                }
            }
            __syncwarp();
            if (w == 0) stamp(k, 7);
            __threadfence_block();
            S.st_cnt[w] = k + 1; // the buffer may be refilled (every lane stores the same value)
        }
    }
}

// ---- border strips (patolette.pyx:215-239): rows [0, bt), rows [rows-bt-1, rows-1), cols [0, bt), cols [cols-bt-1, cols-1)
struct SalStrips { int bt; };
__device__ __forceinline__ unsigned strip_mask(uint32_t x, uint32_t y, uint32_t rows, uint32_t cols, uint32_t bt) {
    unsigned m = 0;
    if (x < bt) m |= 1u;
    if (x + bt + 1 >= rows && x + 1 < rows) m |= 2u;
    if (y < bt) m |= 4u;
    if (y + bt + 1 >= cols && y + 1 < cols) m |= 8u;
    return m;
}
constexpr int SS_CTAS = 512, SS_THREADS = 256;
// PASS 0: sums of the Lab values per strip; PASS 1: sums of the centred products (00 01 02 11 12 22) per strip.
// Per-CTA partials, combined in CTA order by k_strip_finish: the same bits in every run.
template <int PASS>
__global__ void __launch_bounds__(SS_THREADS) k_strip_partial(const double *__restrict__ l0, const double *__restrict__ l1,
                                                              const double *__restrict__ l2, uint32_t rows, uint32_t cols, uint32_t bt,
                                                              const double *__restrict__ means /* [4][3] (PASS 1) */,
                                                              double *__restrict__ partial /* [CTA][4][6] */) {
    constexpr int NV = PASS == 0 ? 3 : 6;
    __shared__ double red[SS_THREADS / 32][4][6];
    double acc[4][NV];
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int v = 0; v < NV; v++) acc[k][v] = 0.0;
    const size_t n = (size_t)rows * cols;
    for (size_t p = (size_t)blockIdx.x * SS_THREADS + threadIdx.x; p < n; p += (size_t)gridDim.x * SS_THREADS) {
        const uint32_t x = (uint32_t)(p / cols), y = (uint32_t)(p % cols);
        const unsigned m = strip_mask(x, y, rows, cols, bt);
        if (!m) continue;
        const double a = l0[p], b = l1[p], c = l2[p];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (!(m >> k & 1u)) continue;
            if (PASS == 0) {
                acc[k][0] += a; acc[k][1] += b; acc[k][2] += c;
            } else {
                const double da = a - means[k * 3], db = b - means[k * 3 + 1], dc = c - means[k * 3 + 2];
                acc[k][0] += da * da; acc[k][1] += da * db; acc[k][2] += da * dc;
                acc[k][3] += db * db; acc[k][4] += db * dc; acc[k][5] += dc * dc;
            }
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int v = 0; v < NV; v++) {
            double t = acc[k][v];
            for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(FULL, t, o);
            if (lane == 0) red[warp][k][v] = t;
        }
    __syncthreads();
    if (threadIdx.x < 4 * NV) {
        const int k = threadIdx.x / NV, v = threadIdx.x % NV;
        double t = 0.0;
        for (int w = 0; w < SS_THREADS / 32; w++) t += red[w][k][v];
        partial[((size_t)blockIdx.x * 4 + k) * 6 + v] = t;
    }
}
__global__ void k_strip_finish(const double *__restrict__ partial, int nctas, int nv, double *__restrict__ out /* [4][6] */) {
    const int k = threadIdx.x / 6, v = threadIdx.x % 6;
    if (threadIdx.x >= 24 || v >= nv) return;
    double t = 0.0;
    for (int c = 0; c < nctas; c++) t += partial[((size_t)c * 4 + k) * 6 + v];
    out[k * 6 + v] = t;
}

// ---- the per-pixel chain of get_weights (patolette.pyx:241-313), cut where a global maximum is needed --------------
struct SalParams {
    double mean[4][3];
    double vi[4][9];
    double umax[4];      // float32-rounded maxima of the four Mahalanobis maps
    double u_max_final;  // float32-rounded maximum of u_final
    float sal_max;       // maximum of the distance map
    double m1, m2;       // maxima of the two later stages
    double w2, h2, diag; // rows / 2, cols / 2, sqrt(w2^2 + h2^2)
    double scale;        // rows * cols (as the wrapper forms it), tile_size^2
    double tile2;
};
__device__ __forceinline__ double mahalanobis(const SalParams &P, int k, double a, double b, double c) {
    const double d0 = a - P.mean[k][0], d1 = b - P.mean[k][1], d2 = c - P.mean[k][2];
    const double *V = P.vi[k];
    const double t0 = d0 * V[0] + d1 * V[3] + d2 * V[6], t1 = d0 * V[1] + d1 * V[4] + d2 * V[7], t2 = d0 * V[2] + d1 * V[5] + d2 * V[8];
    return sqrt(t0 * d0 + t1 * d1 + t2 * d2);
}
__device__ __forceinline__ double u_final_of(const SalParams &P, double a, double b, double c) {
    double u[4], um = 0.0, sum = 0.0;
#pragma unroll
    for (int k = 0; k < 4; k++) u[k] = mahalanobis(P, k, a, b, c) / P.umax[k];
    um = fmax(fmax(fmax(u[0], u[1]), u[2]), u[3]);
    sum = ((u[0] + u[1]) + u[2]) + u[3];
    return sum - um;
}
__device__ __forceinline__ double s1_of(const SalParams &P, float d, double a, double b, double c) {
    return (double)(d / P.sal_max) + u_final_of(P, a, b, c) / P.u_max_final; // float32 / float32, then f64
}
__device__ __forceinline__ double s2_of(const SalParams &P, float d, double a, double b, double c, uint32_t x, uint32_t y) {
    const double s = s1_of(P, d, a, b, c) / P.m1;
    const double dx = (double)y - P.h2, dy = (double)x - P.w2;
    const double C = 1.0 - sqrt(dx * dx + dy * dy) / P.diag;
    return s * C;
}
__device__ __forceinline__ void atomic_max_pos(unsigned long long *slot, double v) { // v >= 0: bit patterns order like the values
    if (v == v) atomicMax(slot, (unsigned long long)__double_as_longlong(v));
}
// STAGE 0: maxima of the four Mahalanobis maps -> out[0..3]; 1: max u_final -> out[0], max D -> out[1];
// 2: max s1 -> out[0]; 3: max s2 -> out[0]; 4: the weights
template <int STAGE>
__global__ void __launch_bounds__(256) k_sal_stage(SalParams P, const double *__restrict__ l0, const double *__restrict__ l1,
                                                   const double *__restrict__ l2, const float *__restrict__ Dm, uint32_t rows,
                                                   uint32_t cols, unsigned long long *__restrict__ out, double *__restrict__ weights) {
    const size_t n = (size_t)rows * cols;
    double m[4] = {0.0, 0.0, 0.0, 0.0};
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
        const double a = l0[p], b = l1[p], c = l2[p];
        if (STAGE == 0) {
#pragma unroll
            for (int k = 0; k < 4; k++) m[k] = fmax(m[k], mahalanobis(P, k, a, b, c));
        } else if (STAGE == 1) {
            m[0] = fmax(m[0], u_final_of(P, a, b, c));
            m[1] = fmax(m[1], (double)Dm[p]);
        } else if (STAGE == 2) {
            m[0] = fmax(m[0], s1_of(P, Dm[p], a, b, c));
        } else {
            const uint32_t x = (uint32_t)(p / cols), y = (uint32_t)(p % cols);
            const double s2 = s2_of(P, Dm[p], a, b, c, x, y);
            if (STAGE == 3) m[0] = fmax(m[0], s2);
            else {
                const double v = s2 / P.m2;
                const double f = 1.0 / (1.0 + exp(-10.0 * (v - 0.5)));
                weights[p] = 1.0 + f * f * P.scale / P.tile2;
            }
        }
    }
    if (STAGE == 4) return;
    constexpr int NM = STAGE == 0 ? 4 : (STAGE == 1 ? 2 : 1);
#pragma unroll
    for (int k = 0; k < NM; k++) {
        double t = m[k];
        for (int o = 16; o; o >>= 1) t = fmax(t, __shfl_xor_sync(FULL, t, o));
        if ((threadIdx.x & 31) == 0) atomic_max_pos(out + k, t);
    }
}

// The scans run three CTAs of 68 KB shared memory per SM, which makes the driver move the SMs' L1 / shared-memory split to
// the largest shared-memory carve-out, and later kernels without a preference of their own may inherit it.  An empty
// kernel that asks for the largest L1 hands the SMs back in the state a run without saliency finds them in; the kernels
// that follow grow the carve-out as they need it.  Measured effect (4096^2, K = 256, dither): the dither stage of a
// saliency run 12.9 -> 11.6 ms, whole call 53.7 -> 50.6 ms (tools/ab_carveout.py; PB_SAL_KEEP_CARVEOUT=1 skips the
// kernel for such A/B runs).  It does NOT explain why that stage is slower than on an unweighted run (4.5 ms) - see
// profiles/r02_saliency.md, open question.
__global__ void k_prefer_l1() {}
void restore_l1_preference(int sm_count, cudaStream_t st) {
    const char *keep = getenv("PB_SAL_KEEP_CARVEOUT");
    if (keep && keep[0] == '1') return;
    PB_CUDA_OK(cudaFuncSetAttribute(k_prefer_l1, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxL1));
    k_prefer_l1<<<16 * (sm_count > 0 ? sm_count : 148), 32, 0, st>>>();
    PB_CUDA_OK(cudaGetLastError());
}

// mbd() of the wrapper (patolette.pyx:183-199): inverse, raster, inverse scan over img / L / U / D (device, n floats
// each, prepared by k_sal_prepare).  Returns 0, or -1 if a scan gave up waiting for its predecessor (a bug guard).
int mbd_scans(const float *img, float *Lm, float *Um, float *Dm, uint32_t rows, uint32_t cols, int sm_count, cudaStream_t st) {
    // (per call: the attribute belongs to the function on the CURRENT device, and patolette_b200_set_device may have moved us)
    PB_CUDA_OK(cudaFuncSetAttribute(k_mbd_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MbdSmem)));
    const int groups = (int)((rows + 31) / 32);
    int *ctl = nullptr;
    float2 *edge = nullptr;
    unsigned long long *dbg = nullptr;
    const char *dbg_path = getenv("PB_MBD_DEBUG"); // debug: time stamps of the first scan's first four CTAs, dumped to this file
    const size_t dbg_words = dbg_path ? (size_t)4 * (((size_t)cols + 62) / 32 + 1) * 8 : 0;
    int err = 0;
    try {
        ctl = (int *)pb_pool_alloc(2 * sizeof(int));
        edge = (float2 *)pb_pool_alloc((size_t)groups * cols * sizeof(float2));
        PB_CUDA_OK(cudaMemsetAsync(ctl, 0, 2 * sizeof(int), st));
        if (dbg_words) {
            dbg = (unsigned long long *)pb_pool_alloc(dbg_words * sizeof(unsigned long long));
            PB_CUDA_OK(cudaMemsetAsync(dbg, 0, dbg_words * sizeof(unsigned long long), st));
        }
        for (int it = 0; it < 3; it++) {
            const int inverse = it % 2 == 0;
            const int R = inverse ? (int)rows - 3 : (int)rows - 2;
            if (R <= 0 || (inverse ? (int)cols - 3 : (int)cols - 2) <= 0) continue;
            PB_CUDA_OK(cudaMemsetAsync(ctl, 0, sizeof(int), st));                                          // the ticket
            PB_CUDA_OK(cudaMemsetAsync(edge, 0xff, (size_t)((R + 31) / 32) * cols * sizeof(float2), st)); // "not yet produced"
            PbProfScope p("k_mbd_pass", st);
            k_mbd_pass<<<(R + 31) / 32, 32 * MBD_WARPS, sizeof(MbdSmem), st>>>(img, Lm, Um, Dm, (int)rows, (int)cols, inverse, ctl, edge,
                                                                                 it == 0 ? dbg : nullptr);
            PB_CUDA_OK(cudaGetLastError());
        }
        if (dbg_words) {
            std::vector<unsigned long long> h(dbg_words);
            PB_CUDA_OK(cudaMemcpyAsync(h.data(), dbg, dbg_words * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
            PB_CUDA_OK(cudaStreamSynchronize(st));
            if (FILE *f = fopen(dbg_path, "wb")) { fwrite(h.data(), sizeof(unsigned long long), dbg_words, f); fclose(f); }
        }
        restore_l1_preference(sm_count, st);
        PB_CUDA_OK(cudaMemcpyAsync(&err, ctl + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        PB_CUDA_OK(cudaStreamSynchronize(st));
    } catch (...) {
        cudaDeviceSynchronize();
        pb_pool_free(ctl);
        pb_pool_free(edge);
        pb_pool_free(dbg);
        throw;
    }
    pb_pool_free(ctl);
    pb_pool_free(edge);
    pb_pool_free(dbg);
    return err ? -1 : 0;
}

bool invert3(const double c[6] /* 00 01 02 11 12 22 */, double vi[9]) {
    const double a = c[0], b = c[1], cc = c[2], d = c[3], e = c[4], f = c[5];
    const double A = d * f - e * e, B = -(b * f - cc * e), C = b * e - cc * d;
    const double det = a * A + b * B + cc * C;
    if (!(fabs(det) > 0.0) || !std::isfinite(det)) return false;
    const double D = a * f - cc * cc, E = -(a * e - b * cc), F = a * d - b * b;
    const double inv[9] = {A / det, B / det, C / det, B / det, D / det, E / det, C / det, E / det, F / det};
    for (int i = 0; i < 9; i++) vi[i] = inv[i];
    return true;
}

} // namespace

// planes: device sRGB planes in [0, 1], pixel p = row * width + col (rows = height, cols = width as the wrapper
// reshapes, patolette.pyx:411).  Returns 0, -7 for images the scans cannot run on (a side <= 3, :154-155: the reference
// raises there), -1 when a border strip has a singular covariance (np.linalg.inv raises in the reference).
int pb_saliency_weights(const double *const planes[3], size_t width, size_t height, double tile_size, double *d_weights, int sm_count,
                        cudaStream_t st) {
    const uint32_t rows = (uint32_t)height, cols = (uint32_t)width;
    if (rows <= 3 || cols <= 3) return -7;
    const size_t n = (size_t)rows * cols;
    const uint32_t bt = (uint32_t)floor(0.1 * sqrt((double)(rows * (double)cols)));
    // patolette.pyx:215-239: the four strips are reshaped to exactly bt rows (columns); a strip that does not fit makes
    // the reference raise, and bt = 0 gives it empty strips (NaN weights)
    if (bt < 1 || bt + 1 > rows || bt + 1 > cols) return -7;
    float *f32 = nullptr;
    double *lab = nullptr, *small = nullptr;
    int rc = 0;
    auto cleanup = [&]() { pb_pool_free(f32); pb_pool_free(lab); pb_pool_free(small); };
    try {
        f32 = (float *)pb_pool_alloc(4 * n * sizeof(float));
        lab = (double *)pb_pool_alloc(3 * n * sizeof(double));
        small = (double *)pb_pool_alloc(((size_t)SS_CTAS * 24 + 64) * sizeof(double));
        float *img = f32, *Lm = f32 + n, *Um = f32 + 2 * n, *Dm = f32 + 3 * n;
        double *l0 = lab, *l1 = lab + n, *l2 = lab + 2 * n;
        double *partial = small, *sums = small + (size_t)SS_CTAS * 24;            // [24]
        unsigned long long *maxima = reinterpret_cast<unsigned long long *>(small + (size_t)SS_CTAS * 24 + 32); // [8]
        const size_t want = (n + 255) / 256, cap = (size_t)sm_count * 8;
        const int grid = (int)(want < cap ? want : cap);
        { PbProfScope p("k_sal_prepare", st);
          k_sal_prepare<<<grid, 256, 0, st>>>(planes[0], planes[1], planes[2], rows, cols, img, Lm, Um, Dm, l0, l1, l2); }
        if (mbd_scans(img, Lm, Um, Dm, rows, cols, sm_count, st) != 0) { cleanup(); return -1; }
        // strip means and covariances (np.mean, np.cov with ddof = 1), inverses on the host
        double h[24];
        SalParams P{};
        const double cnt[4] = {(double)bt * cols, (double)bt * cols, (double)bt * rows, (double)bt * rows};
        { PbProfScope p("k_strip_partial", st);
          k_strip_partial<0><<<SS_CTAS, SS_THREADS, 0, st>>>(l0, l1, l2, rows, cols, bt, nullptr, partial); }
        { PbProfScope p("k_strip_finish", st, false);
          k_strip_finish<<<1, 32, 0, st>>>(partial, SS_CTAS, 3, sums); }
        PB_CUDA_OK(cudaMemcpyAsync(h, sums, sizeof h, cudaMemcpyDeviceToHost, st));
        PB_CUDA_OK(cudaStreamSynchronize(st));
        double means[12];
        for (int k = 0; k < 4; k++)
            for (int v = 0; v < 3; v++) { means[k * 3 + v] = h[k * 6 + v] / cnt[k]; P.mean[k][v] = means[k * 3 + v]; }
        double *d_means = small + (size_t)SS_CTAS * 24 + 48; // [12]
        PB_CUDA_OK(cudaMemcpyAsync(d_means, means, sizeof means, cudaMemcpyHostToDevice, st));
        { PbProfScope p("k_strip_partial", st);
          k_strip_partial<1><<<SS_CTAS, SS_THREADS, 0, st>>>(l0, l1, l2, rows, cols, bt, d_means, partial); }
        { PbProfScope p("k_strip_finish", st, false);
          k_strip_finish<<<1, 32, 0, st>>>(partial, SS_CTAS, 6, sums); }
        PB_CUDA_OK(cudaMemcpyAsync(h, sums, sizeof h, cudaMemcpyDeviceToHost, st));
        PB_CUDA_OK(cudaStreamSynchronize(st));
        for (int k = 0; k < 4 && rc == 0; k++) {
            double cov[6];
            for (int v = 0; v < 6; v++) cov[v] = h[k * 6 + v] / (cnt[k] - 1.0);
            if (!(cnt[k] > 1.0) || !invert3(cov, P.vi[k])) rc = -1;
        }
        if (rc == 0) {
            auto read_max = [&](int count, double *dst) {
                unsigned long long hm[8];
                PB_CUDA_OK(cudaMemcpyAsync(hm, maxima, sizeof hm, cudaMemcpyDeviceToHost, st));
                PB_CUDA_OK(cudaStreamSynchronize(st));
                for (int i = 0; i < count; i++) memcpy(&dst[i], &hm[i], 8);
            };
            auto zero_max = [&]() { PB_CUDA_OK(cudaMemsetAsync(maxima, 0, 8 * sizeof(unsigned long long), st)); };
            double m[4];
            zero_max();
            { PbProfScope p("k_sal_stage", st); k_sal_stage<0><<<grid, 256, 0, st>>>(P, l0, l1, l2, Dm, rows, cols, maxima, nullptr); }
            read_max(4, m);
            for (int k = 0; k < 4; k++) P.umax[k] = (double)(float)m[k]; // cdef float max_u_* (patolette.pyx:268-271)
            zero_max();
            { PbProfScope p("k_sal_stage", st); k_sal_stage<1><<<grid, 256, 0, st>>>(P, l0, l1, l2, Dm, rows, cols, maxima, nullptr); }
            read_max(2, m);
            P.u_max_final = (double)(float)m[0]; // :282
            P.sal_max = (float)m[1];             // :283
            zero_max();
            { PbProfScope p("k_sal_stage", st); k_sal_stage<2><<<grid, 256, 0, st>>>(P, l0, l1, l2, Dm, rows, cols, maxima, nullptr); }
            read_max(1, m);
            P.m1 = m[0]; // :286
            P.w2 = rows / 2.0; P.h2 = cols / 2.0; P.diag = sqrt(P.w2 * P.w2 + P.h2 * P.h2); // :288-294
            zero_max();
            { PbProfScope p("k_sal_stage", st); k_sal_stage<3><<<grid, 256, 0, st>>>(P, l0, l1, l2, Dm, rows, cols, maxima, nullptr); }
            read_max(1, m);
            P.m2 = m[0]; // :303
            P.scale = (double)((size_t)rows * cols);
            P.tile2 = tile_size * tile_size;
            { PbProfScope p("k_sal_stage", st); k_sal_stage<4><<<grid, 256, 0, st>>>(P, l0, l1, l2, Dm, rows, cols, maxima, d_weights); }
            PB_CUDA_OK(cudaGetLastError());
            PB_CUDA_OK(cudaStreamSynchronize(st));
        }
    } catch (...) {
        cudaDeviceSynchronize();
        cleanup();
        throw;
    }
    cleanup();
    return rc;
}

// the distance map alone (tests: it is the part with bit parity)
int pb_saliency_mbd(const double *const planes[3], size_t width, size_t height, float *d_out, int sm_count, cudaStream_t st) {
    const uint32_t rows = (uint32_t)height, cols = (uint32_t)width;
    if (rows <= 3 || cols <= 3) return -7;
    const size_t n = (size_t)rows * cols;
    float *f32 = nullptr;
    double *lab = nullptr;
    auto cleanup = [&]() { pb_pool_free(f32); pb_pool_free(lab); };
    try {
        f32 = (float *)pb_pool_alloc(4 * n * sizeof(float));
        lab = (double *)pb_pool_alloc(3 * n * sizeof(double));
        float *img = f32, *Lm = f32 + n, *Um = f32 + 2 * n, *Dm = f32 + 3 * n;
        const size_t want = (n + 255) / 256, cap = (size_t)sm_count * 8;
        const int grid = (int)(want < cap ? want : cap);
        { PbProfScope p("k_sal_prepare", st);
          k_sal_prepare<<<grid, 256, 0, st>>>(planes[0], planes[1], planes[2], rows, cols, img, Lm, Um, Dm, lab, lab + n, lab + 2 * n); }
        if (mbd_scans(img, Lm, Um, Dm, rows, cols, sm_count, st) != 0) { cleanup(); return -1; }
        PB_CUDA_OK(cudaMemcpyAsync(d_out, Dm, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
        PB_CUDA_OK(cudaStreamSynchronize(st));
    } catch (...) {
        cudaDeviceSynchronize();
        cleanup();
        throw;
    }
    cleanup();
    return 0;
}
