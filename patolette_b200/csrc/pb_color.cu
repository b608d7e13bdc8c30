// pb_color.cu - per-pixel colour transforms (reference: lib/src/color/*.c).
//
// One thread per pixel, planar f64 in / out in place: 24 B read + 24 B written per
// pixel, ~9 pow() each, so these kernels are FP64-ALU bound rather than HBM bound.
// Every arithmetic step keeps the reference's association and rounding points
// (no FMA contraction - the TU is built with --fmad=false) and pow() is the
// bit-exact glibc restatement in pow_glibc.h with its tables staged in shared
// memory (lane-divergent table look-ups would serialise in constant memory).
#include "pb_common.cuh"
#include "pb_kernels.h"
#include "pb_prof.h"
#include "pow_glibc.h"

namespace {

__device__ uint64_t g_log_tab[128 * 3];
__device__ uint64_t g_exp_tab[256];
bool g_tabs_ready[64] = {false};

struct PowTabs {
    const uint64_t *lg, *ex;
    __device__ __forceinline__ double pw(double x, double y) const { return pb_pow(x, y, lg, ex); }
};

// color/sRGB.c:70-89
__device__ __forceinline__ double gamma_decode(const PowTabs &t, double c) {
    double r = (c <= 0.0404500) ? c / 12.92 : t.pw((c + 0.055) / 1.055, 2.4);
    return fmin(fmax(r, 0.0), 1.0);
}
// color/sRGB.c:91-110
__device__ __forceinline__ double gamma_encode(const PowTabs &t, double c) {
    double r = (c <= 0.0031308) ? c * 12.92 : 1.055 * t.pw(c, 1.0 / 2.4) - 0.055;
    return fmin(fmax(r, 0.0), 1.0);
}
// color/eotf.c:14-19
#define PQ_Lp 10000.0
#define PQ_m1 0.1593017578125
#define PQ_m2 78.84375
#define PQ_c1 0.8359375
#define PQ_c2 18.8515625
#define PQ_c3 18.6875
// color/eotf.c:29-42
__device__ __forceinline__ double pq_eotf(const PowTabs &t, double c) {
    const double m1d = 1 / PQ_m1, m2d = 1 / PQ_m2;
    double Vp = t.pw(c, m2d);
    double n = fmax(0.0, Vp - PQ_c1);
    double L = t.pw(n / (PQ_c2 - PQ_c3 * Vp), m1d);
    return PQ_Lp * L;
}
// color/eotf.c:44-57
__device__ __forceinline__ double pq_inverse_eotf(const PowTabs &t, double c) {
    double y = t.pw(c / PQ_Lp, PQ_m1);
    return t.pw((PQ_c1 + PQ_c2 * y) / (1 + PQ_c3 * y), PQ_m2);
}
// color/xyz.c:14-40
__device__ __forceinline__ void srgb_to_xyz(const PowTabs &t, double r, double g, double b,
                                            double &x, double &y, double &z) {
    double R = gamma_decode(t, r), G = gamma_decode(t, g), B = gamma_decode(t, b);
    x = R * 0.4124564 + G * 0.3575761 + B * 0.1804375;
    y = R * 0.2126729 + G * 0.7151522 + B * 0.0721750;
    z = R * 0.0193339 + G * 0.1191920 + B * 0.9503041;
}
// color/xyz.c:42-64
__device__ __forceinline__ void rec2020_to_xyz(double r, double g, double b, double &x, double &y,
                                               double &z) {
    x = r * 0.63695351 + g * 0.14461919 + b * 0.16885585;
    y = r * 0.26269834 + g * 0.67800877 + b * 0.0592929;
    z = g * 0.02807314 + b * 1.06082723;
}
// color/rec2020.c:75-102
__device__ __forceinline__ void xyz_to_rec2020(double x, double y, double z, double &r, double &g,
                                               double &b) {
    r = x * 1.71666343 + y * -0.35567332 + z * -0.25336809;
    g = x * -0.66667384 + y * 1.61645574 + z * 0.0157683;
    b = x * 0.01764248 + y * -0.04277698 + z * 0.94224328;
}
// color/ICtCp.c:41-79 (the stored Ct is halved, :78)
__device__ __forceinline__ void rec2020_to_ictcp(const PowTabs &t, double r, double g, double b,
                                                 double &I, double &Ct, double &Cp) {
    double L = (r * 1688 + g * 2146 + b * 262) / 4096;
    double M = (r * 683 + g * 2951 + b * 462) / 4096;
    double S = (r * 99 + g * 309 + b * 3688) / 4096;
    double L_ = pq_inverse_eotf(t, L), M_ = pq_inverse_eotf(t, M), S_ = pq_inverse_eotf(t, S);
    I = L_ * 0.5 + M_ * 0.5;
    Ct = (L_ * 6610 - M_ * 13613 + S_ * 7003) / 4096;
    Cp = (L_ * 17933 - M_ * 17390 - S_ * 543) / 4096;
    Ct = Ct * 0.5;
}
// color/rec2020.c:32-69
__device__ __forceinline__ void ictcp_to_rec2020(const PowTabs &t, double I, double Ct, double Cp,
                                                 double &r, double &g, double &b) {
    Ct = Ct * 2;
    double L_ = I + 0.00860904 * Ct + 0.11102963 * Cp;
    double M_ = I - 0.00860904 * Ct - 0.11102963 * Cp;
    double S_ = I + 0.56003134 * Ct - 0.32062717 * Cp;
    double L = pq_eotf(t, L_), M = pq_eotf(t, M_), S = pq_eotf(t, S_);
    r = L * 3.43660669 - M * 2.50645212 + S * 0.06984542;
    g = -L * 0.79132956 + M * 1.98360045 - S * 0.1922709;
    b = -L * 0.0259499 - M * 0.09891371 + S * 1.12486361;
}
// color/CIELuv.c:19-24
#define LUV_rwx 0.95047
#define LUV_rwy 1.0
#define LUV_rwz 1.08883
#define LUV_kE (216.0 / 24389.0)
#define LUV_kK (24389.0 / 27.0)
#define LUV_kKE 8.0
// color/CIELuv.c:54-89
__device__ __forceinline__ void xyz_to_cieluv(const PowTabs &t, double x, double y, double z,
                                              double &L, double &u, double &v) {
    double den = x + 15.0 * y + 3.0 * z;
    double up = (den > 0.0) ? ((4.0 * x) / (x + 15.0 * y + 3.0 * z)) : 0.0;
    double vp = (den > 0.0) ? ((9.0 * y) / (x + 15.0 * y + 3.0 * z)) : 0.0;
    const double urp = (4.0 * LUV_rwx) / (LUV_rwx + 15.0 * LUV_rwy + 3.0 * LUV_rwz);
    const double vrp = (9.0 * LUV_rwy) / (LUV_rwx + 15.0 * LUV_rwy + 3.0 * LUV_rwz);
    double yr = y / LUV_rwy;
    double L_ = (yr > LUV_kE) ? (116.0 * t.pw(yr, 1.0 / 3.0) - 16.0) : (LUV_kK * yr);
    L = L_;
    u = 13.0 * L_ * (up - urp);
    v = 13.0 * L_ * (vp - vrp);
}
// color/CIELuv.c:100-164
__device__ __forceinline__ void cieluv_to_xyz(const PowTabs &t, double L, double u, double v,
                                              double &x, double &y, double &z) {
    double y_ = (L > LUV_kKE) ? t.pw((L + 16.0) / 116.0, 3.0) : (L / LUV_kK);
    const double u0 = (4.0 * LUV_rwx) / (LUV_rwx + 15.0 * LUV_rwy + 3.0 * LUV_rwz);
    const double v0 = (9.0 * LUV_rwy) / (LUV_rwx + 15.0 * LUV_rwy + 3.0 * LUV_rwz);
    double a_den = u + 13.0 * L * u0;
    double a = (a_den == 0.0) ? 0.0 : (((52.0 * L) / a_den) - 1.0) / 3.0;
    double b = -5.0 * y_;
    const double c = -1.0 / 3.0;
    double d_den = v + 13.0 * L * v0;
    double d = (d_den == 0.0) ? 0.0 : y_ * (((39.0 * L) / d_den) - 5.0);
    double x_den = a - c;
    double x_ = (x_den == 0.0) ? 0.0 : (d - b) / x_den;
    double z_ = x_ * a + b;
    x = x_;
    y = y_;
    z = z_;
}
// color/sRGB.c:32-59
__device__ __forceinline__ void rec2020_to_srgb(const PowTabs &t, double r2, double g2, double b2,
                                                double &r, double &g, double &b) {
    double x, y, z;
    rec2020_to_xyz(r2, g2, b2, x, y, z);
    r = x * 3.2404542 - y * 1.5371385 - z * 0.4985314;
    g = -x * 0.9692660 + y * 1.8760108 + z * 0.0415560;
    b = x * 0.0556434 - y * 0.2040259 + z * 1.0572252;
    r = gamma_encode(t, r);
    g = gamma_encode(t, g);
    b = gamma_encode(t, b);
}

template <int WHICH>
__device__ __forceinline__ void transform_one(const PowTabs &t, double a, double b, double d,
                                              double &o0, double &o1, double &o2) {
    double x, y, z;
    if (WHICH == PB_T_SRGB_TO_ICTCP) { // ICtCp.c:120-146
        srgb_to_xyz(t, a, b, d, x, y, z);
        xyz_to_rec2020(x, y, z, o0, o1, o2);
        rec2020_to_ictcp(t, o0, o1, o2, o0, o1, o2);
    } else if (WHICH == PB_T_SRGB_TO_CIELUV) { // CIELuv.c:166-197
        a = gamma_decode(t, a); b = gamma_decode(t, b); d = gamma_decode(t, d);
        x = a * 0.4124564 + b * 0.3575761 + d * 0.1804375;
        y = a * 0.2126729 + b * 0.7151522 + d * 0.0721750;
        z = a * 0.0193339 + b * 0.1191920 + d * 0.9503041;
        xyz_to_cieluv(t, x, y, z, o0, o1, o2);
    } else if (WHICH == PB_T_ICTCP_TO_REC2020) { // rec2020.c:128-148
        ictcp_to_rec2020(t, a, b, d, o0, o1, o2);
    } else if (WHICH == PB_T_CIELUV_TO_REC2020) { // rec2020.c:150-173
        cieluv_to_xyz(t, a, b, d, x, y, z);
        xyz_to_rec2020(x, y, z, o0, o1, o2);
    } else if (WHICH == PB_T_SRGB_TO_REC2020) { // rec2020.c:175-195
        srgb_to_xyz(t, a, b, d, x, y, z);
        xyz_to_rec2020(x, y, z, o0, o1, o2);
    } else if (WHICH == PB_T_REC2020_TO_SRGB) { // sRGB.c:112-132
        rec2020_to_srgb(t, a, b, d, o0, o1, o2);
    } else { // PB_T_CIELUV_TO_ICTCP: the NN-map detour of patolette.c:305-314, fused
        cieluv_to_xyz(t, a, b, d, x, y, z);
        xyz_to_rec2020(x, y, z, o0, o1, o2);
        rec2020_to_srgb(t, o0, o1, o2, a, b, d);
        srgb_to_xyz(t, a, b, d, x, y, z);
        xyz_to_rec2020(x, y, z, o0, o1, o2);
        rec2020_to_ictcp(t, o0, o1, o2, o0, o1, o2);
    }
}

template <int WHICH>
__global__ void __launch_bounds__(256) k_color(const double *__restrict__ s0, const double *__restrict__ s1,
                                               const double *__restrict__ s2, double *__restrict__ d0,
                                               double *__restrict__ d1, double *__restrict__ d2, size_t n) {
    __shared__ uint64_t s_log[128 * 3];
    __shared__ uint64_t s_exp[256];
    for (int i = threadIdx.x; i < 128 * 3; i += blockDim.x) s_log[i] = g_log_tab[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_exp[i] = g_exp_tab[i];
    __syncthreads();
    PowTabs t{s_log, s_exp};
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double o0, o1, o2;
        transform_one<WHICH>(t, s0[i], s1[i], s2[i], o0, o1, o2);
        d0[i] = o0; d1[i] = o1; d2[i] = o2;
    }
}

// N x 3 row-major (interleaved RGB, numpy's default layout) -> three planes.  The reference's Python
// wrapper does this on the host with np.asfortranarray (patolette.pyx:388-391), ~0.2 s at 4096^2.
__global__ void __launch_bounds__(256) k_deinterleave(const double *__restrict__ rgb, size_t n, double *__restrict__ d0,
                                                      double *__restrict__ d1, double *__restrict__ d2) {
    __shared__ double tile[256 * 3];
    for (size_t base = (size_t)blockIdx.x * 256; base < n; base += (size_t)gridDim.x * 256) {
        const size_t cnt = min((size_t)256, n - base);
        for (size_t i = threadIdx.x; i < cnt * 3; i += 256) tile[i] = rgb[base * 3 + i]; // coalesced
        __syncthreads();
        if (threadIdx.x < cnt) {
            d0[base + threadIdx.x] = tile[3 * threadIdx.x];
            d1[base + threadIdx.x] = tile[3 * threadIdx.x + 1];
            d2[base + threadIdx.x] = tile[3 * threadIdx.x + 2];
        }
        __syncthreads();
    }
}

// N1: N x 3 interleaved uint8 RGB -> three f64 planes of value / 255 - the division the README's
// workflow performs on the host (README.md:155-158: astype(float64), then /= 255), IEEE-exact, so the planes hold
// the very doubles the f64 ABI would have been given.  24x fewer bytes cross PCIe.
__global__ void __launch_bounds__(256) k_u8_to_planes(const uint8_t *__restrict__ rgb, size_t n, double *__restrict__ d0,
                                                      double *__restrict__ d1, double *__restrict__ d2) {
    __shared__ uint8_t tile[1024 * 3];
    for (size_t base = (size_t)blockIdx.x * 1024; base < n; base += (size_t)gridDim.x * 1024) {
        const size_t cnt = min((size_t)1024, n - base);
        for (size_t i = threadIdx.x; i < cnt * 3; i += 256) tile[i] = rgb[base * 3 + i]; // coalesced bytes
        __syncthreads();
        for (size_t i = threadIdx.x; i < cnt; i += 256) {
            d0[base + i] = __ddiv_rn((double)tile[3 * i], 255.0);
            d1[base + i] = __ddiv_rn((double)tile[3 * i + 1], 255.0);
            d2[base + i] = __ddiv_rn((double)tile[3 * i + 2], 255.0);
        }
        __syncthreads();
    }
}

// palette_map as the index type a palette image stores: uint8 (K <= 256) or uint16 (K <= 65536) instead of size_t
template <typename T>
__global__ void __launch_bounds__(256) k_narrow_map(const unsigned long long *__restrict__ map, size_t n, T *__restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) out[i] = (T)map[i];
}

__global__ void k_pow(const double *x, double y, double *out, size_t n) {
    __shared__ uint64_t s_log[128 * 3];
    __shared__ uint64_t s_exp[256];
    for (int i = threadIdx.x; i < 128 * 3; i += blockDim.x) s_log[i] = g_log_tab[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_exp[i] = g_exp_tab[i];
    __syncthreads();
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = pb_pow(x[i], y, s_log, s_exp);
}

void ensure_tabs() {
    int dev = 0;
    PB_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 64 && g_tabs_ready[dev]) return;
    PB_CUDA_OK(cudaMemcpyToSymbol(g_log_tab, GLIBC_POW_LOG_TAB, sizeof(GLIBC_POW_LOG_TAB)));
    PB_CUDA_OK(cudaMemcpyToSymbol(g_exp_tab, GLIBC_EXP_TAB, sizeof(GLIBC_EXP_TAB)));
    if (dev < 64) g_tabs_ready[dev] = true;
}

int grid_for(size_t n, int threads, int sm_count) {
    size_t want = (n + threads - 1) / threads;
    size_t cap = (size_t)sm_count * 8; // persistent grid-stride: 8 CTAs of 256 per SM
    return (int)(want < cap ? (want ? want : 1) : cap);
}

} // namespace

void pb_launch_color(int which, const double *const src[3], double *const dst[3], size_t n,
                     int sm_count, cudaStream_t st) {
    if (n == 0) return;
    ensure_tabs();
    int grid = grid_for(n, 256, sm_count);
    PbProfScope _prof("k_color", st);
#define PB_CASE(W) \
    case W: k_color<W><<<grid, 256, 0, st>>>(src[0], src[1], src[2], dst[0], dst[1], dst[2], n); break;
    switch (which) {
        PB_CASE(PB_T_SRGB_TO_ICTCP)
        PB_CASE(PB_T_SRGB_TO_CIELUV)
        PB_CASE(PB_T_ICTCP_TO_REC2020)
        PB_CASE(PB_T_CIELUV_TO_REC2020)
        PB_CASE(PB_T_SRGB_TO_REC2020)
        PB_CASE(PB_T_REC2020_TO_SRGB)
        PB_CASE(PB_T_CIELUV_TO_ICTCP)
    default: break;
    }
#undef PB_CASE
    PB_CUDA_OK(cudaGetLastError());
}

void pb_launch_deinterleave(const double *d_rgb, size_t n, double *const dst[3], int sm_count, cudaStream_t st) {
    if (n == 0) return;
    PbProfScope _prof("k_deinterleave", st);
    k_deinterleave<<<grid_for(n, 256, sm_count), 256, 0, st>>>(d_rgb, n, dst[0], dst[1], dst[2]);
    PB_CUDA_OK(cudaGetLastError());
}

void pb_launch_u8_to_planes(const uint8_t *d_rgb, size_t n, double *const dst[3], int sm_count, cudaStream_t st) {
    if (n == 0) return;
    PbProfScope _prof("k_u8_to_planes", st);
    k_u8_to_planes<<<grid_for(n, 1024, sm_count), 256, 0, st>>>(d_rgb, n, dst[0], dst[1], dst[2]);
    PB_CUDA_OK(cudaGetLastError());
}

void pb_launch_narrow_map(const unsigned long long *d_map, size_t n, void *d_out, int bytes, int sm_count, cudaStream_t st) {
    if (n == 0) return;
    PbProfScope _prof("k_narrow_map", st, false);
    if (bytes == 1) k_narrow_map<uint8_t><<<grid_for(n, 256, sm_count), 256, 0, st>>>(d_map, n, (uint8_t *)d_out);
    else k_narrow_map<uint16_t><<<grid_for(n, 256, sm_count), 256, 0, st>>>(d_map, n, (uint16_t *)d_out);
    PB_CUDA_OK(cudaGetLastError());
}

void pb_launch_pow(const double *x, double y, double *out, size_t n, int sm_count, cudaStream_t st) {
    if (n == 0) return;
    ensure_tabs();
    { PbProfScope _prof("k_pow", st);
    k_pow<<<grid_for(n, 256, sm_count), 256, 0, st>>>(x, y, out, n);
    }
    PB_CUDA_OK(cudaGetLastError());
}
