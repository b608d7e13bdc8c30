/*
 * patolette_oracle.c - CPU restatement of patolette's pixel-array hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker for patolette_b200's CUDA
 * path.  It is compiled by oracle/build_oracle.py into
 * oracle/_build/libpatolette_oracle.so and may be loaded only by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 * Nothing under patolette_b200/ includes, links or dlopens it.
 *
 * Parity status: PINNED against the reference's own code.  The reference ships
 * no tests or golden vectors, so the pin is
 * oracle/_ref/libpatolette_ref.so - the reference's C/C++ compiled in place -
 * checked bit-for-bit against this restatement in tests/test_oracle_vs_ref.py,
 * plus the frozen outputs in tests/golden/.
 *
 * It is a restatement, not a copy: one translation unit, its own data layout
 * (explicit index lists + planar pixel arrays), every function citing the
 * reference file:line whose ARITHMETIC (operation order, rounding points,
 * tie-breaks) it reproduces.  All paths are relative to /root/reference.
 *
 * Third-party arithmetic the reference delegates and how it is restated:
 *   - LAPACK dsyev_ (math/eigen.c:50)     -> the same entry point of the OpenBLAS
 *     bundled with scipy (scipy_dsyev_), because the eigenvector SIGN follows no
 *     closed-form rule.
 *   - cblas_dgemv (quantize/sort.c:43)    -> per-row formula
 *     fma(a0,x0, a1*x1) + a2*x2, which is what OpenBLAS 0.3.31 Haswell/SkylakeX
 *     dgemv_n computes for a 3-column matrix; orc_selftest_dgemv() re-checks it
 *     against the live BLAS.
 *   - FLANN (palette/nearest.c)           -> exact f64 squared-L2 1-NN, lowest
 *     index on ties (FLANN's tie order is unspecified).
 *   - faiss 1.10.0 Clustering / IndexFlatL2 (generic build, sgemm path)
 *     -> restated in orc_kmeans() below.
 *   - libm pow/exp/log                      -> the host libm (glibc).
 *
 * Build: gcc -O2 -ffp-contract=off (no implicit FMA: the reference is a generic
 * x86-64 build without -mfma, CMakeLists.txt has no arch flags).
 */
#include <math.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))
#define ORC_DELTA 1e-16 /* math/misc.h:5 */
#define ORC_BUCKETS 512 /* quantize/global.c:22, quantize/local.c:15 */

/* ------------------------------------------------------------------------- */
/* public ABI types (lib/include/patolette.h:7-20)                            */
/* ------------------------------------------------------------------------- */
typedef enum { ORC_sRGB = 0, ORC_CIELuv = 1, ORC_ICtCp = 2 } orc_color_space;
typedef struct {
    bool dither;
    bool palette_only;
    int color_space;
    int kmeans_niter;
    size_t kmeans_max_samples;
    bool verbose;
} orc_options;

/* LAPACK from the scipy-bundled OpenBLAS (same entry point as math/eigen.c:50) */
extern void scipy_dsyev_(const char *jobz, const char *uplo, const int *n, double *a,
                         const int *lda, double *w, double *work, const int *lwork, int *info,
                         size_t, size_t);
extern void scipy_cblas_dgemv(int order, int trans, int m, int n, double alpha, const double *a,
                              int lda, const double *x, int incx, double beta, double *y,
                              int incy);

static inline double sq(double x) { return x * x; } /* math/misc.h:8 SQ */

/* ========================================================================= */
/* 1. colour transforms (lib/src/color/)                                      */
/* ========================================================================= */

/* color/sRGB.c:70-89 */
static double gamma_decode(double c) {
    double r = (c <= 0.0404500) ? c / 12.92 : pow((c + 0.055) / 1.055, 2.4);
    return fmin(fmax(r, 0.0), 1.0);
}
/* color/sRGB.c:91-110 */
static double gamma_encode(double c) {
    double r = (c <= 0.0031308) ? c * 12.92 : 1.055 * pow(c, 1.0 / 2.4) - 0.055;
    return fmin(fmax(r, 0.0), 1.0);
}
/* color/eotf.c:14-19 constants, :29-42 EOTF, :44-57 inverse EOTF */
static const double PQ_Lp = 10000, PQ_m1 = 0.1593017578125, PQ_m2 = 78.84375,
                    PQ_c1 = 0.8359375, PQ_c2 = 18.8515625, PQ_c3 = 18.6875;
static double pq_eotf(double c) {
    double m1d = 1 / PQ_m1, m2d = 1 / PQ_m2;
    double Vp = pow(c, m2d);
    double n = fmax(0, Vp - PQ_c1);
    double L = pow(n / (PQ_c2 - PQ_c3 * Vp), m1d);
    return PQ_Lp * L;
}
static double pq_inverse_eotf(double c) {
    double y = pow(c / PQ_Lp, PQ_m1);
    return pow((PQ_c1 + PQ_c2 * y) / (1 + PQ_c3 * y), PQ_m2);
}
/* color/xyz.c:14-40 (sRGB -> XYZ, D65 matrix after gamma decode) */
static void srgb_to_xyz(double r, double g, double b, double *x, double *y, double *z) {
    double R = gamma_decode(r), G = gamma_decode(g), B = gamma_decode(b);
    *x = R * 0.4124564 + G * 0.3575761 + B * 0.1804375;
    *y = R * 0.2126729 + G * 0.7151522 + B * 0.0721750;
    *z = R * 0.0193339 + G * 0.1191920 + B * 0.9503041;
}
/* color/xyz.c:42-64 */
static void rec2020_to_xyz(double r, double g, double b, double *x, double *y, double *z) {
    *x = r * 0.63695351 + g * 0.14461919 + b * 0.16885585;
    *y = r * 0.26269834 + g * 0.67800877 + b * 0.0592929;
    *z = g * 0.02807314 + b * 1.06082723;
}
/* color/rec2020.c:75-102 */
static void xyz_to_rec2020(double x, double y, double z, double *r, double *g, double *b) {
    *r = x * 1.71666343 + y * -0.35567332 + z * -0.25336809;
    *g = x * -0.66667384 + y * 1.61645574 + z * 0.0157683;
    *b = x * 0.01764248 + y * -0.04277698 + z * 0.94224328;
}
/* color/ICtCp.c:41-79 (Ct is halved on purpose) */
static void rec2020_to_ictcp(double r, double g, double b, double *I, double *Ct, double *Cp) {
    double L = (r * 1688 + g * 2146 + b * 262) / 4096;
    double M = (r * 683 + g * 2951 + b * 462) / 4096;
    double S = (r * 99 + g * 309 + b * 3688) / 4096;
    double L_ = pq_inverse_eotf(L), M_ = pq_inverse_eotf(M), S_ = pq_inverse_eotf(S);
    *I = L_ * 0.5 + M_ * 0.5;
    *Ct = (L_ * 6610 - M_ * 13613 + S_ * 7003) / 4096;
    *Cp = (L_ * 17933 - M_ * 17390 - S_ * 543) / 4096;
    /* ICtCp.c:78: the stored Ct is halved */
    *Ct = *Ct * 0.5;
}
/* color/rec2020.c:32-69 */
static void ictcp_to_rec2020(double I, double Ct, double Cp, double *r, double *g, double *b) {
    Ct *= 2;
    double L_ = I + 0.00860904 * Ct + 0.11102963 * Cp;
    double M_ = I - 0.00860904 * Ct - 0.11102963 * Cp;
    double S_ = I + 0.56003134 * Ct - 0.32062717 * Cp;
    double L = pq_eotf(L_), M = pq_eotf(M_), S = pq_eotf(S_);
    *r = L * 3.43660669 - M * 2.50645212 + S * 0.06984542;
    *g = -L * 0.79132956 + M * 1.98360045 - S * 0.1922709;
    *b = -L * 0.0259499 - M * 0.09891371 + S * 1.12486361;
}
/* color/CIELuv.c:19-24 constants */
static const double LUV_rwx = 0.95047, LUV_rwy = 1.0, LUV_rwz = 1.08883;
#define LUV_kE (216.0 / 24389.0)
#define LUV_kK (24389.0 / 27.0)
#define LUV_kKE 8.0
/* color/CIELuv.c:54-89 */
static void xyz_to_cieluv(double x, double y, double z, double *L, double *u, double *v) {
    double den = x + 15.0 * y + 3.0 * z;
    double up = (den > 0.0) ? ((4.0 * x) / (x + 15.0 * y + 3.0 * z)) : 0.0;
    double vp = (den > 0.0) ? ((9.0 * y) / (x + 15.0 * y + 3.0 * z)) : 0.0;
    double urp = (4.0 * LUV_rwx) / (LUV_rwx + 15.0 * LUV_rwy + 3.0 * LUV_rwz);
    double vrp = (9.0 * LUV_rwy) / (LUV_rwx + 15.0 * LUV_rwy + 3.0 * LUV_rwz);
    double yr = y / LUV_rwy;
    double L_ = (yr > LUV_kE) ? (116.0 * pow(yr, 1.0 / 3.0) - 16.0) : (LUV_kK * yr);
    *L = L_;
    *u = 13.0 * L_ * (up - urp);
    *v = 13.0 * L_ * (vp - vrp);
}
/* color/CIELuv.c:100-164 */
static void cieluv_to_xyz(double L, double u, double v, double *x, double *y, double *z) {
    double y_ = (L > LUV_kKE) ? pow((L + 16.0) / 116.0, 3.0) : (L / LUV_kK);
    double u0 = (4.0 * LUV_rwx) / (LUV_rwx + 15.0 * LUV_rwy + 3.0 * LUV_rwz);
    double v0 = (9.0 * LUV_rwy) / (LUV_rwx + 15.0 * LUV_rwy + 3.0 * LUV_rwz);
    double a, a_den = u + 13.0 * L * u0;
    a = (!a_den) ? 0 : (((52.0 * L) / a_den) - 1.0) / 3.0;
    double b = -5.0 * y_;
    double c = -1.0 / 3.0;
    double d, d_den = v + 13.0 * L * v0;
    d = (!d_den) ? 0 : y_ * (((39.0 * L) / d_den) - 5.0);
    double x_, x_den = a - c;
    x_ = (!x_den) ? 0 : (d - b) / x_den;
    double z_ = x_ * a + b;
    *x = x_;
    *y = y_;
    *z = z_;
}
/* color/sRGB.c:32-59 */
static void rec2020_to_srgb(double r2, double g2, double b2, double *r, double *g, double *b) {
    double x, y, z;
    rec2020_to_xyz(r2, g2, b2, &x, &y, &z);
    *r = x * 3.2404542 - y * 1.5371385 - z * 0.4985314;
    *g = -x * 0.9692660 + y * 1.8760108 + z * 0.0415560;
    *b = x * 0.0556434 - y * 0.2040259 + z * 1.0572252;
    *r = gamma_encode(*r);
    *g = gamma_encode(*g);
    *b = gamma_encode(*b);
}

/* Matrix-level transforms.  c = planar N x 3 (column-major, matrix2D.h:29). */
enum { ORC_T_SRGB_TO_ICTCP = 0, ORC_T_SRGB_TO_CIELUV = 1, ORC_T_ICTCP_TO_REC2020 = 2,
       ORC_T_CIELUV_TO_REC2020 = 3, ORC_T_SRGB_TO_REC2020 = 4, ORC_T_REC2020_TO_SRGB = 5 };

ORC_API void orc_color_transform(int which, double *c, size_t n) {
    double *c0 = c, *c1 = c + n, *c2 = c + 2 * n;
    for (size_t i = 0; i < n; i++) {
        double a = c0[i], b = c1[i], d = c2[i], x, y, z, o0, o1, o2;
        switch (which) {
        case ORC_T_SRGB_TO_ICTCP: /* ICtCp.c:120-146 -> rec2020.c:104-126 */
            srgb_to_xyz(a, b, d, &x, &y, &z);
            xyz_to_rec2020(x, y, z, &o0, &o1, &o2);
            rec2020_to_ictcp(o0, o1, o2, &o0, &o1, &o2);
            break;
        case ORC_T_SRGB_TO_CIELUV: /* CIELuv.c:166-197 */
            a = gamma_decode(a); b = gamma_decode(b); d = gamma_decode(d);
            x = a * 0.4124564 + b * 0.3575761 + d * 0.1804375;
            y = a * 0.2126729 + b * 0.7151522 + d * 0.0721750;
            z = a * 0.0193339 + b * 0.1191920 + d * 0.9503041;
            xyz_to_cieluv(x, y, z, &o0, &o1, &o2);
            break;
        case ORC_T_ICTCP_TO_REC2020: /* rec2020.c:128-148 */
            ictcp_to_rec2020(a, b, d, &o0, &o1, &o2);
            break;
        case ORC_T_CIELUV_TO_REC2020: /* rec2020.c:150-173 */
            cieluv_to_xyz(a, b, d, &x, &y, &z);
            xyz_to_rec2020(x, y, z, &o0, &o1, &o2);
            break;
        case ORC_T_SRGB_TO_REC2020: /* rec2020.c:175-195 */
            srgb_to_xyz(a, b, d, &x, &y, &z);
            xyz_to_rec2020(x, y, z, &o0, &o1, &o2);
            break;
        default: /* ORC_T_REC2020_TO_SRGB, sRGB.c:112-132 */
            rec2020_to_srgb(a, b, d, &o0, &o1, &o2);
            break;
        }
        c0[i] = o0; c1[i] = o1; c2[i] = o2;
    }
}

/* ========================================================================= */
/* 2. PCA, eigen solve, axis sort (lib/src/math/, lib/src/quantize/sort.c)    */
/* ========================================================================= */

/* A cluster is an ascending list of pixel indices into the planar dataset
 * (quantize/cluster.h:26-69).  The reference gathers rows lazily
 * (cluster.c:219 -> matrix2D.c:163); gathering is value-preserving, so the
 * restatement reads through the index list directly. */
typedef struct {
    const double *c0, *c1, *c2; /* planes of the dataset */
    const double *w;            /* per-pixel weights or NULL */
} orc_dataset;

/* array/matrix2D.c:200-233: mean_j = (sum_i c_ij * w_i) * (1 / sum_i w_i),
 * left-to-right sums, product rounded before the add, scale by the RECIPROCAL. */
static void weighted_mean(const orc_dataset *d, const uint32_t *idx, size_t n, bool use_w,
                          double mean[3], double *wsum_out) {
    const double *pl[3] = { d->c0, d->c1, d->c2 };
    for (int j = 0; j < 3; j++) {
        double m = 0;
        for (size_t i = 0; i < n; i++) {
            size_t p = idx ? idx[i] : i;
            double w = use_w ? d->w[p] : 1;
            double v = pl[j][p] * w;
            m += v;
        }
        mean[j] = m;
    }
    double s, wsum;
    if (!use_w) {
        wsum = (double)n;
        s = 1 / (double)n;
    } else {
        wsum = 0; /* array/vector.c:97-109 */
        for (size_t i = 0; i < n; i++) wsum += d->w[idx ? idx[i] : i];
        s = 1 / wsum;
    }
    for (int j = 0; j < 3; j++) mean[j] *= s; /* vector.c:111-121 */
    if (wsum_out) *wsum_out = wsum;
}

/* math/pca.c:62-101: centred copy (pca.c:33-60), then for every (j,k):
 * V_jk = (sum_i (w_i * c^_ij) * c^_ik) / wsum.  Column-major 3x3, V[k*3+j]. */
static void weighted_vcov(const orc_dataset *d, const uint32_t *idx, size_t n, bool use_w,
                          double vcov[9], double mean_out[3]) {
    const double *pl[3] = { d->c0, d->c1, d->c2 };
    double mean[3], wsum;
    weighted_mean(d, idx, n, use_w, mean, &wsum);
    if (mean_out) memcpy(mean_out, mean, sizeof mean);
    for (int j = 0; j < 3; j++) {
        for (int k = 0; k < 3; k++) {
            double value = 0;
            for (size_t i = 0; i < n; i++) {
                size_t p = idx ? idx[i] : i;
                double w = use_w ? d->w[p] : 1;
                double cij = pl[j][p] - mean[j];
                double cik = pl[k][p] - mean[k];
                value += w * cij * cik;
            }
            vcov[k * 3 + j] = value / wsum;
        }
    }
}

/* math/eigen.c:83-140: dsyev_('V','L',3): workspace query then solve; the
 * eigenvalues come back ascending and the eigenvectors overwrite the columns.
 * Returns false when the QUERY reports info != 0 (eigen.c:115-118). */
static bool eigen_solve3(double a[9], double evals[3]) {
    char jobz = 'V', uplo = 'L';
    int n = 3, lda = 3, lwork = -1, info = 0;
    double q[1];
    scipy_dsyev_(&jobz, &uplo, &n, a, &lda, NULL, q, &lwork, &info, 1, 1);
    if (info != 0) return false;
    lwork = (int)q[0];
    /* eigen.c:125 allocates lwork BYTES (bug B3); LAPACK needs lwork doubles.
     * We give it what it asked for - the results are the same when the
     * reference's heap overrun happens to be harmless. */
    double *work = malloc(sizeof(double) * (size_t)lwork);
    scipy_dsyev_(&jobz, &uplo, &n, a, &lda, evals, work, &lwork, &info, 1, 1);
    free(work);
    return true;
}

/* math/pca.c:122-149: the principal axis is the LAST eigenvector column. */
static bool pca_from_vcov(double vcov[9], double axis[3]) {
    double evals[3];
    if (!eigen_solve3(vcov, evals)) return false;
    axis[0] = vcov[6]; axis[1] = vcov[7]; axis[2] = vcov[8];
    return true;
}

/* Stage-level export: math/pca.c:151 patolette__PCA_perform_PCA on a cluster. */
ORC_API int orc_pca(const double *planar, size_t n_total, const double *weights,
                    const uint32_t *idx, size_t n, double mean[3], double vcov[9],
                    double axis[3]) {
    orc_dataset d = { planar, planar + n_total, planar + 2 * n_total, weights };
    double v[9];
    weighted_vcov(&d, idx, n, weights != NULL, v, mean);
    if (vcov) memcpy(vcov, v, sizeof v);
    return pca_from_vcov(v, axis) ? 0 : -1;
}

/* The per-row arithmetic of cblas_dgemv(ColMajor, NoTrans, n, 3, 1, A, n, x, 1, 0, y, 1)
 * as OpenBLAS 0.3.31 (x86-64 Haswell / SkylakeX dgemv_n, single thread) evaluates it:
 *   - rows are processed four at a time: the 2-column micro-kernel fuses a0*x0 onto the
 *     rounded a1*x1, the 1-column tail then adds the separately rounded a2*x2;
 *   - the last (n mod 4) rows go through the scalar tail loop temp += a[j]*x[j], which the
 *     FMA build contracts into fma(a2,x2, fma(a1,x1, a0*x0)).
 * Re-derived against the live BLAS by orc_selftest_dgemv(). */
static inline double dgemv_row3(double a0, double a1, double a2, const double x[3], size_t row,
                                size_t n) {
    if (row >= n - (n & 3)) return fma(a2, x[2], fma(a1, x[1], a0 * x[0]));
    return fma(a0, x[0], a1 * x[1]) + a2 * x[2];
}

/* quantize/sort.c:12-91.  bucket[i] for the i-th member of the cluster. */
static void axis_sort(const orc_dataset *d, const uint32_t *idx, size_t n, const double axis[3],
                      uint16_t *bucket) {
    double *dots = malloc(sizeof(double) * (n ? n : 1));
    for (size_t i = 0; i < n; i++) {
        size_t p = idx ? idx[i] : i;
        dots[i] = dgemv_row3(d->c0[p], d->c1[p], d->c2[p], axis, i, n);
    }
    /* array/vector.c:26-46: strict comparisons, first extremum wins */
    double mn = dots[0], mx = dots[0];
    for (size_t i = 0; i < n; i++) {
        if (dots[i] < mn) mn = dots[i];
        if (dots[i] > mx) mx = dots[i];
    }
    if (mx - mn < ORC_DELTA) { /* sort.c:61-79 round-robin */
        size_t j = 0;
        for (size_t i = 0; i < n; i++) {
            bucket[i] = (uint16_t)j;
            if (j >= ORC_BUCKETS - 1) j = 0; else j++;
        }
        free(dots);
        return;
    }
    double s = 1 / (mx - mn);
    for (size_t i = 0; i < n; i++) {
        double ratio = (dots[i] - mn) * s;
        size_t b = (size_t)((double)ORC_BUCKETS * ratio);
        bucket[i] = (uint16_t)(b < ORC_BUCKETS - 1 ? b : ORC_BUCKETS - 1);
    }
    free(dots);
}

ORC_API void orc_axis_sort(const double *planar, size_t n_total, const uint32_t *idx, size_t n,
                           const double axis[3], uint16_t *bucket) {
    orc_dataset d = { planar, planar + n_total, planar + 2 * n_total, NULL };
    axis_sort(&d, idx, n, axis, bucket);
}

/* Compares dgemv_row3 with the live BLAS; returns the number of mismatching rows. */
ORC_API long orc_selftest_dgemv(const double *planar, size_t n, const double axis[3]) {
    double *y = malloc(sizeof(double) * (n ? n : 1));
    scipy_cblas_dgemv(102, 111, (int)n, 3, 1.0, planar, (int)n, axis, 1, 0.0, y, 1);
    long bad = 0;
    for (size_t i = 0; i < n; i++) {
        double r = dgemv_row3(planar[i], planar[n + i], planar[2 * n + i], axis, i, n);
        if (memcmp(&r, &y[i], 8) != 0) bad++;
    }
    free(y);
    return bad;
}

/* ========================================================================= */
/* 3. clusters (lib/src/quantize/cluster.c)                                   */
/* ========================================================================= */
typedef struct orc_cluster {
    uint32_t *idx; /* ascending pixel indices, owned */
    size_t n;
    bool has_center, has_dist;
    double center[3], dist;
} orc_cluster;

static orc_cluster *cluster_new(uint32_t *idx, size_t n) {
    orc_cluster *c = calloc(1, sizeof *c);
    c->idx = idx; c->n = n;
    return c;
}
static void cluster_free(orc_cluster *c) {
    if (!c) return;
    free(c->idx); free(c);
}
/* cluster.c:171-189 -> matrix2D.c:200 (an empty cluster yields 0 * inf = NaN) */
static const double *cluster_center(const orc_dataset *d, orc_cluster *c) {
    if (!c->has_center) {
        weighted_mean(d, c->idx, c->n, d->w != NULL, c->center, NULL);
        c->has_center = true;
    }
    return c->center;
}
/* cluster.c:111-152: D = sum_i ((cx-x)^2 + (cy-y)^2 + (cz-z)^2) * w_i */
static double cluster_distortion(const orc_dataset *d, orc_cluster *c) {
    if (c->has_dist) return c->dist;
    const double *m = cluster_center(d, c);
    double x = m[0], y = m[1], z = m[2], dist = 0;
    for (size_t i = 0; i < c->n; i++) {
        size_t p = c->idx[i];
        double w = d->w ? d->w[p] : 1;
        double e = (sq(d->c0[p] - x) + sq(d->c1[p] - y) + sq(d->c2[p] - z)) * w;
        dist += e;
    }
    c->dist = dist; c->has_dist = true;
    return dist;
}

/* ========================================================================= */
/* 4. local quantizer (lib/src/quantize/local.c)                              */
/* ========================================================================= */
typedef struct { orc_cluster *left, *right; } orc_pair;

/* local.c:102-177 */
static size_t optimal_bucket(const orc_dataset *d, const orc_cluster *c, const uint16_t *bucket) {
    size_t sizes[ORC_BUCKETS];
    double sums[3][ORC_BUCKETS];
    memset(sizes, 0, sizeof sizes);
    memset(sums, 0, sizeof sums);
    for (size_t i = 0; i < c->n; i++) {
        size_t p = c->idx[i], b = bucket[i];
        double w = d->w ? d->w[p] : 1;
        sums[0][b] += d->c0[p] * w;
        sums[1][b] += d->c1[p] * w;
        sums[2][b] += d->c2[p] * w;
        sizes[b] += w; /* local.c:133: size_t += double (bug B4: truncating) */
    }
    for (size_t i = 1; i < ORC_BUCKETS; i++)
        for (int j = 0; j < 3; j++) sums[j][i] += sums[j][i - 1];
    for (size_t i = 1; i < ORC_BUCKETS; i++) sizes[i] += sizes[i - 1];
    double best = 0; size_t loc = 0;
    for (size_t i = 0; i < ORC_BUCKETS; i++) {
        double obj = 0;
        for (int j = 0; j < 3; j++) {
            double csl = sums[j][i];
            double csr = sums[j][ORC_BUCKETS - 1] - csl;
            double sl = (double)sizes[i];
            double sr = (double)(sizes[ORC_BUCKETS - 1] - sizes[i]);
            double v = 0;
            if (sl != 0) v += sq(csl) / sl;
            if (sr != 0) v += sq(csr) / sr;
            obj += v;
        }
        /* vector.c:26-46 maxloc: starts from element 0, strict '>' */
        if (i == 0 || obj > best) { best = obj; loc = i; }
    }
    return loc;
}

/* local.c:179-254.  NULL when the cluster cannot be split. */
static orc_pair *split_cluster(const orc_dataset *d, orc_cluster *c) {
    if (c->n <= 1) return NULL;
    double vcov[9], axis[3];
    weighted_vcov(d, c->idx, c->n, d->w != NULL, vcov, NULL);
    if (!pca_from_vcov(vcov, axis)) return NULL;
    uint16_t *bucket = malloc(sizeof(uint16_t) * c->n);
    axis_sort(d, c->idx, c->n, axis, bucket);
    size_t split = optimal_bucket(d, c, bucket);
    size_t nl = 0;
    for (size_t i = 0; i < c->n; i++) nl += (bucket[i] <= split);
    size_t nr = c->n - nl;
    uint32_t *li = malloc(sizeof(uint32_t) * (nl ? nl : 1));
    uint32_t *ri = malloc(sizeof(uint32_t) * (nr ? nr : 1));
    size_t pl = 0, pr = 0;
    for (size_t i = 0; i < c->n; i++) {
        if (bucket[i] <= split) li[pl++] = c->idx[i]; else ri[pr++] = c->idx[i];
    }
    free(bucket);
    orc_pair *pair = malloc(sizeof *pair);
    pair->left = cluster_new(li, nl);
    pair->right = cluster_new(ri, nr);
    return pair;
}

/* local.c:256-275 */
static double split_benefit(const orc_dataset *d, orc_cluster *c, orc_pair *ch) {
    if (!ch) return 0;
    double dd = cluster_distortion(d, c);
    double dl = cluster_distortion(d, ch->left);
    double dr = cluster_distortion(d, ch->right);
    return dd - (dl + dr);
}

/* local.c:318-404.  clusters[0..*count) in, up to K out (array has room for K). */
static void lq_quantize(const orc_dataset *d, orc_cluster **clusters, size_t *count, size_t K) {
    size_t len = *count;
    if (len >= K) return;
    orc_pair **children = calloc(K, sizeof *children);
    for (size_t i = 0; i < len; i++) children[i] = split_cluster(d, clusters[i]);
    size_t i;
    for (i = len; i < K; i++) {
        /* local.c:277-307 + vector.c:26-46: first maximum of the benefits */
        size_t best = 0; double bb = 0;
        for (size_t j = 0; j < i; j++) {
            double b = children[j] ? split_benefit(d, clusters[j], children[j]) : 0;
            if (j == 0 || b > bb) { bb = b; best = j; }
        }
        if (bb < ORC_DELTA) break; /* local.c:365-370 */
        orc_cluster *left = children[best]->left, *right = children[best]->right;
        cluster_free(clusters[best]);
        free(children[best]);
        clusters[i] = left;        /* local.c:375 */
        clusters[best] = right;    /* local.c:376 */
        children[i] = split_cluster(d, left);
        children[best] = split_cluster(d, right);
    }
    for (size_t j = 0; j < K; j++) {
        if (children[j]) {
            cluster_free(children[j]->left); cluster_free(children[j]->right); free(children[j]);
        }
    }
    free(children);
    *count = i;
}

/* ========================================================================= */
/* 5. global quantizer (lib/src/quantize/cells.c, global.c)                   */
/* ========================================================================= */
#define ORC_CELLS (ORC_BUCKETS + 1) /* 1-based buckets, cells.c:72-76 */
typedef struct {
    uint64_t w0[ORC_CELLS];
    double w1[3][ORC_CELLS];
    double w2[ORC_CELLS];
    double wrs[3][3][ORC_CELLS]; /* [r][s], r <= s */
} orc_cells;

/* cells.c:53-139 */
static void cells_preprocess(const orc_dataset *d, size_t n, const uint16_t *bucket, orc_cells *m) {
    memset(m, 0, sizeof *m);
    const double *pl[3] = { d->c0, d->c1, d->c2 };
    for (size_t i = 0; i < n; i++) {
        size_t j = (size_t)bucket[i] + 1;
        double cx = d->c0[i], cy = d->c1[i], cz = d->c2[i];
        m->w0[j] += 1;
        m->w1[0][j] += cx;
        m->w1[1][j] += cy;
        m->w1[2][j] += cz;
        m->w2[j] += (sq(cx) + sq(cy) + sq(cz));
    }
    for (size_t i = 0; i < n; i++) {
        size_t j = (size_t)bucket[i] + 1;
        for (int s = 0; s < 3; s++)
            for (int r = 0; r <= s; r++) m->wrs[r][s][j] += pl[r][i] * pl[s][i];
    }
    for (size_t i = 1; i < ORC_CELLS; i++) {
        m->w0[i] += m->w0[i - 1];
        m->w2[i] += m->w2[i - 1];
        for (int j = 0; j < 3; j++) m->w1[j][i] += m->w1[j][i - 1];
        for (int s = 0; s < 3; s++)
            for (int r = 0; r <= s; r++) m->wrs[r][s][i] += m->wrs[r][s][i - 1];
    }
}
/* cells.c:141-182 */
static double cell_distortion(size_t a, size_t b, const orc_cells *m) {
    if (m->w0[a] == m->w0[b]) return 0;
    return m->w2[b] - m->w2[a] -
           (sq(m->w1[0][b] - m->w1[0][a]) + sq(m->w1[1][b] - m->w1[1][a]) +
            sq(m->w1[2][b] - m->w1[2][a])) / (double)(m->w0[b] - m->w0[a]);
}
/* cells.c:184-259: cell covariance from the cumulative moments, then PCA */
static bool cell_pca(size_t a, size_t b, const orc_cells *m, double axis[3]) {
    double v[9] = { 0 };
    for (int s = 0; s < 3; s++) {
        for (int r = 0; r <= s; r++) {
            double e = 0;
            if (m->w0[a] != m->w0[b]) {
                double cnt = (double)(m->w0[b] - m->w0[a]);
                e = (m->wrs[r][s][b] - m->wrs[r][s][a]) / cnt -
                    (m->w1[r][b] - m->w1[r][a]) * (m->w1[s][b] - m->w1[s][a]) / sq(cnt);
            }
            v[s * 3 + r] = e; /* (row r, col s) */
        }
    }
    v[0 * 3 + 2] = v[2 * 3 + 0]; /* (2,0) = (0,2) */
    v[0 * 3 + 1] = v[1 * 3 + 0]; /* (1,0) = (0,1) */
    v[1 * 3 + 2] = v[2 * 3 + 1]; /* (2,1) = (1,2) */
    return pca_from_vcov(v, axis);
}
/* array/vector.c:135-159 norm = sqrt(sum pow(v_i, 2)) */
static double norm3(const double v[3]) {
    double s = 0;
    for (int i = 0; i < 3; i++) s += pow(v[i], 2);
    return sqrt(s);
}
/* cells.c:280-328; -1 on PCA failure */
static double cell_bias(size_t a, size_t b, const double axis[3], const orc_cells *m) {
    double ca[3];
    if (!cell_pca(a, b, m, ca)) return -1;
    double norms = norm3(axis) * norm3(ca);
    if (norms < ORC_DELTA) return 0;
    double dot = ca[0] * axis[0] + ca[1] * axis[1] + ca[2] * axis[2];
    return fmin(1, fabs(dot / norms));
}
/* global.c:99-187 */
static bool gq_should_terminate(const size_t *q, size_t qlen, const double axis[3],
                                const orc_cells *m, bool *error) {
    double distortion = 0;
    for (size_t j = 0; j + 1 < qlen; j++) distortion += cell_distortion(q[j], q[j + 1], m);
    if (distortion < ORC_DELTA) return true;
    double bias = 0;
    for (size_t i = 0; i + 1 < qlen; i++) {
        double cd = cell_distortion(q[i], q[i + 1], m);
        double cb = cell_bias(q[i], q[i + 1], axis, m);
        if (cb < 0) { *error = true; return true; }
        if (cb < 0.9) continue;            /* global.c:21 cell_bias_threshold */
        bias += (cd / distortion) * cb;
    }
    return bias < 0.1;                      /* global.c:20 bias_threshold */
}
/* global.c:72-97 */
static void l_chain(const double *L, size_t ld, size_t k, size_t N, size_t *chain) {
    size_t t = N;
    for (size_t j = k - 1; j >= 1; j--) {
        t = (size_t)L[t * ld + (j + 1)]; /* column-major (row j+1, col t) */
        chain[j] = t;
    }
    chain[0] = 0;
    chain[k] = N;
}
/* global.c:189-298.  Returns the cut count k (cells) and fills q[0..k]; 0 on error. */
static size_t principal_quantizer(size_t K, const orc_cells *m, size_t *q) {
    bool error = false;
    const size_t N = ORC_CELLS - 1, max_k = 12;
    double axis[3];
    if (!cell_pca(0, N, m, axis)) return 0;
    double *E = calloc(N + 1, sizeof(double)), *E2 = calloc(N + 1, sizeof(double));
    size_t ls = (K > N ? K : N) + 1;
    double *L = calloc(ls * ls, sizeof(double));
    for (size_t i = 1; i <= N; i++) E[i] = cell_distortion(0, i, m);
    for (size_t i = 1; i <= K; i++) L[i * ls + i] = (double)i;
    size_t k_out = 1;
    l_chain(L, ls, 1, N, q);
    size_t kmax = max_k < K ? max_k : K;
    for (size_t k = 2; k <= kmax; k++) {
        if (gq_should_terminate(q, k_out + 1, axis, m, &error)) break;
        memcpy(E2, E, sizeof(double) * (N + 1));
        for (size_t n = k + 1; n <= N; n++) {
            double cut = (double)(n - 1);
            double e = E2[n - 1];
            for (size_t t = n - 2; t >= k - 1; t--) {
                double c = E2[t] + cell_distortion(t, n, m);
                if (c < e) { cut = (double)t; e = c; }
            }
            L[n * ls + k] = cut;
            E[n] = e;
        }
        l_chain(L, ls, k, N, q);
        k_out = k;
    }
    free(E); free(E2); free(L);
    (void)error; /* global.c:250-262: the error flag can never be observed (break comes first) */
    return k_out;
}

/* global.c:388-443 + get_color_clusters :300-377.  Returns cluster count (0 = error). */
static size_t gq_quantize(const orc_dataset *d, size_t n, size_t K, orc_cluster **out) {
    double vcov[9], axis[3];
    weighted_vcov(d, NULL, n, false, vcov, NULL); /* global.c:407: UNWEIGHTED */
    if (!pca_from_vcov(vcov, axis)) return 0;
    uint16_t *bucket = malloc(sizeof(uint16_t) * n);
    axis_sort(d, NULL, n, axis, bucket);
    orc_cells *m = malloc(sizeof *m);
    cells_preprocess(d, n, bucket, m);
    size_t q[16];
    size_t cells = principal_quantizer(K, m, q);
    free(m);
    if (cells == 0) { free(bucket); return 0; }
    /* bucket -> cell: first j with bucket + 1 <= q[j + 1] (global.c:322-332) */
    uint8_t lut[ORC_BUCKETS];
    for (size_t b = 0; b < ORC_BUCKETS; b++) {
        lut[b] = 0;
        for (size_t j = 0; j < cells; j++) if (b + 1 <= q[j + 1]) { lut[b] = (uint8_t)j; break; }
    }
    size_t sizes[16] = { 0 }, piv[16] = { 0 };
    for (size_t i = 0; i < n; i++) sizes[lut[bucket[i]]]++;
    uint32_t *lists[16];
    for (size_t j = 0; j < cells; j++) lists[j] = malloc(sizeof(uint32_t) * (sizes[j] ? sizes[j] : 1));
    for (size_t i = 0; i < n; i++) { size_t j = lut[bucket[i]]; lists[j][piv[j]++] = (uint32_t)i; }
    for (size_t j = 0; j < cells; j++) out[j] = cluster_new(lists[j], sizes[j]);
    free(bucket);
    return cells;
}

/* Stage-level export: GQ + LQ.  labels[i] = slot of the cluster holding pixel i,
 * centers = K x 3 row-major cluster centres (palette/create.c:11-33). */
ORC_API int orc_quantize_clusters(const double *planar, size_t n, const double *weights, size_t K,
                                  uint32_t *labels, double *centers, size_t *count_out,
                                  size_t *gq_count_out) {
    orc_dataset d = { planar, planar + n, planar + 2 * n, weights };
    size_t cap = K > 16 ? K : 16;
    orc_cluster **cl = calloc(cap, sizeof *cl);
    size_t count = gq_quantize(&d, n, K, cl);
    if (gq_count_out) *gq_count_out = count;
    if (count == 0) { free(cl); return -1; }
    lq_quantize(&d, cl, &count, K);
    for (size_t j = 0; j < count; j++) {
        const double *c = cluster_center(&d, cl[j]);
        if (centers) memcpy(centers + 3 * j, c, 3 * sizeof(double));
        if (labels) for (size_t i = 0; i < cl[j]->n; i++) labels[cl[j]->idx[i]] = (uint32_t)j;
        cluster_free(cl[j]);
    }
    free(cl);
    *count_out = count;
    return 0;
}

/* ========================================================================= */
/* 6. nearest-palette map (lib/src/palette/nearest.c + FLANN contract)        */
/* ========================================================================= */
/* Exact squared-L2 1-NN in f64; dims summed 0,1,2; lowest index on ties. */
static size_t nearest_palette(const double *pal /* K x 3 row-major */, size_t K, double x,
                              double y, double z) {
    size_t best = 0; double bd = 0;
    for (size_t j = 0; j < K; j++) {
        double dx = x - pal[3 * j], dy = y - pal[3 * j + 1], dz = z - pal[3 * j + 2];
        double dd = dx * dx; dd += dy * dy; dd += dz * dz;
        if (j == 0 || dd < bd) { bd = dd; best = j; }
    }
    return best;
}
/* nearest.c:150-209 */
ORC_API void orc_fill_palette_map_nearest(const double *planar, size_t n, const double *pal_rm,
                                          size_t K, size_t *map) {
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; i++)
        map[i] = nearest_palette(pal_rm, K, planar[i], planar[n + i], planar[2 * n + i]);
}

/* ========================================================================= */
/* 7. KMeans refinement (lib/src/palette/refine.c + vendored faiss 1.10.0)    */
/* ========================================================================= */
/* std::mt19937 as used by faiss::RandomGenerator (faiss/utils/random.cpp:35-55) */
typedef struct { uint32_t s[624]; int i; } orc_mt;
static void mt_seed(orc_mt *m, uint32_t seed) {
    m->s[0] = seed;
    for (int i = 1; i < 624; i++) m->s[i] = 1812433253u * (m->s[i - 1] ^ (m->s[i - 1] >> 30)) + (uint32_t)i;
    m->i = 624;
}
static uint32_t mt_next(orc_mt *m) {
    if (m->i >= 624) {
        for (int k = 0; k < 624; k++) {
            uint32_t y = (m->s[k] & 0x80000000u) | (m->s[(k + 1) % 624] & 0x7fffffffu);
            m->s[k] = m->s[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        m->i = 0;
    }
    uint32_t y = m->s[m->i++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
}
/* random.cpp:184-194 rand_perm: forward Fisher-Yates, rand_int(max) = mt() % max */
static void rand_perm(int *perm, size_t n, int64_t seed) {
    for (size_t i = 0; i < n; i++) perm[i] = (int)i;
    orc_mt rng; mt_seed(&rng, (uint32_t)seed);
    for (size_t i = 0; i + 1 < n; i++) {
        int i2 = (int)(i + (size_t)((uint64_t)mt_next(&rng) % (uint64_t)(int)(n - i)));
        int t = perm[i]; perm[i] = perm[i2]; perm[i2] = t;
    }
}

/* IndexFlatL2::search k=1 on the GENERIC faiss build (utils/distances.cpp:259-343,
 * impl/ResultHandler.h Top1): dis = (|x|^2 + |y|^2) - 2*ip, clamped at 0, with
 * ip as OpenBLAS sgemm_ evaluates a k=3 dot: fmaf(x2,y2, fmaf(x1,y1, x0*y0));
 * norms ((x0^2 + x1^2) + x2^2); strict '<' over ascending centroid index.
 * For fewer than 20 queries faiss takes the sequential path instead
 * (distances.cpp:813-818 -> fvec_L2sqr: sum of (x-y)^2 in order). */
static void kmeans_assign(const float *x, size_t nx, const float *cen, size_t k, int64_t *assign,
                          float *dis) {
    if (nx < 20) {
        for (size_t i = 0; i < nx; i++) {
            size_t best = 0; float bd = 0;
            for (size_t j = 0; j < k; j++) {
                float s = 0;
                for (int t = 0; t < 3; t++) { float df = x[3 * i + t] - cen[3 * j + t]; s += df * df; }
                if (j == 0 || s < bd) { bd = s; best = j; }
            }
            assign[i] = (int64_t)best; dis[i] = bd;
        }
        return;
    }
    float *yn = malloc(sizeof(float) * k);
    for (size_t j = 0; j < k; j++) {
        const float *y = cen + 3 * j;
        float s = y[0] * y[0]; s += y[1] * y[1]; s += y[2] * y[2];
        yn[j] = s;
    }
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)nx; i++) {
        const float *xi = x + 3 * i;
        float xn = xi[0] * xi[0]; xn += xi[1] * xi[1]; xn += xi[2] * xi[2];
        size_t best = 0; float bd = 0;
        for (size_t j = 0; j < k; j++) {
            const float *y = cen + 3 * j;
            float ip = fmaf(xi[2], y[2], fmaf(xi[1], y[1], xi[0] * y[0]));
            float dd = xn + yn[j] - 2 * ip;
            if (dd < 0) dd = 0;
            if (j == 0 || dd < bd) { bd = dd; best = j; }
        }
        assign[i] = (int64_t)best; dis[i] = bd;
    }
    free(yn);
}

/* faiss/Clustering.cpp:587-603 kmeans_clustering -> :267-554 train_encoded with the
 * parameters of refine.c:77-89 (nredo 1, min_points 1, seed 1234, d = 3).
 * x: n x 3 row-major f32 (refine.c:122-146), cen: k x 3 in/out, w: n or NULL.
 * Returns 0, or -1 when faiss would throw (n < k): centres stay untouched (bug B8). */
ORC_API int orc_kmeans(const float *x_in, size_t n, size_t k, float *cen, const float *w_in,
                       int niter, int max_points_per_centroid) {
    if (n < k) return -1; /* Clustering.cpp:273-279 */
    for (size_t i = 0; i < n * 3; i++) /* :295-304 */
        if (!isfinite(x_in[i])) return -1;
    const float *x = x_in, *w = w_in;
    float *xs = NULL, *ws = NULL;
    size_t nx = n;
    if (nx > k * (size_t)max_points_per_centroid) { /* :311-319 -> :70-120 */
        int *perm = malloc(sizeof(int) * nx);
        rand_perm(perm, nx, 1234);
        nx = k * (size_t)max_points_per_centroid;
        xs = malloc(sizeof(float) * 3 * nx);
        for (size_t i = 0; i < nx; i++) memcpy(xs + 3 * i, x_in + 3 * (size_t)perm[i], 12);
        if (w_in) {
            ws = malloc(sizeof(float) * nx);
            for (size_t i = 0; i < nx; i++) ws[i] = w_in[perm[i]];
        }
        free(perm);
        x = xs; w = ws;
    }
    if (nx == k) { /* :330-352: "just copying" the (original) training set */
        memcpy(cen, x_in, sizeof(float) * 3 * k);
        free(xs); free(ws);
        return 0;
    }
    /* :413 rand_perm(seed + 1) only feeds centroids beyond the provided ones: none here. */
    int64_t *assign = malloc(sizeof(int64_t) * nx);
    float *dis = malloc(sizeof(float) * nx);
    float *hassign = malloc(sizeof(float) * k);
    for (int it = 0; it < niter; it++) {
        kmeans_assign(x, nx, cen, k, assign, dis);
        /* compute_centroids :135-204: per centroid, sequential f32 sums in sample order */
        memset(hassign, 0, sizeof(float) * k);
        memset(cen, 0, sizeof(float) * 3 * k);
        for (size_t i = 0; i < nx; i++) {
            int64_t ci = assign[i];
            float *c = cen + 3 * ci;
            const float *xi = x + 3 * i;
            if (w) {
                float wi = w[i];
                hassign[ci] += wi;
                for (int j = 0; j < 3; j++) c[j] += xi[j] * wi;
            } else {
                hassign[ci] += 1.0;
                for (int j = 0; j < 3; j++) c[j] += xi[j];
            }
        }
        for (size_t ci = 0; ci < k; ci++) {
            if (hassign[ci] == 0) continue;
            float norm = 1 / hassign[ci];
            for (int j = 0; j < 3; j++) cen[3 * ci + j] *= norm;
        }
        /* split_clusters :216-263: refill empty clusters */
        orc_mt rng; mt_seed(&rng, 1234u);
        for (size_t ci = 0; ci < k; ci++) {
            if (hassign[ci] != 0) continue;
            size_t cj;
            for (cj = 0; 1; cj = (cj + 1) % k) {
                float p = (hassign[cj] - 1.0) / (float)(nx - k);
                float r = (float)(uint64_t)mt_next(&rng) / 4294967295.0f; /* mt() / float(mt.max()) */
                if (r < p) break;
            }
            memcpy(cen + 3 * ci, cen + 3 * cj, 12);
            for (int j = 0; j < 3; j++) {
                if (j % 2 == 0) { cen[3 * ci + j] *= 1 + (1 / 1024.); cen[3 * cj + j] *= 1 - (1 / 1024.); }
                else { cen[3 * ci + j] *= 1 - (1 / 1024.); cen[3 * cj + j] *= 1 + (1 / 1024.); }
            }
            hassign[ci] = hassign[cj] / 2;
            hassign[cj] -= hassign[ci];
        }
    }
    free(assign); free(dis); free(hassign); free(xs); free(ws);
    return 0;
}

/* ========================================================================= */
/* 8. Riemersma dither (lib/src/dither/riemersma.c)                           */
/* ========================================================================= */
/* The recursion of riemersma.c:176-257 started at (0,0) heading UP visits the
 * 2^level square in textbook Hilbert order (d2xy with x = column, y = row);
 * cells outside the image are skipped without touching the queue (:146-156).
 * It is restated here as the same recursion, iteratively unrolled per cell. */
typedef struct {
    size_t x, y, W, H;
    const double *c0, *c1, *c2;
    const double *pal;      /* K x 3 row-major, linear Rec2020 */
    double *palw;           /* K x 3 scaled by (float)sqrt-luma weights (:419-425) */
    size_t K;
    double q[16][3], qw[16];
    size_t *map;
} orc_dither;
static const double DW_R = 0.51254268114958, DW_G = 0.8234075540095561, DW_B = 0.2435159132377184;

static void dither_pixel(orc_dither *s) { /* riemersma.c:275-341 */
    double eR = 0, eG = 0, eB = 0;
    for (int i = 0; i < 16; i++) {
        eR += s->q[i][0] * s->qw[i];
        eG += s->q[i][1] * s->qw[i];
        eB += s->q[i][2] * s->qw[i];
    }
    size_t p = s->y * s->W + s->x;
    double R = s->c0[p], G = s->c1[p], B = s->c2[p];
    double cR = R + eR, cG = G + eG, cB = B + eB;
    size_t idx = nearest_palette(s->palw, s->K, DW_R * cR, DW_G * cG, DW_B * cB);
    s->map[p] = idx;
    memmove(&s->q[0], &s->q[1], sizeof(double) * 3 * 15);
    s->q[15][0] = R - s->pal[3 * idx];
    s->q[15][1] = G - s->pal[3 * idx + 1];
    s->q[15][2] = B - s->pal[3 * idx + 2];
}
enum { D_NONE, D_UP, D_LEFT, D_RIGHT, D_DOWN };
static void dmove(orc_dither *s, int dir) { /* :146-174 (size_t wrap-around kept) */
    if (s->x < s->W && s->y < s->H) dither_pixel(s);
    switch (dir) {
    case D_LEFT: s->x--; break;
    case D_RIGHT: s->x++; break;
    case D_UP: s->y--; break;
    case D_DOWN: s->y++; break;
    default: break;
    }
}
static void traverse(orc_dither *s, int level, int dir) { /* :176-257 */
    static const int first[5] = { 0, D_LEFT, D_UP, D_DOWN, D_RIGHT };   /* sub-curve 1 */
    static const int last[5] = { 0, D_RIGHT, D_DOWN, D_UP, D_LEFT };    /* sub-curve 4 */
    static const int m1[5] = { 0, D_DOWN, D_RIGHT, D_LEFT, D_UP };
    static const int m2[5] = { 0, D_RIGHT, D_DOWN, D_UP, D_LEFT };
    static const int m3[5] = { 0, D_UP, D_LEFT, D_RIGHT, D_DOWN };
    if (dir == D_NONE) return;
    if (level == 1) {
        dmove(s, m1[dir]); dmove(s, m2[dir]); dmove(s, m3[dir]);
        return;
    }
    traverse(s, level - 1, first[dir]);
    dmove(s, m1[dir]);
    traverse(s, level - 1, dir);
    dmove(s, m2[dir]);
    traverse(s, level - 1, dir);
    dmove(s, m3[dir]);
    traverse(s, level - 1, last[dir]);
}
/* riemersma.c:437-459.  planar = colours in linear Rec2020, pal row-major K x 3. */
ORC_API void orc_dither_riemersma(const double *planar, size_t W, size_t H, const double *pal,
                                  size_t K, size_t *map) {
    orc_dither s;
    memset(&s, 0, sizeof s);
    size_t n = W * H;
    s.W = W; s.H = H; s.c0 = planar; s.c1 = planar + n; s.c2 = planar + 2 * n;
    s.pal = pal; s.K = K; s.map = map;
    s.palw = malloc(sizeof(double) * 3 * (K ? K : 1));
    double fx = (double)(float)DW_R, fy = (double)(float)DW_G, fz = (double)(float)DW_B;
    for (size_t j = 0; j < K; j++) { /* nearest.c:32-61 build_index_data */
        s.palw[3 * j] = pal[3 * j] * fx;
        s.palw[3 * j + 1] = pal[3 * j + 1] * fy;
        s.palw[3 * j + 2] = pal[3 * j + 2] * fz;
    }
    double m = exp(log((double)16) / ((double)16 - 1)), v = 1; /* :360-373 */
    for (int i = 0; i < 16; i++) { s.qw[i] = v / (double)16; v *= m; }
    int level = 0; /* :124-144 */
    size_t mx = W > H ? W : H, value = mx;
    while (value > 1) { value >>= 1; level++; }
    if (((size_t)1 << level) < mx) level++;
    if (level > 0) { traverse(&s, level, D_UP); dmove(&s, D_NONE); }
    free(s.palw);
}

/* ========================================================================= */
/* 9. public C ABI (lib/src/patolette.c)                                      */
/* ========================================================================= */
static const char *orc_messages[5] = { /* patolette.c:32-38 */
    "Quantization successful.", "Internal quantization error.",
    "Image dimensions should be greater than 0.", "Palette size should be greater than 0.",
    "Image dimensions are too big.",
};
ORC_API const char *get_patolette_exit_code_info_message(int exit_code) { /* :97-105 */
    return orc_messages[-1 * exit_code];
}
ORC_API orc_options *patolette_create_default_options(void) { /* :107-119 */
    orc_options *o = malloc(sizeof *o);
    o->dither = true; o->palette_only = false; o->color_space = ORC_ICtCp;
    o->kmeans_niter = 32; o->kmeans_max_samples = 512 * 512; o->verbose = false;
    return o;
}

/* K x 3 row-major palette helpers (the reference keeps it as a column-major
 * Matrix2D; the per-row arithmetic is identical). */
static void palette_transform(int which, double *pal_rm, size_t K) {
    double *t = malloc(sizeof(double) * 3 * (K ? K : 1));
    for (size_t j = 0; j < K; j++) { t[j] = pal_rm[3 * j]; t[K + j] = pal_rm[3 * j + 1]; t[2 * K + j] = pal_rm[3 * j + 2]; }
    orc_color_transform(which, t, K);
    for (size_t j = 0; j < K; j++) { pal_rm[3 * j] = t[j]; pal_rm[3 * j + 1] = t[K + j]; pal_rm[3 * j + 2] = t[2 * K + j]; }
    free(t);
}

/* patolette.c:157-343 */
ORC_API void patolette(size_t width, size_t height, const double *color_data,
                       const double *weight_data, size_t palette_size, const orc_options *opt,
                       double *palette, size_t *palette_map, int *exit_code) {
    *exit_code = 0; /* validate_arguments :61-95 */
    size_t n = width * height;
    if (n == 0) { *exit_code = -2; return; }
    if (palette_size < 1) { *exit_code = -3; return; }
    if (width * height > (size_t)40000 * 40000) { *exit_code = -4; return; }

    double *colors = malloc(sizeof(double) * 3 * n); /* :187 copy, input never mutated */
    memcpy(colors, color_data, sizeof(double) * 3 * n);
    double *weights = NULL;
    if (weight_data) { weights = malloc(sizeof(double) * n); memcpy(weights, weight_data, sizeof(double) * n); }

    if (opt->color_space == ORC_CIELuv) orc_color_transform(ORC_T_SRGB_TO_CIELUV, colors, n);
    else if (opt->color_space == ORC_ICtCp) orc_color_transform(ORC_T_SRGB_TO_ICTCP, colors, n);
    if (opt->verbose) printf("patolette ======== Palette generation \n");

    orc_dataset d = { colors, colors + n, colors + 2 * n, weights };
    size_t cap = palette_size > 16 ? palette_size : 16;
    orc_cluster **cl = calloc(cap, sizeof *cl);
    size_t count = gq_quantize(&d, n, palette_size, cl);
    if (count == 0) { *exit_code = -1; free(cl); free(colors); free(weights); return; }
    if (opt->verbose) printf("patolette ======== Base cluster count: %zu\n", count);
    lq_quantize(&d, cl, &count, palette_size);

    /* palette: cluster centres (create.c) or KMeans-refined (refine.c:165-221) */
    double *pal = malloc(sizeof(double) * 3 * count);
    for (size_t j = 0; j < count; j++) memcpy(pal + 3 * j, cluster_center(&d, cl[j]), 24);
    if (opt->kmeans_niter > 0) {
        if (opt->verbose) printf("patolette ======== KMeans refinement\n");
        float *xs = malloc(sizeof(float) * 3 * n), *cen = malloc(sizeof(float) * 3 * count), *ws = NULL;
        for (size_t i = 0; i < n; i++) { xs[3 * i] = (float)d.c0[i]; xs[3 * i + 1] = (float)d.c1[i]; xs[3 * i + 2] = (float)d.c2[i]; }
        for (size_t j = 0; j < 3 * count; j++) cen[j] = (float)pal[j];
        if (weights) { ws = malloc(sizeof(float) * n); for (size_t i = 0; i < n; i++) ws[i] = (float)weights[i]; }
        size_t ms = opt->kmeans_max_samples > 65536 ? opt->kmeans_max_samples : 65536; /* refine.c:21,87 */
        orc_kmeans(xs, n, count, cen, ws, opt->kmeans_niter, (int)(ms / count));
        for (size_t j = 0; j < 3 * count; j++) pal[j] = (double)cen[j];
        free(xs); free(cen); free(ws);
    }

    if (!opt->palette_only) {
        if (opt->dither) { /* :268-299 */
            if (opt->verbose) printf("patolette ======== Dithering\n");
            int t = opt->color_space == ORC_CIELuv ? ORC_T_CIELUV_TO_REC2020
                  : opt->color_space == ORC_ICtCp ? ORC_T_ICTCP_TO_REC2020 : ORC_T_SRGB_TO_REC2020;
            orc_color_transform(t, colors, n);
            palette_transform(t, pal, count);
            orc_dither_riemersma(colors, width, height, pal, count, palette_map);
            palette_transform(ORC_T_REC2020_TO_SRGB, pal, count);
        } else { /* :300-324 */
            if (opt->verbose) printf("patolette ======== NN mapping\n");
            if (opt->color_space == ORC_CIELuv) {
                orc_color_transform(ORC_T_CIELUV_TO_REC2020, colors, n);
                palette_transform(ORC_T_CIELUV_TO_REC2020, pal, count);
                orc_color_transform(ORC_T_REC2020_TO_SRGB, colors, n);
                palette_transform(ORC_T_REC2020_TO_SRGB, pal, count);
                orc_color_transform(ORC_T_SRGB_TO_ICTCP, colors, n);
                palette_transform(ORC_T_SRGB_TO_ICTCP, pal, count);
            }
            orc_fill_palette_map_nearest(colors, n, pal, count, palette_map);
            /* :322-323 applied even for ColorSpace_sRGB (bug B1) */
            palette_transform(ORC_T_ICTCP_TO_REC2020, pal, count);
            palette_transform(ORC_T_REC2020_TO_SRGB, pal, count);
        }
    }
    for (size_t j = 0; j < palette_size * 3; j++) palette[j] = -1.0; /* :328-330 */
    for (int c = 0; c < 3; c++)
        for (size_t j = 0; j < count; j++) palette[palette_size * c + j] = pal[3 * j + c];
    for (size_t j = 0; j < count; j++) cluster_free(cl[j]);
    free(cl); free(pal); free(colors); free(weights);
    *exit_code = 0;
}
