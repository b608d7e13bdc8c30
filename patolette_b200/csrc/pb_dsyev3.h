// pb_dsyev3.h - row N2: LAPACK's dsyev('V', 'L', n = 3) restated operation for operation, for host AND device.
//
// The reference solves every cluster's 3 x 3 covariance with dsyev_ (lib/src/math/eigen.c:83-140) and takes the last
// eigenvector as the principal axis (math/pca.c:136-138).  Its SIGN decides which child of a split is "left"
// (quantize/local.c:375-376), hence the order of the palette, and its last BIT decides bucket boundaries - and
// neither follows a closed form: they fall out of the dsytd2 -> dorgtr -> dsteqr sequence.  This header follows that
// sequence for n = 3 (LAPACK 3.12.0 as bundled with OpenBLAS 0.3.31 / scipy 1.18, the LAPACK the oracle's reference
// build links): same control flow, same operation order, every rounding in the same place.
//
//   dsyev   -> dlansy('M'), dlascl (out-of-range norms only), dsytrd (n = 3 < crossover: dsytd2), dorgtr, dsteqr, dscal
//   dsytd2  -> dlarfg (dnrm2 of ONE element = |x|, dlapy2, dscal), dsymv, ddot, daxpy, dsyr2 on 2 x 2 / length 2
//   dorgtr  -> dorgqr (2 x 2: dorg2r) -> dlarf (iladlc, dgemv 'T' 2 x 1, dger 2 x 1), dscal
//   dsteqr  -> dlanst, dlascl, dlaev2, dlartg (3.10+ la_constants version), dlapy2, dlasr ('R','V','F'|'B'), dswap
//
// The LAPACK routines are Fortran compiled without contraction (no FMA instruction in the shipped objects); the BLAS
// level-1/2 calls land in OpenBLAS's x86-64 AVX2/AVX-512 kernels, whose scalar tails are C compiled WITH contraction.
// Which products are fused was established by probing every routine at the lengths that occur here (tools/probe_blas.c,
// 200 000 random inputs each, one formula matches all of them):
//   ddot(2)      fma(x1, y1, x0 * y0)                      daxpy       fma(a, x_i, y_i)
//   dsymv('L',2) y0 = fma(alpha, a21 * x1, (alpha x0) a11), y1 = fma(alpha x1, a22, (alpha x0) a21)
//   dsyr2('L',2) a_ij = fma(x_i, alpha y_j, fma(y_i, alpha x_j, a_ij))
//   dgemv('T',2x1) fma(c0, v0, c1 * v1)                   dger(2x1)   fma(x_i, alpha * y, c_i)
//   dnrm2(1) = |x|, dscal = plain product.
// tests/test_eigen.py compares this header, compiled for the host, with the real dsyev_ on millions of matrices
// (random, rank-deficient, luminance-dominated, tiny / huge norms, exact ties): eigenvalues and eigenvectors bit for
// bit.  The GPU test runs the same comparison with the device instantiation.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define PB_EIG_HD __host__ __device__ inline
#else
#define PB_EIG_HD inline
#endif

namespace pb_eig {

PB_EIG_HD double fma_(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}
PB_EIG_HD double sign_(double a, double b) { return copysign(a, b); } // Fortran SIGN(a, b)
PB_EIG_HD double max_(double a, double b) { return a > b ? a : b; }   // Fortran MAX / MIN on non-NaN data
PB_EIG_HD double min_(double a, double b) { return a < b ? a : b; }

constexpr double SAFMIN = 2.2250738585072014e-308; // dlamch('S') = 2^-1022
constexpr double EPS_E = 1.1102230246251565e-16;   // dlamch('E') = 2^-53
constexpr double EPS_P = 2.2204460492503131e-16;   // dlamch('P') = 2^-52
constexpr double HUGEVAL = 1.7976931348623157e308; // dlamch('O')

// dlapy2.f (3.7+)
PB_EIG_HD double dlapy2(double x, double y) {
    const bool xn = x != x, yn = y != y;
    double r = 0.0;
    if (xn) r = x;
    if (yn) r = y;
    if (!(xn || yn)) {
        const double xabs = fabs(x), yabs = fabs(y);
        const double w = max_(xabs, yabs), z = min_(xabs, yabs);
        if (z == 0.0 || w > HUGEVAL) r = w;
        else {
            const double q = z / w;
            r = w * sqrt(1.0 + q * q);
        }
    }
    return r;
}

// dlascl.f: the multiplier sequence that takes a quantity from scale cfrom to scale cto without over/underflow;
// mul[] receives up to 3 factors, returns their count (0: nothing to do)
PB_EIG_HD int dlascl_factors(double cfrom, double cto, double mul[4]) {
    const double smlnum = SAFMIN, bignum = 1.0 / smlnum;
    double cfromc = cfrom, ctoc = cto;
    int k = 0;
    for (;;) {
        const double cfrom1 = cfromc * smlnum;
        double m;
        bool done;
        if (cfrom1 == cfromc) { m = ctoc / cfromc; done = true; }
        else {
            const double cto1 = ctoc / bignum;
            if (cto1 == ctoc) { m = ctoc; done = true; cfromc = 1.0; }
            else if (fabs(cfrom1) > fabs(ctoc) && ctoc != 0.0) { m = smlnum; done = false; cfromc = cfrom1; }
            else if (fabs(cto1) > fabs(cfromc)) { m = bignum; done = false; ctoc = cto1; }
            else {
                m = ctoc / cfromc;
                done = true;
                if (m == 1.0) return k;
            }
        }
        mul[k++] = m;
        if (done || k == 4) return k;
    }
}

// dlartg.f90 (3.10+)
PB_EIG_HD void dlartg(double f, double g, double &c, double &s, double &r) {
    const double safmin = SAFMIN, safmax = 1.0 / safmin;
    const double rtmin = sqrt(safmin), rtmax = sqrt(safmax / 2);
    const double f1 = fabs(f), g1 = fabs(g);
    if (g == 0.0) { c = 1.0; s = 0.0; r = f; }
    else if (f == 0.0) { c = 0.0; s = sign_(1.0, g); r = g1; }
    else if (f1 > rtmin && f1 < rtmax && g1 > rtmin && g1 < rtmax) {
        const double d = sqrt(f * f + g * g);
        c = f1 / d;
        r = sign_(d, f);
        s = g / r;
    } else {
        const double u = min_(safmax, max_(safmin, max_(f1, g1)));
        const double fs = f / u, gs = g / u;
        const double d = sqrt(fs * fs + gs * gs);
        c = fabs(fs) / d;
        r = sign_(d, f);
        s = gs / r;
        r = r * u;
    }
}

// dlaev2.f
PB_EIG_HD void dlaev2(double a, double b, double c, double &rt1, double &rt2, double &cs1, double &sn1) {
    const double sm = a + c, df = a - c, adf = fabs(df), tb = b + b, ab = fabs(tb);
    double acmx, acmn, rt;
    if (fabs(a) > fabs(c)) { acmx = a; acmn = c; } else { acmx = c; acmn = a; }
    if (adf > ab) { const double q = ab / adf; rt = adf * sqrt(1.0 + q * q); }
    else if (adf < ab) { const double q = adf / ab; rt = ab * sqrt(1.0 + q * q); }
    else rt = ab * sqrt(2.0);
    int sgn1, sgn2;
    if (sm < 0.0) { rt1 = 0.5 * (sm - rt); sgn1 = -1; rt2 = (acmx / rt1) * acmn - (b / rt1) * b; }
    else if (sm > 0.0) { rt1 = 0.5 * (sm + rt); sgn1 = 1; rt2 = (acmx / rt1) * acmn - (b / rt1) * b; }
    else { rt1 = 0.5 * rt; rt2 = -0.5 * rt; sgn1 = 1; }
    double cs;
    if (df >= 0.0) { cs = df + rt; sgn2 = 1; } else { cs = df - rt; sgn2 = -1; }
    const double acs = fabs(cs);
    if (acs > ab) { const double ct = -tb / cs; sn1 = 1.0 / sqrt(1.0 + ct * ct); cs1 = ct * sn1; }
    else if (ab == 0.0) { cs1 = 1.0; sn1 = 0.0; }
    else { const double tn = -cs / tb; cs1 = 1.0 / sqrt(1.0 + tn * tn); sn1 = tn * cs1; }
    if (sgn1 == sgn2) { const double tn = cs1; cs1 = -sn1; sn1 = tn; }
}

// dlasr('R', 'V', direct, m = 3, n = cnt, c, s, Z(:, first), 3): Z is 3 x 3 column-major, columns first .. first+cnt-1
// (0-based `first`); c[j], s[j] j = 0 .. cnt-2
PB_EIG_HD void dlasr_rv(bool forward, int cnt, const double *c, const double *s, double *z, int first) {
    for (int jj = 0; jj < cnt - 1; jj++) {
        const int j = forward ? jj : cnt - 2 - jj;
        const double ct = c[j], st = s[j];
        if (ct != 1.0 || st != 0.0) {
            double *cj = z + 3 * (first + j), *cj1 = z + 3 * (first + j + 1);
            for (int i = 0; i < 3; i++) {
                const double temp = cj1[i];
                cj1[i] = ct * temp - st * cj[i];
                cj[i] = st * temp + ct * cj[i];
            }
        }
    }
}

// dsteqr('V', 3, d, e, z): d[0..2], e[0..1] (destroyed), z = the orthogonal matrix of dorgtr (3 x 3 column-major).
// Returns LAPACK's info (0 = converged).
PB_EIG_HD int dsteqr3(double *d0, double *e0, double *z) {
    constexpr int n = 3, maxit = 30;
    double *d = d0 - 1, *e = e0 - 1; // 1-based as in the Fortran
    double wc[3] = {0, 0, 0}, ws[3] = {0, 0, 0}; // work(1:n-1), work(n:2n-2), 1-based
    const double eps = EPS_E, eps2 = eps * eps, safmin = SAFMIN, safmax = 1.0 / safmin;
    const double ssfmax = sqrt(safmax) / 3.0, ssfmin = sqrt(safmin) / eps2;
    const int nmaxit = n * maxit;
    int jtot = 0, l1 = 1;
    const int nm1 = n - 1;
    int l = 0, m = 0, lsv = 0, lend = 0, lendsv = 0, iscale = 0;
    double anorm = 0.0;
    double mul[4];
    for (;;) { // label 10
        if (l1 > n) break;
        if (l1 > 1) e[l1 - 1] = 0.0;
        m = n;
        if (l1 <= nm1) {
            for (int mm = l1; mm <= nm1; mm++) {
                const double tst = fabs(e[mm]);
                if (tst == 0.0) { m = mm; break; }
                if (tst <= (sqrt(fabs(d[mm])) * sqrt(fabs(d[mm + 1]))) * eps) { e[mm] = 0.0; m = mm; break; }
            }
        }
        l = l1; lsv = l; lend = m; lendsv = lend; l1 = m + 1;
        if (lend == l) continue;
        // scale the submatrix in rows and columns l .. lend (dlanst 'M')
        anorm = fabs(d[lend]);
        for (int i = l; i <= lend - 1; i++) {
            double sum = fabs(d[i]);
            if (anorm < sum || sum != sum) anorm = sum;
            sum = fabs(e[i]);
            if (anorm < sum || sum != sum) anorm = sum;
        }
        iscale = 0;
        if (anorm == 0.0) continue;
        if (anorm > ssfmax || anorm < ssfmin) {
            iscale = anorm > ssfmax ? 1 : 2;
            const int k = dlascl_factors(anorm, iscale == 1 ? ssfmax : ssfmin, mul);
            for (int q = 0; q < k; q++) {
                for (int i = l; i <= lend; i++) d[i] = d[i] * mul[q];
            }
            for (int q = 0; q < k; q++) {
                for (int i = l; i <= lend - 1; i++) e[i] = e[i] * mul[q];
            }
        }
        // choose between QL and QR iteration
        if (fabs(d[lend]) < fabs(d[l])) { lend = lsv; l = lendsv; }
        if (lend > l) {
            // ---- QL iteration
            for (;;) { // label 40
                m = lend;
                if (l != lend) {
                    for (int mm = l; mm <= lend - 1; mm++) {
                        const double ae = fabs(e[mm]);
                        const double tst = ae * ae;
                        if (tst <= (eps2 * fabs(d[mm])) * fabs(d[mm + 1]) + safmin) { m = mm; break; }
                    }
                }
                if (m < lend) e[m] = 0.0;
                double p = d[l];
                if (m == l) { // label 80: eigenvalue found
                    d[l] = p;
                    l = l + 1;
                    if (l <= lend) continue;
                    break;
                }
                if (m == l + 1) { // 2 x 2 block
                    double rt1, rt2, c, s;
                    dlaev2(d[l], e[l], d[l + 1], rt1, rt2, c, s);
                    wc[l] = c; ws[l] = s;
                    dlasr_rv(false, 2, &wc[l], &ws[l], z, l - 1);
                    d[l] = rt1; d[l + 1] = rt2; e[l] = 0.0;
                    l = l + 2;
                    if (l <= lend) continue;
                    break;
                }
                if (jtot == nmaxit) break;
                jtot = jtot + 1;
                // form shift
                double g = (d[l + 1] - p) / (2.0 * e[l]);
                double r = dlapy2(g, 1.0);
                g = d[m] - p + (e[l] / (g + sign_(r, g)));
                double s = 1.0, c = 1.0;
                p = 0.0;
                for (int i = m - 1; i >= l; i--) {
                    const double f = s * e[i], b = c * e[i];
                    dlartg(g, f, c, s, r);
                    if (i != m - 1) e[i + 1] = r;
                    g = d[i + 1] - p;
                    r = (d[i] - g) * s + 2.0 * c * b;
                    p = s * r;
                    d[i + 1] = g + p;
                    g = c * r - b;
                    wc[i] = c; ws[i] = -s;
                }
                dlasr_rv(false, m - l + 1, &wc[l], &ws[l], z, l - 1);
                d[l] = d[l] - p;
                e[l] = g;
            }
        } else {
            // ---- QR iteration
            for (;;) { // label 90
                m = lend;
                if (l != lend) {
                    for (int mm = l; mm >= lend + 1; mm--) {
                        const double ae = fabs(e[mm - 1]);
                        const double tst = ae * ae;
                        if (tst <= (eps2 * fabs(d[mm])) * fabs(d[mm - 1]) + safmin) { m = mm; break; }
                    }
                }
                if (m > lend) e[m - 1] = 0.0;
                double p = d[l];
                if (m == l) { // label 130
                    d[l] = p;
                    l = l - 1;
                    if (l >= lend) continue;
                    break;
                }
                if (m == l - 1) {
                    double rt1, rt2, c, s;
                    dlaev2(d[l - 1], e[l - 1], d[l], rt1, rt2, c, s);
                    wc[m] = c; ws[m] = s;
                    dlasr_rv(true, 2, &wc[m], &ws[m], z, l - 2);
                    d[l - 1] = rt1; d[l] = rt2; e[l - 1] = 0.0;
                    l = l - 2;
                    if (l >= lend) continue;
                    break;
                }
                if (jtot == nmaxit) break;
                jtot = jtot + 1;
                double g = (d[l - 1] - p) / (2.0 * e[l - 1]);
                double r = dlapy2(g, 1.0);
                g = d[m] - p + (e[l - 1] / (g + sign_(r, g)));
                double s = 1.0, c = 1.0;
                p = 0.0;
                const int lm1 = l - 1;
                for (int i = m; i <= lm1; i++) {
                    const double f = s * e[i], b = c * e[i];
                    dlartg(g, f, c, s, r);
                    if (i != m) e[i - 1] = r;
                    g = d[i] - p;
                    r = (d[i + 1] - g) * s + 2.0 * c * b;
                    p = s * r;
                    d[i] = g + p;
                    g = c * r - b;
                    wc[i] = c; ws[i] = s;
                }
                dlasr_rv(true, l - m + 1, &wc[m], &ws[m], z, m - 1);
                d[l] = d[l] - p;
                e[lm1] = g;
            }
        }
        // label 140: undo scaling
        if (iscale != 0) {
            const int k = dlascl_factors(iscale == 1 ? ssfmax : ssfmin, anorm, mul);
            for (int q = 0; q < k; q++)
                for (int i = lsv; i <= lendsv; i++) d[i] = d[i] * mul[q];
            for (int q = 0; q < k; q++)
                for (int i = lsv; i <= lendsv - 1; i++) e[i] = e[i] * mul[q];
        }
        if (jtot < nmaxit) continue;
        int info = 0; // no convergence after n * maxit iterations
        for (int i = 1; i <= n - 1; i++)
            if (e[i] != 0.0) info++;
        return info;
    }
    // label 160: selection sort (ascending), swapping eigenvector columns
    for (int ii = 2; ii <= n; ii++) {
        const int i = ii - 1;
        int k = i;
        double p = d[i];
        for (int j = ii; j <= n; j++)
            if (d[j] < p) { k = j; p = d[j]; }
        if (k != i) {
            d[k] = d[i];
            d[i] = p;
            for (int r = 0; r < 3; r++) { const double t = z[3 * (i - 1) + r]; z[3 * (i - 1) + r] = z[3 * (k - 1) + r]; z[3 * (k - 1) + r] = t; }
        }
    }
    return 0;
}

// dsyev('V', 'L', 3, a, 3, w, ...): a is 3 x 3 column-major with the lower triangle significant; on return the columns
// of a are the eigenvectors for the ascending eigenvalues w.  Returns LAPACK's info.
PB_EIG_HD int dsyev3(double *a, double *w) {
#define PB_A(i, j) a[((j)-1) * 3 + ((i)-1)]
    // ---- dsyev: scale the matrix to the allowable range if necessary
    const double smlnum = SAFMIN / EPS_P, bignum = 1.0 / smlnum, rmin = sqrt(smlnum), rmax = sqrt(bignum);
    double anrm = 0.0; // dlansy('M', 'L')
    for (int j = 1; j <= 3; j++)
        for (int i = j; i <= 3; i++) {
            const double sum = fabs(PB_A(i, j));
            if (anrm < sum || sum != sum) anrm = sum;
        }
    int iscale = 0;
    double sigma = 1.0;
    if (anrm > 0.0 && anrm < rmin) { iscale = 1; sigma = rmin / anrm; }
    else if (anrm > rmax) { iscale = 1; sigma = rmax / anrm; }
    if (iscale == 1) {
        double mul[4];
        const int k = dlascl_factors(1.0, sigma, mul);
        for (int q = 0; q < k; q++)
            for (int j = 1; j <= 3; j++)
                for (int i = j; i <= 3; i++) PB_A(i, j) = PB_A(i, j) * mul[q];
    }
    double d[3], e[2], tau[2];
    // ---- dsytd2('L'), i = 1: dlarfg(2, a21, a31, 1, taui)
    {
        double alpha = PB_A(2, 1), taui;
        const double xnorm = fabs(PB_A(3, 1)); // dnrm2 of one element
        if (xnorm == 0.0) taui = 0.0;
        else {
            double beta = -sign_(dlapy2(alpha, xnorm), alpha);
            const double safmin = SAFMIN / EPS_E, rsafmn = 1.0 / safmin;
            int knt = 0;
            double x = PB_A(3, 1);
            if (fabs(beta) < safmin) { // xnorm, beta may be inaccurate: scale x and recompute them
                do {
                    knt++;
                    x = x * rsafmn;
                    beta = beta * rsafmn;
                    alpha = alpha * rsafmn;
                } while (fabs(beta) < safmin && knt < 20);
                beta = -sign_(dlapy2(alpha, fabs(x)), alpha);
            }
            taui = (beta - alpha) / beta;
            x = x * (1.0 / (alpha - beta)); // dscal
            for (int j = 0; j < knt; j++) beta = beta * safmin;
            alpha = beta;
            PB_A(3, 1) = x;
        }
        PB_A(2, 1) = alpha;
        e[0] = PB_A(2, 1);
        if (taui != 0.0) {
            PB_A(2, 1) = 1.0;
            const double v1 = 1.0, v2 = PB_A(3, 1);
            // dsymv('L', 2, taui, A(2:3, 2:3), v) -> tau(1:2)
            const double t1 = taui * v1, t1b = taui * v2;
            double y0 = t1 * PB_A(2, 2), y1 = t1 * PB_A(3, 2);
            const double t2 = PB_A(3, 2) * v2;
            y0 = fma_(taui, t2, y0);
            y1 = fma_(t1b, PB_A(3, 3), y1);
            // alpha = -half * taui * ddot(2, tau, v)
            const double dot = fma_(y1, v2, y0 * v1);
            const double al = -0.5 * taui * dot;
            // daxpy(2, alpha, v, tau)
            y0 = fma_(al, v1, y0);
            y1 = fma_(al, v2, y1);
            // dsyr2('L', 2, -1, v, tau, A(2:3, 2:3)): a_ij = fma(x_i, alpha y_j, fma(y_i, alpha x_j, a_ij))
            {
                double tx = -1.0 * v1, ty = -1.0 * y0;
                PB_A(2, 2) = fma_(v1, ty, fma_(y0, tx, PB_A(2, 2)));
                PB_A(3, 2) = fma_(v2, ty, fma_(y1, tx, PB_A(3, 2)));
                tx = -1.0 * v2; ty = -1.0 * y1;
                PB_A(3, 3) = fma_(v2, ty, fma_(y1, tx, PB_A(3, 3)));
            }
            PB_A(2, 1) = e[0];
        }
        d[0] = PB_A(1, 1);
        tau[0] = taui;
    }
    // i = 2: dlarfg(1, ...) gives tau = 0
    e[1] = PB_A(3, 2);
    d[1] = PB_A(2, 2);
    tau[1] = 0.0;
    d[2] = PB_A(3, 3);
    // ---- dorgtr('L'): shift the reflector vectors one column to the right, unit first row / column
    PB_A(1, 3) = 0.0;
    PB_A(1, 2) = 0.0;
    PB_A(3, 2) = PB_A(3, 1);
    PB_A(1, 1) = 1.0;
    PB_A(2, 1) = 0.0;
    PB_A(3, 1) = 0.0;
    // dorg2r(2, 2, 2, B = A(2:3, 2:3), tau): i = 2
    PB_A(3, 3) = 1.0 - tau[1];
    PB_A(2, 3) = 0.0;
    // i = 1
    PB_A(2, 2) = 1.0;
    if (tau[0] != 0.0) { // dlarf('L', 2, 1, v = B(:, 1), tau1, C = B(:, 2))
        const double v1 = PB_A(2, 2), v2 = PB_A(3, 2);
        int lastv = 2;
        if (v2 == 0.0) lastv = 1; // (v1 = 1 is never zero)
        // iladlc(lastv, 1, C): the last non-zero column of C(1:lastv, 1)
        int lastc;
        if (PB_A(2, 3) != 0.0 || (lastv == 2 && PB_A(3, 3) != 0.0)) lastc = 1;
        else lastc = 0;
        // (iladlc looks at C(1, n) and C(lastv, n) first, then scans; for one column the answer is the same)
        if (lastc > 0) {
            double wv;
            if (lastv == 2) wv = fma_(PB_A(2, 3), v1, PB_A(3, 3) * v2); // dgemv('T')
            else wv = PB_A(2, 3) * v1;
            const double tt = -tau[0] * wv; // dger
            PB_A(2, 3) = fma_(v1, tt, PB_A(2, 3));
            if (lastv == 2) PB_A(3, 3) = fma_(v2, tt, PB_A(3, 3));
        }
    }
    PB_A(3, 2) = PB_A(3, 2) * (-tau[0]); // dscal(1, -tau1, B(2, 1))
    PB_A(2, 2) = 1.0 - tau[0];
    // ---- dsteqr
    int info = dsteqr3(d, e, a);
    w[0] = d[0]; w[1] = d[1]; w[2] = d[2];
    if (iscale == 1) {
        const int imax = info == 0 ? 3 : info - 1;
        const double rs = 1.0 / sigma;
        for (int i = 0; i < imax; i++) w[i] = w[i] * rs; // dscal
    }
#undef PB_A
    return info;
}

} // namespace pb_eig
