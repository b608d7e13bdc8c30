// Host check of patolette_b200/csrc/pb_dsyev3.h against the real LAPACK dsyev_ (scipy's OpenBLAS, the library the
// reference build links): eigenvalues and eigenvectors bit for bit on adversarial families of 3 x 3 covariances.
//   test_dsyev3 <millions>      prints "OK <count>" or the first mismatches
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pb_dsyev3.h"

extern "C" void scipy_dsyev_(const char *, const char *, const int *, double *, const int *, double *, double *, const int *, int *,
                             size_t, size_t);

static uint64_t rs = 0x9E3779B97F4A7C15ull;
static inline uint64_t nxt() { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return rs; }
static inline double u01() { return (double)(nxt() >> 11) * (1.0 / 9007199254740992.0); }
static inline double gauss() { return sqrt(-2.0 * log(u01() + 1e-300)) * cos(6.283185307179586 * u01()); }

static void cov_from_samples(int cnt, int mode, double a[9]) {
    // weighted covariance of cnt colour samples, computed like math/pca.c: mean, then centred products / sum w
    double x[64][3], w[64], sw = 0, m[3] = {0, 0, 0};
    const double ax[3] = {0.57, 0.59, 0.57};
    for (int i = 0; i < cnt; i++) {
        const double t = u01();
        for (int c = 0; c < 3; c++) {
            double v;
            if (mode == 0) v = u01();
            else if (mode == 1) v = floor(u01() * 256.0) / 255.0;                       // 8-bit colours
            else if (mode == 2) v = t * ax[c] + 0.01 * u01();                           // luminance-dominated
            else if (mode == 3) v = floor((t * ax[c] + 0.02 * u01()) * 255.0) / 255.0;   // ... on 8 bits
            else v = t * ax[c];                                                         // exactly rank 1 (up to rounding)
            x[i][c] = v;
        }
        w[i] = (mode & 1) ? 1.0 + floor(u01() * 1000.0) : 1.0;
        sw += w[i];
        for (int c = 0; c < 3; c++) m[c] += w[i] * x[i][c];
    }
    for (int c = 0; c < 3; c++) m[c] /= sw;
    for (int s = 0; s < 3; s++)
        for (int r = s; r < 3; r++) {
            double e = 0;
            for (int i = 0; i < cnt; i++) e += w[i] * (x[i][r] - m[r]) * (x[i][s] - m[s]);
            a[s * 3 + r] = e / sw;
        }
    a[3] = a[6] = a[7] = NAN; // the upper triangle must not be referenced
}

static void make(int family, double a[9]) {
    switch (family) {
    case 0: cov_from_samples(2 + (int)(nxt() % 40), (int)(nxt() % 5), a); break;
    case 1: // generic random symmetric, wide dynamic range
        for (int s = 0; s < 3; s++) for (int r = s; r < 3; r++) a[s * 3 + r] = gauss() * exp(gauss() * 3);
        a[3] = a[6] = a[7] = NAN; break;
    case 2: { // structured zeros / ties / diagonal / repeated eigenvalues
        const double v[6] = {0.0, 1.0, -1.0, 0.5, 0.25, u01()};
        for (int s = 0; s < 3; s++) for (int r = s; r < 3; r++) a[s * 3 + r] = v[nxt() % 6] * ((nxt() & 1) ? 1.0 : 1e-3);
        a[3] = a[6] = a[7] = NAN; break; }
    case 3: { // tiny and huge norms (the dlascl branches of dsyev and dsteqr)
        cov_from_samples(2 + (int)(nxt() % 20), (int)(nxt() % 5), a);
        const double sc[6] = {1e-160, 1e-130, 1e-100, 1e100, 1e150, 1e-300};
        const double f = sc[nxt() % 6];
        for (int s = 0; s < 3; s++) for (int r = s; r < 3; r++) a[s * 3 + r] *= f;
        break; }
    default: { // small integers: exact ties and cancellations
        for (int s = 0; s < 3; s++) for (int r = s; r < 3; r++) a[s * 3 + r] = (double)((int)(nxt() % 7) - 3);
        a[3] = a[6] = a[7] = NAN; break; }
    }
}

int main(int argc, char **argv) {
    const long total = (long)((argc > 1 ? atof(argv[1]) : 1.0) * 1e6);
    long bad = 0, infos = 0;
    for (long t = 0; t < total; t++) {
        double a[9], b[9], c[9], w1[3], w2[3], work[256];
        make((int)(t % 5), a);
        memcpy(b, a, sizeof a);
        memcpy(c, a, sizeof a);
        int n = 3, lda = 3, lwork = 256, info = 0;
        scipy_dsyev_("V", "L", &n, b, &lda, w1, work, &lwork, &info, 1, 1);
        const int info2 = pb_eig::dsyev3(c, w2);
        infos += info != 0;
        if (info != info2 || memcmp(w1, w2, sizeof w1) || memcmp(b, c, sizeof b)) {
            if (bad < 5) {
                printf("MISMATCH family %ld info %d/%d\n a = [%a %a %a; %a %a; %a]\n", t % 5, info, info2, a[0], a[1], a[2], a[4], a[5], a[8]);
                for (int i = 0; i < 3; i++) printf("  w %a | %a\n", w1[i], w2[i]);
                for (int i = 0; i < 9; i++) printf("  z[%d] %a | %a%s\n", i, b[i], c[i], memcmp(&b[i], &c[i], 8) ? "  <--" : "");
            }
            bad++;
        }
    }
    if (bad) { printf("FAILED %ld of %ld\n", bad, total); return 1; }
    printf("nonzero info %ld\nOK %ld\n", infos, total);
    return 0;
}
