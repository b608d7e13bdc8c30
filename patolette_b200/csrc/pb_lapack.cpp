// pb_lapack.cpp - the 3x3 symmetric eigen solve (reference: lib/src/math/eigen.c:83-140).
//
// The reference calls LAPACK dsyev_('V','L',n=3) and takes the last eigenvector column
// as the principal axis (math/pca.c:136-138).  The SIGN of that eigenvector decides
// which child of a split is "left" (quantize/local.c:375-376) and therefore the order
// of the palette, and LAPACK's sign follows no closed-form rule (it falls out of the
// dsytd2 -> dorgtr -> dsteqr sequence).
//
// Default (row N2): pb_dsyev3.h - that sequence restated operation for operation for n = 3,
// bit-identical to the dsyev_ of the LAPACK the reference build links (tests/test_eigen.py:
// millions of matrices, 0 mismatches).  No run-time dependency, nothing to resolve, no
// fallback whose signs could differ; the same header runs on the device (pb_eigen.cu).
//
// patolette_b200_set_option("host_lapack", 1) restores the round-1 behaviour - call a real
// dsyev_ resolved at run time:
//   1. $PATOLETTE_B200_LAPACK (path to a shared object), or the path handed to
//      patolette_b200_set_lapack() by the Python wrapper (scipy's bundled OpenBLAS);
//   2. the usual system sonames.
// In that mode a missing LAPACK FAILS CLOSED: pb_eigen_solve3 returns false and
// patolette() ends with exit code -1, unless patolette_b200_set_option("allow_jacobi", 1)
// accepts a cyclic-Jacobi stand-in whose eigenvector signs, hence palette ORDER, may differ.
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <string>

#include "pb_dsyev3.h"
#include "pb_host.h"

namespace {

typedef void (*dsyev_fn)(const char *, const char *, const int *, double *, const int *, double *,
                         double *, const int *, int *, size_t, size_t);

std::mutex g_mu;
dsyev_fn g_dsyev = nullptr;
bool g_tried = false;
std::string g_user_path;
std::string g_source = "unresolved";
bool g_allow_jacobi = false;
std::atomic<bool> g_host_lapack{false}; // false: the built-in restatement (pb_dsyev3.h)

dsyev_fn try_open(const char *path) {
    void *h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) return nullptr;
    static const char *names[] = {"dsyev_", "scipy_dsyev_", "dsyev"};
    for (const char *nm : names) {
        void *s = dlsym(h, nm);
        if (s) {
            g_source = std::string(path) + ":" + nm;
            return (dsyev_fn)s;
        }
    }
    dlclose(h);
    return nullptr;
}

void resolve_locked() {
    if (g_tried) return;
    g_tried = true;
    if (!g_user_path.empty()) g_dsyev = try_open(g_user_path.c_str());
    const char *env = getenv("PATOLETTE_B200_LAPACK");
    if (!g_dsyev && env && *env) g_dsyev = try_open(env);
    static const char *sonames[] = {"libopenblas.so.0", "libopenblas.so", "liblapack.so.3", "liblapack.so",
                                    "libmkl_rt.so"};
    for (const char *so : sonames)
        if (!g_dsyev) g_dsyev = try_open(so);
    if (!g_dsyev) {
        g_source = "builtin-jacobi";
        fprintf(stderr,
                "patolette_b200: no LAPACK dsyev_ found (set PATOLETTE_B200_LAPACK or call patolette_b200_set_lapack); "
                "%s\n", g_allow_jacobi ? "allow_jacobi is set: using the built-in Jacobi solver - eigenvector signs and "
                                         "hence palette order may differ from the reference"
                                       : "failing the call (exit code -1); patolette_b200_set_option(\"allow_jacobi\", 1) "
                                         "accepts a built-in solver whose palette order may differ from the reference");
    }
}

// Cyclic Jacobi for a symmetric 3x3 (fallback only).  Eigenvalues ascending, vectors in columns.
void jacobi3(double a[9], double w[3]) {
    double A[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) A[r][c] = r >= c ? a[c * 3 + r] : a[r * 3 + c];
    for (int sweep = 0; sweep < 64; sweep++) {
        double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        if (off < 1e-300) break;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                if (A[p][q] == 0.0) continue;
                double theta = (A[q][q] - A[p][p]) / (2 * A[p][q]);
                double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
                double c = 1 / sqrt(t * t + 1), s = t * c;
                for (int k = 0; k < 3; k++) {
                    double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; k++) {
                    double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; k++) {
                    double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq;
                }
            }
    }
    int order[3] = {0, 1, 2};
    for (int i = 0; i < 3; i++)
        for (int j = i + 1; j < 3; j++)
            if (A[order[j]][order[j]] < A[order[i]][order[i]]) { int t = order[i]; order[i] = order[j]; order[j] = t; }
    for (int c = 0; c < 3; c++) {
        w[c] = A[order[c]][order[c]];
        for (int r = 0; r < 3; r++) a[c * 3 + r] = V[r][order[c]];
    }
}

} // namespace

void pb_lapack_set_path(const char *path) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_user_path = path ? path : "";
    g_tried = false;
    g_dsyev = nullptr;
}

void pb_lapack_allow_jacobi(bool on) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_allow_jacobi = on;
    g_tried = false; // the warning text depends on it
    g_dsyev = nullptr;
}

void pb_lapack_use_host(bool on) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_host_lapack = on;
}

const char *pb_lapack_source() {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_host_lapack) return "builtin-dsyev3 (LAPACK 3.12.0 dsyev restated for n = 3, pb_dsyev3.h)";
    resolve_locked();
    return g_source.c_str();
}

// a: 3x3 column-major, lower triangle significant.  On success the columns of a are the
// eigenvectors for ascending eigenvalues w.  Returns false when the workspace QUERY fails
// (the only failure the reference observes, eigen.c:115-118).
bool pb_eigen_solve3(double a[9], double w[3]) {
    if (!g_host_lapack.load(std::memory_order_relaxed)) {
        // eigen.c:115-140: only a failing workspace query makes the reference give up; the solve's own info is
        // ignored there, and so it is here
        (void)pb_eig::dsyev3(a, w);
        return true;
    }
    dsyev_fn fn;
    bool jacobi;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        resolve_locked();
        fn = g_dsyev;
        jacobi = g_allow_jacobi;
    }
    if (!fn) {
        if (!jacobi) return false; // fail closed: the caller ends with exit code -1
        jacobi3(a, w);
        return true;
    }
    char jobz = 'V', uplo = 'L';
    int n = 3, lda = 3, lwork = -1, info = 0;
    double query[1] = {0};
    fn(&jobz, &uplo, &n, a, &lda, w, query, &lwork, &info, 1, 1);
    if (info != 0) return false;
    lwork = (int)query[0];
    double work[512];
    double *wk = lwork <= 512 ? work : (double *)malloc(sizeof(double) * (size_t)lwork);
    fn(&jobz, &uplo, &n, a, &lda, w, wk, &lwork, &info, 1, 1);
    if (wk != work) free(wk);
    return true;
}
