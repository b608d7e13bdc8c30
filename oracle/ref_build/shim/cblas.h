/* Shim <cblas.h> for compiling the reference in place (oracle/_ref only).
 * OUR code, not the reference's: declares just what lib/src/quantize/sort.c:43
 * uses and redirects it to the OpenBLAS bundled with scipy (symbols carry a
 * scipy_ prefix there).  Test infrastructure - never linked into the product. */
#pragma once
#include <stddef.h>
typedef int blasint;
enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
void scipy_cblas_dgemv(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE trans,
                       blasint m, blasint n, double alpha, const double *a,
                       blasint lda, const double *x, blasint incx, double beta,
                       double *y, blasint incy);
#define cblas_dgemv scipy_cblas_dgemv
