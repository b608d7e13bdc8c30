// tools/probe_blas.c - which products do the OpenBLAS kernels fuse at the vector lengths of a 3 x 3 dsyev?
// Every BLAS routine dsytd2 / dorg2r / dlarf call for n = 3 (ddot, daxpy, dnrm2, dscal, dsymv, dsyr2, dgemv, dger on
// lengths 1 and 2) is run on 200 000 random inputs and compared with candidate formulas; the one that matches all of
// them is what patolette_b200/csrc/pb_dsyev3.h restates.  Build (links scipy's OpenBLAS, symbols prefixed scipy_):
//   L=$(python -c "import glob,os,scipy;print(glob.glob(os.path.join(os.path.dirname(scipy.__file__),\"..\",\"scipy.libs\",\"libscipy_openblas-*.so\"))[0])")
//   gcc -O1 -ffp-contract=off tools/probe_blas.c -o /tmp/probe_blas $L -Wl,-rpath,$(dirname $L) -lm && OPENBLAS_NUM_THREADS=1 /tmp/probe_blas
// Result on OpenBLAS 0.3.31.dev (SkylakeX kernels), every line 200000 of 200000 for exactly one candidate:
//   ddot2 fma(x1,y1,x0*y0) | daxpy fma(a,x,y) | dnrm2(1) |x| | dscal plain | dsymv2 "refblas-fma" | dsyr2 "two-axpy fma (y first)"
//   | dgemvT 2x1 fma(c0,v0,c1*v1) | dger 2x1 fma(x,al*y,c)
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
extern double scipy_ddot_(const int*, const double*, const int*, const double*, const int*);
extern void scipy_daxpy_(const int*, const double*, const double*, const int*, double*, const int*);
extern void scipy_dscal_(const int*, const double*, double*, const int*);
extern double scipy_dnrm2_(const int*, const double*, const int*);
extern void scipy_dsymv_(const char*, const int*, const double*, const double*, const int*, const double*, const int*, const double*, double*, const int*, size_t);
extern void scipy_dsyr2_(const char*, const int*, const double*, const double*, const int*, const double*, const int*, double*, const int*, size_t);
extern void scipy_dgemv_(const char*, const int*, const int*, const double*, const double*, const int*, const double*, const int*, const double*, double*, const int*, size_t);
extern void scipy_dger_(const int*, const int*, const double*, const double*, const int*, const double*, const int*, double*, const int*);
static double rnd(void) { return (drand48() - 0.5) * exp((drand48() - 0.5) * 8); }
static int eq(double a, double b) { return memcmp(&a, &b, 8) == 0; }
#define NT 200000
int main() {
    int one = 1, two = 2, three = 3;
    // ddot n=2
    { int c[4] = {0}; for (int t = 0; t < NT; t++) { double x[2] = {rnd(), rnd()}, y[2] = {rnd(), rnd()};
        double r = scipy_ddot_(&two, x, &one, y, &one);
        c[0] += eq(r, x[0]*y[0] + x[1]*y[1]); c[1] += eq(r, fma(x[1], y[1], x[0]*y[0])); c[2] += eq(r, fma(x[0], y[0], x[1]*y[1]));
        c[3] += eq(r, fma(x[1],y[1],fma(x[0],y[0],0.0))); }
      printf("ddot2: plain %d fma(x1y1,+x0y0) %d fma(x0y0,+x1y1) %d chain %d\n", c[0], c[1], c[2], c[3]); }
    // daxpy n=2, n=1
    { int c[2] = {0}, d[2]={0}; for (int t = 0; t < NT; t++) { double x[2] = {rnd(), rnd()}, y[2] = {rnd(), rnd()}, a = rnd(), y0[2] = {y[0], y[1]};
        scipy_daxpy_(&two, &a, x, &one, y, &one);
        c[0] += eq(y[0], y0[0] + a*x[0]) && eq(y[1], y0[1] + a*x[1]); c[1] += eq(y[0], fma(a, x[0], y0[0])) && eq(y[1], fma(a, x[1], y0[1]));
        double z = y0[0]; scipy_daxpy_(&one, &a, x, &one, &z, &one); d[0] += eq(z, y0[0]+a*x[0]); d[1] += eq(z, fma(a,x[0],y0[0])); }
      printf("daxpy2: plain %d fma %d ; daxpy1: plain %d fma %d\n", c[0], c[1], d[0], d[1]); }
    // dnrm2 n=1, n=2
    { int c[3] = {0}; for (int t = 0; t < NT; t++) { double x[2] = {rnd(), rnd()};
        double r1 = scipy_dnrm2_(&one, x, &one), r2 = scipy_dnrm2_(&two, x, &one);
        c[0] += eq(r1, fabs(x[0])); c[1] += eq(r2, sqrt(x[0]*x[0] + x[1]*x[1])); c[2] += eq(r2, (double)sqrtl((long double)x[0]*x[0] + (long double)x[1]*x[1])); }
      printf("dnrm2: n1 abs %d ; n2 plain %d x87 %d\n", c[0], c[1], c[2]); }
    // dscal n=1
    { int c = 0; for (int t = 0; t < NT; t++) { double x = rnd(), a = rnd(), x0 = x; scipy_dscal_(&one, &a, &x, &one); c += eq(x, a*x0); } printf("dscal1: %d\n", c); }
    // dsymv lower n=2: y = alpha*A*x, beta = 0
    { int c[6] = {0}; for (int t = 0; t < NT; t++) { double A[4] = {rnd(), rnd(), rnd()*1e300, rnd()}, x[2] = {rnd(), rnd()}, alpha = rnd(), beta = 0, y[2] = {NAN, NAN};
        A[2] = NAN; // upper part must not be referenced
        scipy_dsymv_("L", &two, &alpha, A, &two, x, &one, &beta, y, &one, 1);
        double a11 = A[0], a21 = A[1], a22 = A[3];
        // reference BLAS: temp1 = alpha*x(j); temp2 = 0; y(j) += temp1*a(j,j); for i>j: y(i) += temp1*a(i,j); temp2 += a(i,j)*x(i); y(j) += alpha*temp2
        { double y0 = 0, y1 = 0, t1 = alpha*x[0], t2 = 0; y0 = y0 + t1*a11; y1 = y1 + t1*a21; t2 = t2 + a21*x[1]; y0 = y0 + alpha*t2;
          t1 = alpha*x[1]; y1 = y1 + t1*a22; c[0] += eq(y[0], y0) && eq(y[1], y1); }
        { // alpha applied at the end: y = alpha*(A x)
          double s0 = a11*x[0] + a21*x[1], s1 = a21*x[0] + a22*x[1]; c[1] += eq(y[0], alpha*s0) && eq(y[1], alpha*s1);
          double f0 = fma(a21, x[1], a11*x[0]), f1 = fma(a22, x[1], a21*x[0]); c[2] += eq(y[0], alpha*f0) && eq(y[1], alpha*f1);
          double g0 = fma(a11, x[0], a21*x[1]), g1 = fma(a21, x[0], a22*x[1]); c[3] += eq(y[0], alpha*g0) && eq(y[1], alpha*g1); }
        { double t1 = alpha*x[0], t1b = alpha*x[1]; double y0 = fma(t1, a11, 0.0), y1 = fma(t1, a21, 0.0); double t2 = fma(a21, x[1], 0.0); y0 = fma(alpha, t2, y0); y1 = fma(t1b, a22, y1);
          c[4] += eq(y[0], y0) && eq(y[1], y1); }
      }
      printf("dsymv2: refblas %d alpha*(plain) %d alpha*fmaA %d alpha*fmaB %d refblas-fma %d\n", c[0], c[1], c[2], c[3], c[4]); }
    // dsyr2 lower n=2 alpha=-1: A -= x y' + y x'
    { int c[4] = {0}; for (int t = 0; t < NT; t++) { double A[4] = {rnd(), rnd(), NAN, rnd()}, A0[4], x[2] = {rnd(), rnd()}, y[2] = {rnd(), rnd()}, alpha = -1;
        memcpy(A0, A, sizeof A); scipy_dsyr2_("L", &two, &alpha, x, &one, y, &one, A, &two, 1);
        // ref BLAS: for j: temp1 = alpha*y(j); temp2 = alpha*x(j); for i>=j: a(i,j) += x(i)*temp1 + y(i)*temp2
        double r[4]; { double t1 = alpha*y[0], t2 = alpha*x[0]; r[0] = A0[0] + (x[0]*t1 + y[0]*t2); r[1] = A0[1] + (x[1]*t1 + y[1]*t2); t1 = alpha*y[1]; t2 = alpha*x[1]; r[3] = A0[3] + (x[1]*t1 + y[1]*t2); }
        c[0] += eq(A[0], r[0]) && eq(A[1], r[1]) && eq(A[3], r[3]);
        // openblas: axpy(x * alpha*y[j]) then axpy(y * alpha*x[j])
        { double t1 = alpha*y[0], t2 = alpha*x[0]; r[0] = (A0[0] + x[0]*t1) + y[0]*t2; r[1] = (A0[1] + x[1]*t1) + y[1]*t2; t1 = alpha*y[1]; t2 = alpha*x[1]; r[3] = (A0[3] + x[1]*t1) + y[1]*t2; }
        c[1] += eq(A[0], r[0]) && eq(A[1], r[1]) && eq(A[3], r[3]);
        { double t1 = alpha*y[0], t2 = alpha*x[0]; r[0] = fma(y[0], t2, fma(x[0], t1, A0[0])); r[1] = fma(y[1], t2, fma(x[1], t1, A0[1])); t1 = alpha*y[1]; t2 = alpha*x[1]; r[3] = fma(y[1], t2, fma(x[1], t1, A0[3])); }
        c[2] += eq(A[0], r[0]) && eq(A[1], r[1]) && eq(A[3], r[3]);
        { double t1 = alpha*x[0], t2 = alpha*y[0]; r[0] = fma(x[0], t2, fma(y[0], t1, A0[0])); r[1] = fma(x[1], t2, fma(y[1], t1, A0[1])); t1 = alpha*x[1]; t2 = alpha*y[1]; r[3] = fma(x[1], t2, fma(y[1], t1, A0[3])); }
        c[3] += eq(A[0], r[0]) && eq(A[1], r[1]) && eq(A[3], r[3]); }
      printf("dsyr2: refblas %d two-axpy plain %d two-axpy fma (x first) %d (y first) %d\n", c[0], c[1], c[2], c[3]); }
    // dgemv T 2x1: w = C' v
    { int c[3] = {0}; for (int t = 0; t < NT; t++) { double Cm[2] = {rnd(), rnd()}, v[2] = {rnd(), rnd()}, al = 1, be = 0, w = NAN; int ldc = 3;
        scipy_dgemv_("T", &two, &one, &al, Cm, &ldc, v, &one, &be, &w, &one, 1);
        c[0] += eq(w, Cm[0]*v[0] + Cm[1]*v[1]); c[1] += eq(w, fma(Cm[1], v[1], Cm[0]*v[0])); c[2] += eq(w, fma(Cm[0], v[0], Cm[1]*v[1])); }
      printf("dgemvT 2x1: plain %d fma(c1v1,+) %d fma(c0v0,+) %d\n", c[0], c[1], c[2]); }
    // dger 2x1: C += alpha x y'
    { int c[3] = {0}; for (int t = 0; t < NT; t++) { double Cm[2] = {rnd(), rnd()}, C0[2], x[2] = {rnd(), rnd()}, y = rnd(), al = rnd(); int ldc = 3; memcpy(C0, Cm, sizeof Cm);
        scipy_dger_(&two, &one, &al, x, &one, &y, &one, Cm, &ldc);
        double tt = al*y; c[0] += eq(Cm[0], C0[0] + x[0]*tt) && eq(Cm[1], C0[1] + x[1]*tt); c[1] += eq(Cm[0], fma(x[0], tt, C0[0])) && eq(Cm[1], fma(x[1], tt, C0[1]));
        c[2] += eq(Cm[0], fma(al*x[0], y, C0[0])) && eq(Cm[1], fma(al*x[1], y, C0[1])); }
      printf("dger 2x1: plain %d fma(x,al*y) %d fma(al*x,y) %d\n", c[0], c[1], c[2]); }
    return 0;
}
