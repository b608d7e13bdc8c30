// pb_common.cuh - shared declarations of the patolette_b200 CUDA library.
//
// Compile the whole library with --fmad=false: every fused multiply-add in these
// sources is an explicit __fma_rn().  The reference is a generic x86-64 build
// (no FMA contraction), and bit-exact parity depends on reproducing each
// rounding point.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define PB_BUCKETS 512 /* reference: quantize/global.c:22, quantize/local.c:15 */
#define PB_DELTA 1e-16 /* reference: math/misc.h:5 */

#include "pb_error.h"

// A cluster = a contiguous range of the permuted pixel arrays (ascending original
// pixel index inside the range, which is the order every reference sum runs in).
struct PbSeg {
    uint32_t lo;    // first permuted position
    uint32_t n;     // pixel count
    uint32_t buf;   // which ping-pong buffer (0/1) holds it
    uint32_t tbase; // first scatter tile of this segment in the packed per-batch tile table
    uint32_t bbase; // first ordered-sum block of this segment in the packed per-batch block table
    uint32_t pad;   // in: E with sum(w) < 2^E (weighted certified route); out (children): PB_ROUTE_* of the split that made them
};

// Planar working set: three colour planes, optional weight plane, original index.
struct PbPlanes {
    double *c[3];
    double *w;      // nullptr when unweighted
    uint32_t *idx;  // original pixel index of each permuted position
};

// Per-segment statistics produced by the ordered-sum kernels (device layout).
struct PbStats {
    double wsum;     // sum of weights (or n)
    double mean[3];  // (sum c_j*w) * (1/wsum)          matrix2D.c:200-233
    double cov[6];   // raw sums (j,k) = (0,0)(1,0)(1,1)(2,0)(2,1)(2,2) of (w*c^_j)*c^_k   pca.c:84-97
    double dist;     // sum ((dx^2+dy^2)+dz^2)*w         cluster.c:135-148
    double pad;
};

// Result of evaluating one cluster's split (device layout).
struct PbSplit {
    unsigned long long mn_enc, mx_enc; // order-encoded extrema of the projections
    uint32_t split;      // optimal bucket index          local.c:171
    uint32_t nleft;      // pixels with bucket <= split
    uint32_t degenerate; // max - min < DELTA -> round-robin buckets (sort.c:61-79)
    uint32_t pad;        // PB_ROUTE_*: how `split` was found (copied into the children's PbSeg::pad)
};

// How the optimal bucket of a split was obtained (pb_certify.cu): the exact route reproduces the reference's
// per-bucket sums bit for bit; the certified route forms them in any order and proves that the argmax is the same.
enum { PB_ROUTE_EXACT = 0, PB_ROUTE_CERTIFIED = 1, PB_ROUTE_REFUSED = 2 };

// Per-segment bucket table of the certified route (zeroed before use): sums of w*c_j in ANY order, pixel counts,
// sum floor(w) and the number of pixels whose weight could make size_t += double round up (weighted runs),
// bit patterns of max |c_j| over the segment, and a flag for weights the certificate cannot handle.
struct PbHist {
    double s[3][PB_BUCKETS];
    unsigned long long sz[PB_BUCKETS];
    uint32_t cnt[PB_BUCKETS];
    uint32_t risky[PB_BUCKETS];
    unsigned long long absmax[3];
    uint32_t bad;
    uint32_t pad;
};

static __device__ __forceinline__ double pb_dgemv_row3(double a0, double a1, double a2,
                                                       double x0, double x1, double x2, bool tail_row) {
    // What OpenBLAS' x86-64 dgemv_n (single thread) computes per row of an n x 3 column-major
    // matrix (sort.c:43).  Rows are taken four at a time: the 2-column kernel fuses a0*x0 onto
    // the rounded a1*x1, the 1-column tail adds the rounded a2*x2.  The last (n mod 4) rows go
    // through a scalar loop that the FMA build contracts into a three-deep fma chain.
    if (tail_row) return __fma_rn(a2, x2, __fma_rn(a1, x1, __dmul_rn(a0, x0)));
    return __dadd_rn(__fma_rn(a0, x0, __dmul_rn(a1, x1)), __dmul_rn(a2, x2));
}

// Order-preserving map double -> uint64 (for atomicMin/atomicMax on doubles).
static __device__ __forceinline__ unsigned long long pb_ord_encode(double d) {
    unsigned long long u = (unsigned long long)__double_as_longlong(d);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ULL);
}
static __host__ __device__ __forceinline__ double pb_ord_decode(unsigned long long u) {
    u = (u >> 63) ? (u & 0x7fffffffffffffffULL) : ~u;
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d;
    memcpy(&d, &u, 8);
    return d;
#endif
}
