// pb_nngrid.cuh - geometry of the candidate-list grid (pb_nngrid.cu), shared with the dither kernels.
#pragma once
#include "pb_common.cuh"

// cells per dimension: 32 for palettes up to 512 entries (lists of ~10 candidates), 16 above (the [cell][K] list
// table grows with ng^3 * K)
static inline int pb_grid_ng(int K) { return K <= 512 ? 32 : 16; }

struct PbGridHdr {
    unsigned long long mn[3], mx[3]; // order-encoded extrema of the pixel planes (k_nn_bbox)
    double scale[3];                 // the queries live in scale * pixel space (dither: sqrt-luma weights)
    double expand;                   // ... and may leave the pixels' box by this fraction of its range
    int ng;                          // cells per dimension
};

struct PbGridGeom {
    double lo[3], w[3], inv[3]; // cell c of dimension d covers [lo + c*w, lo + (c+1)*w]
    double cmax[3];             // largest |coordinate| inside the grid
    int ng;
    bool ok;
};

static __device__ __forceinline__ PbGridGeom pb_grid_geom(const PbGridHdr *h) {
    PbGridGeom g;
    g.ok = true;
    g.ng = h->ng;
    for (int d = 0; d < 3; d++) {
        if (h->mn[d] > h->mx[d]) { g.ok = false; g.lo[d] = g.w[d] = g.inv[d] = g.cmax[d] = 0; continue; } // no finite pixel
        double a = pb_ord_decode(h->mn[d]) * h->scale[d], b = pb_ord_decode(h->mx[d]) * h->scale[d];
        double range = b - a;
        if (!(range > 1e-300)) range = 1e-300; // one colour along this axis
        a -= h->expand * range;
        b += h->expand * range;
        range = b - a;
        if (!(range > 1e-300)) range = 1e-300;
        if (!(range < 1e300) || !(a > -1e300) || !(b < 1e300)) g.ok = false;
        g.lo[d] = a;
        g.cmax[d] = fabs(a) > fabs(b) ? fabs(a) : fabs(b);
        g.w[d] = range / g.ng;
        g.inv[d] = g.ng / range;
    }
    return g;
}

// cell of a query, -1 when it is not inside the grid (the comparisons are false for NaN)
static __device__ __forceinline__ int pb_grid_cell(bool ok, const double *lo, const double *inv, int ng, double x, double y, double z) {
    const double fx = (x - lo[0]) * inv[0], fy = (y - lo[1]) * inv[1], fz = (z - lo[2]) * inv[2], top = (double)ng;
    if (!(ok && fx >= 0.0 && fy >= 0.0 && fz >= 0.0 && fx <= top && fy <= top && fz <= top)) return -1;
    const int ix = min((int)fx, ng - 1), iy = min((int)fy, ng - 1), iz = min((int)fz, ng - 1);
    return (ix * ng + iy) * ng + iz;
}
