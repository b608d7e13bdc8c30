/* Exact brute-force stand-in for the four FLANN C entry points the reference
 * calls (nearest.c:104,137,189; riemersma.c:350).  Squared L2 in double, the
 * three dimensions summed in order 0,1,2 with separately rounded products,
 * strict '<' over ascending palette index (lowest index wins ties - FLANN's own
 * tie order is kd-tree traversal order and is unspecified).
 * OUR code; test infrastructure only. */
#include "flann/flann.h"
#include <stdlib.h>
#include <string.h>

struct FLANNParameters DEFAULT_FLANN_PARAMETERS = { FLANN_INDEX_KDTREE, 32, 0.0f, 1, -1, 0, 4, 4 };

typedef struct { double *pts; int rows, cols; } shim_index;

static int nearest_one(const double *pts, int rows, int cols, const double *q, double *dist_out) {
    int best = 0;
    double bestd = 0;
    for (int j = 0; j < rows; j++) {
        double d = 0;
        for (int c = 0; c < cols; c++) {
            double diff = q[c] - pts[(size_t)j * cols + c];
            d += diff * diff;
        }
        if (j == 0 || d < bestd) { bestd = d; best = j; }
    }
    if (dist_out) *dist_out = bestd;
    return best;
}

flann_index_t flann_build_index_double(double *dataset, int rows, int cols, float *speedup,
                                       struct FLANNParameters *p) {
    (void)p;
    shim_index *ix = malloc(sizeof *ix);
    ix->pts = dataset; /* FLANN keeps a reference to the caller's matrix too */
    ix->rows = rows; ix->cols = cols;
    if (speedup) *speedup = 1.0f;
    return ix;
}

int flann_find_nearest_neighbors_index_double(flann_index_t index_id, double *testset, int trows,
                                              int *indices, double *dists, int nn,
                                              struct FLANNParameters *p) {
    (void)p; (void)nn;
    shim_index *ix = index_id;
    for (int i = 0; i < trows; i++)
        indices[i] = nearest_one(ix->pts, ix->rows, ix->cols, testset + (size_t)i * ix->cols,
                                 dists ? dists + i : NULL);
    return 0;
}

int flann_find_nearest_neighbors_double(double *dataset, int rows, int cols, double *testset,
                                        int trows, int *indices, double *dists, int nn,
                                        struct FLANNParameters *p) {
    (void)p; (void)nn;
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)trows; i++)
        indices[i] = nearest_one(dataset, rows, cols, testset + (size_t)i * cols,
                                 dists ? dists + i : NULL);
    return 0;
}

int flann_free_index_double(flann_index_t index_id, struct FLANNParameters *p) {
    (void)p;
    free(index_id);
    return 0;
}
