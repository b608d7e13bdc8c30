/* Shim <flann/flann.h> (oracle/_ref only).  FLANN is an un-vendored, unpinned
 * dependency of the reference (README.md:44) and is not installed in this image.
 * The reference only relies on FLANN_INDEX_KDTREE_SINGLE with eps = 0, i.e. an
 * EXACT squared-L2 1-NN in double precision (lib/src/palette/nearest.c:72-209),
 * so the shim implements that contract by brute force (flann_shim.c).
 * OUR code; test infrastructure only. */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
enum flann_algorithm_t { FLANN_INDEX_LINEAR = 0, FLANN_INDEX_KDTREE = 1, FLANN_INDEX_KDTREE_SINGLE = 4 };
struct FLANNParameters {
    enum flann_algorithm_t algorithm;
    int checks;
    float eps;
    int sorted;
    int max_neighbors;
    int cores;
    int trees;
    int leaf_max_size;
};
typedef void *flann_index_t;
extern struct FLANNParameters DEFAULT_FLANN_PARAMETERS;
flann_index_t flann_build_index_double(double *dataset, int rows, int cols, float *speedup,
                                       struct FLANNParameters *flann_params);
int flann_find_nearest_neighbors_index_double(flann_index_t index_id, double *testset, int trows,
                                              int *indices, double *dists, int nn,
                                              struct FLANNParameters *flann_params);
int flann_find_nearest_neighbors_double(double *dataset, int rows, int cols, double *testset,
                                        int trows, int *indices, double *dists, int nn,
                                        struct FLANNParameters *flann_params);
int flann_free_index_double(flann_index_t index_id, struct FLANNParameters *flann_params);
#ifdef __cplusplus
}
#endif
