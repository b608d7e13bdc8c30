// pb_pipeline.h - stage entry points implemented outside pb_pipeline.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include <vector>

// KMeans refinement of the palette (pb_kmeans.cu).  planes: device f64 colours in the
// quantisation space, original pixel order; d_w: device f64 weights or nullptr.
void pb_kmeans_refine(const double *const planes[3], const double *d_w, size_t n, std::vector<double> &pal_rm,
                      int niter, int max_points_per_centroid, int sm_count, cudaStream_t st, long *launches);
// KMeans on device-resident planar f32 samples (used by the stage test entry point too).
void pb_kmeans_device(const float *d_x0, const float *d_x1, const float *d_x2, const float *d_wf, size_t nx,
                      std::vector<float> &cen, int K, int niter, int sm_count, cudaStream_t st, long *launches);
// Riemersma dither (pb_dither.cu).  planes: device f64 linear-Rec2020 colours; writes d_map.
// shard (image-sharded runs, optional): this rank runs its share of the speculative chains, the choices are
// all-gathered, and only map[out_first, out_first + out_count) is written.
struct PbDitherShard {
    int rank, world;
    size_t out_first, out_count;
};
void pb_dither_riemersma(const double *const planes[3], size_t width, size_t height,
                         const std::vector<double> &pal_rm, unsigned long long *d_map, int sm_count,
                         cudaStream_t st, long *launches, const PbDitherShard *shard = nullptr);
// Saliency weights of the reference's Python wrapper (pb_saliency.cu; patolette.pyx:54-313).  planes: device sRGB
// planes in [0, 1]; d_weights: device, n doubles.  0, -7 (image too small / too elongated for the wrapper's scans and
// border strips: the reference raises there) or -1 (a border strip with a singular covariance).
int pb_saliency_weights(const double *const planes[3], size_t width, size_t height, double tile_size, double *d_weights,
                        int sm_count, cudaStream_t st);
// the minimum-barrier distance map alone (patolette.pyx:153-201), float32, n values
int pb_saliency_mbd(const double *const planes[3], size_t width, size_t height, float *d_out, int sm_count, cudaStream_t st);
// test knob: candidate-list nearest-neighbour search inside the dither (default on)
void pb_dither_set_grid(bool on);
// test knob: 4 lanes per speculative chain (default) or one warp per chain
void pb_dither_set_subwarp(bool on);
// test knobs: tile-wise permutation kernels / one-wave segment sizing (both default on)
void pb_dither_set_tiles(bool on);
void pb_dither_set_one_wave(bool on);

// Largest palette the f32 KMeans slice supports (16-bit assignments; per-warp class counters of the stable
// sort in shared memory, 4 B x K for one warp).  Above it patolette() returns exit code -6 when
// kmeans_niter > 0 instead of silently skipping the refinement.
#define PB_KMEANS_MAX_K 50000
// Dynamic shared memory a kernel may ask for (227 KB opt-in limit minus static shared memory and slack):
// palettes that do not fit are read from global memory instead (same arithmetic, L1/L2-served broadcasts).
#define PB_SMEM_PALETTE_LIMIT (200 * 1024)
// drops the cached walk-position table of the dither (pb_dither.cu)
void pb_dither_release_cache();
