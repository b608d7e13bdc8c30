/*
 * patolette_b200.h - C ABI of libpatolette_b200.so (B200 / sm_100a).
 *
 * Part 1 is the reference's public interface, symbol for symbol: a caller that
 * binds lib/include/patolette.h of big-nacho/patolette can load this library
 * instead and get the same results, computed on the GPU.
 *
 * Part 2 (patolette_b200_*) is our extension surface: device selection, stage
 * entry points used by the parity tests, timing read-back for bench.py.  None of
 * the signatures carries a torch / CUDA type: plain pointers and sizes only.
 */
#pragma once
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#if defined(PATOLETTE_B200_BUILD)
#define PB200_API __attribute__((visibility("default")))
#else
#define PB200_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ---- Part 1: reference ABI ------------------------------------------------ */

/* replaces lib/include/patolette.h:7-11 */
typedef enum patolette__ColorSpace {
    patolette__sRGB,
    patolette__CIELuv,
    patolette__ICtCp
} patolette__ColorSpace;

/* replaces lib/include/patolette.h:13-20 (x86-64: offsets 0,1,4,8,16,24; sizeof 32) */
typedef struct patolette__QuantizationOptions {
    bool dither;
    bool palette_only;
    patolette__ColorSpace color_space;
    int kmeans_niter;
    size_t kmeans_max_samples;
    bool verbose;
} patolette__QuantizationOptions;

/* replaces lib/include/patolette.h:22-32 / lib/src/patolette.c:157-343.
 * data: width*height x 3 f64, column-major (all R, all G, all B), sRGB in [0,1],
 * row-major pixel scan order.  weights: width*height f64 (each >= 1) or NULL.
 * palette: caller-allocated palette_size x 3 f64 column-major; unused rows = -1.
 * palette_map: caller-allocated width*height size_t (may be NULL iff palette_only).
 * exit_code: 0 ok, -1 internal, -2 bad dims, -3 bad palette size, -4 too big;
 * additionally -5 = CUDA failure (no GPU / out of memory) and -6 = palette_size above
 * 50000 together with kmeans_niter > 0 (the reference accepts any size; the f32 KMeans
 * slice here does not), which the reference cannot produce.  -1 also covers "no LAPACK
 * dsyev_ found" (patolette_b200_set_lapack).  Synchronous; inputs are never written;
 * calls from several threads are serialised. */
PB200_API void patolette(size_t width, size_t height, const double *data, const double *weights,
               size_t palette_size, const patolette__QuantizationOptions *options,
               double *palette, size_t *palette_map, int *exit_code);

/* replaces lib/include/patolette.h:34 / lib/src/patolette.c:97-105 */
PB200_API const char *get_patolette_exit_code_info_message(int exit_code);

/* replaces lib/include/patolette.h:35 / lib/src/patolette.c:107-119 (caller frees) */
PB200_API patolette__QuantizationOptions *patolette_create_default_options(void);

/* ---- Part 2: extensions ---------------------------------------------------- */

/* CUDA device used by subsequent calls from this thread's process (default 0). */
PB200_API int patolette_b200_set_device(int device);
/* Number of visible CUDA devices, or a negative cudaError on failure. */
PB200_API int patolette_b200_device_count(void);
/* Shared object that provides LAPACK dsyev_ (the reference links one too, lib/src/math/eigen.c:50) for the
 * "host_lapack" mode (patolette_b200_set_option); NULL restores the default search.  By default no LAPACK is
 * needed: the solve is the built-in restatement (patolette_b200_eigen3). */
PB200_API void patolette_b200_set_lapack(const char *path);
/* The eigen solver in use: "builtin-dsyev3 ..." (default), or in "host_lapack" mode "path:symbol" of the dsyev_ /
 * "builtin-jacobi". */
PB200_API const char *patolette_b200_lapack_source(void);

/* Stage: one of the reference's matrix colour transforms in place on a host
 * planar n x 3 array (lib/src/color/: 0 sRGB->ICtCp, 1 sRGB->CIELuv, 2 ICtCp->Rec2020,
 * 3 CIELuv->Rec2020, 4 sRGB->Rec2020, 5 Rec2020->sRGB, 6 CIELuv->ICtCp via sRGB). */
PB200_API int patolette_b200_color_transform(int which, double *planar, size_t n);
/* Stage: out[i] = pow(x[i], y) with the glibc-exact device pow. */
PB200_API int patolette_b200_pow(const double *x, double y, double *out, size_t n);
/* Stage: GQ + LQ (lib/src/quantize/global.c:388, local.c:318) on colours that are
 * already in the quantisation space.  labels[i] = palette slot of pixel i,
 * centers = count x 3 row-major cluster centres (lib/src/palette/create.c:11-33). */
PB200_API int patolette_b200_quantize_clusters(const double *planar, size_t n, const double *weights,
                                     size_t palette_size, uint32_t *labels, double *centers,
                                     size_t *count, size_t *gq_count);
/* Stage: exact nearest-palette map (lib/src/palette/nearest.c:150-209).
 * palette_rm: K x 3 row-major. */
PB200_API int patolette_b200_nearest(const double *planar, size_t n, const double *palette_rm, size_t K,
                           size_t *map);
/* Stage: weighted KMeans refinement (lib/src/palette/refine.c:56-100 + faiss
 * Clustering.cpp:587).  x: n x 3 row-major f32; centers: K x 3 f32 in/out. */
PB200_API int patolette_b200_kmeans(const float *x, size_t n, size_t K, float *centers, const float *w,
                          int niter, int max_points_per_centroid);
/* Stage: Riemersma dither (lib/src/dither/riemersma.c:437) on linear-Rec2020 colours. */
PB200_API int patolette_b200_dither(const double *planar, size_t width, size_t height,
                          const double *palette_rm, size_t K, size_t *map);

/* patolette() for colours stored N x 3 ROW-major (interleaved RGB - numpy's default layout); the
 * transposition the reference's wrapper does on the host (patolette.pyx:388-391) happens on the GPU. */
PB200_API void patolette_b200_interleaved(size_t width, size_t height, const double *rgb, const double *weights,
                                          size_t palette_size, const patolette__QuantizationOptions *options,
                                          double *palette, size_t *palette_map, int *exit_code);
/* Same pipeline with DEVICE-resident I/O: d_data / d_weights / d_palette_map are CUDA device
 * pointers on the current device (layouts as for patolette()); palette stays a host pointer.
 * Inputs are copied device-to-device first (never written), as patolette.c:187-199 copies. */
PB200_API void patolette_b200_device(size_t width, size_t height, const double *d_data, const double *d_weights,
                                     size_t palette_size, const patolette__QuantizationOptions *options,
                                     double *palette, size_t *d_palette_map, int *exit_code);
/* N1 (uint8 ingest): the workflow of README.md:150-158 starts from an 8-bit image, converts to f64 and divides by
 * 255 on the host before calling quantize().  Here rgb is N x 3 row-major uint8 (PIL / numpy layout); value / 255
 * is evaluated on the device in f64 (the same IEEE division), so palette and map equal what patolette() returns
 * for the image `rgb / 255.0`.  palette_map element size: map_bytes = 1 (palette_size <= 256), 2 (<= 65536) or 8
 * (size_t).  device_io != 0: rgb / weights / palette_map are device pointers.  Exit codes as patolette(); a
 * map_bytes that cannot hold palette_size - 1 gives -3. */
PB200_API void patolette_b200_u8(size_t width, size_t height, const uint8_t *rgb, const double *weights,
                                 size_t palette_size, const patolette__QuantizationOptions *options, double *palette,
                                 void *palette_map, int map_bytes, int device_io, int *exit_code);

/* N3 (saliency weights): quantize() of the reference's Python wrapper with tile_size > 0
 * (src/patolette/patolette.pyx:332-466; weights from get_weights(), :203-313: three minimum-barrier-distance raster
 * scans of the channel mean, four border-strip Mahalanobis maps in CIELab, a centre prior and a sigmoid).  The
 * weights are computed on the device from the sRGB input and feed the pipeline without leaving it.
 * colors: in_fmt 0 = three f64 planes (column-major N x 3, as patolette()), 1 = N x 3 row-major f64, 2 = N x 3
 * row-major uint8 (/ 255 on the device).  tile_size = 0 runs unweighted.  map_bytes / device_io as for
 * patolette_b200_u8.  Extra exit code -7: the image is too small (a side <= 3) or too elongated for the wrapper's
 * scans and border strips (the reference raises a Python exception there).
 * The distance map is the reference's bit for bit (float32 min / max / subtract only); the Lab / Mahalanobis /
 * sigmoid chain runs in f64 with CUDA's pow / cbrt / exp and agrees with numpy + scikit-image to ~1e-13 relative,
 * not to the last bit (numpy's SIMD pow and the BLAS matmul inside rgb2lab are not reproducible from here). */
PB200_API void patolette_b200_quantize(size_t width, size_t height, const void *colors, int in_fmt, double tile_size,
                                       size_t palette_size, const patolette__QuantizationOptions *options,
                                       double *palette, void *palette_map, int map_bytes, int device_io,
                                       int *exit_code);
/* Stage: get_weights(img, tile_size) (patolette.pyx:203-313).  planar: width*height x 3 f64 column-major sRGB,
 * pixel p = row * width + col; weights: width*height f64 out.  Returns 0, -7 (see above), -1 (a border strip with a
 * singular covariance: np.linalg.inv raises in the reference), or a negative cudaError. */
PB200_API int patolette_b200_saliency_weights(size_t width, size_t height, const double *planar, double tile_size,
                                              double *weights, int device_io);
/* Stage: mbd(mean(img, axis = 2).astype(float32), 3) (patolette.pyx:153-201, :204-205): width*height float32 out. */
PB200_API int patolette_b200_saliency_mbd(size_t width, size_t height, const double *planar, float *distance,
                                          int device_io);
/* Milliseconds the saliency stage of the last call took (CUDA events; inside the "color" slot of last_timings). */
PB200_API double patolette_b200_last_saliency_ms(void);

/* N2 (on-chip eigen solve): n symmetric 3 x 3 eigen problems solved exactly as LAPACK's dsyev('V', 'L', 3) solves
 * them (lib/src/math/eigen.c:83-140) - dsytd2 -> dorgtr -> dsteqr restated operation for operation in
 * csrc/pb_dsyev3.h, bit-identical eigenvalues AND eigenvectors (sign included).  a9: n matrices, column-major, lower
 * triangle significant; w3: n x 3 ascending eigenvalues; z9: n x 9 eigenvectors in columns (principal axis = last
 * column, math/pca.c:136-138); info: n LAPACK info values or NULL.  on_device = 0 runs the host instantiation (no
 * GPU needed), 1 the device kernel (k_eigen3).  The pipeline itself uses this solver by default;
 * patolette_b200_set_option("host_lapack", 1) switches back to a run-time resolved dsyev_. */
PB200_API int patolette_b200_eigen3(const double *a9, size_t n, double *w3, double *z9, int *info, int on_device);

/* Image-sharded multi-GPU runs (one process per GPU).  The library owns an NCCL communicator:
 * rank 0 calls patolette_b200_comm_unique_id() and hands the 128 bytes to every rank (any transport), then all
 * ranks call patolette_b200_comm_init() (collective; after patolette_b200_set_device()).  world = 1 drops it.
 * patolette_b200_sharded() is then a collective call: rank r passes pixels [first, first + count) of the image
 * (patolette_b200_shard_range; planes of `count` doubles, column-major count x 3) and receives the map of the same
 * pixels and the whole palette.  Colour planes are all-gathered over NVLink, the split loop is sharded by cluster
 * with one device-side all-gather of 240 B per evaluated cluster and batch; results are bit-identical for every
 * world size.  comm_info returns 1 when a communicator with world > 1 is active. */
PB200_API int patolette_b200_comm_unique_id(char *id128);
PB200_API int patolette_b200_comm_init(int rank, int world, const char *id128);
PB200_API void patolette_b200_comm_destroy(void);
PB200_API int patolette_b200_comm_info(int *rank, int *world, int *nccl_version);
PB200_API int patolette_b200_shard_range(size_t n_pixels, int rank, int world, size_t *first, size_t *count);
PB200_API void patolette_b200_sharded(size_t width, size_t height, const double *slice, const double *weights_slice,
                                      size_t palette_size, const patolette__QuantizationOptions *options,
                                      double *palette, size_t *map_slice, int device_io, int *exit_code);

/* Working buffers (~100 B per pixel) are cached between calls; this returns them to the driver
 * (and reports how many bytes were held). */
PB200_API size_t patolette_b200_release_cache(void);
/* Run subsequent calls on the caller's CUDA stream (a cudaStream_t passed as void*; enable = 0
 * restores the library's private stream).  Lets a harness bracket calls with its own events. */
PB200_API int patolette_b200_set_stream(void *cuda_stream, int enable);
/* Per-kernel CUDA-event profile: enable(1) resets and starts recording; json() resolves pending
 * events and writes {"kernel": {"launches", "ms", "bytes"}} (returns the size needed). */
PB200_API int patolette_b200_profile_enable(int on);
PB200_API size_t patolette_b200_profile_json(char *buf, size_t cap);
/* Timeline of the profiled launches, one line "kernel stream start_ms end_ms" each, relative to the first launch
 * (call before profile_json, which consumes the records).  With patolette_b200_set_option("prof_timeline", 1) the
 * split loop keeps its two streams while profiling, so the lines show what really overlaps. */
PB200_API size_t patolette_b200_profile_timeline(char *buf, size_t cap);

/* Ordered-sum statistics since the last reset (pb_ordered.cu), 16 counters: (chain, block) pairs
 * accepted from their summary record, pairs replayed, replay reasons {unusable record, state not
 * expressible in the block's unit, interval}, replay rounds, element-wise sub-chunks, pairs accepted
 * through a two-parity record, SM cycles of the resolving warps in {scan walk, two-parity records,
 * replays}, record groups loaded, cycles of the slowest resolving warp, 3 spare. */
PB200_API int patolette_b200_ordered_counts(unsigned long long *out16, int reset);
/* Split-selection statistics since the last reset (pb_certify.cu): {clusters whose optimal bucket was certified from
 * unordered per-bucket sums, certificates refused, clusters re-evaluated through the exact route, spare}. */
PB200_API int patolette_b200_split_counts(unsigned long long *out4, int reset);
/* Host-only stage entry (no GPU needed): the GQ dynamic programme (quantize/global.c:189-298) on 512 x 10
 * per-bucket sums {sum c[3], sum |c|^2, sum c_r*c_s (0,0)(0,1)(1,1)(0,2)(1,2)(2,2)} and 513 class starts.
 * Writes the cut list q[0..cells] into cuts16, returns the cell count (0 on failure). */
PB200_API int patolette_b200_gq_cuts(const double *bucket_sums, const unsigned int *class_start, size_t palette_size,
                                     size_t *cuts16);
/* Chain-sharded multi-GPU runs: every rank (process + GPU) is given the SAME image and calls
 * patolette() at the same time; rank r computes the ordered sums of the chains it owns and the ranks exchange
 * their per-cluster moment rows through `allgather`, which must gather `bytes` bytes from every rank into
 * `recv` (world * bytes, rank-major) - host memory, any transport - and return 0; a non-zero return aborts the
 * running patolette() call on this rank with exit code -1.  Results are identical for every world size.
 * world = 1 (default) switches sharding off.  Returns 0, -1 on bad arguments. */
typedef int (*patolette_b200_allgather_fn)(const void *send, void *recv, size_t bytes, void *user);
PB200_API int patolette_b200_set_sharding(int rank, int world, patolette_b200_allgather_fn allgather, void *user);
/* Test / tuning knobs (never change results, only the route taken): "dump_cap" = cap on the term-dump slots
 * of an ordered-sum pass (-1 default; 0 = every replay recomputes its terms from the planes), "overlap" =
 * two-stream half-batch evaluation of the split loop (-1 default, 0 off, 1 on), "gq_threads" = host threads of the GQ
 * dynamic programme (0 default), "nn_grid" / "dither_grid" = nearest
 * map / the dither's per-step search through per-cell candidate lists (1, default) or brute force (0).
 * "split_certify" = how a split finds its optimal bucket (quantize/local.c:102-177): 1 (default) unordered per-bucket
 * sums plus a proof that the reference's argmax is the same, refused clusters re-evaluated exactly; 0 the exact route
 * for every cluster (bucket sort + sequential per-bucket chains); 2 certified route with every certificate refused.
 * "host_lapack" = 1 solves the 3 x 3 eigen problems with a dsyev_ resolved at run time (patolette_b200_set_lapack)
 * instead of the built-in restatement (default 0; both give the same bits).  In that mode "allow_jacobi" = 1 lets a
 * built-in Jacobi solver stand in when no dsyev_ can be resolved (default 0: such a call fails with exit code -1,
 * because eigenvector signs - hence palette order - would differ from the reference).
 * Returns 0, -1 if unknown. */
PB200_API int patolette_b200_set_option(const char *name, long long value);
/* Debug: per chain of the centred pass {cycles scan walk, cycles record walk, cycles replays, replays,
 * records walked one by one}, 7 x 5 counters. */
PB200_API int patolette_b200_ordered_chain_debug(unsigned long long *out35, int reset);

/* FP64 FMA throughput of the current device, TFLOP/s (a register-resident DFMA kernel timed with CUDA events;
 * the denominator of bench.py's FP64 fractions).  Negative cudaError on failure. */
PB200_API double patolette_b200_fp64_peak(void);

/* Timings of the last patolette() call on this process, milliseconds (CUDA events
 * on the library's stream; h2d/d2h include the host copies).  Keys in order:
 * total, h2d, color, gq, lq, kmeans, nearest, dither, d2h, and the number of
 * kernels launched as the 10th value. */
PB200_API int patolette_b200_last_timings(double *out10);

#ifdef __cplusplus
}
#endif
